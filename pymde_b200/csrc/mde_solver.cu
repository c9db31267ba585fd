// mde_solver.cu -- device-resident projected L-BFGS (the MDE.embed solve loop).
//
// Replaces optim.lbfgs (pymde/optim.py:69-184) driving LBFGS.step (pymde/lbfgs.py:390-590),
// _strong_wolfe (pymde/lbfgs.py:44-253), the value_and_grad closure (pymde/optim.py:100-105)
// and the per-iteration callback/statistics (pymde/optim.py:94-96,139-173).
//
// Design (B200-first):
//  * every vector (X, x_init, d, g, g_prev, the S/Y history ring) and every scalar (loss,
//    g.d, Wolfe bracket, Gram matrix of the history, statistics) lives in HBM; the host
//    only enqueues kernels and reads one 32-byte status word per line-search trial
//    (mode 0) or per batch of iterations (mode 1, CUDA graph with a device-side while loop);
//  * the two-loop recursion is done on the (2h+1)^2 Gram matrix ("vector-free" L-BFGS):
//    ONE pass computes all 5h+4 dot products and writes the new (s, y) pair, ONE pass forms
//    d = cg*g + sum cs_j s_j + cy_j y_j (and snapshots g_prev, x_init, g.d, |d|, |X|).
//    The reference does 4h n*m-sized passes and 2h host syncs per iteration;
//  * reductions are two-stage with a fixed order (per-block partials -> one-block finalize),
//    so scalars are bit-reproducible for a given launch shape -- required for the replicated
//    multi-GPU solve where every rank must take the same Wolfe decisions;
//  * reference quirk kept on purpose: the gradient seen by iteration k+1 is
//    the one left by the LAST trial of iteration k's line search, the loss is the ACCEPTED
//    trial's loss.
#include <cstddef>
#include <new>

#include "mde_common.cuh"
#include "mde_logic.h"
#include "mde_project.cuh"

struct mde_edges;
namespace mde {
int distortion_fused(const mde_edges* e, const float* X, int m, float* grad, int* nblocks, cudaStream_t st);
int distortion_fused_flag(const mde_edges* e, const float* X, int m, float* grad, int* nblocks,
                          const int* flag, cudaStream_t st);
const double* loss_partials_ptr(const mde_edges* e);
int64_t edges_n(const mde_edges* e);
int64_t edges_p_total(const mde_edges* e);
}  // namespace mde

using namespace mde;

namespace {

constexpr int kVecThreads = 256;
constexpr int kVecBlocks = kNumSMs * 2;
constexpr int kPairsPerSlice = 8;
constexpr int kDotsPerSlice = 4 + 5 * kPairsPerSlice;  // 44
constexpr int kMaxSlices = kMaxMemory / kPairsPerSlice;  // 4

struct SolverState {
  // ---- status word (first 32 bytes, copied to the host) ----
  int active;      // kernels exit early when 0
  int converged;   // residual test fired (optim.py:165)
  int iter;        // completed iterations
  int error;       // MDE_E_NAN when the reference would raise SolverError
  int need_fresh;  // lbfgs n_iter == 0: evaluate at X before the direction update
  int ls_active;   // line search wants another trial
  int stop_after;  // residual <= eps seen at the start of this iteration
  int pad0;
  // ---- scalars ----
  double eps;
  double loss;     // f at the current iterate (cached loss, lbfgs.py:418-426,550)
  double gg, g1;   // ||g||^2 and ||g||_1 of the gradient buffer
  float gtd;       // g.d
  float dmax;      // max |d|
  double dd, xx;   // ||d||^2, ||X||^2 at iteration start
  double t_last;   // state["t"]
  double t_eval;   // step of the most recent trial evaluation
  long long func_evals;
  int max_stats;
  int world;
  double *avg, *resid, *pct, *steplen;
  LsState ls;
  LbfgsState lb;
};

__device__ __forceinline__ bool off(const int* flag) { return *flag == 0; }

// fixed-order reduction of K sums over nb block partials, into smem out[K]; all threads call
template <bool MAXLAST>
__device__ void reduce_partials(const double* __restrict__ part, int nb, int K, double* out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k = w; k < K; k += nw) {
    double s = 0.0;
    if (MAXLAST && k == K - 1) {
      for (int b = lane; b < nb; b += 32) s = fmax(s, part[(int64_t)b * K + k]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_xor_sync(kFull, s, o));
    } else {
      for (int b = lane; b < nb; b += 32) s += part[(int64_t)b * K + k];
      s = warp_sum(s);
    }
    if (lane == 0) out[k] = s;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------
// P1: candidate pair + all dot products of the history against (y_c, s_c, g)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kVecThreads)
lbfgs_dots_kernel(SolverState* __restrict__ S, const float* __restrict__ g, const float* __restrict__ gprev,
                  const float* __restrict__ d, float* __restrict__ Sb, float* __restrict__ Yb,
                  int64_t npad, double* __restrict__ part) {
  if (off(&S->active) || S->lb.n_iter == 0) return;
  const int slice = blockIdx.y;
  const int count = S->lb.count;
  if (slice > 0 && slice * kPairsPerSlice >= count) return;
  const float t = (float)S->t_last;
  float* sc = Sb + (int64_t)S->lb.cand * npad;
  float* yc = Yb + (int64_t)S->lb.cand * npad;
  const float* sj[kPairsPerSlice];
  const float* yj[kPairsPerSlice];
  bool val[kPairsPerSlice];
#pragma unroll
  for (int j = 0; j < kPairsPerSlice; ++j) {
    int lj = slice * kPairsPerSlice + j;
    val[j] = lj < count;
    int q = val[j] ? S->lb.order[lj] : 0;
    sj[j] = Sb + (int64_t)q * npad;
    yj[j] = Yb + (int64_t)q * npad;
  }
  float acc[kDotsPerSlice];
#pragma unroll
  for (int k = 0; k < kDotsPerSlice; ++k) acc[k] = 0.0f;
  double dacc[kDotsPerSlice];
#pragma unroll
  for (int k = 0; k < kDotsPerSlice; ++k) dacc[k] = 0.0;
  const int64_t n4 = npad >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 G = reinterpret_cast<const float4*>(g)[i];
    float4 P = reinterpret_cast<const float4*>(gprev)[i];
    float4 D = reinterpret_cast<const float4*>(d)[i];
    float gv[4] = {G.x, G.y, G.z, G.w};
    float yv[4] = {G.x - P.x, G.y - P.y, G.z - P.z, G.w - P.w};
    float sv[4] = {D.x * t, D.y * t, D.z * t, D.w * t};
    if (slice == 0) {
      reinterpret_cast<float4*>(yc)[i] = make_float4(yv[0], yv[1], yv[2], yv[3]);
      reinterpret_cast<float4*>(sc)[i] = make_float4(sv[0], sv[1], sv[2], sv[3]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        acc[0] += yv[q] * sv[q]; acc[1] += yv[q] * yv[q];
        acc[2] += sv[q] * gv[q]; acc[3] += yv[q] * gv[q];
      }
    }
#pragma unroll
    for (int j = 0; j < kPairsPerSlice; ++j) {
      if (val[j]) {
        float4 A = reinterpret_cast<const float4*>(sj[j])[i];
        float4 B = reinterpret_cast<const float4*>(yj[j])[i];
        float av[4] = {A.x, A.y, A.z, A.w};
        float bv[4] = {B.x, B.y, B.z, B.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc[4 + 5 * j + 0] += av[q] * yv[q];
          acc[4 + 5 * j + 1] += bv[q] * yv[q];
          acc[4 + 5 * j + 2] += sv[q] * bv[q];
          acc[4 + 5 * j + 3] += av[q] * gv[q];
          acc[4 + 5 * j + 4] += bv[q] * gv[q];
        }
      }
    }
    if (++cnt == 16) {
#pragma unroll
      for (int k = 0; k < kDotsPerSlice; ++k) { dacc[k] += (double)acc[k]; acc[k] = 0.0f; }
      cnt = 0;
    }
  }
#pragma unroll
  for (int k = 0; k < kDotsPerSlice; ++k) dacc[k] += (double)acc[k];
  __shared__ double sm[kDotsPerSlice * 32];
  block_sum<kDotsPerSlice>(dacc, sm);
  if (threadIdx.x == 0) {
    double* o = part + ((int64_t)slice * gridDim.x + blockIdx.x) * kDotsPerSlice;
#pragma unroll
    for (int k = 0; k < kDotsPerSlice; ++k) o[k] = dacc[k];
  }
}

// ---------------------------------------------------------------------------------------
// S1: statistics of the iteration start + history update + two-loop in Gram form
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
direction_scalar_kernel(SolverState* __restrict__ S, const double* __restrict__ part, int nblocks) {
  if (off(&S->active)) return;
  __shared__ double sums[kMaxSlices * kDotsPerSlice];
  __shared__ double sSY[kMaxMemory][kMaxMemory];
  __shared__ double sYY[kMaxMemory][kMaxMemory];
  const int count = S->lb.count;
  const int n_iter = S->lb.n_iter;
  int slices = (count + kPairsPerSlice - 1) / kPairsPerSlice;
  if (slices < 1) slices = 1;
  if (n_iter > 0) {
    for (int s = 0; s < slices; ++s)
      reduce_partials<false>(part + (int64_t)s * nblocks * kDotsPerSlice, nblocks, kDotsPerSlice,
                             sums + s * kDotsPerSlice);
    for (int k = threadIdx.x; k < kMaxMemory * kMaxMemory; k += blockDim.x) {
      (&sSY[0][0])[k] = (&S->lb.SY[0][0])[k];
      (&sYY[0][0])[k] = (&S->lb.YY[0][0])[k];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // callback of LBFGS.step (optim.py:94-96): loss and ||X.grad||_F at the iteration start
    const int it = S->iter;
    const double resid = (double)sqrtf((float)S->gg);
    if (it < S->max_stats) { S->avg[it] = S->loss; S->resid[it] = resid; }
    S->stop_after = (resid <= S->eps) ? 1 : 0;
    double sj_yc[kMaxMemory + 1], yj_yc[kMaxMemory + 1], sc_yj[kMaxMemory + 1], sj_g[kMaxMemory + 1], yj_g[kMaxMemory + 1];
    for (int j = 0; j < count; ++j) {
      const double* b = sums + (j / kPairsPerSlice) * kDotsPerSlice + 4 + 5 * (j % kPairsPerSlice);
      sj_yc[j] = b[0]; yj_yc[j] = b[1]; sc_yj[j] = b[2]; sj_g[j] = b[3]; yj_g[j] = b[4];
    }
    LbfgsState& B = S->lb;
    if (n_iter > 0) lbfgs_direction(B, sSY, sYY, sums[0], sums[1], sums[2], sums[3], sj_yc, yj_yc, sc_yj, sj_g, yj_g);
    else lbfgs_direction(B, sSY, sYY, 0, 0, 0, 0, sj_yc, yj_yc, sc_yj, sj_g, yj_g);
  }
  __syncthreads();
  if (n_iter > 0) {
    for (int k = threadIdx.x; k < kMaxMemory * kMaxMemory; k += blockDim.x) {
      (&S->lb.SY[0][0])[k] = (&sSY[0][0])[k];
      (&S->lb.YY[0][0])[k] = (&sYY[0][0])[k];
    }
  }
}

// ---------------------------------------------------------------------------------------
// P2: d = cg*g + sum_j cs_j S_j + cy_j Y_j ; g_prev = g ; x_init = X ; partial g.d, d.d, X.X, max|d|
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kVecThreads)
direction_apply_kernel(const SolverState* __restrict__ S, const float* __restrict__ g, float* __restrict__ gprev,
                       float* __restrict__ d, const float* __restrict__ X, float* __restrict__ xinit,
                       const float* __restrict__ Sb, const float* __restrict__ Yb, int64_t npad,
                       double* __restrict__ part) {
  if (off(&S->active)) return;
  __shared__ float cs[kMaxMemory], cy[kMaxMemory];
  __shared__ const float* ps[kMaxMemory];
  __shared__ const float* py[kMaxMemory];
  const int count = S->lb.count;
  const float cg = (float)S->lb.cg;
  if (threadIdx.x < count) {
    cs[threadIdx.x] = (float)S->lb.cs[threadIdx.x];
    cy[threadIdx.x] = (float)S->lb.cy[threadIdx.x];
    int q = S->lb.order[threadIdx.x];
    ps[threadIdx.x] = Sb + (int64_t)q * npad;
    py[threadIdx.x] = Yb + (int64_t)q * npad;
  }
  __syncthreads();
  double acc[3] = {0.0, 0.0, 0.0};
  float fa[3] = {0.0f, 0.0f, 0.0f};
  float mx = 0.0f;
  const int64_t n4 = npad >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 G = reinterpret_cast<const float4*>(g)[i];
    float4 Xv = reinterpret_cast<const float4*>(X)[i];
    float r[4] = {cg * G.x, cg * G.y, cg * G.z, cg * G.w};
    for (int j = 0; j < count; ++j) {
      float4 A = reinterpret_cast<const float4*>(ps[j])[i];
      float4 B = reinterpret_cast<const float4*>(py[j])[i];
      r[0] += cs[j] * A.x + cy[j] * B.x; r[1] += cs[j] * A.y + cy[j] * B.y;
      r[2] += cs[j] * A.z + cy[j] * B.z; r[3] += cs[j] * A.w + cy[j] * B.w;
    }
    reinterpret_cast<float4*>(d)[i] = make_float4(r[0], r[1], r[2], r[3]);
    reinterpret_cast<float4*>(gprev)[i] = G;
    reinterpret_cast<float4*>(xinit)[i] = Xv;
    fa[0] += G.x * r[0] + G.y * r[1] + G.z * r[2] + G.w * r[3];
    fa[1] += r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3];
    fa[2] += Xv.x * Xv.x + Xv.y * Xv.y + Xv.z * Xv.z + Xv.w * Xv.w;
    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(r[0]), fabsf(r[1])), fmaxf(fabsf(r[2]), fabsf(r[3]))));
    if (++cnt == 16) {
      for (int k = 0; k < 3; ++k) { acc[k] += (double)fa[k]; fa[k] = 0.0f; }
      cnt = 0;
    }
  }
  for (int k = 0; k < 3; ++k) acc[k] += (double)fa[k];
  __shared__ double sm[3 * 32];
  __shared__ float smx[32];
  block_sum<3>(acc, sm);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) smx[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m2 = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m2 = fmaxf(m2, smx[w]);
    double* o = part + (int64_t)blockIdx.x * 4;
    o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2]; o[3] = (double)m2;
  }
}

// S2: finalize g.d, |d|, |X|; initial step; arm the line search (lbfgs.py:521-549)
__global__ void __launch_bounds__(256)
ls_init_kernel(SolverState* __restrict__ S, const double* __restrict__ part, int nblocks) {
  if (off(&S->active)) return;
  __shared__ double out[4];
  reduce_partials<true>(part, nblocks, 4, out);
  if (threadIdx.x == 0) {
    S->gtd = (float)out[0]; S->dd = out[1]; S->xx = out[2]; S->dmax = (float)out[3];
    double t0 = 1.0;
    if (S->lb.n_iter == 1) {  // t = min(1, 1/||g||_1) * lr
      float inv = 1.0f / (float)S->g1;
      t0 = (inv < 1.0f) ? (double)inv : 1.0;
    }
    ls_begin(S->ls, t0, S->loss, S->gtd, S->dmax);
    S->ls_active = 1;
  }
}

// ---------------------------------------------------------------------------------------
// trial point: X = x_init + t*d  (LBFGS._add_grad, lbfgs.py:350-357); FINAL uses t_accept
// ---------------------------------------------------------------------------------------
template <bool FINAL>
__global__ void __launch_bounds__(kVecThreads)
trial_axpy_kernel(SolverState* __restrict__ S, const float* __restrict__ xinit, const float* __restrict__ d,
                  float* __restrict__ X, int64_t npad) {
  if (off(&S->active)) return;
  if (!FINAL && off(&S->ls_active)) return;
  const float t = FINAL ? (float)S->ls.t_accept : (float)S->ls.t;
  const int64_t n4 = npad >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 A = reinterpret_cast<const float4*>(xinit)[i];
    float4 D = reinterpret_cast<const float4*>(d)[i];
    reinterpret_cast<float4*>(X)[i] = make_float4(fmaf(t, D.x, A.x), fmaf(t, D.y, A.y), fmaf(t, D.z, A.z), fmaf(t, D.w, A.w));
  }
}

__global__ void __launch_bounds__(kVecThreads)
zero_kernel(const int* flag, float* __restrict__ p, int64_t n4) {
  if (off(flag)) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
    reinterpret_cast<float4*>(p)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// anchored constraint (pymde/constraints.py:114-164): overwrite / zero anchor rows
__global__ void anchor_rows_kernel(const int* flag, float* __restrict__ Z, const int64_t* __restrict__ anchors,
                                   const float* __restrict__ values, int64_t na, int m) {
  if (off(flag)) return;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= na * m) return;
  int64_t a = anchors[i / m];
  Z[a * m + (i % m)] = values ? values[i] : 0.0f;
}

// multi-GPU: pack this rank's loss sum behind the gradient as (hi, lo) floats
__global__ void __launch_bounds__(256)
pack_loss_kernel(const int* flag, const double* __restrict__ lpart, int nl, float* __restrict__ tail) {
  if (off(flag)) return;
  __shared__ double out[1];
  reduce_partials<false>(lpart, nl, 1, out);
  if (threadIdx.x == 0) {
    float hi = (float)out[0];
    tail[0] = hi;
    tail[1] = (float)(out[0] - (double)hi);
  }
}

// T5: partial g.d, g.g, |g|_1
__global__ void __launch_bounds__(kVecThreads)
grad_dots_kernel(const int* flag, const float* __restrict__ g, const float* __restrict__ d, int64_t npad,
                 double* __restrict__ part) {
  if (off(flag)) return;
  double acc[3] = {0.0, 0.0, 0.0};
  float fa[3] = {0.0f, 0.0f, 0.0f};
  const int64_t n4 = npad >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 G = reinterpret_cast<const float4*>(g)[i];
    float4 D = reinterpret_cast<const float4*>(d)[i];
    fa[0] += G.x * D.x + G.y * D.y + G.z * D.z + G.w * D.w;
    fa[1] += G.x * G.x + G.y * G.y + G.z * G.z + G.w * G.w;
    fa[2] += fabsf(G.x) + fabsf(G.y) + fabsf(G.z) + fabsf(G.w);
    if (++cnt == 16) {
      for (int k = 0; k < 3; ++k) { acc[k] += (double)fa[k]; fa[k] = 0.0f; }
      cnt = 0;
    }
  }
  for (int k = 0; k < 3; ++k) acc[k] += (double)fa[k];
  __shared__ double sm[3 * 32];
  block_sum<3>(acc, sm);
  if (threadIdx.x == 0) {
    double* o = part + (int64_t)blockIdx.x * 3;
    o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
  }
}

// loss of one evaluation as the reference sees it: fp32 mean, then float(...)
__device__ double eval_loss(const SolverState* S, const double* lpart, int nl, const float* tail, double* smem1,
                            double p_total) {
  double sum;
  if (S->world > 1) {
    sum = (double)tail[0] + (double)tail[1];
    __syncthreads();
  } else {
    reduce_partials<false>(lpart, nl, 1, smem1);
    sum = smem1[0];
  }
  return (double)(float)(sum / p_total);
}

// S(fresh): closure() at the current iterate (lbfgs.py:426), no line search involved
__global__ void __launch_bounds__(256)
fresh_finish_kernel(SolverState* __restrict__ S, const double* __restrict__ lpart, int nl,
                    const float* __restrict__ tail, const double* __restrict__ dpart, int nd, double p_total) {
  if (off(&S->active) || off(&S->need_fresh)) return;
  __shared__ double out[3];
  __shared__ double l1[1];
  double loss = eval_loss(S, lpart, nl, tail, l1, p_total);
  reduce_partials<false>(dpart, nd, 3, out);
  if (threadIdx.x == 0) {
    S->loss = loss; S->gg = out[1]; S->g1 = out[2];
    S->func_evals += 1;
  }
}

// S(trial): feed (f_new, g.d) to the Wolfe state machine; decide the next step or finish
__global__ void __launch_bounds__(256)
ls_update_kernel(SolverState* __restrict__ S, const double* __restrict__ lpart, int nl,
                 const float* __restrict__ tail, const double* __restrict__ dpart, int nd, double p_total) {
  if (off(&S->active) || off(&S->ls_active)) return;
  __shared__ double out[3];
  __shared__ double l1[1];
  double loss = eval_loss(S, lpart, nl, tail, l1, p_total);
  reduce_partials<false>(dpart, nd, 3, out);
  if (threadIdx.x == 0) {
    S->gg = out[1]; S->g1 = out[2];
    S->func_evals += 1;
    S->t_eval = S->ls.t;
    const bool finite = isfinite(out[1]);
    ls_on_result(S->ls, loss, (float)out[0], finite);
    if (S->ls.phase == LS_DONE) {
      S->ls_active = 0;
      if (S->ls.error) { S->error = MDE_E_NAN; S->active = 0; }
      S->t_last = S->ls.t_accept;
      S->loss = (double)(float)S->ls.f_accept;  // _cached_loss is an fp32 tensor (lbfgs.py:550)
    }
  }
}

// S5: end of iteration (optim.py:135-173)
__global__ void iter_end_kernel(SolverState* __restrict__ S) {
  if (off(&S->active)) return;
  if (threadIdx.x != 0) return;
  const int it = S->iter;
  const double h = S->ls.t_accept;
  const double norm_x = (double)sqrtf((float)S->xx);
  const double pc = 100.0 * h * (double)sqrtf((float)S->dd) / norm_x;
  if (it < S->max_stats) { S->pct[it] = (double)(float)pc; S->steplen[it] = h; }
  S->iter = it + 1;
  if (S->stop_after) { S->converged = 1; S->active = 0; }
  else if (h == 0.0) { lbfgs_reset(S->lb, S->lb.memory); S->need_fresh = 1; }  // opt.reset()
  else S->need_fresh = 0;
  if (S->iter >= S->max_stats) S->active = 0;
}

__global__ void init_state_kernel(SolverState* S, double eps, int memory, int max_stats, int world,
                                  double* avg, double* resid, double* pct, double* steplen) {
  if (threadIdx.x != 0) return;
  S->active = 1; S->converged = 0; S->iter = 0; S->error = 0; S->need_fresh = 1; S->ls_active = 0;
  S->stop_after = 0; S->pad0 = 0; S->eps = eps; S->loss = 0.0; S->gg = 0.0; S->g1 = 0.0; S->gtd = 0.0f;
  S->dmax = 0.0f; S->dd = 0.0; S->xx = 0.0; S->t_last = 0.0; S->t_eval = 0.0; S->func_evals = 0;
  S->max_stats = max_stats; S->world = world;
  S->avg = avg; S->resid = resid; S->pct = pct; S->steplen = steplen;
  lbfgs_reset(S->lb, memory);
  ls_begin(S->ls, 0.0, 0.0, 0.0f, 0.0f);
  S->ls.phase = LS_DONE;
}

int vec_blocks(int64_t n4) {
  int64_t nb = (n4 + kVecThreads - 1) / kVecThreads;
  if (nb < 1) nb = 1;
  if (nb > kVecBlocks) nb = kVecBlocks;
  return (int)nb;
}

}  // namespace

struct mde_solver {
  const mde_edges* edges = nullptr;
  int64_t n = 0, N = 0, npad = 0;
  int m = 0;
  mde_solver_opts_t opts{};
  SolverState* S = nullptr;          // device
  int* status_host = nullptr;        // pinned, 8 ints
  float *X = nullptr, *xinit = nullptr, *d = nullptr, *g = nullptr, *gprev = nullptr, *Sb = nullptr, *Yb = nullptr;
  double *dpart = nullptr;           // dot partials
  double *stats = nullptr;           // 4 * max_iter doubles
  void* projws = nullptr;
  ProjWs pw{};
  int nl = 0;                        // loss-partial blocks of the distortion launch
  int nvb = 0;                       // vector-pass blocks
  int host_need_fresh = 1;
  int host_active = 0;
  mde_allreduce_fn allreduce = nullptr;
  void* allreduce_user = nullptr;
  int64_t* anchors = nullptr;
  float* anchor_values = nullptr;
};

namespace {

int read_status(mde_solver* s, cudaStream_t st) {
  MDE_CUDA_TRY(cudaMemcpyAsync(s->status_host, s->S, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
  MDE_CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

// project the iterate onto the constraint set (project_callback, lbfgs.py:368-372)
int enqueue_project(mde_solver* s, cudaStream_t st) {
  const int* act = &s->S->active;
  switch (s->opts.constraint) {
    case MDE_CONSTRAINT_CENTERED: return enqueue_project_centered(s->X, s->n, s->m, s->pw, act, st);
    case MDE_CONSTRAINT_STANDARDIZED: return enqueue_project_standardized(s->X, s->n, s->m, s->pw, act, st);
    case MDE_CONSTRAINT_ANCHORED: {
      int64_t tot = s->opts.n_anchors * s->m;
      if (tot > 0) {
        anchor_rows_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(act, s->X, s->anchors, s->anchor_values,
                                                                    s->opts.n_anchors, s->m);
        MDE_LAUNCH_CHECK();
      }
      return 0;
    }
  }
  return MDE_E_INVALID;
}

// closure: value_and_grad at s->X (optim.py:100-105); `flag` gates the kernels
int enqueue_eval(mde_solver* s, const int* flag, cudaStream_t st) {
  const int64_t n4 = (s->npad + 4) >> 2;  // gradient + (hi, lo) tail
  zero_kernel<<<vec_blocks(n4), kVecThreads, 0, st>>>(flag, s->g, n4);
  MDE_LAUNCH_CHECK();
  int rc = distortion_fused_flag(s->edges, s->X, s->m, s->g, &s->nl, flag, st);
  if (rc) return rc;
  if (s->opts.world_size > 1) {
    pack_loss_kernel<<<1, 256, 0, st>>>(flag, loss_partials_ptr(s->edges), s->nl, s->g + s->npad);
    MDE_LAUNCH_CHECK();
    if (!s->allreduce) return MDE_E_INVALID;
    rc = s->allreduce(s->allreduce_user, s->g, s->npad + 4, (void*)st);
    if (rc) return rc;
  }
  const int* act = flag;
  if (s->opts.constraint == MDE_CONSTRAINT_STANDARDIZED) {
    rc = enqueue_tangent_standardized(s->X, s->g, s->n, s->m, s->pw, act, st);
    if (rc) return rc;
  } else if (s->opts.constraint == MDE_CONSTRAINT_ANCHORED) {
    int64_t tot = s->opts.n_anchors * s->m;
    if (tot > 0) {
      anchor_rows_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(act, s->g, s->anchors, nullptr,
                                                                  s->opts.n_anchors, s->m);
      MDE_LAUNCH_CHECK();
    }
  }
  grad_dots_kernel<<<s->nvb, kVecThreads, 0, st>>>(flag, s->g, s->d, s->npad, s->dpart);
  MDE_LAUNCH_CHECK();
  return 0;
}

int enqueue_fresh(mde_solver* s, cudaStream_t st) {
  int rc = enqueue_eval(s, &s->S->need_fresh, st);
  if (rc) return rc;
  fresh_finish_kernel<<<1, 256, 0, st>>>(s->S, loss_partials_ptr(s->edges), s->nl, s->g + s->npad, s->dpart,
                                         s->nvb, (double)edges_p_total(s->edges));
  MDE_LAUNCH_CHECK();
  return 0;
}

int enqueue_direction(mde_solver* s, cudaStream_t st) {
  int slices = (s->opts.memory_size + kPairsPerSlice - 1) / kPairsPerSlice;
  dim3 grid(s->nvb, slices);
  lbfgs_dots_kernel<<<grid, kVecThreads, 0, st>>>(s->S, s->g, s->gprev, s->d, s->Sb, s->Yb, s->npad, s->dpart);
  MDE_LAUNCH_CHECK();
  direction_scalar_kernel<<<1, 256, 0, st>>>(s->S, s->dpart, s->nvb);
  MDE_LAUNCH_CHECK();
  direction_apply_kernel<<<s->nvb, kVecThreads, 0, st>>>(s->S, s->g, s->gprev, s->d, s->X, s->xinit, s->Sb,
                                                         s->Yb, s->npad, s->dpart);
  MDE_LAUNCH_CHECK();
  ls_init_kernel<<<1, 256, 0, st>>>(s->S, s->dpart, s->nvb);
  MDE_LAUNCH_CHECK();
  return 0;
}

int enqueue_trial(mde_solver* s, cudaStream_t st) {
  trial_axpy_kernel<false><<<s->nvb, kVecThreads, 0, st>>>(s->S, s->xinit, s->d, s->X, s->npad);
  MDE_LAUNCH_CHECK();
  int rc = enqueue_project(s, st);
  if (rc) return rc;
  rc = enqueue_eval(s, &s->S->ls_active, st);
  if (rc) return rc;
  ls_update_kernel<<<1, 256, 0, st>>>(s->S, loss_partials_ptr(s->edges), s->nl, s->g + s->npad, s->dpart, s->nvb,
                                      (double)edges_p_total(s->edges));
  MDE_LAUNCH_CHECK();
  return 0;
}

int enqueue_finish(mde_solver* s, cudaStream_t st) {
  trial_axpy_kernel<true><<<s->nvb, kVecThreads, 0, st>>>(s->S, s->xinit, s->d, s->X, s->npad);
  MDE_LAUNCH_CHECK();
  int rc = enqueue_project(s, st);
  if (rc) return rc;
  iter_end_kernel<<<1, 32, 0, st>>>(s->S);
  MDE_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

int mde_solver_create(mde_solver_t** out, const mde_edges_t* e, int64_t n, int m, const mde_solver_opts_t* opts,
                      void* stream) {
  if (!out || !e || !opts || n < 1 || m < 1) return MDE_E_INVALID;
  if (opts->memory_size < 1 || opts->memory_size > kMaxMemory) return MDE_E_UNSUPPORTED;
  if (opts->constraint == MDE_CONSTRAINT_STANDARDIZED && m > kProjMaxM) return MDE_E_UNSUPPORTED;
  if (opts->constraint < 0 || opts->constraint > MDE_CONSTRAINT_ANCHORED) return MDE_E_INVALID;
  if (opts->max_iter < 1) return MDE_E_INVALID;
  if (opts->mode != 0) return MDE_E_UNSUPPORTED;
  if (n != edges_n(e)) return MDE_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  mde_solver* s = new (std::nothrow) mde_solver();
  if (!s) return MDE_E_ALLOC;
  s->edges = e; s->n = n; s->m = m; s->N = n * m; s->opts = *opts;
  s->npad = ((s->N + 31) / 32) * 32;
  const int64_t vb = (s->npad + 32) * sizeof(float);  // room for the (hi, lo) tail
  int rc = 0;
#define TRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { rc = (int)_e; goto fail; } } while (0)
  TRY(cudaMalloc(&s->S, sizeof(SolverState)));
  TRY(cudaMallocHost(&s->status_host, 8 * sizeof(int)));
  TRY(cudaMalloc(&s->X, vb)); TRY(cudaMalloc(&s->xinit, vb)); TRY(cudaMalloc(&s->d, vb));
  TRY(cudaMalloc(&s->g, vb)); TRY(cudaMalloc(&s->gprev, vb));
  TRY(cudaMalloc(&s->Sb, (int64_t)(opts->memory_size + 1) * s->npad * sizeof(float)));
  TRY(cudaMalloc(&s->Yb, (int64_t)(opts->memory_size + 1) * s->npad * sizeof(float)));
  TRY(cudaMalloc(&s->dpart, sizeof(double) * (int64_t)kVecBlocks * kDotsPerSlice * kMaxSlices));
  TRY(cudaMalloc(&s->stats, sizeof(double) * 4 * (int64_t)opts->max_iter));
  TRY(cudaMalloc(&s->projws, mde_project_ws_bytes(n, m)));
  s->pw = proj_ws_carve(s->projws, m);
  TRY(cudaMemsetAsync(s->X, 0, vb, st)); TRY(cudaMemsetAsync(s->xinit, 0, vb, st));
  TRY(cudaMemsetAsync(s->d, 0, vb, st)); TRY(cudaMemsetAsync(s->g, 0, vb, st));
  TRY(cudaMemsetAsync(s->gprev, 0, vb, st));
  TRY(cudaMemsetAsync(s->Sb, 0, (int64_t)(opts->memory_size + 1) * s->npad * sizeof(float), st));
  TRY(cudaMemsetAsync(s->Yb, 0, (int64_t)(opts->memory_size + 1) * s->npad * sizeof(float), st));
  if (opts->constraint == MDE_CONSTRAINT_ANCHORED && opts->n_anchors > 0) {
    if (!opts->anchors || !opts->anchor_values) { rc = MDE_E_INVALID; goto fail; }
    TRY(cudaMalloc(&s->anchors, sizeof(int64_t) * opts->n_anchors));
    TRY(cudaMalloc(&s->anchor_values, sizeof(float) * opts->n_anchors * m));
    TRY(cudaMemcpyAsync(s->anchors, opts->anchors, sizeof(int64_t) * opts->n_anchors, cudaMemcpyDeviceToDevice, st));
    TRY(cudaMemcpyAsync(s->anchor_values, opts->anchor_values, sizeof(float) * opts->n_anchors * m,
                        cudaMemcpyDeviceToDevice, st));
  }
  s->nvb = vec_blocks(s->npad >> 2);
  *out = s;
  return 0;
fail:
  mde_solver_destroy(s);
  return rc;
#undef TRY
}

int mde_solver_destroy(mde_solver_t* s) {
  if (!s) return 0;
  cudaFree(s->S); cudaFreeHost(s->status_host);
  cudaFree(s->X); cudaFree(s->xinit); cudaFree(s->d); cudaFree(s->g); cudaFree(s->gprev);
  cudaFree(s->Sb); cudaFree(s->Yb); cudaFree(s->dpart); cudaFree(s->stats); cudaFree(s->projws);
  cudaFree(s->anchors); cudaFree(s->anchor_values);
  delete s;
  return 0;
}

int mde_solver_set_allreduce(mde_solver_t* s, mde_allreduce_fn fn, void* user) {
  if (!s) return MDE_E_INVALID;
  s->allreduce = fn; s->allreduce_user = user;
  return 0;
}

int mde_solver_begin(mde_solver_t* s, const float* X0, double eps, void* stream) {
  if (!s || !X0) return MDE_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  MDE_CUDA_TRY(cudaMemcpyAsync(s->X, X0, sizeof(float) * s->N, cudaMemcpyDeviceToDevice, st));
  const int mi = s->opts.max_iter;
  init_state_kernel<<<1, 32, 0, st>>>(s->S, eps, s->opts.memory_size, mi, s->opts.world_size, s->stats,
                                      s->stats + mi, s->stats + 2 * mi, s->stats + 3 * mi);
  MDE_LAUNCH_CHECK();
  s->host_need_fresh = 1;
  s->host_active = 1;
  return 0;
}

int mde_solver_run(mde_solver_t* s, int iters, int* iters_done, int* converged, void* stream) {
  if (!s) return MDE_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = 0;
  for (int it = 0; it < iters && s->host_active; ++it) {
    if (s->host_need_fresh) { if ((rc = enqueue_fresh(s, st))) return rc; }
    if ((rc = enqueue_direction(s, st))) return rc;
    for (int trial = 0;; ++trial) {  // <= 10 back-offs + 25 search steps + ~85 fallback steps
      if (trial > 256) return MDE_E_INVALID;
      if ((rc = enqueue_trial(s, st))) return rc;
      if ((rc = read_status(s, st))) return rc;
      if (!s->status_host[5] || !s->status_host[0]) break;  // ls_active / active
    }
    if (s->status_host[3]) { s->host_active = 0; if (iters_done) *iters_done = s->status_host[2]; return s->status_host[3]; }
    if ((rc = enqueue_finish(s, st))) return rc;
    if ((rc = read_status(s, st))) return rc;
    s->host_need_fresh = s->status_host[4];
    s->host_active = s->status_host[0];
  }
  if ((rc = read_status(s, st))) return rc;
  if (iters_done) *iters_done = s->status_host[2];
  if (converged) *converged = s->status_host[1];
  return 0;
}

float* mde_solver_x(mde_solver_t* s) { return s ? s->X : nullptr; }

int mde_solver_stats(mde_solver_t* s, double* avg, double* resid, double* pct, double* steplen,
                     int64_t* func_evals, void* stream) {
  if (!s) return MDE_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = read_status(s, st);
  if (rc) return rc;
  const int it = s->status_host[2];
  const int mi = s->opts.max_iter;
  if (avg) MDE_CUDA_TRY(cudaMemcpyAsync(avg, s->stats, sizeof(double) * it, cudaMemcpyDeviceToHost, st));
  if (resid) MDE_CUDA_TRY(cudaMemcpyAsync(resid, s->stats + mi, sizeof(double) * it, cudaMemcpyDeviceToHost, st));
  if (pct) MDE_CUDA_TRY(cudaMemcpyAsync(pct, s->stats + 2 * mi, sizeof(double) * it, cudaMemcpyDeviceToHost, st));
  if (steplen) MDE_CUDA_TRY(cudaMemcpyAsync(steplen, s->stats + 3 * mi, sizeof(double) * it, cudaMemcpyDeviceToHost, st));
  if (func_evals) {
    long long fe = 0;
    MDE_CUDA_TRY(cudaMemcpyAsync(&fe, (const char*)s->S + offsetof(SolverState, func_evals), sizeof(long long),
                                 cudaMemcpyDeviceToHost, st));
    MDE_CUDA_TRY(cudaStreamSynchronize(st));
    *func_evals = fe;
  }
  MDE_CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

// ---- host-side debug entry points: the scalar logic above, runnable without a GPU ----------
void* mde_dbg_ls_new(double t0, double f0, float gtd0, float d_norm) {
  LsState* L = new LsState();
  ls_begin(*L, t0, f0, gtd0, d_norm);
  return L;
}
void mde_dbg_ls_free(void* p) { delete (LsState*)p; }
double mde_dbg_ls_t(void* p) { return ((LsState*)p)->t; }
int mde_dbg_ls_step(void* p, double f_new, float gtd_new, int grad_finite) {
  LsState* L = (LsState*)p;
  ls_on_result(*L, f_new, gtd_new, grad_finite != 0);
  return L->phase;
}
void mde_dbg_ls_result(void* p, double* t_accept, double* f_accept, int* func_evals, int* error) {
  LsState* L = (LsState*)p;
  *t_accept = L->t_accept; *f_accept = L->f_accept; *func_evals = L->func_evals; *error = L->error;
}
void* mde_dbg_lbfgs_new(int memory) {
  LbfgsState* B = new LbfgsState();
  lbfgs_reset(*B, memory);
  return B;
}
void mde_dbg_lbfgs_free(void* p) { delete (LbfgsState*)p; }
// one direction update; arrays have kMaxMemory+1 entries.  Outputs: count, cand, order, coefficients.
void mde_dbg_lbfgs_step(void* p, double ys, double yy, double sc_g, double yc_g, double* sj_yc, double* yj_yc,
                        double* sc_yj, double* sj_g, double* yj_g, int* count, int* cand, int* order, double* cg,
                        double* cs, double* cy) {
  LbfgsState* B = (LbfgsState*)p;
  lbfgs_direction(*B, B->SY, B->YY, ys, yy, sc_g, yc_g, sj_yc, yj_yc, sc_yj, sj_g, yj_g);
  *count = B->count; *cand = B->cand; *cg = B->cg;
  for (int j = 0; j < B->count; ++j) { order[j] = B->order[j]; cs[j] = B->cs[j]; cy[j] = B->cy[j]; }
}
void mde_dbg_lbfgs_reset(void* p) { LbfgsState* B = (LbfgsState*)p; lbfgs_reset(*B, B->memory); }
int mde_dbg_lbfgs_cand(void* p) { return ((LbfgsState*)p)->cand; }

}  // extern "C"

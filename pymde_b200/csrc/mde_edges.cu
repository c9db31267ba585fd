// mde_edges.cu -- edge layout + the fused average-distortion kernel (forward + backward).
//
// Replaces pymde/average_distortion.py:36-80 (gather X[lhs], X[rhs]; row norms; per-edge
// penalty; mean; scatter-add of +-g*diff) and pymde/problem.py:246-307 (per-edge outputs).
//
// Data layout in HBM (one shard):
//   src[p], dst[p]  int32   canonical (src < dst) endpoints, sorted by (class, src, dst)
//                           class 0 = attractive/ordinary, 1 = repulsive (PushAndPull w < 0)
//   par0[p]         fp32    weight | deviation, permuted alongside
//   par1[p]         fp32    optional second array (WeightedQuadratic weights)
//   perm[p]         int32   original position of sorted edge k (per-edge outputs only)
// Algorithmic bytes per fused evaluation: p*(8+4k) + 2*n*m*4 + 8  (SURVEY section 8d).
//
// Kernel shapes
//   m <= 4 : one thread per edge, 32 consecutive sorted edges per warp round; the lhs
//            contributions (sorted => runs of equal src) are summed with a warp segmented
//            reduction and issued as ONE vector red per run; rhs contributions go out as
//            vector reds (REDG.E.ADD.F32x2/x4).
//   m >= 5 : a group of G lanes (8/16/32) walks a contiguous slice of edges; the lhs row and
//            its gradient accumulator stay in registers across a run.
#include <cub/cub.cuh>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>

#include "mde_edges.cuh"

namespace mde {
unsigned long long g_launch_count = 0;
}

using namespace mde;

static constexpr int kSmallThreads = 256;
static constexpr int kRounds = 4;  // 32-edge rounds per warp iteration

// ------------------------------------------------------------------------------------------
// layout build
// ------------------------------------------------------------------------------------------
__global__ void make_keys_kernel(const int64_t* __restrict__ edges, const float* __restrict__ par0,
                                 int push_pull, int64_t p, uint64_t* __restrict__ keys,
                                 int32_t* __restrict__ vals) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p) return;
  int64_t i = edges[2 * k], j = edges[2 * k + 1];
  uint64_t lo = (uint64_t)(i < j ? i : j), hi = (uint64_t)(i < j ? j : i);
  uint64_t cls = (push_pull && !(par0[k] >= 0.0f)) ? 1ull : 0ull;
  keys[k] = (cls << 63) | (lo << 32) | hi;  // n < 2^31
  vals[k] = (int32_t)k;
}

__global__ void unpack_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals,
                              const float* __restrict__ par0, const float* __restrict__ par1, int64_t p,
                              int32_t* __restrict__ src, int32_t* __restrict__ dst,
                              float* __restrict__ p0, float* __restrict__ p1, int32_t* __restrict__ perm) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t p4 = (p + 3) & ~(int64_t)3;  // arrays are padded to a multiple of 4 edges
  if (k >= p4) return;
  const int64_t kk = k < p ? k : p - 1;     // pad entries repeat the last edge (masked in the kernels)
  uint64_t key = keys[kk];
  src[k] = (int32_t)((key >> 32) & 0x7fffffffu);
  dst[k] = (int32_t)(key & 0xffffffffu);
  int32_t o = vals[kk];
  p0[k] = par0[o];
  if (par1) p1[k] = par1[o];
  perm[k] = o;  // padded like the other arrays: the quad kernel loads it as int4 (MODE 2)
}

// ------------------------------------------------------------------------------------------
// m <= 4: thread-per-edge kernel
// ------------------------------------------------------------------------------------------
template <int M> struct Row { float v[M]; };

template <int M>
__device__ __forceinline__ Row<M> load_row(const float* __restrict__ X, int r) {
  Row<M> o;
  if constexpr (M == 1) { o.v[0] = __ldg(X + r); }
  else if constexpr (M == 2) { float2 t = __ldg(reinterpret_cast<const float2*>(X) + r); o.v[0] = t.x; o.v[1] = t.y; }
  else if constexpr (M == 4) { float4 t = __ldg(reinterpret_cast<const float4*>(X) + r); o.v[0] = t.x; o.v[1] = t.y; o.v[2] = t.z; o.v[3] = t.w; }
  else {
#pragma unroll
    for (int c = 0; c < M; ++c) o.v[c] = __ldg(X + (int64_t)r * M + c);
  }
  return o;
}

// Deterministic mode: contributions are accumulated as 64-bit fixed point (scale 2^40, resolution 9e-13, range
// +-8.4e6 -- the entries are f'/p-sized).  Integer addition is associative, so the sums do not depend on the order in
// which the reds land (the float reds of the default mode, like the reference's scatter_add_, do:
// pymde/average_distortion.py:75-76).
constexpr float kFxScale = 1099511627776.0f;  // 2^40
template <int M>
__device__ __forceinline__ void red_row_fx(long long* __restrict__ F, int r, const float (&v)[M], float sgn) {
#pragma unroll
  for (int c = 0; c < M; ++c) {
    const long long q = __float2ll_rn(sgn * v[c] * kFxScale);
    atomicAdd(reinterpret_cast<unsigned long long*>(F + (int64_t)r * M + c), (unsigned long long)q);
  }
}
__global__ void fx_zero_kernel(const int* __restrict__ flag, long long* __restrict__ F, int64_t count) {
  if (flag != nullptr && *flag == 0) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) F[i] = 0ll;
}
__global__ void fx_apply_kernel(const int* __restrict__ flag, const long long* __restrict__ F, float* __restrict__ grad,
                                int64_t count) {
  if (flag != nullptr && *flag == 0) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    grad[i] += (float)((double)F[i] * (1.0 / 1099511627776.0));
}

template <int M>
__device__ __forceinline__ void red_row(float* __restrict__ G, int r, const float (&v)[M], float sgn) {
  if constexpr (M == 1) red_add(G + r, sgn * v[0]);
  else if constexpr (M == 2) red_add_v2(G + 2 * (int64_t)r, sgn * v[0], sgn * v[1]);
  else if constexpr (M == 4) red_add_v4(G + 4 * (int64_t)r, sgn * v[0], sgn * v[1], sgn * v[2], sgn * v[3]);
  else {
#pragma unroll
    for (int c = 0; c < M; ++c) red_add(G + (int64_t)r * M + c, sgn * v[c]);
  }
}

// MODE 0: fused value + gradient; 1: value only; 2: external per-edge g (gradient only)
template <int M, int MODE, int FA, int FR>
__global__ void __launch_bounds__(kSmallThreads)
distortion_small_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                        const float* __restrict__ par0, const float* __restrict__ par1,
                        const int32_t* __restrict__ perm, const float* __restrict__ gext,
                        int64_t p, const float* __restrict__ X, float* __restrict__ grad,
                        double* __restrict__ loss_partials, FnDev fn, float inv_p,
                        const int* __restrict__ flag) {
  if (flag != nullptr && *flag == 0) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int64_t kPerWarp = 32 * kRounds;
  double lsum = 0.0;

  for (int64_t base = warp0 * kPerWarp; base < p; base += nwarps * kPerWarp) {
    int s[kRounds], t[kRounds];
    float a[kRounds], b[kRounds];
    bool ok[kRounds];
    // Out-of-range lanes re-read the last edge (index clamp) instead of selecting zeros: every lane
    // then holds valid row indices and no address is formed from a predicated value.  (A zero-select
    // on the 64-bit row offset compiled to `CS2R Rd, SRZ` + predicated IMAD.WIDE, which ptxas 12.9
    // under-stalls on sm_100a -- see profiles/r01_ptxas_cs2r_hazard.md.)
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      int64_t k = base + r * 32 + lane;
      ok[r] = k < p;
      const int64_t kk = ok[r] ? k : (p - 1);
      s[r] = __ldg(src + kk);
      t[r] = __ldg(dst + kk);
      if (MODE == 2) a[r] = __ldg(gext + __ldg(perm + kk));
      else a[r] = __ldg(par0 + kk);
      b[r] = (MODE != 2 && par1 != nullptr) ? __ldg(par1 + kk) : 0.0f;
    }
    Row<M> xi[kRounds], xj[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      xi[r] = load_row<M>(X, s[r]);
      xj[r] = load_row<M>(X, t[r]);
    }
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      float diff[M];
      float d2 = 0.0f;
#pragma unroll
      for (int c = 0; c < M; ++c) { diff[c] = xi[r].v[c] - xj[r].v[c]; d2 += diff[c] * diff[c]; }
      float g;
      if (MODE == 2) {
        g = a[r];
      } else {
        float d = sqrtf(d2), f;
        if (MODE == 0) edge_coeff<FA, FR>(fn, d, a[r], b[r], inv_p, f, g);
        else { edge_value<FA, FR>(fn, d, a[r], b[r], f); g = 0.0f; }
        if (ok[r]) lsum += (double)f;
      }
      if (MODE != 1) {
        float v[M];
#pragma unroll
        for (int c = 0; c < M; ++c) v[c] = ok[r] ? g * diff[c] : 0.0f;
        if (ok[r]) red_row<M>(grad, t[r], v, -1.0f);
        // lhs: sorted => equal keys are adjacent lanes; segmented warp reduction
        int key = ok[r] ? s[r] : (-1 - lane);
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          int k2 = __shfl_down_sync(kFull, key, off);
          bool take = (lane + off < 32) && (k2 == key);
#pragma unroll
          for (int c = 0; c < M; ++c) {
            float o = __shfl_down_sync(kFull, v[c], off);
            if (take) v[c] += o;
          }
        }
        int kprev = __shfl_up_sync(kFull, key, 1);
        bool head = (lane == 0) || (kprev != key);
        if (head && ok[r]) red_row<M>(grad, s[r], v, 1.0f);
      }
    }
  }
  if (MODE != 2) {
    __shared__ double sm[32];
    double v1[1] = {lsum};
    block_sum<1>(v1, sm);
    if (threadIdx.x == 0) loss_partials[blockIdx.x] = v1[0];
  }
}

// ------------------------------------------------------------------------------------------
// m <= 4, thread-contiguous variant: each thread owns 4 CONSECUTIVE sorted edges (three 16-byte
// loads for src/dst/par0), sums the lhs contributions of equal-src runs in registers and issues one
// red per run; no warp shuffles.  FAST selects the MUFU-based math for PushAndPull(Log1p(1.5), Log(1)).
// ------------------------------------------------------------------------------------------
static constexpr int kQuadThreads = 256;

template <int M, int MODE, int FA, int FR, bool FAST, int NQ>
__global__ void __launch_bounds__(kQuadThreads)
distortion_quad_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                       const float* __restrict__ par0, const float* __restrict__ par1,
                       const int32_t* __restrict__ perm, const float* __restrict__ gext,
                       int64_t p, const float* __restrict__ X, float* __restrict__ grad,
                       double* __restrict__ loss_partials, FnDev fn, float inv_p,
                       const int* __restrict__ flag, long long* __restrict__ fx) {
  if (flag != nullptr && *flag == 0) return;
  constexpr int E = 4 * NQ;  // consecutive edges owned by a thread per iteration
  const int64_t nquads = (p + 3) >> 2;
  const int64_t nunits = (nquads + NQ - 1) / NQ;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float lsum_f = 0.0f;
  double lsum = 0.0;
  // software pipeline (NQ == 1): the edge record of the NEXT grid-stride iteration is requested before the
  // vertex rows of the current one are gathered, so its latency overlaps the gathers and the math
  int4 s4n = make_int4(0, 0, 0, 0), t4n = make_int4(0, 0, 0, 0);
  float4 a4n = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool pipe = (NQ == 1 && MODE != 2);
  {
    const int64_t u0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pipe && u0 < nunits) {
      s4n = __ldg(reinterpret_cast<const int4*>(src) + u0);
      t4n = __ldg(reinterpret_cast<const int4*>(dst) + u0);
      a4n = __ldg(reinterpret_cast<const float4*>(par0) + u0);
    }
  }
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < nunits; u += stride) {
    int s[E], t[E];
    float a[E], b[E];
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      const int64_t q = u * NQ + j;
      const int64_t qc = q < nquads ? q : nquads - 1;  // clamped: masked below through `ok`
      int4 s4, t4;
      if (pipe) {
        s4 = s4n; t4 = t4n;
        const int64_t un = u + stride;
        if (un < nunits) {
          s4n = __ldg(reinterpret_cast<const int4*>(src) + un);
          t4n = __ldg(reinterpret_cast<const int4*>(dst) + un);
        }
      } else {
        s4 = __ldg(reinterpret_cast<const int4*>(src) + qc);
        t4 = __ldg(reinterpret_cast<const int4*>(dst) + qc);
      }
      s[4 * j] = s4.x; s[4 * j + 1] = s4.y; s[4 * j + 2] = s4.z; s[4 * j + 3] = s4.w;
      t[4 * j] = t4.x; t[4 * j + 1] = t4.y; t[4 * j + 2] = t4.z; t[4 * j + 3] = t4.w;
      if (MODE == 2) {
        const int4 o4 = __ldg(reinterpret_cast<const int4*>(perm) + qc);  // pad entries repeat the last edge
        a[4 * j] = __ldg(gext + o4.x); a[4 * j + 1] = __ldg(gext + o4.y);
        a[4 * j + 2] = __ldg(gext + o4.z); a[4 * j + 3] = __ldg(gext + o4.w);
      } else {
        float4 a4;
        if (pipe) {
          a4 = a4n;
          const int64_t un = u + stride;
          if (un < nunits) a4n = __ldg(reinterpret_cast<const float4*>(par0) + un);
        } else {
          a4 = __ldg(reinterpret_cast<const float4*>(par0) + qc);
        }
        a[4 * j] = a4.x; a[4 * j + 1] = a4.y; a[4 * j + 2] = a4.z; a[4 * j + 3] = a4.w;
      }
      if (MODE != 2 && par1 != nullptr) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(par1) + qc);
        b[4 * j] = b4.x; b[4 * j + 1] = b4.y; b[4 * j + 2] = b4.z; b[4 * j + 3] = b4.w;
      } else { b[4 * j] = b[4 * j + 1] = b[4 * j + 2] = b[4 * j + 3] = 0.0f; }
    }
    Row<M> xi[E], xj[E];
#pragma unroll
    for (int e = 0; e < E; ++e) { xi[e] = load_row<M>(X, s[e]); xj[e] = load_row<M>(X, t[e]); }
    float acc[M];
#pragma unroll
    for (int c = 0; c < M; ++c) acc[c] = 0.0f;
    int cur = s[0];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const bool ok = (4 * (u * NQ) + e) < p;
      float diff[M];
      float d2 = 0.0f;
#pragma unroll
      for (int c = 0; c < M; ++c) { diff[c] = xi[e].v[c] - xj[e].v[c]; d2 += diff[c] * diff[c]; }
      float g, f = 0.0f;
      if (MODE == 2) {
        g = a[e];
      } else if (FAST) {
        edge_coeff_fast_log1p_log(d2, a[e], inv_p, f, g);
      } else {
        const float d = sqrtf(d2);
        if (MODE == 0) edge_coeff<FA, FR>(fn, d, a[e], b[e], inv_p, f, g);
        else { edge_value<FA, FR>(fn, d, a[e], b[e], f); g = 0.0f; }
      }
      if (MODE != 2 && ok) { if (FAST) lsum_f += f; else lsum += (double)f; }
      if (MODE != 1) {
        // d = 0: the reference replaces the non-finite g by 1 and the difference vector is 0
        const bool live = ok && (FAST ? (d2 > 0.0f) : true);
        float v[M];
#pragma unroll
        for (int c = 0; c < M; ++c) v[c] = live ? g * diff[c] : 0.0f;
        if (live) { if (fx) red_row_fx<M>(fx, t[e], v, -1.0f); else red_row<M>(grad, t[e], v, -1.0f); }
        if (s[e] != cur) {  // run of equal src ended: flush its sum
          if (fx) red_row_fx<M>(fx, cur, acc, 1.0f); else red_row<M>(grad, cur, acc, 1.0f);
          cur = s[e];
#pragma unroll
          for (int c = 0; c < M; ++c) acc[c] = 0.0f;
        }
#pragma unroll
        for (int c = 0; c < M; ++c) acc[c] += v[c];
      }
    }
    if (MODE != 1) { if (fx) red_row_fx<M>(fx, cur, acc, 1.0f); else red_row<M>(grad, cur, acc, 1.0f); }
    if (FAST) { lsum += (double)lsum_f; lsum_f = 0.0f; }
  }
  if (MODE != 2) {
    __shared__ double sm[32];
    double v1[1] = {lsum};
    block_sum<1>(v1, sm);
    if (threadIdx.x == 0) loss_partials[blockIdx.x] = v1[0];
  }
}

// ------------------------------------------------------------------------------------------
// m >= 5: group-per-edge kernel.  G lanes share one edge; lane l owns columns l, l+G, ...
// (CPL of them).  VW = 4 treats the row as m/4 float4 columns.
// ------------------------------------------------------------------------------------------
template <int VW> struct Vec { float v[VW]; };

template <int VW>
__device__ __forceinline__ Vec<VW> ldv(const float* __restrict__ p) {
  Vec<VW> o;
  if constexpr (VW == 4) { float4 t = __ldg(reinterpret_cast<const float4*>(p)); o.v[0] = t.x; o.v[1] = t.y; o.v[2] = t.z; o.v[3] = t.w; }
  else o.v[0] = __ldg(p);
  return o;
}
template <int VW>
__device__ __forceinline__ void redv(float* p, const Vec<VW>& x, float sgn) {
  if constexpr (VW == 4) red_add_v4(p, sgn * x.v[0], sgn * x.v[1], sgn * x.v[2], sgn * x.v[3]);
  else red_add(p, sgn * x.v[0]);
}

static constexpr int kWideThreads = 256;
static constexpr int kWideSlice = 64;  // consecutive edges walked by one group

template <int G, int CPL, int VW, int MODE>
__global__ void __launch_bounds__(kWideThreads)
distortion_wide_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                       const float* __restrict__ par0, const float* __restrict__ par1,
                       const int32_t* __restrict__ perm, const float* __restrict__ gext,
                       int64_t p, int m, const float* __restrict__ X, float* __restrict__ grad,
                       double* __restrict__ loss_partials, FnDev fn, float inv_p,
                       const int* __restrict__ flag) {
  if (flag != nullptr && *flag == 0) return;
  const int lg = threadIdx.x % G;
  // shuffles stay inside the G-lane group: groups of one warp may run different trip counts
  const unsigned gmask = (G == 32) ? kFull : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
  const int64_t group0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / G;
  const int mv = m / VW;  // columns in units of VW floats
  double lsum = 0.0;

  for (int64_t k0 = group0 * kWideSlice; k0 < p; k0 += ngroups * kWideSlice) {
    int cur = -1;
    Vec<VW> xi[CPL], acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
      for (int q = 0; q < VW; ++q) { xi[c].v[q] = 0.0f; acc[c].v[q] = 0.0f; }
    // trip count is uniform across the group (shuffles inside); out-of-range iterations re-read the
    // last edge / last column (clamped indices) and are masked out of every sum and store
    for (int it = 0; it < kWideSlice; ++it) {
      const int64_t k = k0 + it;
      const bool ok = k < p;
      const int64_t kk = ok ? k : (p - 1);
      const int s = __ldg(src + kk);
      const int t = __ldg(dst + kk);
      if (ok && s != cur) {
        if (cur >= 0 && MODE != 1) {
#pragma unroll
          for (int c = 0; c < CPL; ++c) {
            int col = lg + c * G;
            if (col < mv) redv<VW>(grad + (int64_t)cur * m + col * VW, acc[c], 1.0f);
          }
        }
        cur = s;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const int col = lg + c * G;
          const int colc = col < mv ? col : (mv - 1);
          xi[c] = ldv<VW>(X + (int64_t)s * m + colc * VW);
#pragma unroll
          for (int q = 0; q < VW; ++q) acc[c].v[q] = 0.0f;
        }
      }
      Vec<VW> diff[CPL];
      float d2 = 0.0f;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int col = lg + c * G;
        const int colc = col < mv ? col : (mv - 1);
        const Vec<VW> xj = ldv<VW>(X + (int64_t)t * m + colc * VW);
        const float keep = (col < mv) ? 1.0f : 0.0f;
#pragma unroll
        for (int q = 0; q < VW; ++q) { diff[c].v[q] = (xi[c].v[q] - xj.v[q]) * keep; d2 += diff[c].v[q] * diff[c].v[q]; }
      }
#pragma unroll
      for (int off = G / 2; off > 0; off >>= 1) d2 += __shfl_xor_sync(gmask, d2, off);
      float g, f = 0.0f;
      if (MODE == 2) {
        g = __ldg(gext + __ldg(perm + kk));
      } else {
        const float a = __ldg(par0 + kk);
        const float b = (par1 != nullptr) ? __ldg(par1 + kk) : 1.0f;
        const float d = sqrtf(d2);
        if (MODE == 0) edge_coeff<-1, -1>(fn, d, a, b, inv_p, f, g);
        else { edge_value<-1, -1>(fn, d, a, b, f); g = 0.0f; }
        if (ok && lg == 0) lsum += (double)f;
      }
      if (MODE != 1 && ok) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          int col = lg + c * G;
          if (col < mv) {
            Vec<VW> v;
#pragma unroll
            for (int q = 0; q < VW; ++q) { v.v[q] = g * diff[c].v[q]; acc[c].v[q] += v.v[q]; }
            redv<VW>(grad + (int64_t)t * m + col * VW, v, -1.0f);
          }
        }
      }
    }
    if (cur >= 0 && MODE != 1) {
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        int col = lg + c * G;
        if (col < mv) redv<VW>(grad + (int64_t)cur * m + col * VW, acc[c], 1.0f);
      }
    }
  }
  if (MODE != 2) {
    __shared__ double sm[32];
    double v1[1] = {lsum};
    block_sum<1>(v1, sm);
    if (threadIdx.x == 0) loss_partials[blockIdx.x] = v1[0];
  }
}

// sum of per-block partials, added to *out (single block; fixed order => deterministic)
__global__ void add_partials_kernel(const double* __restrict__ partials, int nblocks, double* out) {
  __shared__ double sm[32];
  double v[1] = {0.0};
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) v[0] += partials[i];
  block_sum<1>(v, sm);
  if (threadIdx.x == 0) out[0] += v[0];
}

// per-edge outputs in the caller's order
template <int MFIX>
__global__ void edge_outputs_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                    const float* __restrict__ par0, const float* __restrict__ par1,
                                    const int32_t* __restrict__ perm, int64_t p, int m,
                                    const float* __restrict__ X, float* __restrict__ distances,
                                    float* __restrict__ distortions, FnDev fn) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p) return;
  int s = src[k], t = dst[k];
  float d2 = 0.0f;
  for (int c = 0; c < m; ++c) {
    float df = __ldg(X + (int64_t)s * m + c) - __ldg(X + (int64_t)t * m + c);
    d2 += df * df;
  }
  float d = sqrtf(d2);
  int o = perm[k];
  if (distances) distances[o] = d;
  if (distortions) {
    float f;
    edge_value<-1, -1>(fn, d, par0[k], par1 ? par1[k] : 0.0f, f);
    distortions[o] = f;
  }
}

__global__ void function_eval_kernel(FnDev fn, const float* __restrict__ par0, int64_t par0_len,
                                     const float* __restrict__ par1, const float* __restrict__ dist,
                                     int64_t p, float* __restrict__ f_out, float* __restrict__ fp_out) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p) return;
  float a = par0[par0_len == 1 ? 0 : k];
  float b = par1 ? par1[k] : 0.0f;
  float d = dist[k];
  float f, fp;
  edge_f_fp<-1, -1>(fn, d, a, b, f, fp);
  if (f_out) f_out[k] = f;
  if (fp_out) fp_out[k] = fp;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
namespace mde {

int loss_blocks_small(int64_t p) {
  int64_t per_block = (int64_t)(kSmallThreads / 32) * 32 * kRounds;
  int64_t nb = (p + per_block - 1) / per_block;
  if (nb < 1) nb = 1;
  if (nb > kNumSMs * 8) nb = kNumSMs * 8;
  return (int)nb;
}

int quad_nq() {  // consecutive quads per thread: MDE_B200_NQ=1|2 (experimental A/B switch; default 1)
  static int v = -1;
  if (v < 0) { const char* e = getenv("MDE_B200_NQ"); v = (e && e[0] == '2') ? 2 : 1; }
  return v;
}

int quad_grid_cap() {  // blocks per SM in the grid cap: MDE_B200_QUAD_BPS (default 4 = one resident wave)
  static int v = -1;
  if (v < 0) { const char* e = getenv("MDE_B200_QUAD_BPS"); v = e ? atoi(e) : 4; if (v < 1) v = 1; if (v > 16) v = 16; }
  return v;
}

int loss_blocks_quad(int64_t p, int nq) {
  int64_t per_block = (int64_t)kQuadThreads * 4 * nq;
  int64_t nb = (p + per_block - 1) / per_block;
  if (nb < 1) nb = 1;
  if (nb > kNumSMs * quad_grid_cap()) nb = kNumSMs * quad_grid_cap();
  return (int)nb;
}

// A/B switch for measurements: MDE_B200_KERNEL=strided selects the lane-strided kernel with the warp
// segmented reduction; MDE_B200_KERNEL=precise keeps the thread-contiguous kernel but IEEE math.
int small_kernel_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MDE_B200_KERNEL");
    v = 0;
    if (e && !strcmp(e, "strided")) v = 1;
    if (e && !strcmp(e, "precise")) v = 2;
  }
  return v;
}

int loss_blocks_wide(int64_t p, int G) {
  int64_t groups_per_block = kWideThreads / G;
  int64_t per_block = groups_per_block * kWideSlice;
  int64_t nb = (p + per_block - 1) / per_block;
  if (nb < 1) nb = 1;
  if (nb > kNumSMs * 8) nb = kNumSMs * 8;
  return (int)nb;
}

template <int MODE>
int launch_distortion(const mde_edges* e, const float* X, int m, float* grad, const float* gext,
                      int* nblocks_out, const int* flag, cudaStream_t st) {
  const float inv_p = 1.0f / (float)e->p_total;
  const int64_t p = e->p;
  int nb;
  // deterministic mode (m <= 4, sorted-SoA layout): zero the fixed-point buffer, accumulate into it, add it to grad
  long long* fxp = nullptr;
  if (e->det && MODE != 1 && m <= 4 && m <= e->m_hint && small_kernel_variant() != 1) {
    fxp = e->fx;
    const int64_t cnt = e->n * m;
    int zb = (int)((cnt + 255) / 256); if (zb > kNumSMs * 8) zb = kNumSMs * 8;
    fx_zero_kernel<<<zb, 256, 0, st>>>(flag, fxp, cnt);
    ++g_launch_count;
  }
#define SMALLK(MM, FA, FR)                                                                            \
  distortion_small_kernel<MM, MODE, FA, FR><<<nb, kSmallThreads, 0, st>>>(                            \
      e->src, e->dst, e->par0, e->has_par1 ? e->par1 : nullptr, e->perm, gext, p, X, grad,           \
      e->loss_partials, e->fn, inv_p, flag)
  // hot function combinations get compile-time ids (fused mode, m = 2 / 3); the rest use the table
#define QUADK_(MM, FA, FR, FAST, NQV)                                                                 \
  distortion_quad_kernel<MM, MODE, FA, FR, FAST, NQV><<<nb, kQuadThreads, 0, st>>>(                   \
      e->src, e->dst, e->par0, e->has_par1 ? e->par1 : nullptr, e->perm, gext, p, X, grad,           \
      e->loss_partials, e->fn, inv_p, flag, fxp)
#define QUADK(MM, FA, FR, FAST)                                                                       \
  if (FAST && quad_nq() == 2) { nb = loss_blocks_quad(p, 2); QUADK_(MM, FA, FR, FAST, 2); }           \
  else { nb = loss_blocks_quad(p, 1); QUADK_(MM, FA, FR, FAST, 1); }
#define SMALL(MM)                                                                                     \
  if (small_kernel_variant() != 1) {                                                                  \
    nb = loss_blocks_quad(p, 1);                                                                      \
    const int fa = e->fn.fn_att, fr = e->fn.fn_rep, pp = e->fn.push_pull;                             \
    const bool hot = pp && fa == MDE_FN_P_LOG1P && fr == MDE_FN_P_LOG && e->fn.a0 == 1.5f &&          \
                     e->fn.r0 == 1.0f && small_kernel_variant() == 0;                                 \
    if constexpr (MODE == 0 && (MM == 2 || MM == 3)) {                                                \
      if (hot) { QUADK(MM, MDE_FN_P_LOG1P, MDE_FN_P_LOG, true); }                                     \
      else if (pp && fa == MDE_FN_P_LOG1P && fr == MDE_FN_P_LOG) { QUADK(MM, MDE_FN_P_LOG1P, MDE_FN_P_LOG, false); } \
      else if (pp && fa == MDE_FN_P_LOG1P && fr == MDE_FN_P_LOGRATIO) { QUADK(MM, MDE_FN_P_LOG1P, MDE_FN_P_LOGRATIO, false); } \
      else if (!pp && fa == MDE_FN_P_QUADRATIC) { QUADK(MM, MDE_FN_P_QUADRATIC, MDE_FN_P_QUADRATIC, false); }     \
      else if (!pp && fa == MDE_FN_L_ABSOLUTE) { QUADK(MM, MDE_FN_L_ABSOLUTE, MDE_FN_L_ABSOLUTE, false); }        \
      else if (!pp && fa == MDE_FN_L_QUADRATIC) { QUADK(MM, MDE_FN_L_QUADRATIC, MDE_FN_L_QUADRATIC, false); }     \
      else if (!pp && fa == MDE_FN_L_WEIGHTED_QUADRATIC) { QUADK(MM, MDE_FN_L_WEIGHTED_QUADRATIC, MDE_FN_L_WEIGHTED_QUADRATIC, false); } \
      else if (!pp && fa == MDE_FN_L_HUBER) { QUADK(MM, MDE_FN_L_HUBER, MDE_FN_L_HUBER, false); }                 \
      else { QUADK(MM, -1, -1, false); }                                                              \
    } else { QUADK(MM, -1, -1, false); }                                                              \
  } else {                                                                                            \
    SMALL_STRIDED(MM);                                                                                \
  }
#define SMALL_STRIDED(MM)                                                                             \
  nb = loss_blocks_small(p);                                                                          \
  if constexpr (MODE == 0 && (MM == 2 || MM == 3)) {                                                  \
    const int fa = e->fn.fn_att, fr = e->fn.fn_rep, pp = e->fn.push_pull;                             \
    if (pp && fa == MDE_FN_P_LOG1P && fr == MDE_FN_P_LOG) { SMALLK(MM, MDE_FN_P_LOG1P, MDE_FN_P_LOG); }            \
    else if (pp && fa == MDE_FN_P_LOG1P && fr == MDE_FN_P_LOGRATIO) { SMALLK(MM, MDE_FN_P_LOG1P, MDE_FN_P_LOGRATIO); } \
    else if (!pp && fa == MDE_FN_P_QUADRATIC) { SMALLK(MM, MDE_FN_P_QUADRATIC, MDE_FN_P_QUADRATIC); }             \
    else if (!pp && fa == MDE_FN_P_LOG1P) { SMALLK(MM, MDE_FN_P_LOG1P, MDE_FN_P_LOG1P); }                         \
    else if (!pp && fa == MDE_FN_L_ABSOLUTE) { SMALLK(MM, MDE_FN_L_ABSOLUTE, MDE_FN_L_ABSOLUTE); }                \
    else if (!pp && fa == MDE_FN_L_QUADRATIC) { SMALLK(MM, MDE_FN_L_QUADRATIC, MDE_FN_L_QUADRATIC); }             \
    else if (!pp && fa == MDE_FN_L_WEIGHTED_QUADRATIC) { SMALLK(MM, MDE_FN_L_WEIGHTED_QUADRATIC, MDE_FN_L_WEIGHTED_QUADRATIC); } \
    else if (!pp && fa == MDE_FN_L_HUBER) { SMALLK(MM, MDE_FN_L_HUBER, MDE_FN_L_HUBER); }                         \
    else { SMALLK(MM, -1, -1); }                                                                      \
  } else { SMALLK(MM, -1, -1); }
#define WIDE(GG, CC, VV)                                                                              \
  nb = loss_blocks_wide(p, GG);                                                                       \
  distortion_wide_kernel<GG, CC, VV, MODE><<<nb, kWideThreads, 0, st>>>(                              \
      e->src, e->dst, e->par0, e->has_par1 ? e->par1 : nullptr, e->perm, gext, p, m, X, grad,        \
      e->loss_partials, e->fn, inv_p, flag)
  if (m == 1) { SMALL(1); }
  else if (m == 2) { SMALL(2); }
  else if (m == 3) { SMALL(3); }
  else if (m == 4) { SMALL(4); }
  else if (m % 4 == 0) {
    int mv = m / 4;
    if (mv <= 8) { WIDE(8, 1, 4); }
    else if (mv <= 16) { WIDE(16, 1, 4); }
    else if (mv <= 32) { WIDE(32, 1, 4); }
    else if (mv <= 64) { WIDE(32, 2, 4); }
    else if (mv <= 128) { WIDE(32, 4, 4); }
    else if (mv <= 256) { WIDE(32, 8, 4); }
    else return MDE_E_UNSUPPORTED;
  } else {
    if (m <= 8) { WIDE(8, 1, 1); }
    else if (m <= 16) { WIDE(16, 1, 1); }
    else if (m <= 32) { WIDE(32, 1, 1); }
    else if (m <= 64) { WIDE(32, 2, 1); }
    else if (m <= 128) { WIDE(32, 4, 1); }
    else if (m <= 256) { WIDE(32, 8, 1); }
    else if (m <= 512) { WIDE(32, 16, 1); }
    else return MDE_E_UNSUPPORTED;
  }
#undef SMALL
#undef SMALL_STRIDED
#undef QUADK
#undef QUADK_
#undef SMALLK
#undef WIDE
  MDE_LAUNCH_CHECK();
  if (fxp) {
    const int64_t cnt = e->n * m;
    int zb = (int)((cnt + 255) / 256); if (zb > kNumSMs * 8) zb = kNumSMs * 8;
    fx_apply_kernel<<<zb, 256, 0, st>>>(flag, fxp, grad, cnt);
    MDE_LAUNCH_CHECK();
  }
  if (nblocks_out) *nblocks_out = nb;
  return 0;
}

// used by the solver: fused launch leaving per-block loss partials in e->loss_partials
int distortion_fused(const mde_edges* e, const float* X, int m, float* grad, int* nblocks, cudaStream_t st) {
  if (e->kind == 3 && m == e->m_hint) return ell_launch(e, X, m, grad, nblocks, nullptr, st);
  if (e->kind == 2) return pull_launch(0, e, X, m, grad, nullptr, nblocks, nullptr, st);
  if (e->kind == 1) return tiled_launch(0, e, X, m, grad, nullptr, nblocks, nullptr, st);
  return launch_distortion<0>(e, X, m, grad, nullptr, nblocks, nullptr, st);
}
int distortion_fused_flag(const mde_edges* e, const float* X, int m, float* grad, int* nblocks,
                          const int* flag, cudaStream_t st) {
  if (e->kind == 3 && m == e->m_hint) return ell_launch(e, X, m, grad, nblocks, flag, st);
  if (e->kind == 2) return pull_launch(0, e, X, m, grad, nullptr, nblocks, flag, st);
  if (e->kind == 1) return tiled_launch(0, e, X, m, grad, nullptr, nblocks, flag, st);
  return launch_distortion<0>(e, X, m, grad, nullptr, nblocks, flag, st);
}
int64_t edges_p_total(const mde_edges* e) { return e->p_total; }
const double* loss_partials_ptr(const mde_edges* e) { return e->loss_partials; }
int64_t edges_n(const mde_edges* e) { return e->n; }

}  // namespace mde

extern "C" {

int mde_abi_version(void) { return MDE_ABI_VERSION; }

uint64_t mde_launch_count(void) { return (uint64_t)mde::g_launch_count; }

const char* mde_error_string(int code) {
  if (code == 0) return "ok";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  switch (code) {
    case MDE_E_INVALID: return "mde: invalid argument";
    case MDE_E_UNSUPPORTED: return "mde: unsupported configuration";
    case MDE_E_NAN: return "mde: function or gradient evaluation returned NaN/Inf";
    case MDE_E_ALLOC: return "mde: allocation failed";
    case MDE_E_COMM: return "mde: multi-GPU handshake timed out (a peer rank never arrived)";
  }
  return "mde: unknown error";
}

int mde_edges_create(mde_edges_t** out, const int64_t* edges, int64_t p, int64_t n_items,
                     const float* par0, const float* par1, const mde_fn_t* fn, int64_t p_total,
                     void* stream) {
  return mde_edges_create_ex(out, edges, p, n_items, par0, par1, fn, p_total, 2, stream);
}

// layout choice: MDE_B200_LAYOUT=soa forces the sorted-SoA layout (quad / strided kernels), =tiles insists on
// the tile-record layout whenever it can be built; default: tiles for m <= 4 without a second parameter array
static int layout_pref() {  // read at every create: A/B runs build both layouts in one process
  const char* ev = getenv("MDE_B200_LAYOUT");
  if (ev && !strcmp(ev, "soa")) return 1;
  if (ev && !strcmp(ev, "tiles")) return 2;
  if (ev && !strcmp(ev, "pull")) return 3;
  if (ev && !strcmp(ev, "ell")) return 4;
  return 0;
}

int mde_edges_create_ex(mde_edges_t** out, const int64_t* edges, int64_t p, int64_t n_items,
                        const float* par0, const float* par1, const mde_fn_t* fn, int64_t p_total,
                        int embedding_dim, void* stream) {
  if (!out || !edges || !par0 || !fn || p <= 0 || n_items <= 0 || p >= (1ll << 31) ||
      n_items >= (1ll << 31) || p_total < p)
    return MDE_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  mde_edges* e = new (std::nothrow) mde_edges();
  if (!e) return MDE_E_ALLOC;
  e->p = p; e->n = n_items; e->p_total = p_total; e->fn = to_dev(*fn); e->has_par1 = par1 != nullptr;

  uint64_t *keys_in = nullptr, *keys_out = nullptr;
  int32_t *vals_in = nullptr, *vals_out = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  int rc = 0, pref = 0;
  bool dense = false, want_ell = false;
#define TRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { rc = (int)_e; goto fail; } } while (0)
  TRY(cudaMalloc(&e->loss_partials, sizeof(double) * kMaxLossBlocks));
  // Layout choice (measured on B200, profiles/r02_kernels.md): the SM's L1 -> L2 request path (~1 sector request per
  // cycle) bounds the gather / scatter; the tile kernels trade requests for instructions (pull: every edge is
  // evaluated from both ends) and win once an owner's run inside a tile is long, i.e. on dense graphs
  // (C3: 447 edges per node, 2.0x faster); on sparse ones (C2: 22, C5: 10) the sorted-SoA quad kernel is as fast or
  // faster, and needs half the layout memory.
  dense = (p / n_items) >= 64;
  pref = layout_pref();
  {  // MDE_B200_DETERMINISTIC=1: order-independent (fixed-point) gradient accumulation; sorted-SoA layout, m <= 4
    const char* ev = getenv("MDE_B200_DETERMINISTIC");
    if (ev && ev[0] == '1' && embedding_dim >= 1 && embedding_dim <= 4) {
      e->det = 1; e->m_hint = embedding_dim; pref = 1;
      TRY(cudaMalloc(&e->fx, sizeof(long long) * n_items * embedding_dim));
    }
  }
  // dense graphs: ELL pull records next to the sorted-SoA arrays (C3-shaped: 110 us against 129 us for the flat pull
  // records, profiles/r02_kernels.md) whenever the ELL builder accepts the shape (n < 2^24, <= 32 neighbour tiles);
  // MDE_B200_LAYOUT=ell asks for them on any graph
  want_ell = embedding_dim >= 1 && embedding_dim <= 4 && !par1 && !e->det && (pref == 4 || (pref == 0 && dense)) &&
             ell_supported(n_items, embedding_dim);
  if (embedding_dim >= 1 && embedding_dim <= 4 && !par1 && pref != 1 && pref != 4 && (pref != 0 || dense) && !want_ell) {
    if (pref != 2) {  // pull records
      rc = pull_build(e, edges, par0, fn, embedding_dim, st);
      if (rc == 0) { *out = e; return 0; }
      if (rc != MDE_E_UNSUPPORTED) goto fail;
    }
    if (pref != 3) {
      rc = tiled_build(e, edges, par0, fn, embedding_dim, st);
      if (rc == 0) { *out = e; return 0; }
      if (rc != MDE_E_UNSUPPORTED) goto fail;
    }
    rc = 0;  // not suited to tiles (very sparse / huge): sorted-SoA layout below
  }
  TRY(cudaMalloc(&keys_in, sizeof(uint64_t) * p));
  TRY(cudaMalloc(&keys_out, sizeof(uint64_t) * p));
  TRY(cudaMalloc(&vals_in, sizeof(int32_t) * p));
  TRY(cudaMalloc(&vals_out, sizeof(int32_t) * p));
  TRY(cudaMalloc(&e->src, sizeof(int32_t) * (p + 4)));
  TRY(cudaMalloc(&e->dst, sizeof(int32_t) * (p + 4)));
  TRY(cudaMalloc(&e->par0, sizeof(float) * (p + 4)));
  TRY(cudaMalloc(&e->perm, sizeof(int32_t) * (p + 4)));
  if (par1) TRY(cudaMalloc(&e->par1, sizeof(float) * (p + 4)));
  e->nbytes = p * (4 + 4 + 4 + 4 + (par1 ? 4 : 0)) + 8 * kMaxLossBlocks;
  {
    int tb = 256;
    int nb = ceil_div_i64(p, tb);
    make_keys_kernel<<<nb, tb, 0, st>>>(edges, par0, fn->push_pull, p, keys_in, vals_in);
    ++g_launch_count;
    TRY(cudaPeekAtLastError());
    TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int)p, 0, 64, st));
    TRY(cudaMalloc(&tmp, tmp_bytes));
    TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int)p, 0, 64, st));
    unpack_kernel<<<ceil_div_i64(p + 4, tb), tb, 0, st>>>(keys_out, vals_out, par0, par1, p, e->src, e->dst, e->par0,
                                                         e->par1, e->perm);
    ++g_launch_count;
    TRY(cudaPeekAtLastError());
    TRY(cudaStreamSynchronize(st));
  }
  cudaFree(keys_in); cudaFree(keys_out); cudaFree(vals_in); cudaFree(vals_out); cudaFree(tmp);
  keys_in = keys_out = nullptr; vals_in = vals_out = nullptr; tmp = nullptr;
  if (want_ell) {
    // ELL pull records next to the sorted-SoA arrays (kind 3); MDE_E_UNSUPPORTED leaves the layout at kind 0
    rc = ell_build(e, fn, embedding_dim, st);
    if (rc == (int)cudaErrorMemoryAllocation) {  // no room for the second layout: the sorted-SoA one is complete
      (void)cudaGetLastError();
      rc = MDE_E_UNSUPPORTED;
    }
    if (rc != 0 && rc != MDE_E_UNSUPPORTED) goto fail;
    rc = 0;
  }
  *out = e;
  return 0;
fail:
  cudaFree(keys_in); cudaFree(keys_out); cudaFree(vals_in); cudaFree(vals_out); cudaFree(tmp);
  mde_edges_destroy(e);
  return rc;
#undef TRY
}

int mde_edges_destroy(mde_edges_t* e) {
  if (!e) return 0;
  cudaFree(e->src); cudaFree(e->dst); cudaFree(e->perm); cudaFree(e->par0); cudaFree(e->par1);
  cudaFree(e->loss_partials); cudaFree(e->fx);
  tiled_free(e);
  pull_free(e);
  ell_free(e);
  delete e;
  return 0;
}

int64_t mde_edges_count(const mde_edges_t* e) { return e ? e->p : 0; }
int mde_edges_kind(const mde_edges_t* e) { return e ? e->kind : -1; }
int mde_edges_deterministic(const mde_edges_t* e) { return e ? e->det : 0; }
int64_t mde_edges_nbytes(const mde_edges_t* e) { return e ? e->nbytes : 0; }

int mde_distortion(const mde_edges_t* e, const float* X, int m, float* grad, double* loss_sum,
                   void* stream) {
  if (!e || !X || m < 1) return MDE_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  int nb = 0, rc;
  if (e->kind == 3 && grad && m == e->m_hint) rc = ell_launch(e, X, m, grad, &nb, nullptr, st);
  else if (e->kind == 2) rc = pull_launch(grad ? 0 : 1, e, X, m, grad, nullptr, &nb, nullptr, st);
  else if (e->kind == 1) rc = tiled_launch(grad ? 0 : 1, e, X, m, grad, nullptr, &nb, nullptr, st);
  else if (grad) rc = launch_distortion<0>(e, X, m, grad, nullptr, &nb, nullptr, st);
  else rc = launch_distortion<1>(e, X, m, nullptr, nullptr, &nb, nullptr, st);
  if (rc) return rc;
  if (loss_sum) {  // NULL: leave the per-block partials (kernel-only timing)
    add_partials_kernel<<<1, 256, 0, st>>>(e->loss_partials, nb, loss_sum);
    MDE_LAUNCH_CHECK();
  }
  return 0;
}

int mde_function_eval(const mde_fn_t* fn, const float* par0, int64_t par0_len, const float* par1,
                      const float* distances, int64_t p, float* f, float* fprime, void* stream) {
  if (!fn || !par0 || !distances || p < 0 || (par0_len != 1 && par0_len != p)) return MDE_E_INVALID;
  if (p == 0) return 0;
  int tb = 256, nb = ceil_div_i64(p, tb);
  function_eval_kernel<<<nb, tb, 0, (cudaStream_t)stream>>>(to_dev(*fn), par0, par0_len, par1, distances, p, f, fprime);
  MDE_LAUNCH_CHECK();
  return 0;
}

int mde_scatter_external(const mde_edges_t* e, const float* X, int m, const float* g, float* grad,
                         void* stream) {
  if (!e || !X || !g || !grad || m < 1) return MDE_E_INVALID;
  if (e->kind == 2) return pull_launch(2, e, X, m, grad, g, nullptr, nullptr, (cudaStream_t)stream);
  if (e->kind == 1) return tiled_launch(2, e, X, m, grad, g, nullptr, nullptr, (cudaStream_t)stream);
  return launch_distortion<2>(e, X, m, grad, g, nullptr, nullptr, (cudaStream_t)stream);
}

int mde_edge_outputs(const mde_edges_t* e, const float* X, int m, float* distances, float* distortions,
                     void* stream) {
  if (!e || !X || m < 1) return MDE_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (e->kind == 2) return pull_edge_outputs(e, X, m, distances, distortions, st);
  if (e->kind == 1) return tiled_edge_outputs(e, X, m, distances, distortions, st);
  int tb = 256, nb = ceil_div_i64(e->p, tb);
  edge_outputs_kernel<0><<<nb, tb, 0, st>>>(e->src, e->dst, e->par0, e->has_par1 ? e->par1 : nullptr,
                                            e->perm, e->p, m, X, distances, distortions, e->fn);
  MDE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

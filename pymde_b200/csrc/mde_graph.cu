// mde_graph.cu -- hop-count shortest paths of an unweighted graph on the device.
//
// Replaces pymde/preprocess/graph.py:310-474 for unweighted graphs: the reference runs ONE breadth-first search
// per node (Cython, pymde/preprocess/_graph.pyx:10-52) in a multiprocessing pool and keeps, for node s, the
// distances to the nodes v > s, each with probability `retain_fraction`.
//
// Here the searches are bit-parallel: a batch of 256 sources advances together, node v holding 4 x 64-bit words of
// "reached by source b" bits.  One level is one pass over the CSR adjacency:
//     next[v] = (OR over neighbours u of frontier[u]) & ~visited[v]
// (64 sources per 8-byte load, no atomics on the frontier), and every newly set bit (s, v) with v > s is a
// finished shortest path of `level` hops; it is kept when a counter-based hash of (seed, s, v) falls under
// `retain` and appended through one warp-aggregated atomic per warp.  The output order depends on the schedule
// (the SET of triples does not); the caller sorts by (s, v).
#include <cstdint>
#include <cstdlib>

#include "mde_common.cuh"

using namespace mde;

namespace {

constexpr int kWords = 4;               // 64-bit words per node: 256 sources per batch
constexpr int kBatch = 64 * kWords;

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void hops_init_kernel(uint64_t* __restrict__ visited, uint64_t* __restrict__ front, int64_t n,
                                 int64_t s0, int nsrc) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n * kWords) return;
  const int64_t v = k / kWords;
  const int w = (int)(k % kWords);
  uint64_t bits = 0ull;
  const int64_t b = v - s0;  // source b of the batch is node s0 + b
  if (b >= 0 && b < nsrc && (b >> 6) == w) bits = 1ull << (b & 63);
  visited[k] = bits;
  front[k] = bits;
}

__global__ void __launch_bounds__(256)
hops_level_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                  const uint64_t* __restrict__ front_in, uint64_t* __restrict__ visited,
                  uint64_t* __restrict__ front_out, int64_t n, int64_t s0, int level, int emit, uint64_t seed,
                  uint64_t thresh, int32_t* __restrict__ out_src, int32_t* __restrict__ out_dst,
                  float* __restrict__ out_len, int64_t cap, unsigned long long* __restrict__ count,
                  int* __restrict__ any) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  uint64_t fresh = 0ull;
  int64_t v = 0;
  int w = 0;
  if (k < n * kWords) {
    v = k / kWords;
    w = (int)(k % kWords);
    uint64_t acc = 0ull;
    const int e0 = indptr[v], e1 = indptr[v + 1];
    for (int e = e0; e < e1; ++e) acc |= front_in[(int64_t)indices[e] * kWords + w];
    fresh = acc & ~visited[k];
    front_out[k] = fresh;
    if (fresh) { visited[k] |= fresh; *any = 1; }
  }
  if (!emit) return;
  // finished paths (s, v), v > s, kept with probability thresh / 2^64
  uint64_t keep = 0ull;
  uint64_t bits = fresh;
  while (bits) {
    const int b = __ffsll((long long)bits) - 1;
    bits &= bits - 1;
    const int64_t s = s0 + 64 * w + b;
    if (v > s && (thresh == ~0ull || splitmix64(seed ^ ((uint64_t)s * (uint64_t)n + (uint64_t)v)) < thresh))
      keep |= 1ull << b;
  }
  const int cnt = __popcll(keep);
  // warp-aggregated append
  int incl = cnt;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int o = __shfl_up_sync(kFull, incl, off);
    if (lane >= off) incl += o;
  }
  const int total = __shfl_sync(kFull, incl, 31);
  unsigned long long base = 0ull;
  if (lane == 31 && total > 0) base = atomicAdd(count, (unsigned long long)total);
  base = __shfl_sync(kFull, base, 31);
  unsigned long long pos = base + (unsigned long long)(incl - cnt);
  while (keep) {
    const int b = __ffsll((long long)keep) - 1;
    keep &= keep - 1;
    if ((int64_t)pos < cap) {
      out_src[pos] = (int32_t)(s0 + 64 * w + b);
      out_dst[pos] = (int32_t)v;
      out_len[pos] = (float)level;
    }
    ++pos;
  }
}

}  // namespace

extern "C" {

int64_t mde_graph_hops_ws_bytes(int64_t n) { return 3 * n * kWords * (int64_t)sizeof(uint64_t) + 64; }

int mde_graph_hops(const int32_t* indptr, const int32_t* indices, int64_t n, int64_t s_begin, int64_t s_end,
                   int max_length, double retain, uint64_t seed, int32_t* out_src, int32_t* out_dst,
                   float* out_len, int64_t cap, unsigned long long* count_dev, void* ws, int64_t ws_bytes,
                   void* stream) {
  if (!indptr || !indices || n < 1 || s_begin < 0 || s_end > n || s_begin > s_end || !count_dev || !ws ||
      ws_bytes < mde_graph_hops_ws_bytes(n) || n >= (1ll << 31) || (cap > 0 && (!out_src || !out_dst || !out_len)))
    return MDE_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  uint64_t* visited = reinterpret_cast<uint64_t*>(ws);
  uint64_t* fa = visited + n * kWords;
  uint64_t* fb = fa + n * kWords;
  int* any_d = reinterpret_cast<int*>(fb + n * kWords);
  const uint64_t thresh = (retain >= 1.0) ? ~0ull : (uint64_t)(retain * 18446744073709551616.0);
  const int limit = (max_length <= 0) ? 0x7fffffff : max_length;
  const int tb = 256;
  const int nb = ceil_div_i64(n * kWords, tb);
  int* any_h = nullptr;
  MDE_CUDA_TRY(cudaMallocHost(&any_h, sizeof(int)));
  int rc = 0;
  for (int64_t s0 = s_begin; s0 < s_end && !rc; s0 += kBatch) {
    const int nsrc = (int)((s_end - s0) < kBatch ? (s_end - s0) : kBatch);
    hops_init_kernel<<<nb, tb, 0, st>>>(visited, fa, n, s0, nsrc);
    ++g_launch_count;
    uint64_t *fin = fa, *fout = fb;
    for (int level = 1; level <= limit; ++level) {
      cudaError_t err = cudaMemsetAsync(any_d, 0, sizeof(int), st);
      if (err != cudaSuccess) { rc = (int)err; break; }
      hops_level_kernel<<<nb, tb, 0, st>>>(indptr, indices, fin, visited, fout, n, s0, level, 1, seed, thresh, out_src,
                                          out_dst, out_len, cap, count_dev, any_d);
      ++g_launch_count;
      err = cudaMemcpyAsync(any_h, any_d, sizeof(int), cudaMemcpyDeviceToHost, st);
      if (err == cudaSuccess) err = cudaStreamSynchronize(st);
      if (err != cudaSuccess) { rc = (int)err; break; }
      if (!*any_h) break;  // every search of the batch has stopped growing
      uint64_t* tmp = fin; fin = fout; fout = tmp;
    }
  }
  cudaFreeHost(any_h);
  if (!rc) { cudaError_t e = cudaPeekAtLastError(); if (e != cudaSuccess) rc = (int)e; }
  return rc;
}

}  // extern "C"

// mde_knn.cu -- exact k-nearest neighbours of the rows of a data matrix.
//
// Replaces the neighbour search of pymde/preprocess/data_matrix.py:91-178 (scikit-learn brute force below 10 000
// rows, pynndescent above) by an exact search whose cross terms run on the 5th-generation tensor cores:
//
//   prep     X (n x d fp32) -> Xh, Xl (n_pad x K_pad bf16, x = hi + lo to 2^-16), ||x||^2 (fp32; +inf on padding)
//   tiles    one CTA per 128 query rows sweeps ALL candidates in tiles of 256:
//              TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) stages 64-wide K blocks of the hi and lo parts of both
//              operands in shared memory (2 stages x 96 KB), one elected thread issues
//              tcgen05.mma.cta_group::1.kind::f16  D[128 x 256] += Ah Bh^T + Ah Bl^T + Al Bh^T   (fp32 in TMEM),
//              and four epilogue warps read the finished accumulator with tcgen05.ld (the second accumulator is
//              being filled meanwhile: 2 x 256 TMEM columns), form ||x||^2 - 2 q.x and keep, per query row, the
//              KK = 32 smallest in a thread-private list.  The n x n distance matrix never exists.
//   re-rank  exact fp32 sum (q - x)^2 of the KK candidates of a row (one warp per row), k smallest, ascending.
//
// The bf16 x 3 split leaves an error of ~2^-16 |q||x| on a cross term, far below the gap between the k-th and the
// (k + 8)-th neighbour of real data, and the re-rank removes it from the result: the neighbour lists are those of a
// brute-force fp32 search (ties and fp32 rounding aside).  k <= 24.
//
// Hangs are not an option on a shared GPU: every mbarrier wait is bounded (mde_tma.cuh) and traps.
#include <cuda.h>  // CUtensorMap and its enums (types only: the encoder is fetched with cudaGetDriverEntryPoint)
#include <cuda_bf16.h>

#include <cstdint>
#include <cstdlib>

#include "mde_common.cuh"
#include "mde_tma.cuh"

using namespace mde;

namespace {

constexpr int kTileM = 128;                 // query rows per CTA = TMEM lanes
constexpr int kTileN = 256;                 // candidates per accumulator = TMEM columns
constexpr int kBlockK = 64;                 // bf16 elements per 128-byte swizzle row
constexpr int kUmmaK = 16;                  // K of one tcgen05.mma.kind::f16
constexpr int kStages = 2;
constexpr int kKK = 32;                     // candidates kept per row before the exact re-rank
constexpr int kMaxK = 24;
constexpr int kRowBytes = kBlockK * 2;      // 128
constexpr int kABytes = kTileM * kRowBytes; // 16 KB: one 128-row operand block (hi or lo)
constexpr int kStageBytes = 6 * kABytes;    // A hi, A lo, B hi (2 x 128 rows), B lo (2 x 128 rows) = 96 KB
constexpr int kThreads = 256;               // warp 0: TMA, warp 1: MMA, warp 2: TMEM allocation, warps 4-7: epilogue
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /* alignment slack */ + 2 * kTileN * 4 /* norms */ + 128;

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers (tcgen05 / tensor TMA); mbarriers come from mde_tma.cuh
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major; issued by ONE thread for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor of a K-major operand block staged by TMA with the 128-byte swizzle: rows of 128
// bytes, 8-row groups 1024 bytes apart (SBO), one swizzle atom along K (LBO unused), descriptor version 1 (sm_100),
// layout type 2 = SWIZZLE_128B.  Field layout: cute/arch/mma_sm100_desc.hpp (UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (UMMA::InstrDescriptor): fp32 accumulate, bf16 x bf16, both K-major, M = 128, N = 256
constexpr uint32_t kInstrDesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

// ---------------------------------------------------------------------------------------------------------------
// prep: bf16 hi / lo split (zero padded to n_pad x k_pad) and squared norms (+inf on padded rows)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
knn_prep_kernel(const float* __restrict__ X, int64_t n, int d, int64_t n_pad, int k_pad, __nv_bfloat16* __restrict__ Xh,
                __nv_bfloat16* __restrict__ Xl, float* __restrict__ norms) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_pad) return;
  float acc = 0.0f;
  for (int c = lane; c < k_pad; c += 32) {
    const float x = (row < n && c < d) ? X[row * d + c] : 0.0f;
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    Xh[row * k_pad + c] = h;
    Xl[row * k_pad + c] = l;
    acc += x * x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
  if (lane == 0) norms[row] = (row < n) ? acc : __int_as_float(0x7f800000);
}

// ---------------------------------------------------------------------------------------------------------------
// tiles: tensor-core cross terms + running top-KK per query row
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
knn_tile_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_l,
                const float* __restrict__ norms, int64_t n, int64_t n_pad, int k_pad, int32_t* __restrict__ cand_idx,
                float* __restrict__ cand_val) {
  extern __shared__ uint8_t smem_raw[];
  // carve: [stages x 96 KB, 1024-aligned] | norms[2][256] | barriers | tmem base
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* s_norm = reinterpret_cast<float*>(gen + kStages * kStageBytes);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(gen + kStages * kStageBytes + 2 * kTileN * 4);
  const uint32_t bar0 = smem_u32(s_bar);
  // barriers: full[s] = bar0 + 8 s, empty[s] = bar0 + 16 + 8 s, tfull[a] = bar0 + 32 + 8 a, tempty[a] = bar0 + 48 + 8 a
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = k_pad / kBlockK;
  const int num_tiles = (int)(n_pad / kTileN);
  const int row0 = blockIdx.x * kTileM;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar0 + 16 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar0 + 32 + 8 * a, 1); mbar_init(bar0 + 48 + 8 * a, 4); }
    fence_mbar_init();
  } else if (warp == 2) {
    tmem_alloc(smem_u32(s_tmem), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    int stage = 0; uint32_t phase = 0;
    for (int t = 0; t < num_tiles; ++t) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(bar0 + 16 + 8 * stage, phase ^ 1);  // slot released by the MMA thread
        const uint32_t full = bar0 + 8 * stage;
        const uint32_t dst = base + stage * kStageBytes;
        mbar_expect_tx(full, kStageBytes);
        tma_load_2d(dst, &map_h, kb * kBlockK, row0, full);
        tma_load_2d(dst + kABytes, &map_l, kb * kBlockK, row0, full);
        tma_load_2d(dst + 2 * kABytes, &map_h, kb * kBlockK, t * kTileN, full);
        tma_load_2d(dst + 3 * kABytes, &map_h, kb * kBlockK, t * kTileN + kTileM, full);
        tma_load_2d(dst + 4 * kABytes, &map_l, kb * kBlockK, t * kTileN, full);
        tma_load_2d(dst + 5 * kABytes, &map_l, kb * kBlockK, t * kTileN + kTileM, full);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    int stage = 0; uint32_t phase = 0;
    for (int t = 0; t < num_tiles; ++t) {
      const int a = t & 1;
      mbar_wait(bar0 + 48 + 8 * a, (uint32_t)((t >> 1) & 1) ^ 1);  // accumulator drained by the epilogue
      tc_fence_after();
      const uint32_t tacc = tmem + (uint32_t)(a * kTileN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(bar0 + 8 * stage, phase);  // operands landed
        tc_fence_after();
        const uint32_t sa = base + stage * kStageBytes;
        const uint64_t ah = smem_desc_sw128(sa), al = smem_desc_sw128(sa + kABytes);
        const uint64_t bh = smem_desc_sw128(sa + 2 * kABytes), bl = smem_desc_sw128(sa + 4 * kABytes);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          const uint64_t adv = (uint64_t)((k * kUmmaK * 2) >> 4);  // 32 bytes per K step inside the swizzle atom
          umma_bf16(tacc, ah + adv, bh + adv, kInstrDesc, (kb | k) != 0);
          umma_bf16(tacc, ah + adv, bl + adv, kInstrDesc, 1u);
          umma_bf16(tacc, al + adv, bh + adv, kInstrDesc, 1u);
        }
        umma_commit(bar0 + 16 + 8 * stage);                 // frees the smem slot when these MMAs are done
        if (kb == num_kb - 1) umma_commit(bar0 + 32 + 8 * a);  // accumulator complete
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread <-> query row (TMEM lane) =====
    const int et = threadIdx.x - 128;            // 0..127
    const int row = row0 + et;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    float bd[kKK];
    int bi[kKK];
#pragma unroll
    for (int q = 0; q < kKK; ++q) { bd[q] = __int_as_float(0x7f800000); bi[q] = -1; }
    float thr = __int_as_float(0x7f800000);
    int worst = 0;
    for (int t = 0; t < num_tiles; ++t) {
      const int a = t & 1;
      float* sn = s_norm + a * kTileN;
      // norms of this tile's candidates (the buffer of tile t - 2 was released by the barrier below)
      sn[et] = __ldg(norms + (int64_t)t * kTileN + et);
      sn[et + 128] = __ldg(norms + (int64_t)t * kTileN + 128 + et);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(bar0 + 32 + 8 * a, (uint32_t)((t >> 1) & 1));
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < kTileN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_base + (uint32_t)(a * kTileN + c * 32), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float dist = fmaf(-2.0f, __uint_as_float(v[i]), sn[c * 32 + i]);
          if (dist < thr) {
            const int col = t * kTileN + c * 32 + i;
            if (col != row) {
              bd[worst] = dist; bi[worst] = col;
              float m = bd[0]; int w = 0;
#pragma unroll
              for (int q = 1; q < kKK; ++q) { if (bd[q] > m) { m = bd[q]; w = q; } }
              thr = m; worst = w;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + 48 + 8 * a);  // 4 warps -> accumulator free
    }
    if (row < n) {
#pragma unroll
      for (int q = 0; q < kKK; ++q) {
        cand_idx[(int64_t)row * kKK + q] = bi[q];
        cand_val[(int64_t)row * kKK + q] = bd[q];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_free(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// re-rank: exact fp32 squared distances of a row's candidates, the k smallest in ascending order
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
knn_rerank_kernel(const float* __restrict__ X, int64_t n, int d, const int32_t* __restrict__ cand_idx, int k,
                  int32_t* __restrict__ out_idx, float* __restrict__ out_d2) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int mine = cand_idx[row * kKK + lane];  // lane q owns candidate q
  float my_d = __int_as_float(0x7f800000);
  const float* xq = X + row * d;
  for (int q = 0; q < kKK; ++q) {
    const int c = __shfl_sync(kFull, mine, q);
    if (c < 0) continue;  // (warp-uniform)
    const float* xc = X + (int64_t)c * d;
    float acc = 0.0f;
    for (int j = lane; j < d; j += 32) { const float t = xq[j] - xc[j]; acc = fmaf(t, t, acc); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
    if (lane == q) my_d = acc;
  }
  // rank of (my_d, mine) among the 32 candidates: ties broken by index, missing candidates last
  int rank = 0;
  for (int q = 0; q < kKK; ++q) {
    const float od = __shfl_sync(kFull, my_d, q);
    const int oi = __shfl_sync(kFull, mine, q);
    if (q != lane && (od < my_d || (od == my_d && (unsigned)oi < (unsigned)mine))) ++rank;
  }
  if (rank < k) {
    out_idx[row * k + rank] = mine;
    out_d2[row * k + rank] = my_d;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(EncodeTiledFn enc, CUtensorMap* map, void* ptr, int64_t n_pad, int k_pad) {
  const cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)n_pad};
  const cuuint64_t strides[1] = {(cuuint64_t)k_pad * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)kTileM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MDE_E_INVALID;
}

struct KnnLayout {
  int64_t n_pad; int k_pad;
  size_t off_h, off_l, off_norm, off_ci, off_cv, total;
};

KnnLayout knn_layout(int64_t n, int d) {
  KnnLayout L;
  L.n_pad = (n + kTileN - 1) / kTileN * kTileN;
  L.k_pad = (d + kBlockK - 1) / kBlockK * kBlockK;
  auto up = [](size_t x) { return (x + 1023) / 1024 * 1024; };
  size_t o = 0;
  L.off_h = o; o = up(o + (size_t)L.n_pad * L.k_pad * 2);
  L.off_l = o; o = up(o + (size_t)L.n_pad * L.k_pad * 2);
  L.off_norm = o; o = up(o + (size_t)L.n_pad * 4);
  L.off_ci = o; o = up(o + (size_t)n * kKK * 4);
  L.off_cv = o; o = up(o + (size_t)n * kKK * 4);
  L.total = o;
  return L;
}

}  // namespace

extern "C" {

int mde_knn_max_k(void) { return kMaxK; }

int mde_knn_ws_bytes(int64_t n, int d, size_t* bytes) {
  if (!bytes || n < 2 || d < 1) return MDE_E_INVALID;
  *bytes = knn_layout(n, d).total;
  return 0;
}

int mde_knn(const float* X, int64_t n, int d, int k, int32_t* idx_out, float* d2_out, void* ws, size_t ws_bytes,
            void* stream) {
  if (!X || !idx_out || !d2_out || !ws || n < 2 || d < 1 || k < 1 || k > kMaxK || k > n - 1) return MDE_E_INVALID;
  if (n > (1ll << 31) - kTileN) return MDE_E_UNSUPPORTED;
  const KnnLayout L = knn_layout(n, d);
  if (ws_bytes < L.total || (reinterpret_cast<uintptr_t>(ws) & 1023)) return MDE_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MDE_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return MDE_E_UNSUPPORTED;
    enc = (EncodeTiledFn)fn;
  }
  uint8_t* w = static_cast<uint8_t*>(ws);
  __nv_bfloat16* Xh = reinterpret_cast<__nv_bfloat16*>(w + L.off_h);
  __nv_bfloat16* Xl = reinterpret_cast<__nv_bfloat16*>(w + L.off_l);
  float* norms = reinterpret_cast<float*>(w + L.off_norm);
  int32_t* ci = reinterpret_cast<int32_t*>(w + L.off_ci);
  float* cv = reinterpret_cast<float*>(w + L.off_cv);
  CUtensorMap mh, ml;
  int rc;
  if ((rc = make_map(enc, &mh, Xh, L.n_pad, L.k_pad))) return rc;
  if ((rc = make_map(enc, &ml, Xl, L.n_pad, L.k_pad))) return rc;
  knn_prep_kernel<<<(unsigned)((L.n_pad + 7) / 8), 256, 0, st>>>(X, n, d, L.n_pad, L.k_pad, Xh, Xl, norms);
  MDE_LAUNCH_CHECK();
  static bool attr_set = false;
  if (!attr_set) {
    MDE_CUDA_TRY(cudaFuncSetAttribute(knn_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const unsigned grid = (unsigned)((n + kTileM - 1) / kTileM);
  knn_tile_kernel<<<grid, kThreads, kSmemBytes, st>>>(mh, ml, norms, n, L.n_pad, L.k_pad, ci, cv);
  MDE_LAUNCH_CHECK();
  knn_rerank_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(X, n, d, ci, k, idx_out, d2_out);
  MDE_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

// mde_ell.cu -- owner-per-lane PULL kernel ("ELL records"): fused average distortion (forward + backward) for
// m <= 4 with no atomics inside a lane, no run detection and ~half the instructions of mde_pull.cu.
//
// Replaces pymde/average_distortion.py:36-80 (gather, norm, per-edge f, mean, scatter-add).  Like mde_pull.cu every
// undirected edge {i, j} is stored as the two directed entries (owner i, neighbour j) and (owner j, neighbour i) and an
// entry only contributes to the gradient row of its OWNER.  What changes is who walks the entries:
//
//   * vertices are cut into neighbour tiles of R = 2^rb rows (64 KB of X, resident in shared memory);
//   * the entries of one (neighbour tile, class, owner) group are cut into LANE-SLOTS of at most 8 entries;
//     the lane-slots of a (tile, class) are sorted by length and packed 32 at a time into one RECORD:
//         int32 W, cls, K, nslots | u32 owner word[K][32] | K x W/2 x ( float2 w[32] | u32 neighbour pair[32] )
//     (owner word: row | count << 24 | duplicate << 31; neighbour pair: two u16 BYTE offsets of the neighbour rows
//     inside the tile; W even <= 8; K = 4 / 2 / 1 lane-slots PER LANE for W = 2 / 4 / >= 6, so that short lane-slots
//     do not pay the per-record work alone; 16 + K (128 + 192 W) bytes; column-major, so lane l reads word l of every
//     column: conflict-free).  6 bytes per directed entry = 12 bytes per edge, the size of the
//     sorted-SoA stream;
//   * ONE LANE owns one lane-slot: it loads its owner row once (LDG through L1), walks its W entries (two LDS for a
//     pair of entries + one LDS gather of the neighbour row each), keeps the gradient sum of the owner in registers
//     and leaves with ONE vector red.  No owner compare, no run flush, no u16 owner offsets, no canonical-direction
//     test (the loss is half the sum over directed entries), the trip count W is warp-uniform;
//   * pads (a lane-slot shorter than W, lanes past the last lane-slot of a group) carry w = 0 and repeat a real
//     neighbour; the generic (non-weight) functions mask them with the per-lane count instead.
//
// Records are streamed with cp.async.bulk (TMA) into two shared-memory slots per warp (mbarrier per slot, the next
// record is in flight while the current one is consumed straight from shared memory); a slot is refilled after the
// warp's last consuming instruction has issued, i.e. after every lane's LDS has returned.
//
// The layout is built on the HOST from the sorted-SoA arrays (ell_build_host: counting sorts, O(p + tiles * n)); the
// same function is exported for the CPU tests (mde_ell_host_layout), which decode the records and compare the pull
// sums with the oracle.  The sorted-SoA layout stays alongside (kind 3 = SoA + ELL): value-only evaluation, per-edge
// outputs and external coefficients run on the SoA kernels.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mde_edges.cuh"
#include "mde_tma.cuh"

using namespace mde;

namespace {

constexpr int kEllWarps = 32;
constexpr int kEllThreads = kEllWarps * 32;
constexpr int kEllWmax = 8;                                            // entries per lane-slot
constexpr int kEllPair = 256 + 128;                                    // one column pair: float2 w[32] | u32 idx[32]
__host__ __device__ constexpr int ell_kmax(int W) { return W <= 2 ? 4 : (W <= 4 ? 2 : 1); }  // lane-slots per lane
__host__ __device__ constexpr int ell_rec_bytes(int W, int K) { return 16 + K * (128 + (W / 2) * kEllPair); }
constexpr int kEllSlotBytes = ell_rec_bytes(2, 4);                     // 2064: the largest record (W 8, K 1: 1680)
static_assert(ell_rec_bytes(4, 2) <= kEllSlotBytes && ell_rec_bytes(8, 1) <= kEllSlotBytes &&
              ell_rec_bytes(6, 1) <= kEllSlotBytes, "slot size");
constexpr uint32_t kOwnMask = 0x00ffffffu;

int ell_pack_enabled() {  // MDE_B200_ELL_PACK=0: one lane-slot per lane in every record (A/B)
  const char* e = getenv("MDE_B200_ELL_PACK");
  return !(e && e[0] == '0');
}

struct EllHost {
  std::vector<unsigned char> rec;
  std::vector<uint32_t> rec_off;  // [nrec + 1], units of 16 bytes
  std::vector<int32_t> bkt_tile, bkt_wt0, cta_wt0, cta_bkt0;
  int rb = 0, ncta = 0;
  int64_t nrec = 0, nslots = 0, nentries = 0, npadded = 0;
};

// Pure host code.  src/dst/par0: p canonical edges in any order.  Returns 0 or MDE_E_UNSUPPORTED.
int ell_build_host(int64_t n, int64_t p, int m, const int32_t* src, const int32_t* dst, const float* par0,
                   int push_pull, int rb, int max_cta, EllHost& out) {
  if (m < 1 || m > 4 || n < 1 || p < 1 || n >= (1ll << 24) || p >= (1ll << 29)) return MDE_E_UNSUPPORTED;
  if (rb < 8 || rb > 15) return MDE_E_UNSUPPORTED;
  const int64_t R = 1ll << rb;
  const int64_t ndt = (n + R - 1) >> rb;
  if (ndt > 32) return MDE_E_UNSUPPORTED;
  const int64_t row_bytes = 4ll * m;
  if ((R - 1) * row_bytes > 65535) return MDE_E_UNSUPPORTED;  // neighbour rows are addressed by u16 byte offsets
  const int64_t ng = ndt * 2 * n;  // groups (tile, class, owner)
  std::vector<uint32_t> start((size_t)ng + 1, 0u);
  auto cls_of = [&](int64_t k) -> int64_t { return (push_pull && !(par0[k] >= 0.0f)) ? 1 : 0; };
  for (int64_t k = 0; k < p; ++k) {
    const int64_t s = src[k], d = dst[k], c = cls_of(k);
    if (s < 0 || d < 0 || s >= n || d >= n) return MDE_E_INVALID;
    ++start[(size_t)((((d >> rb) * 2 + c) * n) + s) + 1];
    ++start[(size_t)((((s >> rb) * 2 + c) * n) + d) + 1];
  }
  for (int64_t g = 0; g < ng; ++g) start[(size_t)g + 1] += start[(size_t)g];
  const int64_t p2 = 2 * p;
  std::vector<float> ew((size_t)p2);
  std::vector<uint16_t> ej((size_t)p2);
  {
    std::vector<uint32_t> fill(start.begin(), start.end() - 1);
    for (int64_t k = 0; k < p; ++k) {
      const int64_t s = src[k], d = dst[k], c = cls_of(k);
      uint32_t& f1 = fill[(size_t)((((d >> rb) * 2 + c) * n) + s)];
      ew[f1] = par0[k]; ej[f1] = (uint16_t)((d & (R - 1)) * row_bytes); ++f1;
      uint32_t& f2 = fill[(size_t)((((s >> rb) * 2 + c) * n) + d)];
      ew[f2] = par0[k]; ej[f2] = (uint16_t)((s & (R - 1)) * row_bytes); ++f2;
    }
  }
  struct Slot { uint32_t first; uint32_t own; uint8_t len; };
  std::vector<Slot> slots, sorted;
  std::vector<int64_t> rec_cost;
  out.rec.clear(); out.rec_off.assign(1, 0u);
  out.bkt_tile.clear(); out.bkt_wt0.clear();
  out.rb = rb; out.nslots = 0; out.nentries = p2; out.npadded = 0;
  for (int64_t tile = 0; tile < ndt; ++tile) {
    bool tile_open = false;
    for (int64_t c = 0; c < 2; ++c) {
      const int64_t g0 = (tile * 2 + c) * n;
      slots.clear();
      for (int64_t own = 0; own < n; ++own) {
        uint32_t a = start[(size_t)(g0 + own)];
        const uint32_t b = start[(size_t)(g0 + own) + 1];
        while (a < b) {
          const uint32_t len = std::min<uint32_t>(kEllWmax, b - a);
          slots.push_back({a, (uint32_t)own, (uint8_t)len});
          a += len;
        }
      }
      if (slots.empty()) continue;
      // stable counting sort by length, longest first (owners stay ascending inside one length)
      size_t cnt[kEllWmax + 2] = {0};
      for (const Slot& s : slots) ++cnt[kEllWmax - s.len + 1];
      for (int i = 0; i <= kEllWmax; ++i) cnt[i + 1] += cnt[i];
      sorted.resize(slots.size());
      for (const Slot& s : slots) sorted[cnt[kEllWmax - s.len]++] = s;
      out.nslots += (int64_t)sorted.size();
      if (!tile_open) {
        out.bkt_tile.push_back((int32_t)tile);
        out.bkt_wt0.push_back((int32_t)(out.rec_off.size() - 1));
        tile_open = true;
      }
      const int pack = ell_pack_enabled();
      for (size_t i0 = 0; i0 < sorted.size();) {
        const int W = (sorted[i0].len + 1) & ~1;
        const size_t left = sorted.size() - i0;
        const int K = (int)std::min<size_t>(pack ? ell_kmax(W) : 1, (left + 31) / 32);
        const int ns = (int)std::min<size_t>((size_t)32 * K, left);
        const size_t bytes = (size_t)ell_rec_bytes(W, K);
        const size_t off = out.rec.size();
        out.rec.resize(off + bytes, 0);
        unsigned char* r = out.rec.data() + off;
        int32_t hdr[4] = {W, (int32_t)c, K, ns};
        memcpy(r, hdr, 16);
        uint32_t* ow = reinterpret_cast<uint32_t*>(r + 16);
        unsigned char* cols = r + 16 + 128 * K;
        for (int i = 0; i < 32 * K; ++i) {
          const int k = i / 32, l = i % 32;
          const bool dup = i >= ns;
          const Slot& s = sorted[i0 + (dup ? 0 : i)];
          ow[i] = s.own | (dup ? 0x80000000u : ((uint32_t)s.len << 24));
          unsigned char* cb = cols + (size_t)k * (W / 2) * kEllPair;
          for (int e = 0; e < W; ++e) {
            const bool real = !dup && e < (int)s.len;
            const uint32_t src_e = s.first + (uint32_t)std::min<int>(e, (int)s.len - 1);
            float* wp = reinterpret_cast<float*>(cb + (e / 2) * kEllPair) + 2 * l + (e & 1);
            uint16_t* ip = reinterpret_cast<uint16_t*>(cb + (e / 2) * kEllPair + 256) + 2 * l + (e & 1);
            *wp = real ? ew[src_e] : 0.0f;
            *ip = ej[src_e];
          }
        }
        out.npadded += 32ll * K * W;
        // warp instructions (ncu source page, C2): per record, per lane-slot row, per entry column
        rec_cost.push_back(60 + (int64_t)K * (30 + (int64_t)W * (c ? 31 : 23)));
        out.rec_off.push_back((uint32_t)(out.rec.size() / 16));
        i0 += (size_t)ns;
      }
    }
  }
  out.nrec = (int64_t)out.rec_off.size() - 1;
  if (out.nrec < 1 || out.rec.size() >= (1ull << 35)) return MDE_E_UNSUPPORTED;
  out.bkt_wt0.push_back((int32_t)out.nrec);
  // persistent grid: contiguous record ranges of equal estimated cost; every tile a CTA has to load (its first one
  // and one per bucket boundary inside its range) is charged like kTileCost warp instructions
  const int ncta = (int)std::min<int64_t>(max_cta, std::max<int64_t>(1, (out.nrec + kEllWarps - 1) / kEllWarps));
  out.ncta = ncta;
  const int64_t kTileCost = 6000;
  std::vector<char> bkt_first((size_t)out.nrec, 0);
  for (size_t b = 0; b + 1 < out.bkt_wt0.size(); ++b) bkt_first[(size_t)out.bkt_wt0[b]] = 1;
  int64_t remaining = kTileCost * ((int64_t)out.bkt_tile.size() - 1 + ncta);
  for (int64_t c : rec_cost) remaining += c;
  out.cta_wt0.assign(ncta + 1, (int32_t)out.nrec);
  {
    int64_t r = 0;
    for (int c = 0; c < ncta; ++c) {
      out.cta_wt0[c] = (int32_t)r;
      if (c == ncta - 1) break;  // the last CTA takes what is left
      const int64_t budget = remaining / (ncta - c);
      int64_t acc = kTileCost;
      while (r < out.nrec) {
        const int64_t add = rec_cost[(size_t)r] + ((bkt_first[(size_t)r] && r != out.cta_wt0[c]) ? kTileCost : 0);
        if (acc + add / 2 > budget && r != out.cta_wt0[c]) break;
        acc += add;
        ++r;
      }
      remaining -= acc;
      if (remaining < 0) remaining = 0;
    }
  }
  out.cta_bkt0.resize(ncta);
  for (int c = 0; c < ncta; ++c) {
    const auto it = std::upper_bound(out.bkt_wt0.begin(), out.bkt_wt0.end(), out.cta_wt0[c]);
    int b = (int)(it - out.bkt_wt0.begin()) - 1;
    const int nb = (int)out.bkt_tile.size();
    out.cta_bkt0[c] = b < 0 ? 0 : (b >= nb ? nb - 1 : b);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
struct EllArgs {
  const unsigned char* rec;
  const uint32_t* rec_off;
  const int32_t* bkt_tile;
  const int32_t* bkt_wt0;
  const int4* cta_desc;  // per CTA: first record, end record, first bucket, neighbour tile of that bucket
  const float* X;
  float* grad;
  double* loss_partials;
  const int* flag;
  FnDev fn;
  float inv_p;
  int64_t n;
  int rb;
  int x_vec_ok;
};

// neighbour row at BYTE offset `off` of the resident tile
template <int M>
__device__ __forceinline__ void e_lds_row(const float* __restrict__ Xt, uint32_t off, float (&o)[M]) {
  const unsigned char* q = reinterpret_cast<const unsigned char*>(Xt) + off;
  if constexpr (M == 2) { const float2 t = *reinterpret_cast<const float2*>(q); o[0] = t.x; o[1] = t.y; }
  else if constexpr (M == 4) { const float4 t = *reinterpret_cast<const float4*>(q); o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w; }
  else {
#pragma unroll
    for (int c = 0; c < M; ++c) o[c] = reinterpret_cast<const float*>(q)[c];
  }
}
template <int M>
__device__ __forceinline__ void e_ldg_row(const float* __restrict__ X, uint32_t r, float (&o)[M]) {
  if constexpr (M == 1) { o[0] = __ldg(X + r); }
  else if constexpr (M == 2) { const float2 t = __ldg(reinterpret_cast<const float2*>(X) + r); o[0] = t.x; o[1] = t.y; }
  else if constexpr (M == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(X) + r); o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w; }
  else {
#pragma unroll
    for (int c = 0; c < M; ++c) o[c] = __ldg(X + (size_t)r * M + c);
  }
}
template <int M>
__device__ __forceinline__ void e_red_row(float* __restrict__ G, uint32_t r, const float (&v)[M]) {
  if constexpr (M == 1) red_add(G + r, v[0]);
  else if constexpr (M == 2) red_add_v2(G + 2 * (size_t)r, v[0], v[1]);
  else if constexpr (M == 4) red_add_v4(G + 4 * (size_t)r, v[0], v[1], v[2], v[3]);
  else {
#pragma unroll
    for (int c = 0; c < M; ++c) red_add(G + (size_t)r * M + c, v[c]);
  }
}

// PushAndPull(Log1p(1.5), Log(1.0)), MUFU math, class known at compile time.  Same formulas as
// mde_common.cuh::edge_coeff_fast_log1p_log; returns the loss in log2 units (flog2 = w lg2(.), the block sum is
// multiplied by ln 2 at the end) and gs = f'/(p d) WITHOUT the class constant (1.5/p attractive, 1/p repulsive: applied
// once per lane-slot to the accumulated row); gs = 0 at d = 0 (the reference replaces the non-finite g, and the
// difference vector is 0).
template <int CLS>
__device__ __forceinline__ void ell_fast_coeff(float d2, float w, float& flog2, float& gs) {
  const float kLog2e = 1.44269504089f;
  if constexpr (CLS == 0) {
    // d2 = 0: rsqrt of the clamp is finite, d = 0 * finite = 0, sqrt(0) = 0 -> gs = 0 without a select
    const float rs = fast_rsqrt(fmaxf(d2, 1e-30f));
    const float d = d2 * rs;
    const float sd = fast_sqrt(d);
    const float one_p = fmaf(d, sd, 1.0f);
    flog2 = w * fast_lg2(one_p);
    gs = (sd * rs) * fast_rcp(one_p) * w;
  } else {
    const float rs = fast_rsqrt(d2);
    const bool pos = d2 > 0.0f;
    const float d = pos ? d2 * rs : 0.0f;
    const float em = fast_ex2(-d * kLog2e);
    float one_m = 1.0f - em;
    const float series = d * (1.0f - d * (0.5f - d * (0.16666667f - d * 0.041666668f)));
    one_m = (d < 0.0625f) ? series : one_m;
    flog2 = w * fast_lg2(one_m);
    const float gg = (rs * em) * fast_rcp(one_m) * w;
    gs = pos ? gg : 0.0f;
  }
}

template <int M, int FA, int FR, bool FAST, int CLS>
__device__ __forceinline__ void ell_entry(const EllArgs& a, const float* __restrict__ Xt, const float (&xi)[M], float w,
                                          uint32_t j, bool valid, float (&acc)[M], float& lf) {
  float xj[M], diff[M];
  e_lds_row<M>(Xt, j, xj);
  float d2 = 0.0f;
#pragma unroll
  for (int c = 0; c < M; ++c) { diff[c] = xi[c] - xj[c]; d2 = fmaf(diff[c], diff[c], d2); }
  float f, g;
  if constexpr (FAST) {
    ell_fast_coeff<CLS>(d2, w, f, g);  // pads: w = 0
  } else {
    // d and 1/d from ONE rsqrt.approx (<= 2 ulp) instead of an IEEE sqrt and an IEEE division (18 instructions);
    // d = 0: g is non-finite -> 1 like the reference (average_distortion.py:55-62), the difference vector is 0
    const float rs = fast_rsqrt(d2);
    const float d = (d2 > 0.0f) ? d2 * rs : 0.0f;
    float fp;
    edge_f_fp<FA, FR>(a.fn, d, w, 0.0f, f, fp);
    g = (fp * a.inv_p) * rs;
    if (!isfinite(g)) g = 1.0f;
    f = valid ? f : 0.0f;
    g = valid ? g : 0.0f;
  }
  lf += f;
#pragma unroll
  for (int c = 0; c < M; ++c) acc[c] = fmaf(g, diff[c], acc[c]);
}

// all W entries of this lane's lane-slot, straight from the shared-memory slot
template <int M, int FA, int FR, bool FAST, int CLS>
__device__ __forceinline__ void ell_columns(const EllArgs& a, const float* __restrict__ Xt, const unsigned char* cols,
                                            int lane, int W, int cnt, const float (&xi)[M], float (&acc)[M],
                                            float& lf) {
  const float2* wp = reinterpret_cast<const float2*>(cols) + lane;
  const uint32_t* ip = reinterpret_cast<const uint32_t*>(cols + 256) + lane;
#pragma unroll 2
  for (int c2 = 0; 2 * c2 < W; ++c2) {
    const float2 w2 = wp[c2 * (kEllPair / 8)];
    const uint32_t ix = ip[c2 * (kEllPair / 4)];
    ell_entry<M, FA, FR, FAST, CLS>(a, Xt, xi, w2.x, ix & 0xffffu, 2 * c2 < cnt, acc, lf);
    ell_entry<M, FA, FR, FAST, CLS>(a, Xt, xi, w2.y, ix >> 16, 2 * c2 + 1 < cnt, acc, lf);
  }
}

template <int M, int FA, int FR, bool FAST>
__global__ void __launch_bounds__(kEllThreads, 1)
distortion_ell_kernel(const EllArgs a) {
  if (a.flag != nullptr && *a.flag == 0) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int R = 1 << a.rb;
  float* Xt = reinterpret_cast<float*>(smem_raw);
  unsigned char* slots = reinterpret_cast<unsigned char*>(Xt + R * M);
  uint64_t* bars = reinterpret_cast<uint64_t*>(slots + 2 * kEllWarps * kEllSlotBytes);
  double* red = reinterpret_cast<double*>(bars + 2 * kEllWarps + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x;
  const int4 dsc = __ldg(a.cta_desc + c);  // ONE dependent load before the first copies can be issued
  const int wt0 = dsc.x, wt1 = dsc.y;
  int bkt = dsc.z;

  if (threadIdx.x == 0) {
#pragma unroll 1
    for (int i = 0; i <= 2 * kEllWarps; ++i) mbar_init(smem_u32(bars + i), 1);
    fence_mbar_init();
  }
  __syncthreads();

  const uint64_t pol = policy_evict_first();
  unsigned char* my_slots = slots + (size_t)warp * 2 * kEllSlotBytes;
  const uint32_t my_slot0 = smem_u32(my_slots);
  const uint32_t my_bar0 = smem_u32(bars + 2 * warp), x_bar = smem_u32(bars + 2 * kEllWarps);
  uint32_t phbits = 0, xph = 0;

  // lane 0 streams this warp's records (wt0 + warp, + 32, ...) two deep
  auto issue = [&](uint32_t o0, uint32_t o1, int s) {
    const uint32_t bytes = (o1 - o0) << 4;
    const uint32_t bar = my_bar0 + 8u * (uint32_t)s;
    mbar_expect_tx(bar, bytes);
    bulk_g2s_hint(my_slot0 + (uint32_t)s * kEllSlotBytes, a.rec + ((size_t)o0 << 4), bytes, bar, pol);
  };
  int tile = -1, seg_end = wt0;
  bool tile_pending = false;

  // CTA-wide: request neighbour tile `tl` (thread 0: bulk copies; everybody: the unaligned tail) / wait for it
  auto tile_issue = [&](int tl) {
    tile = tl;
    const int64_t base = (int64_t)tile << a.rb;
    const int64_t rows_l = a.n - base;
    const int rows = (int)(rows_l < (int64_t)R ? rows_l : (int64_t)R);
    const int nfl = rows * M;
    const float* xsrc = a.X + base * M;
    const uint32_t bytes = a.x_vec_ok ? (((uint32_t)nfl * 4u) & ~15u) : 0u;
    if (threadIdx.x == 0 && bytes > 0) {
      fence_proxy_async();
      mbar_expect_tx(x_bar, bytes);
      for (uint32_t off = 0; off < bytes; off += 32768u) {
        const uint32_t chunk = (bytes - off) < 32768u ? (bytes - off) : 32768u;
        bulk_g2s(smem_u32(Xt) + off, reinterpret_cast<const unsigned char*>(xsrc) + off, chunk, x_bar);
      }
    }
    for (int i = (int)(bytes >> 2) + threadIdx.x; i < nfl; i += kEllThreads) Xt[i] = __ldg(xsrc + i);
    tile_pending = true;
    return bytes;
  };
  uint32_t tile_bytes = 0;
  auto tile_wait = [&]() {
    __syncthreads();  // the plain-load tail of every thread
    if (tile_bytes > 0) { mbar_wait(x_bar, xph); xph ^= 1; }
    tile_pending = false;
  };
  // the tile of the first bucket is requested NOW: its copy overlaps the record-offset loads and the first record copies
  if (wt0 < wt1) tile_bytes = tile_issue(dsc.w);

  int t = wt0 + warp;
  if (lane == 0) {
    if (t < wt1) issue(__ldg(a.rec_off + t), __ldg(a.rec_off + t + 1), 0);
    if (t + kEllWarps < wt1) issue(__ldg(a.rec_off + t + kEllWarps), __ldg(a.rec_off + t + kEllWarps + 1), 1);
  }

  // make the neighbour tile of bucket `bkt` resident (same number of barriers for all warps)
  auto enter_bucket = [&]() {
    const int new_tile = __ldg(a.bkt_tile + bkt);
    const int be = __ldg(a.bkt_wt0 + bkt + 1);
    seg_end = be < wt1 ? be : wt1;
    if (new_tile != tile) {
      __syncthreads();  // every warp is done reading the old tile
      tile_bytes = tile_issue(new_tile);
    }
    if (tile_pending) tile_wait();
  };

  const float c_att = 1.5f * a.inv_p;
  double lsum = 0.0;
  bool first = true;
  int s = 0;

  for (; t < wt1; t += kEllWarps, s ^= 1) {
    // offsets of the record that will re-fill this slot (two records ahead): requested now, needed after the columns
    const int t2 = t + 2 * kEllWarps;
    uint32_t o0 = 0, o1 = 0;
    if (lane == 0 && t2 < wt1) { o0 = __ldg(a.rec_off + t2); o1 = __ldg(a.rec_off + t2 + 1); }
    // bucket change first: at the start of the kernel the tile copy then overlaps the first record copies
    while (t >= seg_end) {  // warp-uniform; CTA-wide barrier inside
      if (!first) ++bkt;
      enter_bucket();
      first = false;
    }
    mbar_wait(my_bar0 + 8u * (uint32_t)s, (phbits >> s) & 1u);
    phbits ^= 1u << s;
    const unsigned char* rec = my_slots + (size_t)s * kEllSlotBytes;
    const int4 hdr = *reinterpret_cast<const int4*>(rec);  // W, class, K, nslots (broadcast)
    const int W = hdr.x, K = hdr.z;
    const unsigned char* cols = rec + 16 + 128 * K;
    const int row_bytes = (W >> 1) * kEllPair;
    float lrec = 0.0f;
#pragma unroll 1
    for (int k = 0; k < K; ++k) {  // K lane-slots per lane (warp-uniform)
      const uint32_t ow = reinterpret_cast<const uint32_t*>(rec + 16)[k * 32 + lane];
      const uint32_t own = ow & kOwnMask;
      const int cnt = (int)((ow >> 24) & 0x7fu);
      float xi[M], acc[M];
      e_ldg_row<M>(a.X, own, xi);
#pragma unroll
      for (int q = 0; q < M; ++q) acc[q] = 0.0f;
      float lf = 0.0f;
      if (FAST) {
        if (hdr.y == 0) ell_columns<M, FA, FR, true, 0>(a, Xt, cols + k * row_bytes, lane, W, cnt, xi, acc, lf);
        else ell_columns<M, FA, FR, true, 1>(a, Xt, cols + k * row_bytes, lane, W, cnt, xi, acc, lf);
        const float cc = hdr.y == 0 ? c_att : a.inv_p;  // the class constant of f'/(p d), once per lane-slot
#pragma unroll
        for (int q = 0; q < M; ++q) acc[q] *= cc;
      } else {
        ell_columns<M, FA, FR, false, 2>(a, Xt, cols + k * row_bytes, lane, W, cnt, xi, acc, lf);
      }
      if (!(ow >> 31)) e_red_row<M>(a.grad, own, acc);
      lrec += lf;
    }
    lsum += (double)lrec;
    // every lane's loads of this slot have returned (their consumers above have issued): refill it
    __syncwarp();
    if (lane == 0 && t2 < wt1) issue(o0, o1, s);
  }
  if (first && wt0 < wt1) { enter_bucket(); first = false; }
  while (seg_end < wt1) { ++bkt; enter_bucket(); }
  {
    double v1[1] = {lsum};
    block_sum<1>(v1, red);
    // every undirected edge was seen from both ends
    if (threadIdx.x == 0) a.loss_partials[blockIdx.x] = v1[0] * (FAST ? 0.5 * 0.6931471805599453 : 0.5);
  }
}

size_t ell_smem_bytes(int rb, int m) {
  return (size_t)((size_t)1 << rb) * m * sizeof(float) + (size_t)2 * kEllWarps * kEllSlotBytes +
         (size_t)(2 * kEllWarps + 2) * sizeof(uint64_t) + 32 * sizeof(double);
}

template <int M>
const void* eselect_m(const FnDev& fn) {
  const int fa = fn.fn_att, fr = fn.fn_rep, pp = fn.push_pull;
#define EK(FA, FR, FAST) reinterpret_cast<const void*>(&distortion_ell_kernel<M, FA, FR, FAST>)
  if constexpr (M == 2 || M == 3) {
    const char* ev = getenv("MDE_B200_KERNEL");
    const bool precise = ev && !strcmp(ev, "precise");
    const bool hot = pp && fa == MDE_FN_P_LOG1P && fr == MDE_FN_P_LOG && fn.a0 == 1.5f && fn.r0 == 1.0f && !precise;
    if (hot) return EK(MDE_FN_P_LOG1P, MDE_FN_P_LOG, true);
    if (pp && fa == MDE_FN_P_LOG1P && fr == MDE_FN_P_LOG) return EK(MDE_FN_P_LOG1P, MDE_FN_P_LOG, false);
    if (!pp && fa == MDE_FN_P_QUADRATIC) return EK(MDE_FN_P_QUADRATIC, MDE_FN_P_QUADRATIC, false);
    if (!pp && fa == MDE_FN_L_QUADRATIC) return EK(MDE_FN_L_QUADRATIC, MDE_FN_L_QUADRATIC, false);
    if (!pp && fa == MDE_FN_L_HUBER) return EK(MDE_FN_L_HUBER, MDE_FN_L_HUBER, false);
  }
  return EK(-1, -1, false);
#undef EK
}
const void* eselect_kernel(const FnDev& fn, int m) {
  switch (m) {
    case 1: return eselect_m<1>(fn);
    case 2: return eselect_m<2>(fn);
    case 3: return eselect_m<3>(fn);
    case 4: return eselect_m<4>(fn);
  }
  return nullptr;
}
int econfigure_kernel(const void* k) {
  static std::vector<const void*> done;
  if (std::find(done.begin(), done.end(), k) != done.end()) return 0;
  cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (err != cudaSuccess) return (int)err;
  done.push_back(k);
  return 0;
}

int ell_default_rb(int m) { return (m <= 2) ? 13 : 12; }  // X tile of 64 KB (m = 1: 32 KB, m = 3: 48 KB)

int eenv_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

}  // namespace

namespace mde {

void ell_free(mde_edges* e) {
  cudaFree(e->ell_rec); cudaFree(e->ell_off); cudaFree(e->ell_bkt_tile); cudaFree(e->ell_bkt_wt0);
  cudaFree(e->ell_cta_desc);
  e->ell_rec = nullptr; e->ell_off = nullptr; e->ell_bkt_tile = nullptr; e->ell_bkt_wt0 = nullptr;
  e->ell_cta_desc = nullptr;
}

bool ell_supported(int64_t n, int m) {
  if (m < 1 || m > 4 || n >= (1ll << 24)) return false;
  int rb = ell_default_rb(m);
  { const int r = eenv_int("MDE_B200_TILE_RB", 0); if (r >= 8 && r <= 15) rb = r; }
  const int64_t R = 1ll << rb;
  return ((n + R - 1) >> rb) <= 32 && (R - 1) * 4 * m <= 65535 && ell_smem_bytes(rb, m) <= 227u * 1024u;
}

// Called on a finished sorted-SoA layout (kind 0): copies src / dst / par0 to the host, builds the ELL records there and
// uploads them; on success the layout becomes kind 3.  Returns 0, MDE_E_UNSUPPORTED (layout stays kind 0) or an error.
int ell_build(mde_edges* e, const mde_fn_t* fn, int m, cudaStream_t st) {
  if (e->kind != 0 || e->has_par1 || e->det || m < 1 || m > 4) return MDE_E_UNSUPPORTED;
  int rb = ell_default_rb(m);
  { const int r = eenv_int("MDE_B200_TILE_RB", 0); if (r >= 8 && r <= 15) rb = r; }
  if (ell_smem_bytes(rb, m) > 227u * 1024u) return MDE_E_UNSUPPORTED;
  const int64_t p = e->p, n = e->n;
  if (n >= (1ll << 24) || ((n + (1ll << rb) - 1) >> rb) > 32) return MDE_E_UNSUPPORTED;
  std::vector<int32_t> hs((size_t)p), hd((size_t)p);
  std::vector<float> hw((size_t)p);
  MDE_CUDA_TRY(cudaMemcpyAsync(hs.data(), e->src, sizeof(int32_t) * p, cudaMemcpyDeviceToHost, st));
  MDE_CUDA_TRY(cudaMemcpyAsync(hd.data(), e->dst, sizeof(int32_t) * p, cudaMemcpyDeviceToHost, st));
  MDE_CUDA_TRY(cudaMemcpyAsync(hw.data(), e->par0, sizeof(float) * p, cudaMemcpyDeviceToHost, st));
  MDE_CUDA_TRY(cudaStreamSynchronize(st));
  EllHost h;
  int rc = ell_build_host(n, p, m, hs.data(), hd.data(), hw.data(), fn->push_pull, rb, kNumSMs, h);
  if (rc) return rc;
  const void* k = eselect_kernel(e->fn, m);
  if (!k) return MDE_E_UNSUPPORTED;
  if ((rc = econfigure_kernel(k))) return rc;
  const int nbkt = (int)h.bkt_tile.size();
#define UP(dst, vec, T)                                                                              \
  do {                                                                                               \
    cudaError_t _e = cudaMalloc(&(dst), sizeof(T) * (vec).size());                                   \
    if (_e == cudaSuccess) _e = cudaMemcpyAsync((dst), (vec).data(), sizeof(T) * (vec).size(), cudaMemcpyHostToDevice, st); \
    if (_e != cudaSuccess) { ell_free(e); return (int)_e; }                                          \
  } while (0)
  UP(e->ell_rec, h.rec, unsigned char);
  UP(e->ell_off, h.rec_off, uint32_t);
  UP(e->ell_bkt_tile, h.bkt_tile, int32_t);
  UP(e->ell_bkt_wt0, h.bkt_wt0, int32_t);
  std::vector<int32_t> desc((size_t)4 * h.ncta);
  for (int c = 0; c < h.ncta; ++c) {
    desc[4 * c] = h.cta_wt0[c]; desc[4 * c + 1] = h.cta_wt0[c + 1]; desc[4 * c + 2] = h.cta_bkt0[c];
    desc[4 * c + 3] = h.bkt_tile[(size_t)h.cta_bkt0[c]];
  }
  UP(e->ell_cta_desc, desc, int32_t);
#undef UP
  cudaError_t se = cudaStreamSynchronize(st);  // the host vectors die with this frame
  if (se != cudaSuccess) { ell_free(e); return (int)se; }
  e->kind = 3; e->m_hint = m; e->rb = rb; e->ell_nrec = h.nrec; e->ell_ncta = h.ncta; e->nbkt = nbkt;
  e->nbytes += (int64_t)h.rec.size() + 4 * (h.nrec + 1) + 4ll * (2 * nbkt + 2 * h.ncta + 2);
  return 0;
}

int ell_launch(const mde_edges* e, const float* X, int m, float* grad, int* nblocks_out, const int* flag,
               cudaStream_t st) {
  if (e->kind != 3 || m != e->m_hint || !grad) return MDE_E_UNSUPPORTED;
  const size_t smem = ell_smem_bytes(e->rb, m);
  EllArgs a;
  a.rec = e->ell_rec; a.rec_off = e->ell_off; a.bkt_tile = e->ell_bkt_tile; a.bkt_wt0 = e->ell_bkt_wt0;
  a.cta_desc = reinterpret_cast<const int4*>(e->ell_cta_desc); a.X = X; a.grad = grad;
  a.loss_partials = e->loss_partials; a.flag = flag; a.fn = e->fn; a.inv_p = 1.0f / (float)e->p_total;
  a.n = e->n; a.rb = e->rb;
  a.x_vec_ok = ((reinterpret_cast<uintptr_t>(X) & 15u) == 0) ? 1 : 0;
  const void* k = eselect_kernel(e->fn, m);
  if (!k) return MDE_E_UNSUPPORTED;
  int rc = econfigure_kernel(k);
  if (rc) return rc;
  void* args[] = {(void*)&a};
  MDE_CUDA_TRY(cudaLaunchKernel(k, dim3(e->ell_ncta), dim3(kEllThreads), args, smem, st));
  MDE_LAUNCH_CHECK();
  if (nblocks_out) *nblocks_out = e->ell_ncta;
  return 0;
}

}  // namespace mde

// ------------------------------------------------------------------------------------------
// host-only export for the CPU tests: build the ELL records from host arrays, no device involved
// ------------------------------------------------------------------------------------------
extern "C" {

int mde_ell_host_layout(int64_t n_items, int64_t p, int embedding_dim, const int32_t* src, const int32_t* dst,
                        const float* par0, int push_pull, int tile_rows_log2, int max_cta, mde_ell_host_t* out) {
  if (!src || !dst || !par0 || !out) return MDE_E_INVALID;
  EllHost h;
  const int rb = tile_rows_log2 > 0 ? tile_rows_log2 : ell_default_rb(embedding_dim);
  const int rc = ell_build_host(n_items, p, embedding_dim, src, dst, par0, push_pull, rb, max_cta > 0 ? max_cta : kNumSMs, h);
  if (rc) return rc;
  memset(out, 0, sizeof(*out));
  auto dup = [](const void* p_, size_t bytes) -> void* {
    void* q = malloc(bytes ? bytes : 1);
    if (q && bytes) memcpy(q, p_, bytes);
    return q;
  };
  out->rec_bytes = (int64_t)h.rec.size();
  out->nrec = h.nrec; out->nbkt = (int32_t)h.bkt_tile.size(); out->ncta = h.ncta; out->tile_rows_log2 = h.rb;
  out->nslots = h.nslots; out->nentries = h.nentries; out->npadded = h.npadded;
  out->rec = (unsigned char*)dup(h.rec.data(), h.rec.size());
  out->rec_off = (uint32_t*)dup(h.rec_off.data(), 4 * h.rec_off.size());
  out->bkt_tile = (int32_t*)dup(h.bkt_tile.data(), 4 * h.bkt_tile.size());
  out->bkt_wt0 = (int32_t*)dup(h.bkt_wt0.data(), 4 * h.bkt_wt0.size());
  out->cta_wt0 = (int32_t*)dup(h.cta_wt0.data(), 4 * h.cta_wt0.size());
  out->cta_bkt0 = (int32_t*)dup(h.cta_bkt0.data(), 4 * h.cta_bkt0.size());
  if (!out->rec || !out->rec_off || !out->bkt_tile || !out->bkt_wt0 || !out->cta_wt0 || !out->cta_bkt0) {
    mde_ell_host_free(out);
    return MDE_E_ALLOC;
  }
  return 0;
}

void mde_ell_host_free(mde_ell_host_t* h) {
  if (!h) return;
  free(h->rec); free(h->rec_off); free(h->bkt_tile); free(h->bkt_wt0); free(h->cta_wt0); free(h->cta_bkt0);
  memset(h, 0, sizeof(*h));
}

}  // extern "C"

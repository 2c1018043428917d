// tcgen05 attention forward for L <= 128, head_dim 64 (modeling_vlbert.py:63-103 + transpose_for_scores
// modeling_bert.py:299-303): the Blackwell-native replacement of the mma.sync kernel in attention.cu.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0   TMA producer: Q, K, V tiles of one (batch, head) straight out of the fused QKV projection
//            [B, L, 3H] through ONE 3-D tensor map (box 64 x 128 x 1, 128-byte swizzle); rows >= L are zero-filled
//            by the TMA bounds check, so ragged L needs no padding pass.  3-stage ring (48 KB per stage).
//   warp 1   MMA issuer: S = Q K^T (tcgen05.mma 128 x N16 x 16, 4 steps over d) and O = P V (128 x 64 x 16,
//            N16 / 16 steps over the keys), accumulators in TMEM (512 columns: two groups x {S 128, O 64}).
//   warp 2   TMEM allocator.
//   warps 4-7, 8-11  two softmax + epilogue warpgroups, each owning every other item (its own S / O columns in TMEM and
//            its own P tile), thread = one query row = one TMEM lane.  The score row is fetched by back-to-back
//            tcgen05.ld and one wait, stays in registers from the row maximum to exp2 / row sum / dropout, and goes
//            as bf16 into the K-major swizzled A-operand tile of the second MMA.  The epilogue reads O, applies
//            1 / sum (and the dropout rescale), and stores the head-merged context [B, L, H] by a 3-D TMA store
//            (staged in the item's dead Q tile) whose bounds check clips the rows >= L.
// The first version had ONE softmax warpgroup and a two-pass row: correct, but the serial chain per item (8 TMEM
// round trips, the PV latency, two barriers) made it latency bound at ~2.6 us per head -- slower than the mma.sync
// kernel (profiles/r2_attention_tc_v4_microbench.txt).  Two groups overlap each other's latencies; S(n + 2) is issued as soon as
// group n & 1 has consumed S(n), so tensor work and TMA latency hide behind the softmax warps.
// Scores never touch HBM; lse [B, nh, L] is saved for the backward kernel; the dropout mask is the stateless hash of
// common.cuh with the SAME element indexing as attention.cu, so its backward regenerates the mask bit for bit.
#include <cudaTypedefs.h>

#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace mvptr {
namespace attn_tc {

constexpr int D = 64;
constexpr int kRows = 128;           // MMA M: query rows (padded)
constexpr int kTileBytes = kRows * 128;  // one [128][64] bf16 tile, 128-byte rows
constexpr int kThreads = 384;
constexpr float kLog2e = 1.4426950408889634f;

struct Params {
  const float* maskadd;  // [B, L] additive mask (1 - mask) * -10000
  float* lse;            // [B, nh, L] or null
  int B, L, nh, H;
  int n16;               // keys padded to a multiple of 16 (MMA N of S, K of PV)
  int items;             // B * nh
  float scale_log2;      // 1/sqrt(64) * log2(e)
  uint32_t keep_thr;
  float inv_keep;
  uint32_t seed;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, SWIZZLE_128B (same bit layout as gemm_tcgen05.cu)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int m, int n, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

constexpr int kGroups = 2;  // softmax warpgroups, each owns every other item

// Forward geometry by tile rows R (= TMA box rows >= L): smaller tiles buy pipeline stages, and the stage count is
// what bounds this kernel -- a stage is held from its TMA issue until O = P V retires (~4-5 us), so bytes in flight,
// not instruction issue, set the rate (measured: 3 stages of 48 KB gave 2.1-2.8 us per item whatever L was).
template <int R>
struct FwdCfg {
  static constexpr int kTile = R * 128;                  // one [R][64] bf16 tile
  static constexpr int kStage = 3 * kTile;               // Q | K | V
  static constexpr int kStages = R == 64 ? 5 : R == 96 ? 3 : 2;
  static constexpr int kPBlocks = R > 64 ? 2 : 1;        // 64-key blocks of the P tile
  static constexpr int kPBytes = kPBlocks * kTileBytes;  // P keeps 128-row blocks: the second MMA reads M = 128 rows
  static constexpr int kSlab = 32 * 128;                 // one warp's 32 context rows: a TMA-store box
  static constexpr int kSmem = kStages * kStage + kGroups * kPBytes + 8 * kSlab + 8 * kRows * 4 /* per-warp mask rows */ + 256 + 1024;
};

// One softmax warp's work on one item.  NCH = 32-column chunks of the score row (n16 <= 32 NCH): the whole row
// is fetched from TMEM by NCH back-to-back tcgen05.ld and ONE wait, and lives in registers from the row maximum to
// the packed bf16 probabilities -- a single pass, no second TMEM read.
template <int NCH>
__device__ __forceinline__ void softmax_row(const Params& p, uint32_t tmem_row, const float* mk, int row, uint32_t rbase,
                                            uint32_t dseed, bool drop, uint8_t* sPg, float& mx_out, float& sum_out) {
  uint32_t r[NCH][32];
#pragma unroll
  for (int c = 0; c < NCH; ++c) tc_ld32(tmem_row + c * 32, r[c]);
  tc_wait_ld();
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      // columns >= L: mask -inf (zero-filled K rows give finite scores; columns >= n16 hold stale TMEM bits, which
      // the select keeps out of the arithmetic)
      const float m = mk[c * 32 + j];
      const float v = m == -INFINITY ? -INFINITY : fmaf(__uint_as_float(r[c][j]), p.scale_log2, m);
      r[c][j] = __float_as_uint(v);
      mx = fmaxf(mx, v);
    }
  float sum = 0.f;
  uint8_t* prow = sPg + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    float e[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      e[j] = ex2_approx(__uint_as_float(r[c][j]) - mx);
      sum += e[j];
    }
    if (drop) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        bool k0, k1;
        dropout_pair(dseed, rbase + c * 32 + j, p.keep_thr, k0, k1);
        e[j] = k0 ? e[j] : 0.f;
        e[j + 1] = k1 ? e[j + 1] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int chunk = c * 4 + u;  // 16-byte unit = 8 keys
      if (chunk * 8 < p.n16) {
        uint4 v;
        v.x = pack2(e[u * 8 + 0], e[u * 8 + 1]);
        v.y = pack2(e[u * 8 + 2], e[u * 8 + 3]);
        v.z = pack2(e[u * 8 + 4], e[u * 8 + 5]);
        v.w = pack2(e[u * 8 + 6], e[u * 8 + 7]);
        *reinterpret_cast<uint4*>(prow + (chunk >> 3) * kTileBytes + (((chunk & 7) ^ (row & 7)) << 4)) = v;
      }
    }
  }
  mx_out = mx;
  sum_out = sum;
}

template <int R>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, const Params p) {
  using C = FwdCfg<R>;
  constexpr int kStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sP = smem + kStages * C::kStage;                                     // [kGroups][kPBlocks][16 KB]
  uint8_t* sSlab = sP + kGroups * C::kPBytes;                                   // [8 warps][32 rows][128 B] store slabs
  float* sMask = reinterpret_cast<float*>(sSlab + 8 * C::kSlab);                // [8 warps][128], log2 domain
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMask + 8 * kRows);
  const uint32_t full_bar = smem_u32(bars);                      // [kStages] TMA -> MMA
  const uint32_t empty_bar = smem_u32(bars + kStages);           // [kStages] O = P V retired -> TMA
  const uint32_t sfull_bar = smem_u32(bars + 2 * kStages);       // [kGroups] S in TMEM -> softmax group
  const uint32_t pfull_bar = smem_u32(bars + 2 * kStages + 2);   // [kGroups] P in smem (and S consumed) -> MMA
  const uint32_t ofull_bar = smem_u32(bars + 2 * kStages + 4);   // [kGroups] O in TMEM -> softmax group
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKV) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(full_bar + 8 * i, 1);
      mbar_init(empty_bar + 8 * i, 1);
    }
    for (int g = 0; g < kGroups; ++g) {
      mbar_init(sfull_bar + 8 * g, 1);
      mbar_init(pfull_bar + 8 * g, 4);
      mbar_init(ofull_bar + 8 * g, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;  // group g: S at columns [256 g, 256 g + 128), O at [256 g + 128, 256 g + 192)
  const int H = p.H;
  const int n_local = p.items > (int)blockIdx.x ? (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int n = 0; n < n_local; ++n) {
        const int it = blockIdx.x + n * gridDim.x;
        const int b = it / p.nh, h = it - b * p.nh;
        mbar_wait(empty_bar + 8 * stage, phase ^ 1);
        const uint32_t sq = smem_u32(smem + stage * C::kStage);
        const uint32_t fb = full_bar + 8 * stage;
        mbar_expect_tx(fb, C::kStage);
        tma_load_3d(sq, &tmQKV, fb, h * D, 0, b);
        tma_load_3d(sq + C::kTile, &tmQKV, fb, H + h * D, 0, b);
        tma_load_3d(sq + 2 * C::kTile, &tmQKV, fb, 2 * H + h * D, 0, b);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      // S[128, n16] = Q . K^T (both K-major; Q rows >= R read the K / V tiles behind it: garbage rows nobody stores)
      const uint32_t idesc_s = make_idesc(kRows, p.n16, false);
      const uint32_t idesc_o = make_idesc(kRows, D, true);       // O[128, 64] = P . V, V is MN-major ([key][d])
      const int ksteps = p.n16 >> 4;
      auto try_wait = [](uint32_t bar, uint32_t parity) -> bool {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        return ok != 0;
      };
      // Two independent event streams feed this thread -- "Q, K of item s have landed" (-> issue S(s)) and "group
      // n & 1 has written P(n)" (-> issue O(n) = P V) -- and neither may wait behind the other: it polls both.
      // S(s) may overwrite its group's score columns once P(s - 2) has been taken, i.e. s < next_pv + kGroups.
      int next_s = 0, next_pv = 0;
      while (next_pv < n_local) {
        if (next_s < n_local && next_s < next_pv + kGroups) {
          const int st = next_s % kStages, g = next_s & 1;
          if (try_wait(full_bar + 8 * st, (uint32_t)(next_s / kStages) & 1u)) {
            tc_fence_after();
            const uint32_t sq = smem_u32(smem + st * C::kStage), sk = sq + C::kTile;
#pragma unroll
            for (int k = 0; k < D / 16; ++k)
              tc_mma(tmem_base + g * 256, make_desc(sq + k * 32, 16, 1024), make_desc(sk + k * 32, 16, 1024), idesc_s,
                     k > 0 ? 1u : 0u);
            tc_commit(sfull_bar + 8 * g);
            ++next_s;
          }
        }
        if (next_pv < next_s) {
          const int st = next_pv % kStages, g = next_pv & 1;
          if (try_wait(pfull_bar + 8 * g, (uint32_t)(next_pv >> 1) & 1u)) {  // P in smem, S read out of TMEM
            tc_fence_after();
            const uint32_t sv = smem_u32(smem + st * C::kStage) + 2 * C::kTile;
            const uint32_t sp = smem_u32(sP + g * C::kPBytes);
            for (int k = 0; k < ksteps; ++k)
              tc_mma(tmem_base + g * 256 + 128, make_desc(sp + (k >> 2) * kTileBytes + (k & 3) * 32, 16, 1024),
                     make_desc(sv + k * 2048, 64 * 128, 1024), idesc_o, k > 0 ? 1u : 0u);
            tc_commit(ofull_bar + 8 * g);
            tc_commit(empty_bar + 8 * st);  // Q, K, V of this stage are dead once O = P V retires
            ++next_pv;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ======================= softmax + epilogue: thread = query row = TMEM lane; warps are independent ==========
    const int g = (warp - 4) >> 2;                  // warpgroup: items n with n & 1 == g
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;                 // query index inside the head
    const int L = p.L;
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    const uint32_t tmem_S = tmem_base + g * 256 + lane_off, tmem_O = tmem_S + 128;
    const uint32_t dseed = site_seed(p.seed);
    const bool drop = p.keep_thr != 0xffffffffu;
    const uint32_t Lp = (uint32_t)(L + 1) & ~1u;    // even row pitch of the dropout index (as attention.cu)
    const int nch = (p.n16 + 31) >> 5;
    uint8_t* sPg = sP + g * C::kPBytes;
    float* mk = sMask + (warp - 4) * kRows;         // this warp's copy of the mask row
    // the mask row of the next item is fetched one item ahead (all heads of a batch element share it)
    float next_mask[4];
    auto fetch_mask = [&](int it) {
      const float* src = p.maskadd + (size_t)(it / p.nh) * L;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = lane + 32 * i;
        next_mask[i] = k < L ? __ldg(src + k) * kLog2e : -INFINITY;
      }
    };
    if (g < n_local) fetch_mask(blockIdx.x + g * gridDim.x);
    int m = 0;  // items this group has processed
    for (int n = g; n < n_local; n += kGroups, ++m) {
      const int it = blockIdx.x + n * gridDim.x;
      const int b = it / p.nh, h = it - b * p.nh;
#pragma unroll
      for (int i = 0; i < 4; ++i) mk[lane + 32 * i] = next_mask[i];
      if (n + kGroups < n_local) fetch_mask(it + kGroups * gridDim.x);
      __syncwarp();
      mbar_wait(sfull_bar + 8 * g, (uint32_t)m & 1u);
      tc_fence_after();
      const uint32_t rbase = (((uint32_t)b * p.nh + h) * L + row) * Lp;
      float mx, sum;
      if constexpr (R == 64) {  // n16 <= 64
        if (nch == 1) softmax_row<1>(p, tmem_S, mk, row, rbase, dseed, drop, sPg, mx, sum);
        else softmax_row<2>(p, tmem_S, mk, row, rbase, dseed, drop, sPg, mx, sum);
      } else if constexpr (R == 96) {  // n16 = 80 or 96
        softmax_row<3>(p, tmem_S, mk, row, rbase, dseed, drop, sPg, mx, sum);
      } else {  // n16 = 112 or 128
        softmax_row<4>(p, tmem_S, mk, row, rbase, dseed, drop, sPg, mx, sum);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // P visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pfull_bar + 8 * g);
      if (p.lse && row < L) p.lse[((size_t)b * p.nh + h) * L + row] = (mx + log2f(sum)) * (1.0f / kLog2e);
      const float inv = p.inv_keep / sum;
      // ---- epilogue: O row * inv -> bf16 -> this warp's swizzled 32-row slab -> its own 3-D TMA store (rows >= L
      //      clipped by the tensor map).  Thread-per-row st.global of the 128-byte rows was tried and is 20-65 %
      //      slower (16-byte pieces at a 1536-byte stride; profiles/r2_attention_tc_v3_direct_stores_microbench.txt)
      mbar_wait(ofull_bar + 8 * g, (uint32_t)m & 1u);
      tc_fence_after();
      uint32_t r[2][32];
      tc_ld32(tmem_O, r[0]);
      tc_ld32(tmem_O + 32, r[1]);
      tc_wait_ld();
      tc_fence_before();
      if (m > 0) {  // this warp's previous store has been read out of the slab
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
      }
      uint8_t* slab = sSlab + (warp - 4) * C::kSlab;
      uint8_t* orow = slab + lane * 128;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 v;
          v.x = pack2(__uint_as_float(r[c][u * 8 + 0]) * inv, __uint_as_float(r[c][u * 8 + 1]) * inv);
          v.y = pack2(__uint_as_float(r[c][u * 8 + 2]) * inv, __uint_as_float(r[c][u * 8 + 3]) * inv);
          v.z = pack2(__uint_as_float(r[c][u * 8 + 4]) * inv, __uint_as_float(r[c][u * 8 + 5]) * inv);
          v.w = pack2(__uint_as_float(r[c][u * 8 + 6]) * inv, __uint_as_float(r[c][u * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + (((c * 4 + u) ^ (lane & 7)) << 4)) = v;
        }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0 && q4 * 32 < L) {
        tma_store_3d(&tmO, smem_u32(slab), h * D, q4 * 32, b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// =====================================================================================================
// Backward, L <= 128.  Per (batch, head) item, all five contractions on the tensor core, accumulators in TMEM:
//   S = Q K^T, dP = dO V^T                      (K-major operands)                      -> columns [0,128), [128,256)
//   softmax warps: p = exp2(s * scale + mask - lse), Pd = dropout(p), dS = p * (dropout(dP) - D),  D = rowsum(Pd * dP)
//                  (D from the scores themselves: O = Pd V  =>  dO . O = sum_k Pd dP -- the forward output is not read)
//   dV = Pd^T dO, dK = dS^T Q * scale           (A MN-major: the SAME swizzled [q][k] tiles read transposed)
//   dQ = dS K * scale                           (A K-major, B = K tile MN-major)        -> columns [256,448)
// Warp roles as in the forward; ONE softmax group of 8 warps, two threads per query row (16-column sub-chunks of the
// row alternate between them; the row sum D is exchanged through shared memory).  Q, K, V, dO arrive by 3-D TMA (rows
// >= L zero-filled), dQ | dK | dV leave by 3-D TMA stores staged in the item's dead Q | K | V tiles, and the QKV-bias
// gradient (column sums of dQ | dK | dV) is accumulated from the staged tiles.  Same dropout hash / indexing as the
// forward.  Replaces attn_bwd_smem_kernel (mma.sync) of attention.cu for L <= 128.
// =====================================================================================================
struct BwdParams {
  const float* maskadd;  // [B, L]
  const float* lse;      // [B, nh, L]
  float* dbias;          // [3H] fp32 (+=) or null
  int B, L, nh, H;
  int n16, items;
  float scale, scale_log2;
  uint32_t keep_thr;
  float inv_keep;
  uint32_t seed;
};
constexpr int kSmemP = 2 * kTileBytes;  // a [128 q][128 k] bf16 tile: two 64-key blocks
constexpr int kBwdStages = 2;
constexpr int kBwdStageBytes = 4 * kTileBytes;  // Q | K | V | dO
constexpr int kBwdMaxH = 1024;                  // bias-gradient accumulators [3][H] live in shared memory
constexpr int kBwdSlab = 32 * 128;              // 32 rows x 64 columns of one gradient tensor: a TMA-store box
constexpr int kBwdSmem = kBwdStages * kBwdStageBytes + 2 * kSmemP /* Pd, dS */ + 4 * kRows * 4 /* mask, lse, D halves */ +
                         3 * kBwdMaxH * 4 + 256 + 1024;

template <int NS>  // 16-column sub-chunks of a score row per thread: ceil(n16 / 32)
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmDQKV, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sPd = smem + kBwdStages * kBwdStageBytes;  // [2 blocks][16 KB]: [q][k] tile, 64 keys per block
  uint8_t* sDS = sPd + kSmemP;
  float* sMask = reinterpret_cast<float*>(sDS + kSmemP);  // [128] log2 domain
  float* sLse = sMask + kRows;                            // [128] log2 domain (+inf for rows >= L)
  float* sD = sLse + kRows;                               // [2][128] partial row sums of the two column halves
  float* sBias = sD + 2 * kRows;                          // [3][H] column sums of dQ | dK | dV over this CTA's items
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 3 * kBwdMaxH);
  const uint32_t full_bar = smem_u32(bars);                       // [2] TMA -> MMA
  const uint32_t empty_bar = smem_u32(bars + kBwdStages);         // [2] gradients stored -> TMA
  const uint32_t sfull_bar = smem_u32(bars + 2 * kBwdStages);     // S, dP in TMEM -> softmax warps
  const uint32_t pfull_bar = smem_u32(bars + 2 * kBwdStages + 1); // Pd, dS in smem (S, dP consumed) -> MMA
  const uint32_t gfull_bar = smem_u32(bars + 2 * kBwdStages + 2); // dQ, dK, dV in TMEM -> softmax warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kBwdStages + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKV) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDO) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDQKV) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kBwdStages; ++i) {
      mbar_init(full_bar + 8 * i, 1);
      mbar_init(empty_bar + 8 * i, 4);  // the four warp pairs, each once its gradient stores have read the stage
    }
    mbar_init(sfull_bar, 1);
    mbar_init(pfull_bar, 8);
    mbar_init(gfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (p.dbias != nullptr)
    for (int i = threadIdx.x; i < 3 * p.H; i += blockDim.x) sBias[(i / p.H) * kBwdMaxH + (i % p.H)] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDQ = tmem_base + 256, tDK = tmem_base + 320, tDV = tmem_base + 384;
  const int H = p.H;
  const int n_local = p.items > (int)blockIdx.x ? (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      for (int n = 0; n < n_local; ++n) {
        const int it = blockIdx.x + n * gridDim.x;
        const int b = it / p.nh, h = it - b * p.nh;
        const int st = n & 1;
        mbar_wait(empty_bar + 8 * st, ((uint32_t)(n >> 1) & 1u) ^ 1u);
        const uint32_t sq = smem_u32(smem + st * kBwdStageBytes);
        const uint32_t fb = full_bar + 8 * st;
        mbar_expect_tx(fb, 4 * kTileBytes);
        tma_load_3d(sq, &tmQKV, fb, h * D, 0, b);
        tma_load_3d(sq + kTileBytes, &tmQKV, fb, H + h * D, 0, b);
        tma_load_3d(sq + 2 * kTileBytes, &tmQKV, fb, 2 * H + h * D, 0, b);
        tma_load_3d(sq + 3 * kTileBytes, &tmDO, fb, h * D, 0, b);
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(kRows, p.n16, false);                      // [q, k] = A[q, d] B[k, d]^T, K-major both
      const uint32_t idesc_dq = make_idesc(kRows, D, true);                          // dQ[q, d] = dS[q, k] K[k, d]
      const uint32_t idesc_t = make_idesc(kRows, D, true) | (1u << 15);              // dK / dV: A MN-major, B MN-major
      const int ksteps_k = p.n16 >> 4;                 // contraction over keys (dQ)
      const int ksteps_q = (p.L + 15) >> 4;            // contraction over queries (dK, dV); rows >= L are zero
      auto issue_s = [&](int n) {
        const int st = n & 1;
        mbar_wait(full_bar + 8 * st, (uint32_t)(n >> 1) & 1u);
        tc_fence_after();
        const uint32_t sq = smem_u32(smem + st * kBwdStageBytes), sk = sq + kTileBytes, sv = sq + 2 * kTileBytes,
                       sdo = sq + 3 * kTileBytes;
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          tc_mma(tS, make_desc(sq + k * 32, 16, 1024), make_desc(sk + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          tc_mma(tDP, make_desc(sdo + k * 32, 16, 1024), make_desc(sv + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        tc_commit(sfull_bar);
      };
      if (n_local > 0) issue_s(0);
      for (int n = 0; n < n_local; ++n) {
        const int st = n & 1;
        mbar_wait(pfull_bar, (uint32_t)n & 1u);  // Pd, dS in shared memory; S, dP read out of TMEM
        tc_fence_after();
        const uint32_t sq = smem_u32(smem + st * kBwdStageBytes), sk = sq + kTileBytes, sdo = sq + 3 * kTileBytes;
        const uint32_t sp = smem_u32(sPd), sds = smem_u32(sDS);
        // A MN-major from the [q][k] tiles: 64-key chunks kTileBytes apart (LBO), 8-query groups 1024 B apart (SBO),
        // 16 queries (one K step) = 2048 B
        for (int k = 0; k < ksteps_q; ++k)
          tc_mma(tDV, make_desc(sp + k * 2048, kTileBytes, 1024), make_desc(sdo + k * 2048, 64 * 128, 1024), idesc_t,
                 k > 0 ? 1u : 0u);
        for (int k = 0; k < ksteps_q; ++k)
          tc_mma(tDK, make_desc(sds + k * 2048, kTileBytes, 1024), make_desc(sq + k * 2048, 64 * 128, 1024), idesc_t,
                 k > 0 ? 1u : 0u);
        for (int k = 0; k < ksteps_k; ++k)
          tc_mma(tDQ, make_desc(sds + (k >> 2) * kTileBytes + (k & 3) * 32, 16, 1024),
                 make_desc(sk + k * 2048, 64 * 128, 1024), idesc_dq, k > 0 ? 1u : 0u);
        tc_commit(gfull_bar);
        if (n + 1 < n_local) issue_s(n + 1);  // overlaps the epilogue of item n
      }
    }
  } else if (warp >= 4) {
    // ======================= softmax / gradient warps: two threads per query row =======================
    const int q4 = warp & 3;
    const int hf = (warp - 4) >> 2;                 // which 16-column sub-chunks (and which 32-column halves of the outputs)
    const int row = q4 * 32 + lane;
    const int tid = threadIdx.x - 128;              // 0..255
    const int L = p.L;
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    const uint32_t dseed = site_seed(p.seed);
    const bool drop = p.keep_thr != 0xffffffffu;
    const uint32_t Lp = (uint32_t)(L + 1) & ~1u;
    const int nsub = p.n16 >> 4;                    // 16-column sub-chunks of the row
    auto group_sync = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    float next_mask = 0.f, next_lse = 0.f;
    auto prefetch = [&](int it) {
      if (tid < kRows) next_mask = tid < L ? p.maskadd[(size_t)(it / p.nh) * L + tid] * kLog2e : -INFINITY;
      else next_lse = (tid - kRows) < L ? p.lse[(size_t)it * L + (tid - kRows)] * kLog2e : INFINITY;
    };
    if (n_local > 0) prefetch(blockIdx.x);
    for (int n = 0; n < n_local; ++n) {
      const int it = blockIdx.x + n * gridDim.x;
      const int b = it / p.nh, h = it - b * p.nh;
      const int st = n & 1;
      if (tid < kRows) sMask[tid] = next_mask;
      else sLse[tid - kRows] = next_lse;
      if (n + 1 < n_local) prefetch(it + gridDim.x);
      group_sync();
      mbar_wait(sfull_bar, (uint32_t)n & 1u);
      tc_fence_after();
      const float lse2 = sLse[row];
      const uint32_t rbase = (((uint32_t)b * p.nh + h) * L + row) * Lp;
      const bool live = row < L;
      // ---- this thread's sub-chunks of the S and dP rows: all TMEM loads back to back, ONE wait; p and the dropped
      //      dP stay in registers from the row sum D to dS (single pass: one exp, one dropout hash per score)
      uint32_t rs[NS][16], rd[NS][16];
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const int sc = 2 * i + hf;
        if (sc < nsub) {
          tc_ld16(tS + lane_off + sc * 16, rs[i]);
          tc_ld16(tDP + lane_off + sc * 16, rd[i]);
        }
      }
      tc_wait_ld();
      float dpart = 0.f;
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const int sc = 2 * i + hf;
        if (sc < nsub) {
          uint32_t bits = 0xffffu;
          if (drop) {
            bits = 0;
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              bool k0, k1;
              dropout_pair(dseed, rbase + sc * 16 + j, p.keep_thr, k0, k1);
              bits |= (k0 ? 1u : 0u) << j;
              bits |= (k1 ? 1u : 0u) << (j + 1);
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float m = sMask[sc * 16 + j];
            const float pr = (!live || m == -INFINITY) ? 0.f : ex2_approx(fmaf(__uint_as_float(rs[i][j]), p.scale_log2, m - lse2));
            const bool keep = (bits >> j) & 1u;
            const float dpd = keep ? __uint_as_float(rd[i][j]) * p.inv_keep : 0.f;  // dropout applied to dP
            dpart = fmaf(keep ? pr * p.inv_keep : 0.f, __uint_as_float(rd[i][j]), dpart);
            rs[i][j] = __float_as_uint(pr);
            rd[i][j] = __float_as_uint(dpd);
          }
          // Pd = dropout(p) as bf16 into the swizzled [q][k] tile (dpd == 0 exactly where the element was dropped)
          uint8_t* prow = sPd + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int chunk = sc * 2 + u;
            float v8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              v8[j] = ((bits >> (u * 8 + j)) & 1u) ? __uint_as_float(rs[i][u * 8 + j]) * p.inv_keep : 0.f;
            uint4 v;
            v.x = pack2(v8[0], v8[1]); v.y = pack2(v8[2], v8[3]); v.z = pack2(v8[4], v8[5]); v.w = pack2(v8[6], v8[7]);
            *reinterpret_cast<uint4*>(prow + (chunk >> 3) * kTileBytes + (((chunk & 7) ^ (row & 7)) << 4)) = v;
          }
        }
      }
      sD[hf * kRows + row] = dpart;
      group_sync();
      const float Drow = sD[row] + sD[kRows + row];
      uint8_t* drow = sDS + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const int sc = 2 * i + hf;
        if (sc < nsub) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int chunk = sc * 2 + u;
            float v8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)  // dS = p * (dropout(dP) - D)
              v8[j] = __uint_as_float(rs[i][u * 8 + j]) * (__uint_as_float(rd[i][u * 8 + j]) - Drow);
            uint4 v;
            v.x = pack2(v8[0], v8[1]); v.y = pack2(v8[2], v8[3]); v.z = pack2(v8[4], v8[5]); v.w = pack2(v8[6], v8[7]);
            *reinterpret_cast<uint4*>(drow + (chunk >> 3) * kTileBytes + (((chunk & 7) ^ (row & 7)) << 4)) = v;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pfull_bar);
      // ---- epilogue: dQ | dK | dV rows (this thread: 32 of the 64 columns) -> bf16 -> the 32-row slab this warp
      //      shares with its partner (the other column half) -> one 3-D TMA store per tensor and row quarter (rows >= L
      //      clipped); the QKV-bias gradient = column sums of the rounded values, reduced across the warp's 32 rows by
      //      the halving butterfly (31 shuffles per 32 columns) into the CTA's shared accumulators
      mbar_wait(gfull_bar, (uint32_t)n & 1u);
      tc_fence_after();
      // store slabs = this warp pair's 32 rows of the item's dead Q | K | V tiles (all MMAs that read them have retired)
      uint8_t* stg = smem + st * kBwdStageBytes;
      {
        uint32_t r[2][32];  // double buffered: tensor w + 1 is in flight while w is packed, stored and column-summed
        tc_ld32(tDQ + lane_off + hf * 32, r[0]);
#pragma unroll
        for (int w = 0; w < 3; ++w) {
          tc_wait_ld();
          if (w == 0) tc_ld32(tDK + lane_off + hf * 32, r[1]);
          if (w == 1) tc_ld32(tDV + lane_off + hf * 32, r[0]);
          const float f = w == 2 ? 1.f : p.scale;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            pk[j] = pack2(__uint_as_float(r[w & 1][2 * j]) * f, __uint_as_float(r[w & 1][2 * j + 1]) * f);
          {
            uint8_t* orow = stg + w * kTileBytes + q4 * kBwdSlab + lane * 128;
#pragma unroll
            for (int u = 0; u < 4; ++u)
              *reinterpret_cast<uint4*>(orow + (((hf * 4 + u) ^ (lane & 7)) << 4)) =
                  make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
          }
          if (p.dbias != nullptr) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              v[2 * j] = live ? __uint_as_float(pk[j] << 16) : 0.f;
              v[2 * j + 1] = live ? __uint_as_float(pk[j] & 0xffff0000u) : 0.f;
            }
            // after step s (offset 16 >> s) lane l holds the partial sums of 16 >> s columns; at the end: column (l)
#pragma unroll
            for (int off = 16, cnt = 16; off >= 1; off >>= 1, cnt >>= 1) {
              const bool upper = (lane & off) != 0;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (j < cnt) {
                  const float mine = upper ? v[j + cnt] : v[j];
                  const float send = upper ? v[j] : v[j + cnt];
                  v[j] = mine + __shfl_xor_sync(0xffffffffu, send, off);
                }
              }
            }
            // lane l now holds the sum over the warp's 32 rows of column bitrev-free index: columns were halved MSB
            // first, so lane l owns column l
            atomicAdd(sBias + w * kBwdMaxH + h * D + hf * 32 + lane, v[0]);
          }
        }
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q4) : "memory");  // both column halves of the slabs are written
        if (hf == 0 && lane == 0) {
          if (q4 * 32 < L) {
#pragma unroll
            for (int w = 0; w < 3; ++w)
              tma_store_3d(&tmDQKV, smem_u32(stg + w * kTileBytes + q4 * kBwdSlab), w * H + h * D, q4 * 32, b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          mbar_arrive(empty_bar + 8 * st);  // this pair is done with the stage
        }
      }
    }
    if (hf == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    group_sync();  // all shared-memory bias accumulations of this CTA are done
    if (p.dbias != nullptr)
      for (int i = tid; i < 3 * p.H; i += 256) {
        const float v = sBias[(i / p.H) * kBwdMaxH + (i % p.H)];
        if (v != 0.f) atomicAdd(p.dbias + i, v);
      }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

// [B, L, cols] bf16 view (row pitch ld elements) with box 64 x 128 x 1: rows >= L are out of bounds -> zero-filled
// on load, clipped on store
static int make_map3(CUtensorMap* map, const void* base, int cols, int L, int B, int ld, int box_rows = kRows) {
  auto enc = get_encode();
  if (!enc) MVPTR_FAIL(MVPTR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)L * ld * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) MVPTR_FAIL(MVPTR_ERR_CUDA, "attention tensor map (cols %d, L %d, B %d, ld %d): error %d", cols, L, B, ld, (int)r);
  return 0;
}

}  // namespace attn_tc
}  // namespace mvptr

// Returns 1 when this path does not apply (the caller falls back to the mma.sync kernel), 0 on success, < 0 on error.
// kernel selection: -1 = follow the environment (MVPTR_ATTN_FWD_TC / MVPTR_ATTN_BWD_TC: unset or "auto" = 0, "1" = tcgen05,
// "0" = mma.sync), otherwise the mode set through mvptr_attn_set_path
static int g_fwd_mode = -1, g_bwd_mode = -1;
static int env_mode(const char* name) {
  const char* v = getenv(name);
  if (!v || !*v || v[0] == 'a') return 0;
  return atoi(v) != 0 ? 1 : 2;
}
extern "C" int mvptr_attn_set_path(int fwd_mode, int bwd_mode) {
  if (fwd_mode < -1 || fwd_mode > 2 || bwd_mode < -1 || bwd_mode > 2)
    MVPTR_FAIL(MVPTR_ERR_ARG, "attn_set_path: modes are -1 (environment), 0 (auto), 1 (tcgen05), 2 (mma.sync)");
  g_fwd_mode = fwd_mode;
  g_bwd_mode = bwd_mode;
  return 0;
}

template <int R>
static int launch_fwd_tc(const CUtensorMap& tq, const CUtensorMap& to, const mvptr::attn_tc::Params& p, cudaStream_t stream) {
  using namespace mvptr;
  using namespace mvptr::attn_tc;
  auto kern = attn_fwd_tc_kernel<R>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdCfg<R>::kSmem);
    if (e != cudaSuccess) MVPTR_FAIL(MVPTR_ERR_CUDA, "attention (tcgen05) smem attr: %s", cudaGetErrorString(e));
    configured = true;
  }
  static_assert(FwdCfg<R>::kSmem <= 227 * 1024, "shared memory budget");
  const int grid = p.items < kNumSMs ? p.items : kNumSMs;
  kern<<<grid, kThreads, FwdCfg<R>::kSmem, stream>>>(tq, to, p);
  MVPTR_CHECK_LAUNCH("attn_fwd_tc");
  return 0;
}

int mvptr_attn_fwd_tc(const void* qkv, int ld_qkv, const float* maskadd, void* ctx, int ld_ctx, float* lse, int B, int L,
                      int nh, int H, float p_drop, uint32_t seed, cudaStream_t stream) {
  using namespace mvptr;
  using namespace mvptr::attn_tc;
  // 0 = auto (the measured choice, profiles/r2_attention_tc_v4_microbench.txt: the tcgen05 kernel where it is at least
  // as fast as the mma.sync one -- the 80 < L <= 96 class, i.e. the cross-modal encoder of the pre-training step),
  // 1 = tcgen05 for every L <= 128, 2 = mma.sync only.  mvptr_attn_set_path / MVPTR_ATTN_FWD_TC select it.
  const int mode = g_fwd_mode >= 0 ? g_fwd_mode : env_mode("MVPTR_ATTN_FWD_TC");
  const bool use = mode == 1 ? true : mode == 2 ? false : (L > 80 && L <= 96);
  if (!use || L > kRows || (reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(ctx) & 15)) return 1;
  const int R = L <= 64 ? 64 : L <= 96 ? 96 : 128;
  CUtensorMap tq, to;
  if (int rc = make_map3(&tq, qkv, 3 * H, L, B, ld_qkv, R)) return rc;
  if (int rc = make_map3(&to, ctx, H, L, B, ld_ctx, 32)) return rc;  // one warp's 32 rows per store box
  Params p;
  p.maskadd = maskadd;
  p.lse = lse;
  p.B = B; p.L = L; p.nh = nh; p.H = H;
  p.n16 = (L + 15) & ~15;
  p.items = B * nh;
  p.scale_log2 = 0.125f * kLog2e;
  p.keep_thr = keep_threshold(p_drop);
  p.inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  p.seed = seed;
  if (R == 64) return launch_fwd_tc<64>(tq, to, p, stream);
  if (R == 96) return launch_fwd_tc<96>(tq, to, p, stream);
  return launch_fwd_tc<128>(tq, to, p, stream);
}

template <int NS>
static int launch_bwd_tc(const CUtensorMap& tq, const CUtensorMap& tdo, const CUtensorMap& tdq,
                         const mvptr::attn_tc::BwdParams& p, cudaStream_t stream) {
  using namespace mvptr;
  using namespace mvptr::attn_tc;
  auto kern = attn_bwd_tc_kernel<NS>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
    if (e != cudaSuccess) MVPTR_FAIL(MVPTR_ERR_CUDA, "attention backward (tcgen05) smem attr: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int grid = p.items < kNumSMs ? p.items : kNumSMs;
  kern<<<grid, kThreads, kBwdSmem, stream>>>(tq, tdo, tdq, p);
  MVPTR_CHECK_LAUNCH("attn_bwd_tc");
  return 0;
}

// Returns 1 when this path does not apply (the caller falls back to the mma.sync kernels), 0 on success, < 0 on error.
int mvptr_attn_bwd_tc(const void* qkv, int ld_qkv, const float* maskadd, const void* dctx, int ld_ctx, const float* lse,
                      void* dqkv, float* dbias, int B, int L, int nh, int H, float p_drop, uint32_t seed,
                      cudaStream_t stream) {
  using namespace mvptr;
  using namespace mvptr::attn_tc;
  // auto = mma.sync: the tcgen05 backward is parity-green but 5-15 % slower at the step's shapes (same profile file)
  const int mode = g_bwd_mode >= 0 ? g_bwd_mode : env_mode("MVPTR_ATTN_BWD_TC");
  if (mode != 1 || L > kRows || ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(dctx) |
                                  reinterpret_cast<uintptr_t>(dqkv)) & 15))
    return 1;
  if (H > kBwdMaxH) return 1;
  CUtensorMap tq, tdo, tdq;
  if (int rc = make_map3(&tq, qkv, 3 * H, L, B, ld_qkv)) return rc;
  if (int rc = make_map3(&tdo, dctx, H, L, B, ld_ctx)) return rc;
  if (int rc = make_map3(&tdq, dqkv, 3 * H, L, B, ld_qkv, 32)) return rc;  // 32-row store boxes
  BwdParams p;
  p.maskadd = maskadd;
  p.lse = lse;
  p.dbias = dbias;
  p.B = B; p.L = L; p.nh = nh; p.H = H;
  p.n16 = (L + 15) & ~15;
  p.items = B * nh;
  p.scale = 0.125f;
  p.scale_log2 = 0.125f * kLog2e;
  p.keep_thr = keep_threshold(p_drop);
  p.inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  p.seed = seed;
  switch ((p.n16 + 31) >> 5) {
    case 1: return launch_bwd_tc<1>(tq, tdo, tdq, p, stream);
    case 2: return launch_bwd_tc<2>(tq, tdo, tdq, p, stream);
    case 3: return launch_bwd_tc<3>(tq, tdo, tdq, p, stream);
    default: return launch_bwd_tc<4>(tq, tdo, tdq, p, stream);
  }
}

// this translation unit's copy of the dropout epoch word (common.cuh): without it the forward here and the mma.sync
// backward in attention.cu would hash different masks as soon as a CUDA-graph replay advances the epoch
MVPTR_DEFINE_EPOCH_SETTER(mvptr_set_epoch_attn_tc)

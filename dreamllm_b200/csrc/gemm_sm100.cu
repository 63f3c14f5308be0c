// Persistent warp-specialised tcgen05 GEMM for sm_100a (bf16 x bf16 -> fp32 in TMEM -> bf16|fp32).
//
//   C[M,N] = op(A) * op(B)      A "K-major":  memory [M, K] row-major        (x, dY)
//                               A "MN-major": memory [K, M] row-major        (dY^T for wgrad)
//                               B "K-major":  memory [N, K] row-major        (nn.Linear weight, fwd)
//                               B "MN-major": memory [K, N] row-major        (weight for dgrad, x for wgrad)
//
// This single kernel serves every dense contraction on the DreamLLM hot path (reference call sites:
// modeling_dreamllm.py:336-338, :395 (q/k/v/o_proj), :237 (gate/up/down_proj), :1452 (lm_head), and their
// autograd dgrad/wgrad), replacing the cuBLAS calls torch makes for nn.Linear.
//
// Structure (one CTA per SM, or one CTA *pair* per two SMs with cta_group::2):
//   warp 0    TMA producer: global -> 128B-swizzled smem stages, mbarrier complete_tx
//   warp 1    MMA issuer (one elected thread): tcgen05.mma, accumulators in TMEM, double-buffered (2 x 256 cols)
//   warp 2    TMEM allocator
//   warps 4-7 epilogue: tcgen05.ld -> convert -> swizzled smem -> per-warp TMA store (overlaps next tile's MMAs)
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "gemm_sm100.h"

namespace dllm {

constexpr int BM = 128;  // A rows per CTA == TMEM lanes
constexpr int BN = 256;  // default UMMA N (kBN = 128 is instantiated for outputs whose width tiles badly by 256: see pick_bn)
constexpr int BK = 64;   // 64 bf16 = one 128-byte swizzle line
constexpr int UMMA_K = 16;
// epilogue warps: 4 (one per TMEM lane quarter; 6/4 operand stages) for mainloop-bound shapes, 8 (two per quarter, each draining half
// of the tile's columns; 5/3 stages) for epilogue-bound ones (small K, fused epilogues).  Same-box A/B: profiles/r01_gemm_smallk_epilogue.md
constexpr int gemm_threads(int ew) { return 128 + 32 * ew; }
constexpr int kEpiWarp0 = 4;

#ifndef DLLM_EPI_BUFS
#define DLLM_EPI_BUFS 2   // staging buffers per epilogue warp (TMA stores in flight per warp)
#endif

template <int kCta, int kEpiWarps = 4, int kBN = BN>
struct GemmCfg {
  static constexpr int kBRows = kBN / kCta;  // rows of B each CTA loads
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiBufs = DLLM_EPI_BUFS;
  static constexpr int kEpiBufBytes = 32 * 128;                // 32 rows x 128 B per warp-store
  static constexpr int kEpiBytes = kEpiWarps * kEpiBufs * kEpiBufBytes;
  static constexpr int kBarBytes = 1024;
  // operand stages + store staging + barriers <= 227 KB:  kBN 256: 6 / 5 (2-CTA, 4 / 8 epilogue warps), 4 / 3 (1-CTA);  kBN 128: 8 / 6, 6 / 5
  static constexpr int kStagesFit = (227 * 1024 - kEpiBytes - kBarBytes - 1024) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kBarBytes + 1024 /*align slack*/;
};

// Fused epilogue:  out = act( bf16(acc + bias[col]) ) (+ residual[row, col])   — rounding points as the reference's
// unfused bf16 sequence (Linear output rounded, activation rounded, residual add rounded).
struct GemmEpi {
  const bf16* bias;      // [N] or nullptr
  const bf16* residual;  // [M, ldr] or nullptr
  long ldr;
  int act;               // 0 none, 1 quick_gelu x*sigmoid(1.702x), 2 gelu (erf), 3 silu
  const bf16* rowbias;   // [M / rows_per_group, N] or nullptr: per-row-group bias (ResnetBlock2D's time-embedding add)
  int rows_per_group;
};

// Implicit-GEMM 3x3 / stride 1 / pad 1 convolution on NHWC activations: A[m, k] = x[n, h + r - 1, w + s - 1, c] with
// m = (n, h, w), k = (r, s, c).  A 128-row tile is the 4-D TMA box {64 c, Wb, Hb, Nb} (Wb*Hb*Nb = 128) fetched at
// coordinates (c0, s - 1, h0 + r - 1, n0): out-of-range rows/columns are zero-filled by TMA == the conv's zero padding.
struct ConvGeom {
  int cpt;       // Cin / 64: K blocks per filter tap
  int W, H;      // plane
  int HW;
};

__device__ __forceinline__ float epi_act(float x, int act) {
  if (act == 1) return x / (1.f + expf(-1.702f * x));
  if (act == 2) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
  if (act == 3) return x / (1.f + expf(-x));
  return x;
}

// kEpi: 0 = plain store, 1 = + bias / row-group bias / residual, 2 = 1 + activation.  Compile-time so that the plain and the
// bias-only epilogues carry none of the (inlined expf / erff) activation code: with a runtime switch every element paid ~78 issue
// slots of predicated-off instructions and bias GEMMs with small K ran at 200 TF/s (profiles/r01_gemm_smallk_epilogue.md).
template <int kCta, bool kAMN, bool kBMN, typename OutT, bool kConv = false, int kEpi = 0, int kEpiWarps = 4, int kBN = BN>
__global__ void __launch_bounds__(gemm_threads(kEpiWarps), 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const __grid_constant__ CUtensorMap tma_c, int M, int N, int K, int group_m, GemmEpi epi, ConvGeom cg,
            int* __restrict__ tile_counter, int ksplit) {
  // ksplit > 1 (split-K, small-M shapes that would leave most SMs idle): a work unit is (output tile, K slice); unit u covers K blocks
  // [ks * kb_per, min(num_kb, (ks + 1) * kb_per)) of tile u / ksplit and stores its fp32 partial at rows ks * Mp + ... of the workspace the
  // C tensor map describes (Mp = M rounded up to whole tiles); splitk_reduce_kernel sums the slices and applies the epilogue.
  using Cfg = GemmCfg<kCta, kEpiWarps, kBN>;
  constexpr int kStages = Cfg::kStages;
  constexpr bool kOutF32 = sizeof(OutT) == 4;
  constexpr int CH = kOutF32 ? 32 : 64;  // output columns per 128-byte store line
  constexpr int kChunks = kBN / CH;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Cfg::kEpiBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full_bar = bars + 2 * kStages;
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  // dynamic tile scheduler: the leader's TMA thread draws tile ids from a global atomic counter and publishes them through a
  // small smem ring to every role of the CTA pair.  A CTA that reaches an SM late (e.g. behind an overlapped NCCL all-reduce
  // kernel) simply draws fewer tiles, instead of serialising a fixed 1/74th of the problem behind everyone else.
  constexpr int kRing = 4;
  uint64_t* ring_full = bars + 2 * kStages + 6;
  uint64_t* ring_empty = ring_full + kRing;
  uint32_t* tile_ring = reinterpret_cast<uint32_t*>(ring_empty + kRing);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kCta == 2) ? cluster_ctarank() : 0u;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_c);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], kEpiWarps * kCta);  // one arrive per epilogue warp of every CTA in the pair
    }
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&ring_full[i], 1);
      mbar_init(&ring_empty[i], (kEpiWarps + 1) * kCta);  // leader: MMA + epilogue warps; peer: TMA thread + epilogue warps
    }
    fence_mbar_init();
  }
  if (warp_idx == 2) {
    tmem_alloc<kCta>(tmem_ptr_smem, 512);
    tmem_relinquish<kCta>();
  }
  tc_fence_before();
  if constexpr (kCta == 2) cluster_sync_all();
  __syncthreads();   // (2-CTA: barrier.cluster already orders the allocator's smem write; the CTA barrier keeps compute-sanitizer's
                     //  racecheck, which does not model cluster barriers, from flagging the tmem_ptr hand-off — profiles/r02f_sanitizer.md)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // ---- persistent tile schedule (identical sequence in every role) ----
  const int tile_m = BM * kCta;
  const int num_m_tiles = (M + tile_m - 1) / tile_m;
  const int num_n_tiles = (N + kBN - 1) / kBN;
  const int num_tiles = num_m_tiles * num_n_tiles * ksplit;       // work units
  const int num_kb = (K + BK - 1) / BK;
  const int kb_per = (num_kb + ksplit - 1) / ksplit;
  const int split_rows = num_m_tiles * tile_m;                      // Mp: row pitch between K slices in the partial workspace
  const int tiles_per_group = group_m * num_n_tiles;
  const uint32_t ring_empty0_leader = (kCta == 2) ? mapa_shared(smem_u32(&ring_empty[0]), 0) : 0u;

  // consumer side of the tile ring (every role except the leader's TMA thread)
  auto ring_pop = [&](int& rslot, uint32_t& rphase, bool arrive) -> int {
    if constexpr (kCta == 2) mbar_wait_cluster(&ring_full[rslot], rphase, 5); else mbar_wait(&ring_full[rslot], rphase, 5);
    const int t = static_cast<int>(tile_ring[rslot]);
    if (arrive) {
      if constexpr (kCta == 2) mbar_arrive_cluster(ring_empty0_leader + rslot * 8); else mbar_arrive(&ring_empty[rslot]);
    }
    if (++rslot == kRing) { rslot = 0; rphase ^= 1; }
    return t;
  };

  auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
    const int g = tile / tiles_per_group;
    const int first_m = g * group_m;
    const int gsize = min(group_m, num_m_tiles - first_m);
    const int r = tile - g * tiles_per_group;
    m_blk = first_m + r % gsize;
    n_blk = r / gsize;
  };

  // producer / issuer: one lane picked with elect.sync.  Unlike `lane == 0` it tells ptxas the region is single-threaded, so the TMA and
  // tcgen05.mma instructions are emitted back to back instead of each inside a per-active-thread ELECT/BRA loop (21 scalar instructions
  // between two UTCHMMAs before: at 64-128 clk of tensor pipe per instruction the issuing thread was the limit for BN = 128 tiles)
  if (warp_idx == 0) { if (elect_one_sync()) {
    // ======================= TMA producer =======================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t full0_cluster = (kCta == 2) ? mapa_shared(smem_u32(&full_bar[0]), 0) : 0u;
    int rslot = 0;
    uint32_t rphase = 0;
    while (true) {
      int tile;
      if (cta_rank == 0) {
        mbar_wait(&ring_empty[rslot], rphase ^ 1, 6);
        tile = atomicAdd(tile_counter, 1);
        tile_ring[rslot] = static_cast<uint32_t>(tile);
        if constexpr (kCta == 2) {
          st_shared_cluster_u32(mapa_shared(smem_u32(&tile_ring[rslot]), 1), static_cast<uint32_t>(tile));
          mbar_arrive_cluster(mapa_shared(smem_u32(&ring_full[rslot]), 1));
        }
        mbar_arrive(&ring_full[rslot]);
        if (++rslot == kRing) { rslot = 0; rphase ^= 1; }
      } else {
        tile = ring_pop(rslot, rphase, true);
      }
      if (tile >= num_tiles) break;
      int m_blk, n_blk;
      tile_coords(tile / ksplit, m_blk, n_blk);
      const int ks = tile % ksplit;
      const int m0 = m_blk * tile_m + static_cast<int>(cta_rank) * BM;
      const int n0 = n_blk * kBN + static_cast<int>(cta_rank) * Cfg::kBRows;
      for (int kb = ks * kb_per; kb < min(num_kb, (ks + 1) * kb_per); ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 1);
        uint8_t* a_s = smem + stage * Cfg::kStageBytes;
        uint8_t* b_s = a_s + Cfg::kABytes;
        const int k0 = kb * BK;
        // implicit-conv coordinates of this CTA's 128-pixel tile and this K block's filter tap
        int cv_c = 0, cv_w = 0, cv_h = 0, cv_n = 0;
        if constexpr (kConv) {
          const int tap = kb / cg.cpt;
          cv_c = (kb - tap * cg.cpt) * BK;
          const int r = tap / 3, sx = tap - 3 * r;
          cv_n = m0 / cg.HW;
          const int rem = m0 - cv_n * cg.HW;
          const int h0 = rem / cg.W;
          cv_h = h0 + r - 1;
          cv_w = (rem - h0 * cg.W) + sx - 1;   // non-zero only for planes wider than one 128-pixel tile
        }
        if constexpr (kCta == 1) {
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if constexpr (kConv) {
            tma_load_4d(a_s, &tma_a, &full_bar[stage], cv_c, cv_w, cv_h, cv_n);
          } else if constexpr (!kAMN) {
            tma_load_2d(a_s, &tma_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d(a_s + i * 8192, &tma_a, &full_bar[stage], m0 + i * 64, k0);
          }
          if constexpr (!kBMN) {
            tma_load_2d(b_s, &tma_b, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int i = 0; i < Cfg::kBRows / 64; ++i)
              tma_load_2d(b_s + i * 8192, &tma_b, &full_bar[stage], n0 + i * 64, k0);
          }
        } else {
          // both CTAs' bytes are counted on the leader's barrier
          if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          const uint32_t fb = full0_cluster + stage * 8;
          if constexpr (kConv) {
            tma_load_4d_2cta(a_s, &tma_a, fb, cv_c, cv_w, cv_h, cv_n);
          } else if constexpr (!kAMN) {
            tma_load_2d_2cta(a_s, &tma_a, fb, k0, m0);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d_2cta(a_s + i * 8192, &tma_a, fb, m0 + i * 64, k0);
          }
          if constexpr (!kBMN) {
            tma_load_2d_2cta(b_s, &tma_b, fb, k0, n0);
          } else {
#pragma unroll
            for (int i = 0; i < Cfg::kBRows / 64; ++i) tma_load_2d_2cta(b_s + i * 8192, &tma_b, fb, n0 + i * 64, k0);
          }
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } } else if (warp_idx == 1) { if (elect_one_sync() && cta_rank == 0) {
    // ======================= MMA issuer (leader CTA only) =======================
    constexpr uint32_t idesc = make_idesc_bf16(BM * kCta, kBN, kAMN, kBMN);
    constexpr uint32_t a_adv = kAMN ? 2048u : 32u;  // bytes per UMMA_K step
    constexpr uint32_t b_adv = kBMN ? 2048u : 32u;
    constexpr uint32_t a_lbo = kAMN ? 8192u : 16u;
    constexpr uint32_t b_lbo = kBMN ? 8192u : 16u;
    const uint64_t adesc0 = make_smem_desc(smem_u32(smem), a_lbo, 1024);
    const uint64_t bdesc0 = make_smem_desc(smem_u32(smem) + Cfg::kABytes, b_lbo, 1024);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    int rslot = 0;
    uint32_t rphase = 0;
    for (;; ++it) {
      const int tile = ring_pop(rslot, rphase, true);
      if (tile >= num_tiles) break;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_empty_bar[as], aphase ^ 1, 2);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * kBN;
      const int kb0 = (tile % ksplit) * kb_per, kb1 = min(num_kb, kb0 + kb_per);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase, 3);
        tc_fence_after();
        // descriptors = stage-0 descriptor + constants (the start-address field counts 16-byte units; no tile crosses the 256 KB it spans)
        const uint64_t adesc = adesc0 + static_cast<uint64_t>(stage * (Cfg::kStageBytes >> 4));
        const uint64_t bdesc = bdesc0 + static_cast<uint64_t>(stage * (Cfg::kStageBytes >> 4));
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_ss<kCta>(tmem_d, adesc + ((k * a_adv) >> 4), bdesc + ((k * b_adv) >> 4), idesc, (kb != kb0 || k != 0) ? 1u : 0u);
        if constexpr (kCta == 1) umma_commit(&empty_bar[stage]); else umma_commit_2cta(&empty_bar[stage], 0b11);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if constexpr (kCta == 1) umma_commit(&tmem_full_bar[as]); else umma_commit_2cta(&tmem_full_bar[as], 0b11);
    }
  } } else if (warp_idx >= kEpiWarp0) {
    // ======================= epilogue warps =======================
    const int ew = warp_idx - kEpiWarp0;  // 0..7
    const int wq = ew & 3;                // TMEM lane quarter: this warp may touch lanes [32*wq, 32*wq+32)
    const int chalf = ew >> 2;            // which part of the tile's column chunks this warp drains
    constexpr int kParts = kEpiWarps / 4;
    uint8_t* my_epi = epi_smem + ew * Cfg::kEpiBufs * Cfg::kEpiBufBytes;
    const uint32_t tmem_empty0_cluster = (kCta == 2) ? mapa_shared(smem_u32(&tmem_empty_bar[0]), 0) : 0u;
    int it = 0;
    int buf = 0;
    int rslot = 0;
    uint32_t rphase = 0;
    for (;; ++it) {
      const int tile = ring_pop(rslot, rphase, false);
      __syncwarp();
      if (lane == 0) {
        const int s_prev = (rslot == 0) ? kRing - 1 : rslot - 1;
        if constexpr (kCta == 2) mbar_arrive_cluster(ring_empty0_leader + s_prev * 8); else mbar_arrive(&ring_empty[s_prev]);
      }
      if (tile >= num_tiles) break;
      int m_blk, n_blk;
      tile_coords(tile / ksplit, m_blk, n_blk);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row0 = (tile % ksplit) * (ksplit > 1 ? split_rows : 0) + m_blk * tile_m + static_cast<int>(cta_rank) * BM + wq * 32;
      const int col0 = n_blk * kBN;
      mbar_wait(&tmem_full_bar[as], aphase, 4);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + as * kBN;
      if constexpr (kEpi == 3) {
        // GEGLU epilogue (diffusers FeedForward: h, gate = proj(x).chunk(2); out = h * gelu(gate)).  The weight rows were permuted at
        // load time so that every 128-column group of this GEMM's N space is [64 value columns | their 64 gate columns]; one warp turns
        // such a pair of accumulator chunks into ONE 64-column chunk of the [M, N/2] output — the [M, N] projection is never written.
        // Rounding points = the unfused sequence (Linear output to bf16, gelu(gate) to bf16, product to bf16): bit-identical results.
        static_assert(kEpi != 3 || (!kOutF32 && (kChunks % (2 * kParts)) == 0), "GEGLU epilogue: bf16 out, whole chunk pairs per warp");
        constexpr int kPairsPerWarp = kChunks / 2 / kParts;
#pragma unroll 1
        for (int pr = chalf * kPairsPerWarp; pr < (chalf + 1) * kPairsPerWarp; ++pr) {
          uint32_t va[64], vg[64];
          tmem_ld32(taddr0 + pr * 128, va);
          tmem_ld32(taddr0 + pr * 128 + 32, va + 32);
          tmem_ld32(taddr0 + pr * 128 + 64, vg);
          tmem_ld32(taddr0 + pr * 128 + 96, vg + 32);
          tmem_ld_wait();
          const int gcol = col0 + pr * 128;   // column of the value chunk in the permuted N space (gate chunk at +64)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int cj = gcol + j * 8;
            float bv[8], bg[8];
            if (epi.bias != nullptr && cj + 128 <= N + 64) {
              const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(epi.bias + cj));
              const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(epi.bias + cj + 64));
              const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&q0);
              const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&q1);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f0 = __bfloat1622float2(h0[e]), f1 = __bfloat1622float2(h1[e]);
                bv[2 * e] = f0.x; bv[2 * e + 1] = f0.y; bg[2 * e] = f1.x; bg[2 * e + 1] = f1.y;
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) bv[e] = bg[e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float a = __bfloat162float(__float2bfloat16_rn(__uint_as_float(va[j * 8 + e]) + bv[e]));
              const float gte = __bfloat162float(__float2bfloat16_rn(__uint_as_float(vg[j * 8 + e]) + bg[e]));
              const float gel = __bfloat162float(__float2bfloat16_rn(0.5f * gte * (1.f + erff(gte * 0.70710678118654752f))));
              va[j * 8 + e] = __float_as_uint(a * gel);
            }
          }
          if (lane == 0) tma_store_wait_read<Cfg::kEpiBufs - 1>();
          __syncwarp();
          uint8_t* dst = my_epi + buf * Cfg::kEpiBufBytes + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 q;
            q.x = pack_bf16(__uint_as_float(va[8 * j + 0]), __uint_as_float(va[8 * j + 1]));
            q.y = pack_bf16(__uint_as_float(va[8 * j + 2]), __uint_as_float(va[8 * j + 3]));
            q.z = pack_bf16(__uint_as_float(va[8 * j + 4]), __uint_as_float(va[8 * j + 5]));
            q.w = pack_bf16(__uint_as_float(va[8 * j + 6]), __uint_as_float(va[8 * j + 7]));
            *reinterpret_cast<uint4*>(dst + ((j ^ (lane & 7)) << 4)) = q;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (row0 < M && gcol < N) tma_store_2d(&tma_c, my_epi + buf * Cfg::kEpiBufBytes, (col0 >> 1) + pr * 64, row0);
            tma_store_commit();
          }
          buf = (buf + 1 == Cfg::kEpiBufs) ? 0 : buf + 1;
        }
      } else
#pragma unroll 1
      for (int c = chalf * (kChunks / kParts); c < (chalf + 1) * (kChunks / kParts); ++c) {
        uint32_t v[kOutF32 ? 32 : 64];
        tmem_ld32(taddr0 + c * CH, v);
        if constexpr (!kOutF32) tmem_ld32(taddr0 + c * CH + 32, v + 32);
        tmem_ld_wait();
        if constexpr (kEpi > 0) {
          const int gcol = col0 + c * CH;
          const int grow = row0 + lane;
#pragma unroll
          for (int j = 0; j < CH / 8; ++j) {
            const int cj = gcol + j * 8;
            const bool col_ok = cj + 8 <= N;
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[j * 8 + e]);
            if (epi.bias != nullptr && col_ok) {
              const uint4 q = __ldg(reinterpret_cast<const uint4*>(epi.bias + cj));
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
              for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h2[e]); x[2 * e] += f.x; x[2 * e + 1] += f.y; }
            }
            if (epi.rowbias != nullptr && col_ok && grow < M) {
              const uint4 q = __ldg(reinterpret_cast<const uint4*>(epi.rowbias + static_cast<size_t>(grow / epi.rows_per_group) * N + cj));
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
              for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h2[e]); x[2 * e] += f.x; x[2 * e + 1] += f.y; }
            }
            if constexpr (kEpi == 2) {
#pragma unroll
              for (int e = 0; e < 8; ++e) x[e] = __bfloat162float(__float2bfloat16_rn(x[e]));
              if (epi.act == 1) {
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = x[e] / (1.f + __expf(-1.702f * x[e]));
              } else if (epi.act == 2) {
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = 0.5f * x[e] * (1.f + erff(x[e] * 0.70710678118654752f));
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = x[e] / (1.f + __expf(-x[e]));
              }
            }
            if (epi.residual != nullptr && col_ok && grow < M) {
              const uint4 q = *reinterpret_cast<const uint4*>(epi.residual + static_cast<size_t>(grow) * epi.ldr + cj);
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h2[e]);
                x[2 * e] = __bfloat162float(__float2bfloat16_rn(x[2 * e])) + f.x;
                x[2 * e + 1] = __bfloat162float(__float2bfloat16_rn(x[2 * e + 1])) + f.y;
              }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) v[j * 8 + e] = __float_as_uint(x[e]);
          }
        }
        // the staging buffer we are about to overwrite was handed to TMA two stores ago
        if (lane == 0) tma_store_wait_read<Cfg::kEpiBufs - 1>();
        __syncwarp();
        uint8_t* dst = my_epi + buf * Cfg::kEpiBufBytes + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 q;
          if constexpr (kOutF32) {
            q = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            q.x = pack_bf16(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
            q.y = pack_bf16(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
            q.z = pack_bf16(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
            q.w = pack_bf16(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
          }
          *reinterpret_cast<uint4*>(dst + ((j ^ (lane & 7)) << 4)) = q;  // 128B swizzle: chunk ^= row % 8
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if ((ksplit > 1 || row0 < M) && col0 + c * CH < N)
            tma_store_2d(&tma_c, my_epi + buf * Cfg::kEpiBufBytes, col0 + c * CH, row0);
          tma_store_commit();
        }
        buf = (buf + 1 == Cfg::kEpiBufs) ? 0 : buf + 1;
      }
      // all TMEM reads of this accumulator are complete -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCta == 1) mbar_arrive(&tmem_empty_bar[as]); else mbar_arrive_cluster(tmem_empty0_cluster + as * 8);
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  // ---- teardown ----
  tc_fence_before();
  if constexpr (kCta == 2) cluster_sync_all(); else __syncthreads();
  if (warp_idx == 2) tmem_dealloc<kCta>(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2D row-major tensor [rows, cols] with row stride `ld` elements; box = [box_rows, box_cols]; 128B swizzle.
int make_tmap_2d(CUtensorMap* map, const void* ptr, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return DLLM_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * elem_bytes) & 15)) return DLLM_ERR_ALIGN;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(map, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    fprintf(stderr, "dllm: cuTensorMapEncodeTiled failed (CUresult %d): ptr=%p elem=%d rows=%llu cols=%llu ld=%llu box=[%u x %u]\n",
            (int)r, ptr, elem_bytes, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
  return r == CUDA_SUCCESS ? 0 : DLLM_ERR_TMAP;
}

// SMs left free for a concurrently running collective (NCCL all-reduce overlapped with backward): a persistent GEMM that
// asks for every SM would have the CTAs that cannot be scheduled wait for the collective kernel to finish and then run
// their whole static tile list alone (measured: 2-GPU step 674 ms vs 622 ms single).  See dreamllm_b200/ddp.py.
static int g_reserved_sms = 0;
void set_reserved_sms(int n) { g_reserved_sms = n < 0 ? 0 : n; }
int reserved_sms() { return g_reserved_sms; }

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      return 148;   // no device visible (build host): plan for a B200 so the workspace queries stay meaningful
    }
  }
  return n;
}

// per-device ring of tile counters for the dynamic scheduler (one-time 4 KiB scratch; each launch zeroes its slot in stream order)
int* tile_counter_slot(cudaStream_t stream) {
  constexpr int kSlots = 1024, kMaxDev = 16;
  static int* base[kMaxDev] = {nullptr};
  static unsigned next[kMaxDev] = {0};
  static std::mutex mu;   // callers may be autograd worker threads of several devices
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDev) return nullptr;
  int* slot;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!base[dev]) {
      if (cudaMalloc(&base[dev], kSlots * sizeof(int)) != cudaSuccess) return nullptr;
      cudaMemset(base[dev], 0, kSlots * sizeof(int));
    }
    slot = base[dev] + (next[dev]++ % kSlots);
  }
  if (cudaMemsetAsync(slot, 0, sizeof(int), stream) != cudaSuccess) return nullptr;
  return slot;
}

// N-tile width.  Measured on the UNet / VAE shapes (profiles/r02d_gemm_bn_{auto,256}.json): a 128-wide N tile reads the same A bytes
// from shared memory for half the MMA work, and the 2-CTA kernel then runs at ~0.55-0.7x the MMA rate of the 256-wide one — which
// more than cancels the padded-work saving at N = 320 / 640 / 960 (conv 320->320: 388 us at 128 vs 275 us at 256).  So 256 everywhere
// except outputs that fit ONE 128-wide tile (VAE C = 128: 459 vs 476 us).  DLLM_GEMM_BN=128|256 forces a width for A/B runs.
static inline int pick_bn(int N) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("DLLM_GEMM_BN");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 128 || forced == 256) return forced;
  if (forced == 1) {   // A/B: the width that pads N the least (N = 320: 384 vs 512 columns of MMA work; 640: 640 vs 768)
    const int w256 = (N + 255) / 256 * 256, w128 = (N + 127) / 128 * 128;
    return w128 < w256 ? 128 : 256;
  }
  return N <= 128 ? 128 : 256;
}

template <int kCta, bool kAMN, bool kBMN, typename OutT, bool kConv = false, int kEpi = 0, int kEW = 4, int kBN = BN>
static int launch_gemm_bn(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, int M, int N, int K,
                          const GemmEpi& epi, cudaStream_t stream, ConvGeom cg = ConvGeom{1, 1, 1, 1}, int ksplit = 1) {
  using Cfg = GemmCfg<kCta, kEW, kBN>;
  auto kern = gemm_kernel<kCta, kAMN, kBMN, OutT, kConv, kEpi, kEW, kBN>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return DLLM_ERR_LAUNCH;
    attr_set = true;
  }
  const int tile_m = BM * kCta;
  const int num_tiles = ((M + tile_m - 1) / tile_m) * ((N + kBN - 1) / kBN) * ksplit;
  int avail = num_sms() - g_reserved_sms;
  if (avail < 2) avail = 2;
  const int max_clusters = avail / kCta;
  const int clusters = num_tiles < max_clusters ? num_tiles : max_clusters;
  const int group_m = (kCta == 2) ? 8 : 16;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * kCta);
  cfg.blockDim = dim3(gemm_threads(kEW));
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int* counter = tile_counter_slot(stream);
  if (!counter) return DLLM_ERR_LAUNCH;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, M, N, K, group_m, epi, cg, counter, ksplit);
  return e == cudaSuccess ? 0 : DLLM_ERR_LAUNCH;
}
// `bn` must be the value the B tensor map's box was built for (pick_bn(N))
template <int kCta, bool kAMN, bool kBMN, typename OutT, bool kConv = false, int kEpi = 0, int kEW = 4>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, int M, int N, int K,
                       const GemmEpi& epi, cudaStream_t stream, ConvGeom cg = ConvGeom{1, 1, 1, 1}, int bn = BN) {
  if (bn == 128) return launch_gemm_bn<kCta, kAMN, kBMN, OutT, kConv, kEpi, kEW, 128>(ta, tb, tc, M, N, K, epi, stream, cg);
  return launch_gemm_bn<kCta, kAMN, kBMN, OutT, kConv, kEpi, kEW, 256>(ta, tb, tc, M, N, K, epi, stream, cg);
}

// ------------------------------------------------------------------------------------------------ split-K (small-M shapes)
// A [256 x 1280] output over K = 11520 (UNet conv 1280 -> 1280 on the 8x8 plane of 4 samples) is 5 tiles: 5 of 74 CTA pairs busy for
// 180 K blocks each.  The stage-1 step (C5) spends 12.8 of its 29.6 ms of GEMM / conv time in launches that fill <= 80 of 148 SMs
// (profiles/r02f_stage1_launch_list_summary.md).  Split-K hands every (tile, K slice) to its own CTA pair, writes fp32 partials to a
// caller-provided workspace and lets one small kernel sum them and apply the epilogue with the usual rounding points.
struct SplitPlan {
  int kcta, bn, ksplit, mp;   // mp = M rounded up to whole tiles
  size_t ws_bytes;
};
static SplitPlan plan_split(int M, int N, int K, int kcta_hint) {
  SplitPlan p;
  p.kcta = kcta_hint;
  p.bn = pick_bn(N);
  const int tile_m = BM * p.kcta;
  const int mt = (M + tile_m - 1) / tile_m, nt = (N + p.bn - 1) / p.bn;
  const int tiles = mt * nt, num_kb = (K + BK - 1) / BK;
  int slots = (num_sms() - g_reserved_sms) / p.kcta;
  if (slots < 1) slots = 1;
  int ks = 1;
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("DLLM_GEMM_NO_SPLITK"); disabled = (e && e[0] == '1') ? 1 : 0; }
  static int min_ks = -1;   // DLLM_GEMM_SPLITK_MIN: smallest slice count worth the fp32 round trip + reduce launch (default 2)
  if (min_ks < 0) { const char* e = getenv("DLLM_GEMM_SPLITK_MIN"); min_ks = (e && atoi(e) >= 2) ? atoi(e) : 2; }
  if (!disabled && tiles * min_ks <= slots && num_kb >= 8) {
    ks = slots / tiles;
    if (ks > num_kb / 4) ks = num_kb / 4;    // at least 4 K blocks per slice
    if (ks > 16) ks = 16;
    if (ks < min_ks) ks = 1;
    // no empty slice: the kernel gives every slice ceil(num_kb / ks) K blocks, so the last of e.g. 16 slices of 90 blocks would start
    // past the end and store an accumulator no MMA ever initialised
    if (ks > 1) { const int per = (num_kb + ks - 1) / ks; ks = (num_kb + per - 1) / per; }
  }
  p.ksplit = ks;
  p.mp = mt * tile_m;
  p.ws_bytes = ks > 1 ? static_cast<size_t>(ks) * p.mp * N * sizeof(float) : 0;
  return p;
}
size_t gemm_splitk_workspace(int M, int N, int K) { return plan_split(M, N, K, M > BM ? 2 : 1).ws_bytes; }

// out[m, n] = epilogue( sum_ks ws[ks * mp + m, n] ): bias / row-group bias added in fp32, rounded to bf16, optional activation (rounded),
// optional residual add (rounded) — the rounding points of gemm_kernel's fused epilogues.  One thread = 8 columns.
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, bf16* __restrict__ out, long ldc, int M, int N, int mp, int ksplit,
                                     GemmEpi epi) {
  const int nvec = N >> 3;
  const long gid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long>(M) * nvec) return;
  const int v = static_cast<int>(gid % nvec);
  const int m = static_cast<int>(gid / nvec);
  float x[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) x[e] = 0.f;
  for (int ks = 0; ks < ksplit; ++ks) {
    const float4* p = reinterpret_cast<const float4*>(ws + (static_cast<size_t>(ks) * mp + m) * N + v * 8);
    const float4 a = p[0], b = p[1];
    x[0] += a.x; x[1] += a.y; x[2] += a.z; x[3] += a.w; x[4] += b.x; x[5] += b.y; x[6] += b.z; x[7] += b.w;
  }
  auto add8 = [&](const bf16* src) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(src));
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h2[e]); x[2 * e] += f.x; x[2 * e + 1] += f.y; }
  };
  if (epi.bias) add8(epi.bias + v * 8);
  if (epi.rowbias) add8(epi.rowbias + static_cast<size_t>(m / epi.rows_per_group) * N + v * 8);
  if (epi.act) {
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = epi_act(__bfloat162float(__float2bfloat16_rn(x[e])), epi.act);
  }
  if (epi.residual) {
    const uint4 q = *reinterpret_cast<const uint4*>(epi.residual + static_cast<size_t>(m) * epi.ldr + v * 8);
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(h2[e]);
      x[2 * e] = __bfloat162float(__float2bfloat16_rn(x[2 * e])) + f.x;
      x[2 * e + 1] = __bfloat162float(__float2bfloat16_rn(x[2 * e + 1])) + f.y;
    }
  }
  uint4 o;
  o.x = pack_bf16(x[0], x[1]); o.y = pack_bf16(x[2], x[3]); o.z = pack_bf16(x[4], x[5]); o.w = pack_bf16(x[6], x[7]);
  *reinterpret_cast<uint4*>(out + static_cast<size_t>(m) * ldc + v * 8) = o;
}
static int launch_splitk_reduce(const void* ws, void* out, long ldc, int M, int N, const SplitPlan& p, const GemmEpi& epi, cudaStream_t st) {
  const long total = static_cast<long>(M) * (N / 8);
  splitk_reduce_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>((const float*)ws, (bf16*)out, ldc, M, N, p.mp, p.ksplit, epi);
  return cudaGetLastError() == cudaSuccess ? 0 : DLLM_ERR_LAUNCH;
}

// bf16-output GEMM (fwd NT or dgrad NN) with the fused-epilogue operands, run as split-K when the shape calls for it and a workspace of
// gemm_splitk_workspace(M, N, K) bytes is given; otherwise identical to gemm_bf16_ex.
int gemm_bf16_ws(const void* A, const void* B, void* C, int M, int N, int K, long lda, long ldb, long ldc, int b_mn, const void* bias,
                 const void* residual, long ldr, int act, void* ws, size_t ws_bytes, cudaStream_t stream) {
  const SplitPlan p = plan_split(M, N, K, M > BM ? 2 : 1);
  if (p.ksplit <= 1 || !ws || ws_bytes < p.ws_bytes || (N % 8) || (reinterpret_cast<uintptr_t>(ws) & 15) || (ldc % 8))
    return gemm_bf16_ex(A, B, C, M, N, K, lda, ldb, ldc, 0, b_mn, 0, -1, bias, residual, ldr, act, stream);
  if (residual && ((reinterpret_cast<uintptr_t>(residual) & 15) || ((ldr * 2) & 15))) return DLLM_ERR_ALIGN;
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return DLLM_ERR_ALIGN;
  CUtensorMap ta, tb, tc;
  int rc;
  if ((rc = make_tmap_2d(&ta, A, 2, M, K, lda, BM, BK))) return rc;
  if (!b_mn) rc = make_tmap_2d(&tb, B, 2, N, K, ldb, p.bn / p.kcta, BK);
  else rc = make_tmap_2d(&tb, B, 2, K, N, ldb, BK, 64);
  if (rc) return rc;
  if ((rc = make_tmap_2d(&tc, ws, 4, static_cast<uint64_t>(p.ksplit) * p.mp, N, N, 32, 32))) return rc;
  GemmEpi none{nullptr, nullptr, 0, 0, nullptr, 1};
  const ConvGeom nocg{1, 1, 1, 1};
#define DLLM_SPLIT_CASE(CTA, BMN_, BNW)                                                                                      \
  if (p.kcta == CTA && (b_mn != 0) == BMN_ && p.bn == BNW)                                                                   \
    rc = launch_gemm_bn<CTA, false, BMN_, float, false, 0, 4, BNW>(ta, tb, tc, M, N, K, none, stream, nocg, p.ksplit);
  rc = DLLM_ERR_SHAPE;
  DLLM_SPLIT_CASE(1, false, 256) DLLM_SPLIT_CASE(1, true, 256) DLLM_SPLIT_CASE(2, false, 256) DLLM_SPLIT_CASE(2, true, 256)
  DLLM_SPLIT_CASE(1, false, 128) DLLM_SPLIT_CASE(1, true, 128) DLLM_SPLIT_CASE(2, false, 128) DLLM_SPLIT_CASE(2, true, 128)
#undef DLLM_SPLIT_CASE
  if (rc) return rc;
  GemmEpi epi{static_cast<const bf16*>(bias), static_cast<const bf16*>(residual), ldr, act, nullptr, 1};
  return launch_splitk_reduce(ws, C, ldc, M, N, p, epi, stream);
}

int gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, long lda, long ldb, long ldc, int a_mn,
              int b_mn, int out_fp32, int cta_pair, cudaStream_t stream) {
  return gemm_bf16_ex(A, B, C, M, N, K, lda, ldb, ldc, a_mn, b_mn, out_fp32, cta_pair, nullptr, nullptr, 0, 0, stream);
}

int gemm_bf16_ex(const void* A, const void* B, void* C, int M, int N, int K, long lda, long ldb, long ldc, int a_mn,
                 int b_mn, int out_fp32, int cta_pair, const void* bias, const void* residual, long ldr, int act,
                 cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || act < 0 || act > 3) return DLLM_ERR_SHAPE;
  if ((bias || residual) && (N % 8)) return DLLM_ERR_SHAPE;
  if (residual && ((reinterpret_cast<uintptr_t>(residual) & 15) || ((ldr * 2) & 15))) return DLLM_ERR_ALIGN;
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return DLLM_ERR_ALIGN;
  GemmEpi epi{static_cast<const bf16*>(bias), static_cast<const bf16*>(residual), ldr, act, nullptr, 1};
  const int kcta = (cta_pair < 0) ? (M > BM ? 2 : 1) : (cta_pair ? 2 : 1);
  const int bn = pick_bn(N);
  const ConvGeom nocg{1, 1, 1, 1};
  CUtensorMap ta, tb, tc;
  int rc;
  // A
  if (!a_mn) rc = make_tmap_2d(&ta, A, 2, M, K, lda, BM, BK);
  else rc = make_tmap_2d(&ta, A, 2, K, M, lda, BK, 64);
  if (rc) return rc;
  // B
  if (!b_mn) rc = make_tmap_2d(&tb, B, 2, N, K, ldb, bn / kcta, BK);
  else rc = make_tmap_2d(&tb, B, 2, K, N, ldb, BK, 64);
  if (rc) return rc;
  // C: per-warp store boxes of 32 rows x 128 bytes
  rc = make_tmap_2d(&tc, C, out_fp32 ? 4 : 2, M, N, ldc, 32, out_fp32 ? 32 : 64);
  if (rc) return rc;

  const int emode = (act != 0) ? 2 : ((bias || residual) ? 1 : 0);
  if (emode != 0) {
    // fused epilogues exist for the forward (NT, bf16 out) contraction only
    if (a_mn || b_mn || out_fp32) return DLLM_ERR_UNSUPPORTED;
    if (kcta == 2) return emode == 2 ? launch_gemm<2, false, false, bf16, false, 2, 8>(ta, tb, tc, M, N, K, epi, stream, nocg, bn)
                                     : launch_gemm<2, false, false, bf16, false, 1, 8>(ta, tb, tc, M, N, K, epi, stream, nocg, bn);
    return emode == 2 ? launch_gemm<1, false, false, bf16, false, 2, 8>(ta, tb, tc, M, N, K, epi, stream, nocg, bn)
                      : launch_gemm<1, false, false, bf16, false, 1, 8>(ta, tb, tc, M, N, K, epi, stream, nocg, bn);
  }
  if (!a_mn && !b_mn && !out_fp32 && K <= 2048) {   // epilogue-bound plain forward GEMMs (UNet / CLIP projections)
    if (kcta == 2) return launch_gemm<2, false, false, bf16, false, 0, 8>(ta, tb, tc, M, N, K, epi, stream, nocg, bn);
    return launch_gemm<1, false, false, bf16, false, 0, 8>(ta, tb, tc, M, N, K, epi, stream, nocg, bn);
  }
#define DLLM_GEMM_CASE(CTA, AMN, BMN)                                                                 \
  if (kcta == CTA && (a_mn != 0) == AMN && (b_mn != 0) == BMN) {                                      \
    return out_fp32 ? launch_gemm<CTA, AMN, BMN, float>(ta, tb, tc, M, N, K, epi, stream, nocg, bn)   \
                    : launch_gemm<CTA, AMN, BMN, bf16>(ta, tb, tc, M, N, K, epi, stream, nocg, bn);   \
  }
  DLLM_GEMM_CASE(1, false, false)
  DLLM_GEMM_CASE(1, false, true)
  DLLM_GEMM_CASE(1, true, true)
  DLLM_GEMM_CASE(1, true, false)
  DLLM_GEMM_CASE(2, false, false)
  DLLM_GEMM_CASE(2, false, true)
  DLLM_GEMM_CASE(2, true, true)
  DLLM_GEMM_CASE(2, true, false)
#undef DLLM_GEMM_CASE
  return DLLM_ERR_SHAPE;
}

// out[M, N/2] = GEGLU(A @ Wp^T + bias_p): Wp / bias_p are the FeedForward's `proj` weight / bias with rows permuted so that every
// 128-row group is [64 value rows | the 64 matching gate rows] (dreamllm_b200/unet.py builds them once per frozen UNet).
int gemm_bf16_geglu(const void* A, const void* Wp, const void* bias_p, void* out, int M, int N, int K, long lda, long ldb, long ldc,
                    cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (N % 128)) return DLLM_ERR_SHAPE;
  if (bias_p && (reinterpret_cast<uintptr_t>(bias_p) & 15)) return DLLM_ERR_ALIGN;
  GemmEpi epi{static_cast<const bf16*>(bias_p), nullptr, 0, 2, nullptr, 1};
  const int kcta = (M > BM) ? 2 : 1;
  CUtensorMap ta, tb, tc;
  int rc;
  if ((rc = make_tmap_2d(&ta, A, 2, M, K, lda, BM, BK))) return rc;
  if ((rc = make_tmap_2d(&tb, Wp, 2, N, K, ldb, BN / kcta, BK))) return rc;
  if ((rc = make_tmap_2d(&tc, out, 2, M, N / 2, ldc, 32, 64))) return rc;
  if (kcta == 2) return launch_gemm_bn<2, false, false, bf16, false, 3, 8, 256>(ta, tb, tc, M, N, K, epi, stream);
  return launch_gemm_bn<1, false, false, bf16, false, 3, 8, 256>(ta, tb, tc, M, N, K, epi, stream);
}

// 4-D NHWC activation map: dims {C, W, H, N}, box {64, Wb, Hb, Nb}, 128B swizzle.
static int make_tmap_nhwc(CUtensorMap* map, const void* ptr, int Nimg, int H, int W, int C, int Wb, int Hb, int Nb) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return DLLM_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (C % 8)) return DLLM_ERR_ALIGN;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nimg};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)Wb, (cuuint32_t)Hb, (cuuint32_t)Nb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fprintf(stderr, "dllm: cuTensorMapEncodeTiled(4D) failed (CUresult %d)\n", (int)r);
  return r == CUDA_SUCCESS ? 0 : DLLM_ERR_TMAP;
}

size_t conv3x3_splitk_workspace(int Nimg, int H, int W, int Cin, int Cout) {
  const int M = Nimg * H * W;
  return plan_split(M, Cout, 9 * Cin, M > BM ? 2 : 1).ws_bytes;
}
int conv3x3_nhwc(const void* x, const void* w, void* y, int Nimg, int H, int W, int Cin, int Cout, const void* bias,
                 const void* rowbias, const void* residual, cudaStream_t stream) {
  return conv3x3_nhwc_ws(x, w, y, Nimg, H, W, Cin, Cout, bias, rowbias, residual, nullptr, 0, stream);
}
// y[n,h,w,:] = conv3x3(x, w) + bias + rowbias[n] (+ residual);  x NHWC [Nimg,H,W,Cin], w [Cout, 3,3,Cin] (K-major), y NHWC.
// With a workspace of conv3x3_splitk_workspace() bytes small planes run split-K (see plan_split).
int conv3x3_nhwc_ws(const void* x, const void* w, void* y, int Nimg, int H, int W, int Cin, int Cout, const void* bias,
                    const void* rowbias, const void* residual, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (Nimg <= 0 || Cin % 64 || Cout % 8) return DLLM_ERR_SHAPE;
  if (W > 128 ? (W % 128) : (128 % W)) return DLLM_ERR_SHAPE;
  const int HW = H * W;
  int Wb = W > 128 ? 128 : W, Hb = W > 128 ? 1 : 128 / W, Nb = 1;
  if (Hb > H) {  // small planes: the 128-pixel tile spans several whole images
    if (Hb % H) return DLLM_ERR_SHAPE;
    Nb = Hb / H;
    Hb = H;
  } else if (H % Hb) {
    return DLLM_ERR_SHAPE;
  }
  const int M = Nimg * HW, K = 9 * Cin;
  const int kcta = (M > BM) ? 2 : 1;
  const int bn = pick_bn(Cout);
  CUtensorMap ta, tb, tc;
  int rc;
  if ((rc = make_tmap_nhwc(&ta, x, Nimg, H, W, Cin, Wb, Hb, Nb))) return rc;
  if ((rc = make_tmap_2d(&tb, w, 2, Cout, K, K, bn / kcta, BK))) return rc;
  GemmEpi epi{static_cast<const bf16*>(bias), static_cast<const bf16*>(residual), Cout, 0, static_cast<const bf16*>(rowbias), HW};
  ConvGeom cg{Cin / 64, W, H, HW};
  const SplitPlan sp = plan_split(M, Cout, K, kcta);
  if (sp.ksplit > 1 && ws && ws_bytes >= sp.ws_bytes && !(reinterpret_cast<uintptr_t>(ws) & 15)) {
    if ((rc = make_tmap_2d(&tc, ws, 4, static_cast<uint64_t>(sp.ksplit) * sp.mp, Cout, Cout, 32, 32))) return rc;
    GemmEpi none{nullptr, nullptr, 0, 0, nullptr, 1};
    if (kcta == 2) rc = (bn == 128) ? launch_gemm_bn<2, false, false, float, true, 0, 4, 128>(ta, tb, tc, M, Cout, K, none, stream, cg, sp.ksplit)
                                    : launch_gemm_bn<2, false, false, float, true, 0, 4, 256>(ta, tb, tc, M, Cout, K, none, stream, cg, sp.ksplit);
    else rc = (bn == 128) ? launch_gemm_bn<1, false, false, float, true, 0, 4, 128>(ta, tb, tc, M, Cout, K, none, stream, cg, sp.ksplit)
                          : launch_gemm_bn<1, false, false, float, true, 0, 4, 256>(ta, tb, tc, M, Cout, K, none, stream, cg, sp.ksplit);
    if (rc) return rc;
    return launch_splitk_reduce(ws, y, Cout, M, Cout, sp, epi, stream);
  }
  if ((rc = make_tmap_2d(&tc, y, 2, M, Cout, Cout, 32, 64))) return rc;
  const bool any_epi = bias || rowbias || residual;
  if (kcta == 2) return any_epi ? launch_gemm<2, false, false, bf16, true, 1, 8>(ta, tb, tc, M, Cout, K, epi, stream, cg, bn)
                                : launch_gemm<2, false, false, bf16, true, 0, 4>(ta, tb, tc, M, Cout, K, epi, stream, cg, bn);
  return any_epi ? launch_gemm<1, false, false, bf16, true, 1, 8>(ta, tb, tc, M, Cout, K, epi, stream, cg, bn)
                 : launch_gemm<1, false, false, bf16, true, 0, 4>(ta, tb, tc, M, Cout, K, epi, stream, cg, bn);
}

}  // namespace dllm

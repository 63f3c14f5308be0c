// tcgen05 flash attention for sm_100a: forward, and backward split in two deterministic kernels (dK/dV, dQ).
//
// Replaces flash_attn_func / flash_attn_varlen_func (reference modeling_dreamllm.py:500-551) and the eager
// softmax(QK^T/sqrt(d) + mask)V path (:357-379).  q/k/v are read straight out of the fused qkv projection buffer
// ([B, S, heads, d] views with a shared token stride) through 3-D TMA maps — no transposes, no .contiguous().
//
// Common structure of the three kernels (192 threads):
//   warps 0-3  "row" warps: thread t of warp w owns TMEM lane 32w+t = one row of the 128-row MMA tile
//              (softmax / dS math, P tile written to swizzled smem as the next MMA's A operand, epilogue)
//   warp 4     TMA producer (one lane)
//   warp 5     TMEM allocator + MMA issuer (one lane): tcgen05.mma, S/dP tiles double-buffered in TMEM
// All inter-role hand-offs are mbarriers; tcgen05.commit signals MMA completion.
#include <stdlib.h>

#include "common.cuh"
#include "gemm_sm100.h"
#include "ops.h"

namespace dllm {

constexpr int kAttnThreads = 192;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

typedef CUresult (*PFN_encodeTiled3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// [B, S, cols] bf16 view: token stride ld (elements), batch stride S*ld; box = [1, box_rows, 64 cols], 128B swizzle
static int make_tmap_bsc(CUtensorMap* map, const void* ptr, int B, int S, int cols, long ld, int box_rows) {
  static PFN_encodeTiled3 enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
      return DLLM_ERR_DRIVER;
    enc = reinterpret_cast<PFN_encodeTiled3>(p);
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15)) return DLLM_ERR_ALIGN;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(S), static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(S) * ld * 2};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : DLLM_ERR_TMAP;
}

// Descriptor of UMMA_K step `ks` of the tile whose step-0 descriptor is d0.  The start-address field counts 16-byte units and no tile
// crosses the 256 KB its 14 bits span, so a plain add is exact; with the loops below fully unrolled every step's descriptor is the
// tile's plus a compile-time constant.
__device__ __forceinline__ uint64_t desc_step(uint64_t d0, bool mn, uint32_t chunk_stride, int ks) {
  return d0 + ((mn ? ks * 2048u : (ks >> 2) * chunk_stride + (ks & 3) * 32u) >> 4);
}
// one full MMA tile: D[tmem] (+)= A * B over kSteps UMMA_K steps.
// Issue cost matters here: the S / dP tile-GEMMs are N = 64 instructions (32-48 clk of tensor pipe each), and the r01 / early-r02 issue
// loop — descriptor rebuilt per step, and ptxas wrapping each UTCHMMA in a per-active-thread ELECT/BRA loop because `lane == 0` does not
// tell it the region is single-threaded — spent 13 dependent scalar instructions (~100 clk) per MMA: the issuing thread, not the tensor
// pipe or the softmax warps, set the iteration time (profiles/r02j_attn_fwd_timeline.md).  The issuer is now chosen with elect.sync and
// the descriptors are hoisted, which lets ptxas emit the kSteps UTCHMMAs back to back.
template <int kSteps>
__device__ __forceinline__ void mma_tile(uint32_t tmem_d, uint32_t a_base, bool a_mn, uint32_t a_cs, uint32_t b_base,
                                         bool b_mn, uint32_t b_cs, uint32_t idesc, bool accumulate) {
  const uint64_t a0 = op_desc(a_base, a_mn, a_cs, 0), b0 = op_desc(b_base, b_mn, b_cs, 0);
#pragma unroll
  for (int ks = 0; ks < kSteps; ++ks)
    umma_ss<1>(tmem_d, desc_step(a0, a_mn, a_cs, ks), desc_step(b0, b_mn, b_cs, ks), idesc, (accumulate || ks > 0) ? 1u : 0u);
}
// same with the A operand in tensor memory: step ks reads the 32-bit columns starting at tmem_a + ks * a_col_stride
template <int kSteps>
__device__ __forceinline__ void mma_tile_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t a_col_stride, uint32_t b_base, bool b_mn,
                                            uint32_t b_cs, uint32_t idesc, bool accumulate) {
  const uint64_t b0 = op_desc(b_base, b_mn, b_cs, 0);
#pragma unroll
  for (int ks = 0; ks < kSteps; ++ks)
    umma_ts(tmem_d, tmem_a + ks * a_col_stride, desc_step(b0, b_mn, b_cs, ks), idesc, (accumulate || ks > 0) ? 1u : 0u);
}

// write 64 bf16 (one 128-byte swizzled line) for tile row `row`; vals come 8 at a time
__device__ __forceinline__ void store_row_chunk(uint8_t* tile, int row, int j, const float* f8) {
  uint4 q;
  q.x = pack_bf16(f8[0], f8[1]);
  q.y = pack_bf16(f8[2], f8[3]);
  q.z = pack_bf16(f8[4], f8[5]);
  q.w = pack_bf16(f8[6], f8[7]);
  *reinterpret_cast<uint4*>(tile + row * 128 + ((j ^ (row & 7)) << 4)) = q;
}

// TMEM accumulator (128 lanes x D fp32 cols) -> *scale -> bf16 -> swizzled staging [D/64][128 rows][128 B] -> per-warp
// TMA store of [32 rows x 64 cols] boxes at (col0 + c*64, row0 + 32*warp, b).
// Only 64-column chunks [ch_begin, ch_end) are handled (lets two warpgroups split the columns).
template <int D>
__device__ __forceinline__ void store_acc_tile(uint32_t tmem_acc, uint8_t* stage, float mul, const CUtensorMap* tm,
                                               int col0, int row0, int b, int warp, int lane, int ch_begin = 0,
                                               int ch_end = D / 64) {
  const int row = warp * 32 + lane;
#pragma unroll 1
  for (int c = 2 * ch_begin; c < 2 * ch_end; ++c) {
    uint32_t v[32];
    tmem_ld32(tmem_acc + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    uint8_t* tile = stage + (c >> 1) * 16384;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]) * mul;
      store_row_chunk(tile, row, (c & 1) * 4 + j, f);
    }
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    for (int c = ch_begin; c < ch_end; ++c) tma_store_3d(tm, stage + c * 16384 + warp * 4096, col0 + c * 64, row0 + warp * 32, b);
    tma_store_commit();
    tma_store_wait_all<0>();
  }
}

// DLLM_ATTN_TRACE (dev builds only, scripts/attn_trace.py): clock64 timestamps of one CTA's softmax warp 0 / MMA thread / TMA thread per KV
// tile, to see which hand-off the iteration time is made of.  Slot layout: [tile j][16].
#ifdef DLLM_ATTN_TRACE
__device__ long long g_attn_trace[64 * 16 + 16];
__device__ long long g_attn_cta_log[8192 * 4];   // per forward CTA: {clock64 at start, at end, smid, n_kv}
#define ATTN_TRACE(j, k) do { if (trace_on && (j) < 64) g_attn_trace[(j) * 16 + (k)] = clock64(); } while (0)
#else
#define ATTN_TRACE(j, k) do { } while (0)
#endif

// ================================================================================================ forward
// kPT = true (default): P never touches shared memory.  The softmax warps write bf16 P back into the TMEM columns S occupied
// (tcgen05.st) and the PV tile-GEMM reads its A operand from tensor memory (tcgen05.mma with [a_tmem]).  Because S is double-buffered, so
// is P: softmax(j+1) no longer waits for PV(j), and the per-tile STS burst + fence.proxy.async + 16 KB of smem are gone (round 1 kept P in
// a single smem buffer and every softmax warp stalled on `pv_done` once per KV tile: 733 TF/s, tensor pipe 31 %).
// kPT = false: the round-1 data path (P staged in swizzled smem), kept selectable with DLLM_ATTN_LEGACY=1 for same-box A/B runs.
template <int D, bool kPT>
struct FwdSmem {
  static constexpr int NCH = D / 64;
  static constexpr int kQ = NCH * 16384;       // [NCH][128][128B]
  static constexpr int kKV = NCH * 8192;       // [NCH][64][128B]
  static constexpr int kP = kPT ? 0 : 16384;   // [128][128B]
  // K ring: the refill of a K stage can only start when the QK^T that read it has retired, and with two stages the next-but-one QK^T
  // then waited out the L2 round trip; the 16 KB P no longer needs pays for a third K stage (2 CTAs / SM: 2 x (112 KB + 256 B) fits 227 KB)
  static constexpr int kKStages = kPT ? 3 : 2;
  static constexpr int oQ = 0, oK = kQ, oV = oK + kKStages * kKV, oP = oV + 2 * kKV, oBar = oP + kP;
  static constexpr int kBytes = oBar + 256;
};

template <int D, bool kCausal, bool kPT>
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap to, float* __restrict__ lse,
                const int* __restrict__ seqlens, int S, int Skv, int nh, float scale_log2,
                const uint8_t* __restrict__ kv_mask, int mask_ld) {
  // kv_mask (nullable): [B, mask_ld] bytes, 0 = padded key position that no query may attend to — the 2-D `attention_mask` of a padded
  // batch living inside a kv-cache (reference :553-583 drops those tokens with `_upad_input`; HF left-pads decoder-only prompts)
  // causal with Skv > S = kv-cache decode / continuation (reference :344-355, :444-449): query i sits at absolute position
  // (Skv - S) + i, i.e. the mask is bottom-right aligned:  kv <= q + coff
  const int coff = kCausal ? (Skv - S) : 0;
  using L = FwdSmem<D, kPT>;
  constexpr int NCH = L::NCH;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::oBar);
  constexpr int KS = L::kKStages;
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;        // [KS]  ring: j % KS, phase (j / KS) & 1
  uint64_t* k_empty = k_full + KS;    // [KS]
  uint64_t* v_full = k_empty + KS;    // [2]
  uint64_t* v_empty = v_full + 2;     // [2]
  uint64_t* s_full = v_empty + 2;     // [2]
  uint64_t* p_full = s_full + 2;      // [2]  (kPT: P(j) lives in S buffer j & 1; legacy: only [0] is used)
  uint64_t* pv_done = p_full + 2;     // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // causal: q tile i has i + 1 KV tiles of work; launch the heaviest first so the grid's tail is made of short CTAs (LPT order)
  const int q0 = (kCausal ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x) * 128, h = blockIdx.y, b = blockIdx.z;
  const int kv_len = seqlens ? min(seqlens[b], Skv) : Skv;   // Skv != S only for cross-attention (non-causal)
  const int q_len = seqlens ? kv_len : S;                    // right-padded self-attention: rows >= len are padding
  const int kv_end = kCausal ? min(kv_len, q0 + 128 + coff) : kv_len;
  const int n_kv = (kv_end + 63) / 64;
#ifdef DLLM_ATTN_TRACE
  const long long cta_t0 = clock64();
  const bool trace_on = (q0 == (gridDim.x - 1) * 128 && blockIdx.y == 1 && blockIdx.z == 0 && lane == 0);
  if (trace_on && warp == 0) g_attn_trace[64 * 16] = clock64();
#endif

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) { printf("attn_fwd: smem misaligned\n"); __trap(); }
    mbar_init(q_full, 1);
    for (int i = 0; i < KS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], kPT ? 4 : 128);      // kPT: one elected arrive per softmax warp (128 same-address arrives serialise)
      mbar_init(&pv_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc<1>(tmem_ptr, 256); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base;        // 2 x 64 columns
  const uint32_t tmem_O = tmem_base + 128;  // D columns

  // producer / issuer: one lane picked with elect.sync — unlike `lane == 0` it lets ptxas treat the region as single-threaded, so the
  // TMA / tcgen05.mma instructions are not wrapped in a per-active-thread loop
  if (warp == 4) { if (elect_one_sync() && n_kv > 0) {
    // ---------------- TMA producer ----------------
    tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv);
    mbar_arrive_expect_tx(q_full, 128 * D * 2);
    for (int c = 0; c < NCH; ++c) tma_load_3d(smem + L::oQ + c * 16384, &tq, q_full, h * D + c * 64, q0, b);
    // K runs ahead of V: all K tiles the ring can hold are requested before the producer blocks on the V stage of tile j
    int jk = 0;
    auto load_k = [&](int jj) {
      const int ks = jj % KS;
      mbar_wait(&k_empty[ks], ((jj / KS) & 1) ^ 1, 10);
      mbar_arrive_expect_tx(&k_full[ks], 64 * D * 2);
      for (int c = 0; c < NCH; ++c) tma_load_3d(smem + L::oK + ks * L::kKV + c * 8192, &tk, &k_full[ks], h * D + c * 64, jj * 64, b);
    };
    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      while (jk < n_kv && jk < j + KS - 1) load_k(jk++);     // K(j) .. K(j + KS - 2) need no stage that V(j)'s consumer frees
      if (jk <= j) load_k(jk++);
      ATTN_TRACE(j, 8);
      mbar_wait(&v_empty[st], ph ^ 1, 11);
      ATTN_TRACE(j, 9);
      mbar_arrive_expect_tx(&v_full[st], 64 * D * 2);
      for (int c = 0; c < NCH; ++c) tma_load_3d(smem + L::oV + st * L::kKV + c * 8192, &tv, &v_full[st], h * D + c * 64, j * 64, b);
    }
  } } else if (warp == 5) { if (elect_one_sync() && n_kv > 0) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, 64, false, false);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, D, false, true);
    const uint32_t sQ = smem_u32(smem + L::oQ), sK = smem_u32(smem + L::oK), sV = smem_u32(smem + L::oV),
                   sP = smem_u32(smem + L::oP);
    (void)sP;
    auto issue_qk = [&](int j) {
      const int ks = j % KS;
      mbar_wait(&k_full[ks], (j / KS) & 1, 12);
      ATTN_TRACE(j, 5);
      tc_fence_after();
      mma_tile<D / 16>(tmem_S + (j & 1) * 64, sQ, false, 16384, sK + ks * L::kKV, false, 8192, idesc_qk, false);
      umma_commit(&k_empty[ks]);
      umma_commit(&s_full[j & 1]);
    };
    mbar_wait(q_full, 0, 13);
    issue_qk(0);
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) issue_qk(j + 1);
      const int st = j & 1;
      mbar_wait(&v_full[st], (j >> 1) & 1, 14);
      ATTN_TRACE(j, 6);
      if constexpr (kPT) {
        mbar_wait(&p_full[j & 1], (j >> 1) & 1, 15);
        ATTN_TRACE(j, 7);
        tc_fence_after();
        // O += P V_j with P read from TMEM: [128 lanes x 64 kv] bf16 = 32 columns at the base of S buffer j & 1, 8 columns per UMMA_K
        mma_tile_ts<4>(tmem_O, tmem_S + (j & 1) * 64, 8, sV + st * L::kKV, true, 8192, idesc_pv, j > 0);
      } else {
        mbar_wait(&p_full[0], j & 1, 15);
        tc_fence_after();
        mma_tile<4>(tmem_O, sP, false, 0, sV + st * L::kKV, true, 8192, idesc_pv, j > 0);
      }
      umma_commit(&v_empty[st]);
      umma_commit(&pv_done[kPT ? (j & 1) : 0]);
    }
  } } else if (warp < 4) {
    // ---------------- softmax rows ----------------
    const int row = warp * 32 + lane;
    const int q_row = q0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      ATTN_TRACE(j, 10);
      mbar_wait(&s_full[j & 1], (j >> 1) & 1, 16);
      if (warp == 0) ATTN_TRACE(j, 0);
      tc_fence_after();
      uint32_t sv[64];
      tmem_ld32(tmem_S + lane_off + (j & 1) * 64, sv);
      tmem_ld32(tmem_S + lane_off + (j & 1) * 64 + 32, sv + 32);
      tmem_ld_wait();
      if (warp == 0) ATTN_TRACE(j, 1);
      const int kv0 = j * 64;
      const bool need_mask = (kCausal && kv0 + 63 > q0 + warp * 32 + coff) || (kv0 + 64 > kv_len) || (kv_mask != nullptr);
      float mx = -INFINITY;
      if (need_mask) {
        uint32_t mw[16];                                      // this tile's 64 key-mask bytes (all-ones when there is no mask)
        if (kv_mask != nullptr) {
          const uint4* mp = reinterpret_cast<const uint4*>(kv_mask + static_cast<size_t>(b) * mask_ld + kv0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 t = __ldg(mp + i);
            mw[4 * i] = t.x; mw[4 * i + 1] = t.y; mw[4 * i + 2] = t.z; mw[4 * i + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) mw[i] = 0x01010101u;
        }
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          const int kvi = kv0 + c;
          const bool ok = (kvi < kv_len) && (!kCausal || kvi <= q_row + coff) && (((mw[c >> 2] >> (8 * (c & 3))) & 0xffu) != 0u);
          float s = ok ? __uint_as_float(sv[c]) : -INFINITY;
          sv[c] = __float_as_uint(s);
          mx = fmaxf(mx, s);
        }
      } else {
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;   // four independent chains (was one 32-deep FMNMX3 chain)
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
          m0 = fmaxf(m0, __uint_as_float(sv[c]));
          m1 = fmaxf(m1, __uint_as_float(sv[c + 1]));
          m2 = fmaxf(m2, __uint_as_float(sv[c + 2]));
          m3 = fmaxf(m3, __uint_as_float(sv[c + 3]));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      const float m_new = fmaxf(m_run, mx * scale_log2);
      // lazy rescale: only move the reference max when it grew by more than 2^8 (keeps P <= 256, exact after 1/l)
      const bool need = (m_new - m_run) > 8.0f;
      const bool any = __any_sync(0xffffffffu, need);
      if (warp == 0) ATTN_TRACE(j, 2);
      if constexpr (!kPT) {
        if (j > 0) mbar_wait(&pv_done[0], (j - 1) & 1, 17);  // P buffer free, O quiescent
      } else {
        // P(j) goes to TMEM buffer j & 1, which PV(j-2) finished reading before QK(j) could overwrite it with S(j) (in-order tensor pipe):
        // nothing to wait for unless O itself is about to be rescaled
        if (any && j > 0) mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1, 17);
      }
      if (any) {
        const float m_tgt = (m_new == -INFINITY) ? m_run : m_new;
        const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_tgt);
        if (j > 0) {
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < D / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(tmem_O + lane_off + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tmem_st32(tmem_O + lane_off + c * 32, o);
          }
          tmem_st_wait();
        }
        l_run *= alpha;
        m_run = m_tgt;
      }
      const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
      float ls0 = 0.f, ls1 = 0.f;
      if constexpr (kPT) {
        uint32_t pw[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float p0 = exp2f(__uint_as_float(sv[2 * c]) * scale_log2 - m_use);
          const float p1 = exp2f(__uint_as_float(sv[2 * c + 1]) * scale_log2 - m_use);
          ls0 += p0;
          ls1 += p1;
          pw[c] = pack_bf16(p0, p1);          // K elements 2c, 2c+1 of this row share one 32-bit TMEM column
        }
        if (warp == 0) ATTN_TRACE(j, 3);
        tmem_st32(tmem_S + lane_off + (j & 1) * 64, pw);
        tmem_st_wait();
        l_run += ls0 + ls1;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[j & 1]);
        if (warp == 0) ATTN_TRACE(j, 4);
      } else {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          float p[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            p[e] = exp2f(__uint_as_float(sv[jj * 8 + e]) * scale_log2 - m_use);
            ls0 += p[e];
          }
          store_row_chunk(smem + L::oP, row, jj, p);
        }
        l_run += ls0;
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&p_full[0]);
      }
    }
    if (n_kv > 0) {
      if constexpr (kPT) mbar_wait(&pv_done[(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1, 18);
      else mbar_wait(&pv_done[0], (n_kv - 1) & 1, 18);
      tc_fence_after();
    }
#ifdef DLLM_ATTN_TRACE
    if (trace_on && warp == 0) g_attn_trace[64 * 16 + 1] = clock64();
#endif
    const bool valid_row = (q_row < q_len) && n_kv > 0 && l_run > 0.f;
    const float inv = valid_row ? 1.f / l_run : 0.f;
    if (n_kv > 0) {
      store_acc_tile<D>(tmem_O, smem + L::oQ, inv, &to, h * D, q0, b, warp, lane);
    } else {
      // no keys at all: write zeros
      for (int c = 0; c < D / 64; ++c)
        for (int jj = 0; jj < 8; ++jj) {
          float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          store_row_chunk(smem + L::oQ + c * 16384, row, jj, z);
        }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        for (int c = 0; c < D / 64; ++c) tma_store_3d(&to, smem + L::oQ + c * 16384 + warp * 4096, h * D + c * 64, q0 + warp * 32, b);
        tma_store_commit();
        tma_store_wait_all<0>();
      }
    }
    if (q_row < S)
      lse[(static_cast<size_t>(b) * nh + h) * S + q_row] = valid_row ? (m_run + log2f(l_run)) * kLn2 : INFINITY;
  }
#ifdef DLLM_ATTN_TRACE
  if (trace_on && warp == 0) g_attn_trace[64 * 16 + 2] = clock64();
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<1>(tmem_base, 256);
#ifdef DLLM_ATTN_TRACE
  if (threadIdx.x == 0) {
    const int lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    if (lin < 8192) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      g_attn_cta_log[lin * 4 + 0] = cta_t0;
      g_attn_cta_log[lin * 4 + 1] = clock64();
      g_attn_cta_log[lin * 4 + 2] = smid;
      g_attn_cta_log[lin * 4 + 3] = n_kv;
    }
  }
#endif
}

// ================================================================================================ forward, persistent
// Same data path as attn_fwd_kernel<D, kCausal, true> (S / P in TMEM, TS-form PV, 3-stage K ring), but the grid is 2 CTAs per SM that
// walk the (q tile, head, batch) items themselves.  Measured on the C2 shape with one CTA per item (profiles/r02n_attn_fwd_cta_log.json):
// a CTA lives 9.7 k clk + 1.53 k clk per KV tile (17 tiles on average), and its SM slot then stays EMPTY for ~4 k clk until the next CTA
// starts — 35 % of every slot went to launch gaps, TMEM alloc, barrier init, the exposed Q/K load latency and the O store.  Here a slot is
// set up once; the producer starts the next item's Q / K loads as soon as the last QK^T of the current item retires (about two KV tiles
// before its softmax / PV / O store finish), and the rings and barrier phases simply keep counting across items.
//   * items come from a global atomic counter (the producer thread draws them and publishes them to the other roles through a 2-slot smem
//     ring).  Their order is windows of ~2 items per CTA: all q tiles of W consecutive (head, batch) pairs, heaviest q tiles first.
//     Heaviest-first balances the causal triangle; the window keeps the K / V of the heads in flight inside L2 — a plain heaviest-first
//     order over ALL heads made every K / V tile an HBM read (2.2 GB per call, 6.1 TB/s: 355 us at B = 8 against 309 us non-persistent,
//     while at B = 2-3, where K / V fit L2 anyway, the same kernel was already 12 % faster: profiles/r02r_attn_fwd_persist.md);
//   * O is staged for its TMA store in the V ring (the Q buffer, which the one-shot kernel used, is already being refilled for the next
//     item): the last PV of the item has retired by then, and the producer holds back the next item's first V loads until the four
//     softmax warps report (o_stored) that their bulk stores have read the staging area;
//   * the next item's first PV overwrites the O accumulator; it waits for P(0) of that item, which the softmax warps produce only after
//     they have drained O — no extra barrier.
template <int D, bool kCausal>
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_fwd_persist_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                        const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap to, bf16* __restrict__ out,
                        long ld_o, float* __restrict__ lse, const int* __restrict__ seqlens, int B, int S, int Skv, int nh,
                        float scale_log2, const uint8_t* __restrict__ kv_mask, int mask_ld, int* __restrict__ item_counter) {
  const int coff = kCausal ? (Skv - S) : 0;
  using L = FwdSmem<D, true>;
  constexpr int NCH = L::NCH;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::oBar);
  constexpr int KS = L::kKStages;
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;       // the item's last QK^T has retired: the Q buffer may be refilled
  uint64_t* k_full = bars + 2;        // [KS]  ring over ALL KV tiles this CTA processes: g % KS, phase (g / KS) & 1
  uint64_t* k_empty = k_full + KS;    // [KS]
  uint64_t* v_full = k_empty + KS;    // [2]
  uint64_t* v_empty = v_full + 2;     // [2]
  uint64_t* s_full = v_empty + 2;     // [2]
  uint64_t* p_full = s_full + 2;      // [2]
  uint64_t* pv_done = p_full + 2;     // [2]
  uint64_t* o_stored = pv_done + 2;   // the item's O staging (V ring) has been read by the bulk stores
  uint64_t* it_full = o_stored + 1;   // [2]  item ring
  uint64_t* it_empty = it_full + 2;   // [2]
  uint32_t* item_ring = reinterpret_cast<uint32_t*>(it_empty + 2);   // [2]
  uint32_t* tmem_ptr = item_ring + 2;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = (S + 127) / 128, n_hb = nh * B, n_items = nq * n_hb;
#ifdef DLLM_ATTN_TRACE
  const bool trace_on = (blockIdx.x == 0 && lane == 0);
  int tr_item = 0;
  if (trace_on && warp == 0) g_attn_trace[64 * 16] = clock64();
#endif

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) { printf("attn_fwd_persist: smem misaligned\n"); __trap(); }
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < KS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
      mbar_init(&it_full[i], 1);
      mbar_init(&it_empty[i], 5);     // MMA thread + one lane of each softmax warp
    }
    mbar_init(o_stored, 4);
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc<1>(tmem_ptr, 256); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base;        // 2 x 64 columns
  const uint32_t tmem_O = tmem_base + 128;  // D columns

  // item w (heaviest first) -> q tile, head, batch, and its KV tile count; identical in every role
  struct Item { int q0, h, b, kv_len, q_len, n_kv; };
  const int win_heads = max(1, min(n_hb, (2 * static_cast<int>(gridDim.x) + nq - 1) / nq));   // ~2 items per CTA per window
  const int full_items = (n_hb / win_heads) * win_heads * nq;
  auto item = [&](int w) {
    Item it;
    const int win = (w < full_items) ? w / (win_heads * nq) : n_hb / win_heads;
    const int idx = (w < full_items) ? w % (win_heads * nq) : w - full_items;
    const int R = (w < full_items) ? win_heads : n_hb % win_heads;
    const int qt = nq - 1 - idx / R, hb = win * win_heads + idx % R;
    it.q0 = qt * 128; it.h = hb % nh; it.b = hb / nh;
    it.kv_len = seqlens ? min(seqlens[it.b], Skv) : Skv;
    it.q_len = seqlens ? it.kv_len : S;
    const int kv_end = kCausal ? min(it.kv_len, it.q0 + 128 + coff) : it.kv_len;
    it.n_kv = (kv_end + 63) / 64;
    return it;
  };

  // consumer side of the item ring: slot n & 1, phase (n >> 1) & 1 for the n-th item this CTA sees (n_items = sentinel: no more work)
  auto next_item = [&](uint32_t n, bool one_lane_arrives) {
    mbar_wait(&it_full[n & 1], (n >> 1) & 1, 19);
    const int w = static_cast<int>(item_ring[n & 1]);
    if (one_lane_arrives) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&it_empty[n & 1]);
    } else {
      mbar_arrive(&it_empty[n & 1]);
    }
    return w;
  };

  if (warp == 4) { if (elect_one_sync()) {
    // ---------------- TMA producer + item scheduler ----------------
    tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv);
    uint32_t kg = 0, vg = 0, tc = 0;     // K tiles / V tiles requested so far, non-empty items started
    for (uint32_t n = 0;; ++n) {
      int w = atomicAdd(item_counter, 1);
      if (w > n_items) w = n_items;
      mbar_wait(&it_empty[n & 1], ((n >> 1) & 1) ^ 1, 19);
      item_ring[n & 1] = static_cast<uint32_t>(w);
      mbar_arrive(&it_full[n & 1]);    // release semantics: the ring write is visible to whoever observes the phase flip
      if (w >= n_items) break;
      const Item it = item(w);
      if (it.n_kv == 0) continue;
      ATTN_TRACE(tc, 6);
      mbar_wait(q_empty, (tc & 1) ^ 1, 10);
      ATTN_TRACE(tc, 7);
      mbar_arrive_expect_tx(q_full, 128 * D * 2);
      for (int c = 0; c < NCH; ++c) tma_load_3d(smem + L::oQ + c * 16384, &tq, q_full, it.h * D + c * 64, it.q0, it.b);
      int jk = 0;
      auto load_k = [&](int jj) {
        const uint32_t ks = kg % KS;
        mbar_wait(&k_empty[ks], ((kg / KS) & 1) ^ 1, 10);
        mbar_arrive_expect_tx(&k_full[ks], 64 * D * 2);
        for (int c = 0; c < NCH; ++c) tma_load_3d(smem + L::oK + ks * L::kKV + c * 8192, &tk, &k_full[ks], it.h * D + c * 64, jj * 64, it.b);
        ++kg;
      };
      for (int j = 0; j < it.n_kv; ++j) {
        while (jk < it.n_kv && jk < j + KS - 1) load_k(jk++);
        if (jk <= j) load_k(jk++);
        const uint32_t st = vg & 1;
        if (j == 0 && tc > 0) mbar_wait(o_stored, (tc - 1) & 1, 11);   // the previous item's O staging lives in the V ring
        mbar_wait(&v_empty[st], ((vg >> 1) & 1) ^ 1, 11);
        mbar_arrive_expect_tx(&v_full[st], 64 * D * 2);
        for (int c = 0; c < NCH; ++c) tma_load_3d(smem + L::oV + st * L::kKV + c * 8192, &tv, &v_full[st], it.h * D + c * 64, j * 64, it.b);
        ++vg;
      }
      ATTN_TRACE(tc, 9);
      ++tc;
    }
  } } else if (warp == 5) { if (elect_one_sync()) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, 64, false, false);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, D, false, true);
    const uint32_t sQ = smem_u32(smem + L::oQ), sK = smem_u32(smem + L::oK), sV = smem_u32(smem + L::oV);
    uint32_t g0 = 0, tc = 0;             // KV tiles of the finished items, non-empty items started
    for (uint32_t n = 0;; ++n) {
      const int w = next_item(n, false);
      if (w >= n_items) break;
      const Item it = item(w);
      if (it.n_kv == 0) continue;
      auto issue_qk = [&](int j) {
        const uint32_t g = g0 + j, ks = g % KS;
        mbar_wait(&k_full[ks], (g / KS) & 1, 12);
        tc_fence_after();
        mma_tile<D / 16>(tmem_S + (g & 1) * 64, sQ, false, 16384, sK + ks * L::kKV, false, 8192, idesc_qk, false);
        umma_commit(&k_empty[ks]);
        umma_commit(&s_full[g & 1]);
        if (j == it.n_kv - 1) umma_commit(q_empty);
      };
      ATTN_TRACE(tc, 10);
      mbar_wait(q_full, tc & 1, 13);
      ATTN_TRACE(tc, 5);
      issue_qk(0);
      for (int j = 0; j < it.n_kv; ++j) {
        if (j + 1 < it.n_kv) issue_qk(j + 1);
        const uint32_t g = g0 + j, st = g & 1, ph = (g >> 1) & 1;
        mbar_wait(&v_full[st], ph, 14);
        mbar_wait(&p_full[st], ph, 15);
        tc_fence_after();
        // O (+)= P V_j, P read from TMEM (32 columns at the base of S buffer g & 1); j == 0 overwrites the previous item's O, which its
        // softmax warps drained before they produced this P
        mma_tile_ts<4>(tmem_O, tmem_S + st * 64, 8, sV + st * L::kKV, true, 8192, idesc_pv, j > 0);
        umma_commit(&v_empty[st]);
        umma_commit(&pv_done[st]);
      }
      g0 += it.n_kv;
      ++tc;
    }
  } } else if (warp < 4) {
    // ---------------- softmax rows ----------------
    const int row = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    uint32_t g0 = 0;
    for (uint32_t n = 0;; ++n) {
      const int w = next_item(n, true);
      if (w >= n_items) break;
      const Item it = item(w);
      const int q_row = it.q0 + row;
      bf16* orow = out + (static_cast<size_t>(it.b) * S + q_row) * ld_o + it.h * D;
      if (it.n_kv == 0) {   // no keys at all: zeros
        if (q_row < S) {
#pragma unroll
          for (int c = 0; c < D / 8; ++c) reinterpret_cast<uint4*>(orow)[c] = make_uint4(0u, 0u, 0u, 0u);
          lse[(static_cast<size_t>(it.b) * nh + it.h) * S + q_row] = INFINITY;
        }
        continue;
      }
      float m_run = -INFINITY, l_run = 0.f;
#ifdef DLLM_ATTN_TRACE
      if (warp == 0) { ATTN_TRACE(tr_item, 0); if (trace_on && tr_item < 64) g_attn_trace[tr_item * 16 + 8] = it.n_kv; }
#endif
      for (int j = 0; j < it.n_kv; ++j) {
        const uint32_t g = g0 + j, sb = g & 1, ph = (g >> 1) & 1;
        mbar_wait(&s_full[sb], ph, 16);
#ifdef DLLM_ATTN_TRACE
        if (warp == 0 && j == 0) ATTN_TRACE(tr_item, 1);
        if (warp == 0 && j == 4) ATTN_TRACE(tr_item, 11);
        if (warp == 0 && j == 8) ATTN_TRACE(tr_item, 12);
#endif
        tc_fence_after();
        uint32_t sv[64];
        tmem_ld32(tmem_S + lane_off + sb * 64, sv);
        tmem_ld32(tmem_S + lane_off + sb * 64 + 32, sv + 32);
        tmem_ld_wait();
        const int kv0 = j * 64;
        const bool need_mask = (kCausal && kv0 + 63 > it.q0 + warp * 32 + coff) || (kv0 + 64 > it.kv_len) || (kv_mask != nullptr);
        // Fast path (no masking, not the item's first tile): the exponentials do not wait for this tile's row max.  They are taken against
        // the running reference m_run — exactly what the lazy rescale below would use unless the max grew by more than 2^8 — while the
        // same pass tracks the max; only if some row's max did grow that much is the tile redone on the slow path (S is still in TMEM:
        // P has not been stored yet).  Removes the separate max pass (~240 clk of dependent FMNMX per tile) from every warp's chain.
        // causal: the second diagonal tile lies entirely above the diagonal for the upper half of the q tile — no key of it is visible to
        // any row of this warp.  P = 0 without touching the exponential unit; the running max / sum do not move.
        if (kCausal && kv_mask == nullptr && kv0 > it.q0 + warp * 32 + 31 + coff) {
          uint32_t pz[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) pz[c] = 0u;
          tmem_st32(tmem_S + lane_off + sb * 64, pz);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[sb]);
          continue;
        }
        bool redo = need_mask || j == 0;
        if (!redo) {
          const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
          float ls0 = 0.f, ls1 = 0.f, mq[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          uint32_t pw[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float s0 = __uint_as_float(sv[2 * c]), s1 = __uint_as_float(sv[2 * c + 1]);
            const float p0 = exp2f(s0 * scale_log2 - m_use);
            const float p1 = exp2f(s1 * scale_log2 - m_use);
            mq[c & 3] = fmaxf(mq[c & 3], fmaxf(s0, s1));
            ls0 += p0;
            ls1 += p1;
            pw[c] = pack_bf16(p0, p1);
          }
          const float mxf = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
          const bool grew = (fmaxf(m_run, mxf * scale_log2) - m_run) > 8.0f;
          redo = __any_sync(0xffffffffu, grew);
          if (!redo) {
            tmem_st32(tmem_S + lane_off + sb * 64, pw);
            tmem_st_wait();
            l_run += ls0 + ls1;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[sb]);
          } else {
            tmem_ld32(tmem_S + lane_off + sb * 64, sv);
            tmem_ld32(tmem_S + lane_off + sb * 64 + 32, sv + 32);
            tmem_ld_wait();
          }
        }
        if (redo) {
          float mx = -INFINITY;
          if (need_mask) {
            // keys [kv0, kv0 + hi) of this tile are visible to this row: sequence end and the causal diagonal fold into one bound
            const int hi = min(it.kv_len - kv0, kCausal ? q_row + coff - kv0 + 1 : 64);
            if (kv_mask != nullptr) {   // + key-padding mask bytes (padded prompt batch inside a kv-cache)
              uint32_t mw[16];
              const uint4* mp = reinterpret_cast<const uint4*>(kv_mask + static_cast<size_t>(it.b) * mask_ld + kv0);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 t = __ldg(mp + i);
                mw[4 * i] = t.x; mw[4 * i + 1] = t.y; mw[4 * i + 2] = t.z; mw[4 * i + 3] = t.w;
              }
#pragma unroll
              for (int c = 0; c < 64; ++c) {
                const bool ok = (c < hi) && (((mw[c >> 2] >> (8 * (c & 3))) & 0xffu) != 0u);
                const float sc = ok ? __uint_as_float(sv[c]) : -INFINITY;
                sv[c] = __float_as_uint(sc);
                mx = fmaxf(mx, sc);
              }
            } else {
#pragma unroll
              for (int c = 0; c < 64; ++c) {
                const float sc = (c < hi) ? __uint_as_float(sv[c]) : -INFINITY;
                sv[c] = __float_as_uint(sc);
                mx = fmaxf(mx, sc);
              }
            }
          } else {
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int c = 0; c < 64; c += 4) {
              m0 = fmaxf(m0, __uint_as_float(sv[c]));
              m1 = fmaxf(m1, __uint_as_float(sv[c + 1]));
              m2 = fmaxf(m2, __uint_as_float(sv[c + 2]));
              m3 = fmaxf(m3, __uint_as_float(sv[c + 3]));
            }
            mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          }
          const float m_new = fmaxf(m_run, mx * scale_log2);
          const bool need = (m_new - m_run) > 8.0f;     // lazy rescale (see attn_fwd_kernel)
          const bool any = __any_sync(0xffffffffu, need);
          if (any) {
            const float m_tgt = (m_new == -INFINITY) ? m_run : m_new;
            const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_tgt);
            if (j > 0) {
              mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1, 17);   // O quiescent
              tc_fence_after();
#pragma unroll 1
              for (int c = 0; c < D / 32; ++c) {
                uint32_t o[32];
                tmem_ld32(tmem_O + lane_off + c * 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                tmem_st32(tmem_O + lane_off + c * 32, o);
              }
              tmem_st_wait();
            }
            l_run *= alpha;
            m_run = m_tgt;
          }
          const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
          float ls0 = 0.f, ls1 = 0.f;
          uint32_t pw[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float p0 = exp2f(__uint_as_float(sv[2 * c]) * scale_log2 - m_use);
            const float p1 = exp2f(__uint_as_float(sv[2 * c + 1]) * scale_log2 - m_use);
            ls0 += p0;
            ls1 += p1;
            pw[c] = pack_bf16(p0, p1);
          }
          tmem_st32(tmem_S + lane_off + sb * 64, pw);
          tmem_st_wait();
          l_run += ls0 + ls1;
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[sb]);
        }
      }
      const uint32_t gl = g0 + it.n_kv - 1;
      if (warp == 0) ATTN_TRACE(tr_item, 2);
      mbar_wait(&pv_done[gl & 1], (gl >> 1) & 1, 18);
      if (warp == 0) ATTN_TRACE(tr_item, 3);
      tc_fence_after();
      const bool valid_row = (q_row < it.q_len) && l_run > 0.f;
      const float inv = valid_row ? 1.f / l_run : 0.f;
      // O -> bf16 -> swizzled staging in the (now idle) V ring -> per-warp TMA store of [32 rows x 64 cols] boxes; rows >= S are clipped
      {
        uint8_t* stage = smem + L::oV;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          uint32_t o[32];
          tmem_ld32(tmem_O + lane_off + c * 32, o);
          tmem_ld_wait();
          uint8_t* tile = stage + (c >> 1) * 16384;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float f[8];
#pragma unroll
            for (int x = 0; x < 8; ++x) f[x] = __uint_as_float(o[8 * e + x]) * inv;
            store_row_chunk(tile, row, (c & 1) * 4 + e, f);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          for (int c = 0; c < NCH; ++c) tma_store_3d(&to, stage + c * 16384 + warp * 4096, it.h * D + c * 64, it.q0 + warp * 32, it.b);
          tma_store_commit();
          tma_store_wait_read<0>();
          mbar_arrive(o_stored);
        }
      }
      if (q_row < S)
        lse[(static_cast<size_t>(it.b) * nh + it.h) * S + q_row] = valid_row ? (m_run + log2f(l_run)) * kLn2 : INFINITY;
      tc_fence_before();   // the TMEM reads above are ordered before the arrive on p_full that releases the next item's first PV
#ifdef DLLM_ATTN_TRACE
      if (warp == 0) ATTN_TRACE(tr_item, 4);
      ++tr_item;
#endif
      g0 += it.n_kv;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<1>(tmem_base, 256);
}

// ================================================================================================ backward prep
// delta[b,h,s] = sum_d dO*O ; lse2 = lse*log2(e).
// Output layout is padded to [B, nh, S_pad] (S_pad = S rounded up to 64, pad entries = 0) so the backward kernels can
// fetch a q tile's 64 values with aligned float4 loads / one 256-byte bulk copy.
// A block owns 32 consecutive tokens of one batch entry and all heads: rows (token, head) are read 16 bytes per thread with four row
// groups in flight (contiguous: heads are adjacent in memory), the dot products are parked in shared memory as [head][token], and the
// outputs leave as 128-byte runs along the token axis.  (First version: one warp per row, 8 bytes per lane, and every result written /
// every lse read as a lone 4-byte access 8 KB from its neighbour's — three 32-byte sectors per 256-byte row, as much sector traffic as
// the payload: 107-117 us under ncu on the C2 shape = 2.4 TB/s, 10 % of the whole backward; profiles/r02u_attn_bwd_kernels.md.)
template <int D>
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ out,
                                                            const float* __restrict__ lse, float* __restrict__ delta,
                                                            float* __restrict__ lse2, int B, int S, int S_pad, int nh, long ld_o) {
  constexpr int GS = D / 8;          // lanes per row
  constexpr int RPP = 256 / GS;      // rows per block pass
  constexpr int U = 4;               // passes in flight
  constexpr int TOK = 32;
  extern __shared__ float sacc[];    // [nh][TOK]
  const int tiles_per_b = S_pad / TOK;
  const int b = blockIdx.x / tiles_per_b, s0 = (blockIdx.x - b * tiles_per_b) * TOK;
  const int ntok = max(0, min(TOK, S - s0));
  const int nrows = ntok * nh;
  const int sub = threadIdx.x / GS, li = threadIdx.x % GS;
  const size_t base = (static_cast<size_t>(b) * S + s0) * ld_o;
  for (int r0 = 0; r0 < nrows; r0 += RPP * U) {
    uint4 a[U], g[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = r0 + u * RPP + sub;
      a[u] = g[u] = make_uint4(0u, 0u, 0u, 0u);
      if (r < nrows) {
        const int tok = r / nh, h = r - tok * nh;
        const size_t off = base + static_cast<size_t>(tok) * ld_o + h * D + li * 8;
        a[u] = __ldg(reinterpret_cast<const uint4*>(out + off));
        g[u] = __ldg(reinterpret_cast<const uint4*>(dout + off));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = r0 + u * RPP + sub;
      const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a[u]);
      const __nv_bfloat162* pg = reinterpret_cast<const __nv_bfloat162*>(&g[u]);
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = __bfloat1622float2(pa[e]), y = __bfloat1622float2(pg[e]);
        acc += x.x * y.x + x.y * y.y;
      }
#pragma unroll
      for (int o = GS / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (li == 0 && r < nrows) {
        const int tok = r / nh, h = r - tok * nh;
        sacc[h * TOK + tok] = acc;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nh * TOK; i += 256) {
    const int h = i / TOK, tk = i - h * TOK, sq = s0 + tk;
    const size_t idxp = (static_cast<size_t>(b) * nh + h) * S_pad + sq;
    const bool real = sq < S;
    delta[idxp] = real ? sacc[i] : 0.f;
    lse2[idxp] = real ? lse[(static_cast<size_t>(b) * nh + h) * S + sq] * kLog2e : 0.f;
  }
}

// ================================================================================================ backward: dK, dV
// CTA = one 128-row kv tile of one (b, h); loops over 64-row q tiles.  Row threads own kv rows, so everything the
// MMAs consume as an A operand ([kv][q] tiles P^T and dS^T) is K-major as written.
//   S^T  = K  Q_i^T      dP^T = V dO_i^T            (M = 128 kv, N = 64 q, K = d)
//   dV  += P^T dO_i      dK  += dS^T Q_i            (M = 128 kv, N = d,    K = 64 q; B operands MN-major)
// 320 threads: warps 0-7 are row warps — warp w owns TMEM lane quarter (w & 3) and q-column half (w >> 2), i.e. two
// warps per SM sub-partition share a row and split its 64 columns (r01's 4-warp version was issue-latency bound:
// tensor pipe 6 %, profiles/r01_ncu_full_summary.csv).  warp 8 = TMA producer, warp 9 = TMEM alloc + MMA issuer.
// P^T / dS^T staging is double-buffered so the row warps never wait on the previous iteration's dV/dK MMAs.
constexpr int kBwdThreads = 320;

template <int D>
struct BwdKVSmem {
  static constexpr int NCH = D / 64;
  static constexpr int kKV = NCH * 16384;  // [NCH][128][128B]
  static constexpr int kQ = NCH * 8192;    // [NCH][64][128B]
  static constexpr int oK = 0, oV = kKV, oQ = 2 * kKV, oDO = oQ + 2 * kQ, oP = oDO + 2 * kQ, oDS = oP + 2 * 16384,
                       oBar = oDS + 2 * 16384;
  static constexpr int kBytes = oBar + 256;
};

template <int D, bool kCausal>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                     const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                     const __grid_constant__ CUtensorMap tdk, const __grid_constant__ CUtensorMap tdv,
                     const float* __restrict__ lse2, const float* __restrict__ delta, const int* __restrict__ seqlens,
                     int S, int Skv, int S_pad, int nh, float scale, float scale_log2) {
  using L = BwdKVSmem<D>;
  constexpr int NCH = L::NCH;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::oBar);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;   // [2]
  uint64_t* qdo_empty = bars + 3;  // [2]
  uint64_t* sdp_full = bars + 5;   // [2]
  uint64_t* pds_full = bars + 7;   // [2] (indexed i & 1: row warps may run one iteration ahead of the MMA warp)
  uint64_t* acc_done = bars + 9;   // [2]: dV/dK MMAs of iteration i retired (indexed i & 1)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int len = seqlens ? min(seqlens[b], S) : S;       // valid q rows
  const int len_kv = seqlens ? len : Skv;                  // valid kv rows (Skv != S only for cross-attention)
  const int i_begin = kCausal ? (kv0 / 64) : 0;
  const int i_end = (len + 63) / 64;  // q rows >= len carry no gradient
  const int n_it = (kv0 < len_kv) ? max(i_end - i_begin, 0) : 0;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) { printf("attn_bwd_dkdv: smem misaligned\n"); __trap(); }
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); mbar_init(&sdp_full[i], 1); mbar_init(&acc_done[i], 1);
      mbar_init(&pds_full[i], 256);
    }
    fence_mbar_init();
  }
  if (warp == 9) { tmem_alloc<1>(tmem_ptr, 512); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_St = tmem_base;         // 2 x 64
  const uint32_t tmem_dPt = tmem_base + 128;  // 2 x 64
  const uint32_t tmem_dV = tmem_base + 256;   // D
  const uint32_t tmem_dK = tmem_base + 256 + D;

  // producer / issuer: one lane picked with elect.sync — unlike `lane == 0` it lets ptxas treat the region as single-threaded, so the
  // TMA / tcgen05.mma instructions are not wrapped in a per-active-thread loop
  if (warp == 8) { if (elect_one_sync() && n_it > 0) {
    tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); tma_prefetch_desc(&tdo);
    mbar_arrive_expect_tx(kv_full, 2 * 128 * D * 2);
    for (int c = 0; c < NCH; ++c) {
      tma_load_3d(smem + L::oK + c * 16384, &tk, kv_full, h * D + c * 64, kv0, b);
      tma_load_3d(smem + L::oV + c * 16384, &tv, kv_full, h * D + c * 64, kv0, b);
    }
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1;
      const int qr0 = (i_begin + it) * 64;
      mbar_wait(&qdo_empty[st], ((it >> 1) & 1) ^ 1, 20);
      mbar_arrive_expect_tx(&qdo_full[st], 2 * 64 * D * 2);
      for (int c = 0; c < NCH; ++c) {
        tma_load_3d(smem + L::oQ + st * L::kQ + c * 8192, &tq, &qdo_full[st], h * D + c * 64, qr0, b);
        tma_load_3d(smem + L::oDO + st * L::kQ + c * 8192, &tdo, &qdo_full[st], h * D + c * 64, qr0, b);
      }
    }
  } } else if (warp == 9) { if (elect_one_sync() && n_it > 0) {
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
    constexpr uint32_t idesc_acc = make_idesc_bf16(128, D, false, true);
    const uint32_t sK = smem_u32(smem + L::oK), sV = smem_u32(smem + L::oV), sQ = smem_u32(smem + L::oQ),
                   sDO = smem_u32(smem + L::oDO), sP = smem_u32(smem + L::oP), sDS = smem_u32(smem + L::oDS);
    auto issue_s = [&](int it) {
      const int st = it & 1;
      mbar_wait(&qdo_full[st], (it >> 1) & 1, 21);
      tc_fence_after();
      mma_tile<D / 16>(tmem_St + st * 64, sK, false, 16384, sQ + st * L::kQ, false, 8192, idesc_s, false);
      mma_tile<D / 16>(tmem_dPt + st * 64, sV, false, 16384, sDO + st * L::kQ, false, 8192, idesc_s, false);
      umma_commit(&sdp_full[st]);
    };
    mbar_wait(kv_full, 0, 22);
    issue_s(0);
    for (int it = 0; it < n_it; ++it) {
      if (it + 1 < n_it) issue_s(it + 1);
      const int st = it & 1;
      mbar_wait(&pds_full[st], (it >> 1) & 1, 23);
      tc_fence_after();
      mma_tile<4>(tmem_dV, sP + st * 16384, false, 0, sDO + st * L::kQ, true, 8192, idesc_acc, it > 0);
      mma_tile<4>(tmem_dK, sDS + st * 16384, false, 0, sQ + st * L::kQ, true, 8192, idesc_acc, it > 0);
      umma_commit(&qdo_empty[st]);
      umma_commit(&acc_done[st]);
    }
  } } else if (warp < 8) {
    const int wq = warp & 3, half = warp >> 2;
    const int row = wq * 32 + lane;  // kv row within tile
    const int kv_row = kv0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const float* lse_bh = lse2 + (static_cast<size_t>(b) * nh + h) * S_pad;
    const float* del_bh = delta + (static_cast<size_t>(b) * nh + h) * S_pad;
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1;
      const int qc0 = (i_begin + it) * 64 + half * 32;  // first q index of my 32 columns
      // my 32 columns' lse / delta: 16 broadcast float4 loads (all lanes read the same addresses)
      float lq[32], dq_[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(lse_bh + qc0) + i);
        const float4 c = __ldg(reinterpret_cast<const float4*>(del_bh + qc0) + i);
        lq[4 * i] = a.x; lq[4 * i + 1] = a.y; lq[4 * i + 2] = a.z; lq[4 * i + 3] = a.w;
        dq_[4 * i] = c.x; dq_[4 * i + 1] = c.y; dq_[4 * i + 2] = c.z; dq_[4 * i + 3] = c.w;
      }
      mbar_wait(&sdp_full[st], (it >> 1) & 1, 24);
      tc_fence_after();
      uint32_t sv[32], dv[32];
      tmem_ld32(tmem_St + lane_off + st * 64 + half * 32, sv);
      tmem_ld32(tmem_dPt + lane_off + st * 64 + half * 32, dv);
      tmem_ld_wait();
      // staging buffer `st` was last read by the dV/dK MMAs of iteration it-2
      if (it >= 2) mbar_wait(&acc_done[st], ((it >> 1) & 1) ^ 1, 25);
      const bool full_tile = (qc0 + 31 < len) && (kv0 + 127 < len_kv) && (!kCausal || kv0 + 127 <= qc0);
      if (full_tile) {  // interior tile (the common case): no per-element predicates
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float p[8], ds[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = jj * 8 + e;
            const float pe = exp2f(__uint_as_float(sv[c]) * scale_log2 - lq[c]);
            p[e] = pe;
            ds[e] = pe * (__uint_as_float(dv[c]) - dq_[c]);   // softmax scale is applied once, to the dK accumulator
          }
          store_row_chunk(smem + L::oP + st * 16384, row, half * 4 + jj, p);
          store_row_chunk(smem + L::oDS + st * 16384, row, half * 4 + jj, ds);
        }
      } else {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float p[8], ds[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = jj * 8 + e;
            const int qi = qc0 + c;
            const bool ok = (qi < len) && (kv_row < len_kv) && (!kCausal || kv_row <= qi);
            const float pe = ok ? exp2f(__uint_as_float(sv[c]) * scale_log2 - lq[c]) : 0.f;
            p[e] = pe;
            ds[e] = ok ? pe * (__uint_as_float(dv[c]) - dq_[c]) : 0.f;
          }
          store_row_chunk(smem + L::oP + st * 16384, row, half * 4 + jj, p);
          store_row_chunk(smem + L::oDS + st * 16384, row, half * 4 + jj, ds);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&pds_full[st]);
    }
    const bool my_store = (D == 128) || (half == 0);
    const int cb = (D == 128) ? half : 0, ce = (D == 128) ? half + 1 : 1;
    if (n_it > 0) {
      mbar_wait(&acc_done[(n_it - 1) & 1], ((n_it - 1) >> 1) & 1, 26);
      tc_fence_after();
      if (my_store) {
        store_acc_tile<D>(tmem_dV, smem + L::oV, 1.f, &tdv, h * D, kv0, b, wq, lane, cb, ce);
        store_acc_tile<D>(tmem_dK, smem + L::oK, scale, &tdk, h * D, kv0, b, wq, lane, cb, ce);
      }
    } else if (my_store) {
      for (int c = cb; c < ce; ++c)
        for (int jj = 0; jj < 8; ++jj) {
          float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          store_row_chunk(smem + L::oK + c * 16384, row, jj, z);
        }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        for (int c = cb; c < ce; ++c) {
          tma_store_3d(&tdk, smem + L::oK + c * 16384 + wq * 4096, h * D + c * 64, kv0 + wq * 32, b);
          tma_store_3d(&tdv, smem + L::oK + c * 16384 + wq * 4096, h * D + c * 64, kv0 + wq * 32, b);
        }
        tma_store_commit();
        tma_store_wait_all<0>();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<1>(tmem_base, 512);
}

// ================================================================================================ backward: dQ
// CTA = one 128-row q tile of one (b, h); loops over 64-row kv tiles.  Same 320-thread role layout as above.
//   S = Q K_j^T    dP = dO V_j^T     (M = 128 q, N = 64 kv, K = d)       dQ += dS K_j   (M = 128 q, N = d, K = 64 kv)
template <int D>
struct BwdQSmem {
  static constexpr int NCH = D / 64;
  static constexpr int kQ = NCH * 16384;
  static constexpr int kKV = NCH * 8192;
  static constexpr int oQ = 0, oDO = kQ, oK = 2 * kQ, oV = oK + 2 * kKV, oDS = oV + 2 * kKV, oBar = oDS + 2 * 16384;
  static constexpr int kBytes = oBar + 256;
};

template <int D, bool kCausal>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                   const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                   const __grid_constant__ CUtensorMap tdq, const float* __restrict__ lse2,
                   const float* __restrict__ delta, const int* __restrict__ seqlens, int S, int Skv, int S_pad, int nh,
                   float scale, float scale_log2) {
  using L = BwdQSmem<D>;
  constexpr int NCH = L::NCH;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::oBar);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* sdp_full = bars + 5;  // [2]
  uint64_t* ds_full = bars + 7;   // [2]
  uint64_t* acc_done = bars + 9;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // causal: q tile i has i + 1 KV tiles of work; launch the heaviest first so the grid's tail is made of short CTAs (LPT order)
  const int q0 = (kCausal ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x) * 128, h = blockIdx.y, b = blockIdx.z;
  const int len = seqlens ? min(seqlens[b], S) : S;
  const int len_kv = seqlens ? len : Skv;
  const int kv_end = kCausal ? min(len_kv, q0 + 128) : len_kv;
  const int n_kv = (q0 < len) ? (kv_end + 63) / 64 : 0;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) { printf("attn_bwd_dq: smem misaligned\n"); __trap(); }
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); mbar_init(&sdp_full[i], 1); mbar_init(&acc_done[i], 1);
      mbar_init(&ds_full[i], 256);
    }
    fence_mbar_init();
  }
  if (warp == 9) { tmem_alloc<1>(tmem_ptr, 512); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 128, tmem_dQ = tmem_base + 256;

  // producer / issuer: one lane picked with elect.sync — unlike `lane == 0` it lets ptxas treat the region as single-threaded, so the
  // TMA / tcgen05.mma instructions are not wrapped in a per-active-thread loop
  if (warp == 8) { if (elect_one_sync() && n_kv > 0) {
    tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); tma_prefetch_desc(&tdo);
    mbar_arrive_expect_tx(q_full, 2 * 128 * D * 2);
    for (int c = 0; c < NCH; ++c) {
      tma_load_3d(smem + L::oQ + c * 16384, &tq, q_full, h * D + c * 64, q0, b);
      tma_load_3d(smem + L::oDO + c * 16384, &tdo, q_full, h * D + c * 64, q0, b);
    }
    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1;
      mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1, 30);
      mbar_arrive_expect_tx(&kv_full[st], 2 * 64 * D * 2);
      for (int c = 0; c < NCH; ++c) {
        tma_load_3d(smem + L::oK + st * L::kKV + c * 8192, &tk, &kv_full[st], h * D + c * 64, j * 64, b);
        tma_load_3d(smem + L::oV + st * L::kKV + c * 8192, &tv, &kv_full[st], h * D + c * 64, j * 64, b);
      }
    }
  } } else if (warp == 9) { if (elect_one_sync() && n_kv > 0) {
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
    constexpr uint32_t idesc_acc = make_idesc_bf16(128, D, false, true);
    const uint32_t sQ = smem_u32(smem + L::oQ), sDO = smem_u32(smem + L::oDO), sK = smem_u32(smem + L::oK),
                   sV = smem_u32(smem + L::oV), sDS = smem_u32(smem + L::oDS);
    auto issue_s = [&](int j) {
      const int st = j & 1;
      mbar_wait(&kv_full[st], (j >> 1) & 1, 31);
      tc_fence_after();
      mma_tile<D / 16>(tmem_S + st * 64, sQ, false, 16384, sK + st * L::kKV, false, 8192, idesc_s, false);
      mma_tile<D / 16>(tmem_dP + st * 64, sDO, false, 16384, sV + st * L::kKV, false, 8192, idesc_s, false);
      umma_commit(&sdp_full[st]);
    };
    mbar_wait(q_full, 0, 32);
    issue_s(0);
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) issue_s(j + 1);
      const int st = j & 1;
      mbar_wait(&ds_full[st], (j >> 1) & 1, 33);
      tc_fence_after();
      mma_tile<4>(tmem_dQ, sDS + st * 16384, false, 0, sK + st * L::kKV, true, 8192, idesc_acc, j > 0);
      umma_commit(&kv_empty[st]);
      umma_commit(&acc_done[st]);
    }
  } } else if (warp < 8) {
    const int wq = warp & 3, half = warp >> 2;
    const int row = wq * 32 + lane;
    const int q_row = q0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const size_t sidx = (static_cast<size_t>(b) * nh + h) * S_pad + min(q_row, S_pad - 1);
    const float my_lse = lse2[sidx], my_delta = delta[sidx];
    const bool row_ok = q_row < len;
    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1;
      const int kc0 = j * 64 + half * 32;
      mbar_wait(&sdp_full[st], (j >> 1) & 1, 34);
      tc_fence_after();
      uint32_t sv[32], dv[32];
      tmem_ld32(tmem_S + lane_off + st * 64 + half * 32, sv);
      tmem_ld32(tmem_dP + lane_off + st * 64 + half * 32, dv);
      tmem_ld_wait();
      if (j >= 2) mbar_wait(&acc_done[st], ((j >> 1) & 1) ^ 1, 35);
      const bool full_tile = (q0 + 127 < len) && (kc0 + 31 < len_kv) && (!kCausal || kc0 + 31 <= q0);
      if (full_tile) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float ds[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = jj * 8 + e;
            const float pe = exp2f(__uint_as_float(sv[c]) * scale_log2 - my_lse);
            ds[e] = pe * (__uint_as_float(dv[c]) - my_delta);   // scale applied to the dQ accumulator
          }
          store_row_chunk(smem + L::oDS + st * 16384, row, half * 4 + jj, ds);
        }
      } else {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float ds[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = jj * 8 + e;
            const int kvi = kc0 + c;
            const bool ok = row_ok && (kvi < len_kv) && (!kCausal || kvi <= q_row);
            const float pe = ok ? exp2f(__uint_as_float(sv[c]) * scale_log2 - my_lse) : 0.f;
            ds[e] = ok ? pe * (__uint_as_float(dv[c]) - my_delta) : 0.f;
          }
          store_row_chunk(smem + L::oDS + st * 16384, row, half * 4 + jj, ds);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&ds_full[st]);
    }
    const bool my_store = (D == 128) || (half == 0);
    const int cb = (D == 128) ? half : 0, ce = (D == 128) ? half + 1 : 1;
    if (n_kv > 0) {
      mbar_wait(&acc_done[(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1, 36);
      tc_fence_after();
      if (my_store) store_acc_tile<D>(tmem_dQ, smem + L::oQ, scale, &tdq, h * D, q0, b, wq, lane, cb, ce);
    } else if (my_store) {
      for (int c = cb; c < ce; ++c)
        for (int jj = 0; jj < 8; ++jj) {
          float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          store_row_chunk(smem + L::oQ + c * 16384, row, jj, z);
        }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        for (int c = cb; c < ce; ++c) tma_store_3d(&tdq, smem + L::oQ + c * 16384 + wq * 4096, h * D + c * 64, q0 + wq * 32, b);
        tma_store_commit();
        tma_store_wait_all<0>();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<1>(tmem_base, 512);
}

// ================================================================================================ backward, TS flavour (default)
// Same two deterministic kernels and the same tile algebra as above, with the three changes the round-1 profile asked for
// (row warps were the critical path: tensor pipe 27 % / 9 %, profiles/r01c_ncu_full_summary.csv):
//   * P^T / dS^T (dkdv) and dS (dq) never touch shared memory: the row warps write the bf16 pairs back into the TMEM columns they just
//     read S / dP from (tcgen05.st) and the accumulating tile-GEMMs take their A operand from tensor memory.  Warp `part` owns fp32
//     columns [16 part, 16 part + 16) of a 64-column tile and stores its 8 packed words at columns [16 part, 16 part + 8): nobody
//     overwrites a column another warp still has to read, and UMMA_K step ks simply starts at column 16 ks.
//     No STS bursts, no fence.proxy.async, no wait for the previous iteration's accumulate MMAs (the in-order tensor pipe orders them
//     against the next S / dP overwrite);
//   * 16 row warps instead of 8 (four per TMEM lane quarter, 16 columns each): the per-iteration row phase is latency-bound, so the
//     work per warp is halved and four warps per scheduler hide each other's LDTM / MUFU latency;
//   * lse / delta of the q tile are staged in shared memory by the TMA producer (one 256-byte bulk copy each, riding on the Q/dO
//     barrier) instead of 16 broadcast __ldg per thread per iteration from L2.
constexpr int kBwdRowWarps = 16;
constexpr int kBwdTsThreads = (kBwdRowWarps + 2) * 32;   // + TMA producer warp + MMA warp

// 1-D bulk copy global -> shared, completion counted on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Operand ring depth.  With two stages the TMA refill of a stage could only start when the accumulate MMAs reading it had retired, and the
// next S / dP GEMMs — issued right behind those MMAs — then sat out a full L2 round trip every iteration (the r01 kernels measured
// ~3200 clk per iteration against ~1100 clk of tensor work; both backward kernels, with 4 vs 3 tile-GEMMs, had the SAME time per
// iteration).  The shared memory freed by moving P / dS to TMEM pays for four stages, so loads run three iterations ahead.
constexpr int kBwdRing = 4;

template <int D>
struct BwdKVTsSmem {
  static constexpr int NCH = D / 64;
  static constexpr int kKV = NCH * 16384;  // [NCH][128][128B]
  static constexpr int kQ = NCH * 8192;    // [NCH][64][128B]
  static constexpr int oK = 0, oV = kKV, oQ = 2 * kKV, oDO = oQ + kBwdRing * kQ, oStat = oDO + kBwdRing * kQ,
                       oBar = oStat + kBwdRing * 512;  // stats: [ring][lse 64 | delta 64] fp32
  static constexpr int kBytes = oBar + 256;
};

template <int D, bool kCausal>
__global__ void __launch_bounds__(kBwdTsThreads, 1)
attn_bwd_dkdv_ts_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                        const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                        const __grid_constant__ CUtensorMap tdk, const __grid_constant__ CUtensorMap tdv,
                        const float* __restrict__ lse2, const float* __restrict__ delta, const int* __restrict__ seqlens,
                        int S, int Skv, int S_pad, int nh, float scale, float scale_log2) {
  using L = BwdKVTsSmem<D>;
  constexpr int NCH = L::NCH;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::oBar);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;                  // [kBwdRing]  smem ring: it % kBwdRing, phase (it / kBwdRing) & 1
  uint64_t* qdo_empty = qdo_full + kBwdRing;      // [kBwdRing]
  uint64_t* sdp_full = qdo_empty + kBwdRing;      // [2]         TMEM double buffer: it & 1, phase (it >> 1) & 1
  uint64_t* pds_full = sdp_full + 2;              // [2]
  uint64_t* acc_done = pds_full + 2;              // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int len = seqlens ? min(seqlens[b], S) : S;
  const int len_kv = seqlens ? len : Skv;
  const int i_begin = kCausal ? (kv0 / 64) : 0;
  const int i_end = (len + 63) / 64;
  const int n_it = (kv0 < len_kv) ? max(i_end - i_begin, 0) : 0;
#ifdef DLLM_ATTN_TRACE
  const bool trace_on = (blockIdx.x == 0 && blockIdx.y == 1 && blockIdx.z == 0 && lane == 0);
  if (trace_on && warp == 0) g_attn_trace[64 * 16 + 3] = clock64();
#endif

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) { printf("attn_bwd_dkdv_ts: smem misaligned\n"); __trap(); }
    mbar_init(kv_full, 1);
    for (int i = 0; i < kBwdRing; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sdp_full[i], 1); mbar_init(&acc_done[i], 1);
      mbar_init(&pds_full[i], kBwdRowWarps);      // one elected arrive per row warp (512 same-address arrives cost ~500 clk per iteration)
    }
    fence_mbar_init();
  }
  if (warp == kBwdRowWarps + 1) { tmem_alloc<1>(tmem_ptr, 512); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_St = tmem_base;         // 2 x 64 (fp32 S^T, then bf16 P^T in the first 8 columns of every 16)
  const uint32_t tmem_dPt = tmem_base + 128;  // 2 x 64 (fp32 dP^T, then bf16 dS^T likewise)
  const uint32_t tmem_dV = tmem_base + 256;   // D
  const uint32_t tmem_dK = tmem_base + 256 + D;

  // producer / issuer: one lane picked with elect.sync — unlike `lane == 0` it lets ptxas treat the region as single-threaded, so the
  // TMA / tcgen05.mma instructions are not wrapped in a per-active-thread loop
  if (warp == kBwdRowWarps) { if (elect_one_sync() && n_it > 0) {
    // ---------------- TMA producer ----------------
    tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); tma_prefetch_desc(&tdo);
    mbar_arrive_expect_tx(kv_full, 2 * 128 * D * 2);
    for (int c = 0; c < NCH; ++c) {
      tma_load_3d(smem + L::oK + c * 16384, &tk, kv_full, h * D + c * 64, kv0, b);
      tma_load_3d(smem + L::oV + c * 16384, &tv, kv_full, h * D + c * 64, kv0, b);
    }
    const float* lse_bh = lse2 + (static_cast<size_t>(b) * nh + h) * S_pad;
    const float* del_bh = delta + (static_cast<size_t>(b) * nh + h) * S_pad;
    for (int it = 0; it < n_it; ++it) {
      const int rs = it % kBwdRing;
      const int qr0 = (i_begin + it) * 64;
      ATTN_TRACE(it, 8);
      mbar_wait(&qdo_empty[rs], ((it / kBwdRing) & 1) ^ 1, 40);
      ATTN_TRACE(it, 9);
      mbar_arrive_expect_tx(&qdo_full[rs], 2 * 64 * D * 2 + 512);
      for (int c = 0; c < NCH; ++c) {
        tma_load_3d(smem + L::oQ + rs * L::kQ + c * 8192, &tq, &qdo_full[rs], h * D + c * 64, qr0, b);
        tma_load_3d(smem + L::oDO + rs * L::kQ + c * 8192, &tdo, &qdo_full[rs], h * D + c * 64, qr0, b);
      }
      bulk_load_1d(smem + L::oStat + rs * 512, lse_bh + qr0, 256, &qdo_full[rs]);          // S_pad is a multiple of 64: always in range
      bulk_load_1d(smem + L::oStat + rs * 512 + 256, del_bh + qr0, 256, &qdo_full[rs]);
    }
  } } else if (warp == kBwdRowWarps + 1) { if (elect_one_sync() && n_it > 0) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
    constexpr uint32_t idesc_acc = make_idesc_bf16(128, D, false, true);
    const uint32_t sK = smem_u32(smem + L::oK), sV = smem_u32(smem + L::oV), sQ = smem_u32(smem + L::oQ),
                   sDO = smem_u32(smem + L::oDO);
    auto issue_s = [&](int it) {
      const int st = it & 1, rs = it % kBwdRing;
      ATTN_TRACE(it, 11);
      mbar_wait(&qdo_full[rs], (it / kBwdRing) & 1, 41);
      ATTN_TRACE(it, 5);
      tc_fence_after();
      mma_tile<D / 16>(tmem_St + st * 64, sK, false, 16384, sQ + rs * L::kQ, false, 8192, idesc_s, false);
      mma_tile<D / 16>(tmem_dPt + st * 64, sV, false, 16384, sDO + rs * L::kQ, false, 8192, idesc_s, false);
      umma_commit(&sdp_full[st]);
    };
    mbar_wait(kv_full, 0, 42);
    issue_s(0);
    for (int it = 0; it < n_it; ++it) {
      if (it + 1 < n_it) issue_s(it + 1);
      const int st = it & 1, rs = it % kBwdRing;
      ATTN_TRACE(it, 6);
      mbar_wait(&pds_full[st], (it >> 1) & 1, 43);
      ATTN_TRACE(it, 7);
      tc_fence_after();
      // dV += P^T dO_i : A = P^T from TMEM (k-step ks = q columns [16 ks, 16 ks + 16) at column 16 ks);  dK += dS^T Q_i
      mma_tile_ts<4>(tmem_dV, tmem_St + st * 64, 16, sDO + rs * L::kQ, true, 8192, idesc_acc, it > 0);
      mma_tile_ts<4>(tmem_dK, tmem_dPt + st * 64, 16, sQ + rs * L::kQ, true, 8192, idesc_acc, it > 0);
      umma_commit(&qdo_empty[rs]);
      umma_commit(&acc_done[st]);
    }
  } } else if (warp < kBwdRowWarps) {
    // ---------------- row warps ----------------
    const int wq = warp & 3, part = warp >> 2;
    const int row = wq * 32 + lane;  // kv row within tile
    const int kv_row = kv0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1, rs = it % kBwdRing;
      const int qc0 = (i_begin + it) * 64 + part * 16;  // first q index of my 16 columns
      if (warp == 0) ATTN_TRACE(it, 10);
      mbar_wait(&qdo_full[rs], (it / kBwdRing) & 1, 45);  // the bulk-copied lse / delta of this q tile are visible to this thread
      mbar_wait(&sdp_full[st], (it >> 1) & 1, 44);
      if (warp == 0) ATTN_TRACE(it, 0);
      tc_fence_after();
      uint32_t sv[16], dv[16];
      tmem_ld16(tmem_St + lane_off + st * 64 + part * 16, sv);
      tmem_ld16(tmem_dPt + lane_off + st * 64 + part * 16, dv);
      float lq[16], dq_[16];
      {
        const float4* ls = reinterpret_cast<const float4*>(smem + L::oStat + rs * 512) + part * 4;
        const float4* ds4 = reinterpret_cast<const float4*>(smem + L::oStat + rs * 512 + 256) + part * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 a = ls[i], c = ds4[i];
          lq[4 * i] = a.x; lq[4 * i + 1] = a.y; lq[4 * i + 2] = a.z; lq[4 * i + 3] = a.w;
          dq_[4 * i] = c.x; dq_[4 * i + 1] = c.y; dq_[4 * i + 2] = c.z; dq_[4 * i + 3] = c.w;
        }
      }
      tmem_ld_wait();
      if (warp == 0) ATTN_TRACE(it, 1);
      uint32_t pw[8], dw[8];
      const bool full_tile = (qc0 + 15 < len) && (kv0 + 127 < len_kv) && (!kCausal || kv0 + 127 <= qc0);
      if (full_tile) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float p0 = exp2f(__uint_as_float(sv[2 * c]) * scale_log2 - lq[2 * c]);
          const float p1 = exp2f(__uint_as_float(sv[2 * c + 1]) * scale_log2 - lq[2 * c + 1]);
          pw[c] = pack_bf16(p0, p1);
          dw[c] = pack_bf16(p0 * (__uint_as_float(dv[2 * c]) - dq_[2 * c]), p1 * (__uint_as_float(dv[2 * c + 1]) - dq_[2 * c + 1]));
        }
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float pe[2], de[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int cc = 2 * c + e;
            const int qi = qc0 + cc;
            const bool ok = (qi < len) && (kv_row < len_kv) && (!kCausal || kv_row <= qi);
            pe[e] = ok ? exp2f(__uint_as_float(sv[cc]) * scale_log2 - lq[cc]) : 0.f;
            de[e] = ok ? pe[e] * (__uint_as_float(dv[cc]) - dq_[cc]) : 0.f;
          }
          pw[c] = pack_bf16(pe[0], pe[1]);
          dw[c] = pack_bf16(de[0], de[1]);
        }
      }
      if (warp == 0) ATTN_TRACE(it, 3);
      tmem_st8(tmem_St + lane_off + st * 64 + part * 16, pw);
      tmem_st8(tmem_dPt + lane_off + st * 64 + part * 16, dw);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full[st]);
      if (warp == 0) ATTN_TRACE(it, 4);
    }
    // epilogue: 4 (D = 128) or 2 (D = 64) 64-column chunk jobs {dV, dK} x chunks, one per `part`
    constexpr int kJobs = 2 * NCH;
    const bool my_store = part < kJobs;
    const int acc = part / NCH, ch = part % NCH;
    if (n_it > 0) {
      mbar_wait(&acc_done[(n_it - 1) & 1], ((n_it - 1) >> 1) & 1, 46);
      tc_fence_after();
      if (my_store) {
        if (acc == 0) store_acc_tile<D>(tmem_dV, smem + L::oV, 1.f, &tdv, h * D, kv0, b, wq, lane, ch, ch + 1);
        else store_acc_tile<D>(tmem_dK, smem + L::oK, scale, &tdk, h * D, kv0, b, wq, lane, ch, ch + 1);
      }
    } else if (my_store) {
      uint8_t* stage = smem + (acc == 0 ? L::oV : L::oK) + ch * 16384;
      for (int jj = 0; jj < 8; ++jj) {
        float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        store_row_chunk(stage, row, jj, z);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(acc == 0 ? &tdv : &tdk, stage + wq * 4096, h * D + ch * 64, kv0 + wq * 32, b);
        tma_store_commit();
        tma_store_wait_all<0>();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kBwdRowWarps + 1) tmem_dealloc<1>(tmem_base, 512);
}

template <int D>
struct BwdQTsSmem {
  static constexpr int NCH = D / 64;
  static constexpr int kQ = NCH * 16384;
  static constexpr int kKV = NCH * 8192;
  static constexpr int oQ = 0, oDO = kQ, oK = 2 * kQ, oV = oK + kBwdRing * kKV, oBar = oV + kBwdRing * kKV;
  static constexpr int kBytes = oBar + 256;
};

template <int D, bool kCausal>
__global__ void __launch_bounds__(kBwdTsThreads, 1)
attn_bwd_dq_ts_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                      const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                      const __grid_constant__ CUtensorMap tdq, const float* __restrict__ lse2, const float* __restrict__ delta,
                      const int* __restrict__ seqlens, int S, int Skv, int S_pad, int nh, float scale, float scale_log2) {
  using L = BwdQTsSmem<D>;
  constexpr int NCH = L::NCH;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::oBar);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;                 // [kBwdRing]
  uint64_t* kv_empty = kv_full + kBwdRing;      // [kBwdRing]
  uint64_t* sdp_full = kv_empty + kBwdRing;     // [2]
  uint64_t* ds_full = sdp_full + 2;             // [2]
  uint64_t* acc_done = ds_full + 2;             // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // causal: q tile i has i + 1 KV tiles of work; launch the heaviest first so the grid's tail is made of short CTAs (LPT order)
  const int q0 = (kCausal ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x) * 128, h = blockIdx.y, b = blockIdx.z;
  const int len = seqlens ? min(seqlens[b], S) : S;
  const int len_kv = seqlens ? len : Skv;
  const int kv_end = kCausal ? min(len_kv, q0 + 128) : len_kv;
  const int n_kv = (q0 < len) ? (kv_end + 63) / 64 : 0;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) { printf("attn_bwd_dq_ts: smem misaligned\n"); __trap(); }
    mbar_init(q_full, 1);
    for (int i = 0; i < kBwdRing; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sdp_full[i], 1); mbar_init(&acc_done[i], 1);
      mbar_init(&ds_full[i], kBwdRowWarps);
    }
    fence_mbar_init();
  }
  if (warp == kBwdRowWarps + 1) { tmem_alloc<1>(tmem_ptr, 512); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 128, tmem_dQ = tmem_base + 256;

  // producer / issuer: one lane picked with elect.sync — unlike `lane == 0` it lets ptxas treat the region as single-threaded, so the
  // TMA / tcgen05.mma instructions are not wrapped in a per-active-thread loop
  if (warp == kBwdRowWarps) { if (elect_one_sync() && n_kv > 0) {
    tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); tma_prefetch_desc(&tdo);
    mbar_arrive_expect_tx(q_full, 2 * 128 * D * 2);
    for (int c = 0; c < NCH; ++c) {
      tma_load_3d(smem + L::oQ + c * 16384, &tq, q_full, h * D + c * 64, q0, b);
      tma_load_3d(smem + L::oDO + c * 16384, &tdo, q_full, h * D + c * 64, q0, b);
    }
    for (int j = 0; j < n_kv; ++j) {
      const int rs = j % kBwdRing;
      mbar_wait(&kv_empty[rs], ((j / kBwdRing) & 1) ^ 1, 50);
      mbar_arrive_expect_tx(&kv_full[rs], 2 * 64 * D * 2);
      for (int c = 0; c < NCH; ++c) {
        tma_load_3d(smem + L::oK + rs * L::kKV + c * 8192, &tk, &kv_full[rs], h * D + c * 64, j * 64, b);
        tma_load_3d(smem + L::oV + rs * L::kKV + c * 8192, &tv, &kv_full[rs], h * D + c * 64, j * 64, b);
      }
    }
  } } else if (warp == kBwdRowWarps + 1) { if (elect_one_sync() && n_kv > 0) {
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
    constexpr uint32_t idesc_acc = make_idesc_bf16(128, D, false, true);
    const uint32_t sQ = smem_u32(smem + L::oQ), sDO = smem_u32(smem + L::oDO), sK = smem_u32(smem + L::oK),
                   sV = smem_u32(smem + L::oV);
    auto issue_s = [&](int j) {
      const int st = j & 1, rs = j % kBwdRing;
      mbar_wait(&kv_full[rs], (j / kBwdRing) & 1, 51);
      tc_fence_after();
      mma_tile<D / 16>(tmem_S + st * 64, sQ, false, 16384, sK + rs * L::kKV, false, 8192, idesc_s, false);
      mma_tile<D / 16>(tmem_dP + st * 64, sDO, false, 16384, sV + rs * L::kKV, false, 8192, idesc_s, false);
      umma_commit(&sdp_full[st]);
    };
    mbar_wait(q_full, 0, 52);
    issue_s(0);
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) issue_s(j + 1);
      const int st = j & 1, rs = j % kBwdRing;
      mbar_wait(&ds_full[st], (j >> 1) & 1, 53);
      tc_fence_after();
      mma_tile_ts<4>(tmem_dQ, tmem_dP + st * 64, 16, sK + rs * L::kKV, true, 8192, idesc_acc, j > 0);   // dQ += dS K_j : A = dS from TMEM
      umma_commit(&kv_empty[rs]);
      umma_commit(&acc_done[st]);
    }
  } } else if (warp < kBwdRowWarps) {
    const int wq = warp & 3, part = warp >> 2;
    const int row = wq * 32 + lane;
    const int q_row = q0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const size_t sidx = (static_cast<size_t>(b) * nh + h) * S_pad + min(q_row, S_pad - 1);
    const float my_lse = lse2[sidx], my_delta = delta[sidx];
    const bool row_ok = q_row < len;
    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1;
      const int kc0 = j * 64 + part * 16;
      mbar_wait(&sdp_full[st], (j >> 1) & 1, 54);
      tc_fence_after();
      uint32_t sv[16], dv[16];
      tmem_ld16(tmem_S + lane_off + st * 64 + part * 16, sv);
      tmem_ld16(tmem_dP + lane_off + st * 64 + part * 16, dv);
      tmem_ld_wait();
      uint32_t dw[8];
      const bool full_tile = (q0 + 127 < len) && (kc0 + 15 < len_kv) && (!kCausal || kc0 + 15 <= q0);
      if (full_tile) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float p0 = exp2f(__uint_as_float(sv[2 * c]) * scale_log2 - my_lse);
          const float p1 = exp2f(__uint_as_float(sv[2 * c + 1]) * scale_log2 - my_lse);
          dw[c] = pack_bf16(p0 * (__uint_as_float(dv[2 * c]) - my_delta), p1 * (__uint_as_float(dv[2 * c + 1]) - my_delta));
        }
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float de[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int cc = 2 * c + e;
            const int kvi = kc0 + cc;
            const bool ok = row_ok && (kvi < len_kv) && (!kCausal || kvi <= q_row);
            const float pe = ok ? exp2f(__uint_as_float(sv[cc]) * scale_log2 - my_lse) : 0.f;
            de[e] = ok ? pe * (__uint_as_float(dv[cc]) - my_delta) : 0.f;
          }
          dw[c] = pack_bf16(de[0], de[1]);
        }
      }
      tmem_st8(tmem_dP + lane_off + st * 64 + part * 16, dw);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full[st]);
    }
    const bool my_store = part < NCH;
    if (n_kv > 0) {
      mbar_wait(&acc_done[(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1, 56);
      tc_fence_after();
      if (my_store) store_acc_tile<D>(tmem_dQ, smem + L::oQ, scale, &tdq, h * D, q0, b, wq, lane, part, part + 1);
    } else if (my_store) {
      for (int jj = 0; jj < 8; ++jj) {
        float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        store_row_chunk(smem + L::oQ + part * 16384, row, jj, z);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&tdq, smem + L::oQ + part * 16384 + wq * 4096, h * D + part * 64, q0 + wq * 32, b);
        tma_store_commit();
        tma_store_wait_all<0>();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kBwdRowWarps + 1) tmem_dealloc<1>(tmem_base, 512);
}

// ================================================================================================ backward, persistent
// The TS-form backward kernels (above) with the same treatment as the forward: ONE CTA per SM walks (tile, head, batch) items drawn from
// an atomic counter in windowed heaviest-first order, instead of one CTA per item.  With one 194 KB CTA per SM nothing hid a CTA's
// fixed costs: the dK/dV trace (profiles/r02m_attn_bwd_timeline.md) shows 9.8 k clk from CTA start to the first S^T (K/V + first Q/dO
// load latency on a cold SM), ~3 k clk of epilogue and ~4 k clk until the next CTA starts on the SM, against 17 iterations x 1.5 k clk of
// work on average — 40 % of the SM's time.  Persistent:
//   * the Q/dO (dK/dV kernel) resp. K/V (dQ kernel) operand ring and every barrier phase keep counting across items, so the producer
//     prefetches the next item's first ring tiles while the current item's last iterations and epilogue run;
//   * the stationary operands (K, V resp. Q, dO) are refilled as soon as the item's last S / dP tile-GEMMs retire (`kv_empty` in the dK/dV kernel, `q_empty` in the dQ kernel);
//   * the accumulators leave through the operand ring: the epilogue stages bf16 tiles in the ring stage(s) the item used last (free once
//     its last accumulate-MMAs retire), and the producer refills those stages only after the `stored_cnt` counter says the bulk stores have read them;
//   * the next item's first accumulate-MMA (accumulate = 0) needs P^T/dS^T of that item from all 16 row warps, which produce it only
//     after their part of the epilogue has drained the accumulators — no extra barrier.
// Monotonic "stores done" counter in shared memory (release add by the storing warps, acquire poll by the waiters).  An mbarrier would
// do for a waiter that observes every phase, but the producer only looks when it is about to refill a staging stage, and with short items
// (two q tiles: the whole item fits the ring) it runs two or more items ahead of the epilogues — a parity wait then aliases and lets it
// overwrite staging that has not been stored yet (seen as wrong dK / dV tiles at window boundaries; profiles/r02x_attn_bwd_race.md).
__device__ __forceinline__ void smem_count_add_release(uint32_t* cnt) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(cnt)) : "memory");
}
__device__ __forceinline__ void smem_count_wait_acquire(const uint32_t* cnt, uint32_t target) {
  uint32_t v;
  do {
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(cnt)) : "memory");
  } while (static_cast<int32_t>(v - target) < 0);
}
struct BwdOut {
  bf16* dq; bf16* dk; bf16* dv;
  long ld_dq, ld_dkv;
};
struct ItemId { int tile, hb; };
// w-th item in windowed order: windows of `win_heads` (head, batch) pairs, inside a window tiles in heaviest-first order
__device__ __forceinline__ ItemId decode_item(int w, int ntiles, int n_hb, int win_heads, bool descending) {
  const int wsz = win_heads * ntiles, n_full = n_hb / win_heads, full_items = n_full * wsz;
  int win, idx, R;
  if (w < full_items) { win = w / wsz; idx = w - win * wsz; R = win_heads; }
  else { win = n_full; idx = w - full_items; R = n_hb - n_full * win_heads; }
  const int tpos = idx / R;
  ItemId r;
  r.tile = descending ? ntiles - 1 - tpos : tpos;
  r.hb = win * win_heads + idx - tpos * R;
  return r;
}
// this warp's 32 accumulator rows x 64 fp32 columns -> * mul -> bf16 -> 128B-swizzled 4 KB staging block -> one [32 x 64] TMA store box;
// returns when the bulk store has read the staging block
__device__ __forceinline__ void store_acc_rows64(uint32_t tmem_acc, uint8_t* stage_warp, float mul, const CUtensorMap* tm, int gcol,
                                                 int grow, int b, int lane) {
#pragma unroll 1
  for (int hlf = 0; hlf < 2; ++hlf) {
    uint32_t v[32];
    tmem_ld32(tmem_acc + hlf * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 q;
      q.x = pack_bf16(__uint_as_float(v[8 * j + 0]) * mul, __uint_as_float(v[8 * j + 1]) * mul);
      q.y = pack_bf16(__uint_as_float(v[8 * j + 2]) * mul, __uint_as_float(v[8 * j + 3]) * mul);
      q.z = pack_bf16(__uint_as_float(v[8 * j + 4]) * mul, __uint_as_float(v[8 * j + 5]) * mul);
      q.w = pack_bf16(__uint_as_float(v[8 * j + 6]) * mul, __uint_as_float(v[8 * j + 7]) * mul);
      *reinterpret_cast<uint4*>(stage_warp + lane * 128 + (((hlf * 4 + j) ^ (lane & 7)) << 4)) = q;
    }
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_3d(tm, stage_warp, gcol, grow, b);
    tma_store_commit();
    tma_store_wait_read<0>();
  }
}
__device__ __forceinline__ void zero_row64(bf16* p) {
#pragma unroll
  for (int c = 0; c < 8; ++c) reinterpret_cast<uint4*>(p)[c] = make_uint4(0u, 0u, 0u, 0u);
}

template <int D, bool kCausal>
__global__ void __launch_bounds__(kBwdTsThreads, 1)
attn_bwd_dkdv_persist_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                             const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                             const __grid_constant__ CUtensorMap tdk, const __grid_constant__ CUtensorMap tdv,
                             const float* __restrict__ lse2, const float* __restrict__ delta, const int* __restrict__ seqlens,
                             BwdOut go, int B, int S, int Skv, int S_pad, int nh, float scale, float scale_log2,
                             int* __restrict__ item_counter) {
  using L = BwdKVTsSmem<D>;
  constexpr int NCH = L::NCH;
  constexpr int R4 = kBwdRing;
  static_assert(R4 == 4, "stage arithmetic below assumes a 4-stage ring");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::oBar);
  uint64_t* kv_full = bars + 0;
  uint64_t* kv_empty = bars + 1;                  // the item's last S^T / dP^T tile-GEMMs have retired: K / V may be refilled
  uint64_t* qdo_full = bars + 2;                  // [4]  ring over ALL q tiles this CTA processes: g % 4, phase (g / 4) & 1
  uint64_t* qdo_empty = qdo_full + R4;            // [4]
  uint64_t* sdp_full = qdo_empty + R4;            // [2]  TMEM double buffer: g & 1, phase (g >> 1) & 1
  uint64_t* pds_full = sdp_full + 2;              // [2]
  uint64_t* acc_done = pds_full + 2;              // [2]
  uint64_t* it_full = acc_done + 2;               // [2]  item ring
  uint64_t* it_empty = it_full + 2;               // [2]
  uint32_t* item_ring = reinterpret_cast<uint32_t*>(it_empty + 2);
  uint32_t* tmem_ptr = item_ring + 2;
  uint32_t* stored_cnt = tmem_ptr + 1;            // storing warps whose bulk stores have read their staging (4 * kJobs per non-empty item)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kJobs = 2 * NCH;                  // 64-column chunk jobs {dV, dK} x chunks, one per `part`
  const int nt = (Skv + 127) / 128, n_hb = nh * B, n_items = nt * n_hb;
  const int win_heads = max(1, min(n_hb, (2 * static_cast<int>(gridDim.x) + nt - 1) / nt));

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) { printf("attn_bwd_dkdv_persist: smem misaligned\n"); __trap(); }
    mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
    for (int i = 0; i < R4; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sdp_full[i], 1); mbar_init(&acc_done[i], 1);
      mbar_init(&pds_full[i], kBwdRowWarps);
      mbar_init(&it_full[i], 1);
      mbar_init(&it_empty[i], kBwdRowWarps + 1);  // MMA thread + one lane of every row warp
    }
    *stored_cnt = 0;
    fence_mbar_init();
  }
  if (warp == kBwdRowWarps + 1) { tmem_alloc<1>(tmem_ptr, 512); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_St = tmem_base, tmem_dPt = tmem_base + 128, tmem_dV = tmem_base + 256, tmem_dK = tmem_base + 256 + D;

  struct Item { int kv0, h, b, len, len_kv, i_begin, n_it; };
  auto item = [&](int w) {
    const ItemId id = decode_item(w, nt, n_hb, win_heads, false);   // causal: kv tile 0 sees every q tile -> heaviest first
    Item it;
    it.kv0 = id.tile * 128; it.h = id.hb % nh; it.b = id.hb / nh;
    it.len = seqlens ? min(seqlens[it.b], S) : S;
    it.len_kv = seqlens ? it.len : Skv;
    it.i_begin = kCausal ? (it.kv0 / 64) : 0;
    const int i_end = (it.len + 63) / 64;
    it.n_it = (it.kv0 < it.len_kv) ? max(i_end - it.i_begin, 0) : 0;
    return it;
  };
  auto next_item = [&](uint32_t n, bool one_lane_arrives) {
    mbar_wait(&it_full[n & 1], (n >> 1) & 1, 39);
    const int w = static_cast<int>(item_ring[n & 1]);
    if (one_lane_arrives) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&it_empty[n & 1]);
    } else {
      mbar_arrive(&it_empty[n & 1]);
    }
    return w;
  };

  if (warp == kBwdRowWarps) { if (elect_one_sync()) {
    // ---------------- TMA producer + item scheduler ----------------
    tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); tma_prefetch_desc(&tdo);
    uint32_t qg = 0, ic = 0, hold = 0;   // q tiles requested, non-empty items started, ring stages holding un-stored dK / dV staging
    for (uint32_t n = 0;; ++n) {
      int w = atomicAdd(item_counter, 1);
      if (w > n_items) w = n_items;
      mbar_wait(&it_empty[n & 1], ((n >> 1) & 1) ^ 1, 39);
      item_ring[n & 1] = static_cast<uint32_t>(w);
      mbar_arrive(&it_full[n & 1]);
      if (w >= n_items) break;
      const Item it = item(w);
      if (it.n_it == 0) continue;
      mbar_wait(kv_empty, (ic & 1) ^ 1, 40);
      mbar_arrive_expect_tx(kv_full, 2 * 128 * D * 2);
      for (int c = 0; c < NCH; ++c) {
        tma_load_3d(smem + L::oK + c * 16384, &tk, kv_full, it.h * D + c * 64, it.kv0, it.b);
        tma_load_3d(smem + L::oV + c * 16384, &tv, kv_full, it.h * D + c * 64, it.kv0, it.b);
      }
      const float* lse_bh = lse2 + (static_cast<size_t>(it.b) * nh + it.h) * S_pad;
      const float* del_bh = delta + (static_cast<size_t>(it.b) * nh + it.h) * S_pad;
      for (int t = 0; t < it.n_it; ++t) {
        const uint32_t rs = qg & 3;
        const int qr0 = (it.i_begin + t) * 64;
        if ((hold >> rs) & 1u) { smem_count_wait_acquire(stored_cnt, 4u * kJobs * ic); hold = 0; }   // every earlier item is stored
        mbar_wait(&qdo_empty[rs], ((qg >> 2) & 1) ^ 1, 40);
        mbar_arrive_expect_tx(&qdo_full[rs], 2 * 64 * D * 2 + 512);
        for (int c = 0; c < NCH; ++c) {
          tma_load_3d(smem + L::oQ + rs * L::kQ + c * 8192, &tq, &qdo_full[rs], it.h * D + c * 64, qr0, it.b);
          tma_load_3d(smem + L::oDO + rs * L::kQ + c * 8192, &tdo, &qdo_full[rs], it.h * D + c * 64, qr0, it.b);
        }
        bulk_load_1d(smem + L::oStat + rs * 512, lse_bh + qr0, 256, &qdo_full[rs]);
        bulk_load_1d(smem + L::oStat + rs * 512 + 256, del_bh + qr0, 256, &qdo_full[rs]);
        ++qg;
      }
      hold |= (1u << ((qg + 3) & 3)) | (1u << ((qg + 2) & 3));   // the stages of the item's last two q tiles become its staging area
      ++ic;
    }
  } } else if (warp == kBwdRowWarps + 1) { if (elect_one_sync()) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
    constexpr uint32_t idesc_acc = make_idesc_bf16(128, D, false, true);
    const uint32_t sK = smem_u32(smem + L::oK), sV = smem_u32(smem + L::oV), sQ = smem_u32(smem + L::oQ),
                   sDO = smem_u32(smem + L::oDO);
    uint32_t g0 = 0, ic = 0;
    for (uint32_t n = 0;; ++n) {
      const int w = next_item(n, false);
      if (w >= n_items) break;
      const Item it = item(w);
      if (it.n_it == 0) continue;
      auto issue_s = [&](int t) {
        const uint32_t g = g0 + t, st = g & 1, rs = g & 3;
        mbar_wait(&qdo_full[rs], (g >> 2) & 1, 41);
        tc_fence_after();
        mma_tile<D / 16>(tmem_St + st * 64, sK, false, 16384, sQ + rs * L::kQ, false, 8192, idesc_s, false);
        mma_tile<D / 16>(tmem_dPt + st * 64, sV, false, 16384, sDO + rs * L::kQ, false, 8192, idesc_s, false);
        umma_commit(&sdp_full[st]);
        if (t == it.n_it - 1) umma_commit(kv_empty);
      };
      mbar_wait(kv_full, ic & 1, 42);
      issue_s(0);
      for (int t = 0; t < it.n_it; ++t) {
        if (t + 1 < it.n_it) issue_s(t + 1);
        const uint32_t g = g0 + t, st = g & 1, rs = g & 3;
        mbar_wait(&pds_full[st], (g >> 1) & 1, 43);
        tc_fence_after();
        // dV (+)= P^T dO_i, dK (+)= dS^T Q_i with P^T / dS^T read from TMEM; t == 0 overwrites the previous item's accumulators, which
        // its row warps drained before they produced this P^T / dS^T
        mma_tile_ts<4>(tmem_dV, tmem_St + st * 64, 16, sDO + rs * L::kQ, true, 8192, idesc_acc, t > 0);
        mma_tile_ts<4>(tmem_dK, tmem_dPt + st * 64, 16, sQ + rs * L::kQ, true, 8192, idesc_acc, t > 0);
        umma_commit(&qdo_empty[rs]);
        umma_commit(&acc_done[st]);
      }
      g0 += it.n_it;
      ++ic;
    }
  } } else if (warp < kBwdRowWarps) {
    // ---------------- row warps ----------------
    const int wq = warp & 3, part = warp >> 2;
    const int row = wq * 32 + lane;  // kv row within tile
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const bool my_store = part < kJobs;
    const int acc = part / NCH, ch = part % NCH;
    uint32_t g0 = 0, ic = 0;
    for (uint32_t n = 0;; ++n) {
      const int w = next_item(n, true);
      if (w >= n_items) break;
      const Item it = item(w);
      const int kv_row = it.kv0 + row;
      if (it.n_it == 0) {   // no query attends to this kv tile: zeros
        if (my_store && kv_row < Skv)
          zero_row64((acc == 0 ? go.dv : go.dk) + (static_cast<size_t>(it.b) * Skv + kv_row) * go.ld_dkv + it.h * D + ch * 64);
        continue;
      }
      for (int t = 0; t < it.n_it; ++t) {
        const uint32_t g = g0 + t, st = g & 1, rs = g & 3;
        const int qc0 = (it.i_begin + t) * 64 + part * 16;  // first q index of my 16 columns
        mbar_wait(&qdo_full[rs], (g >> 2) & 1, 45);  // the bulk-copied lse / delta of this q tile are visible to this thread
        mbar_wait(&sdp_full[st], (g >> 1) & 1, 44);
        tc_fence_after();
        // causal diagonal, per warp rather than per tile: every kv row of this warp lies below every q column of its 16 -> P^T = dS^T = 0
        // without loads or exponentials (the upper half of the kv tile against the first q tile of the item)
        if (kCausal && it.kv0 + wq * 32 > qc0 + 15) {
          uint32_t z[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) z[c] = 0u;
          tmem_st8(tmem_St + lane_off + st * 64 + part * 16, z);
          tmem_st8(tmem_dPt + lane_off + st * 64 + part * 16, z);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&pds_full[st]);
          continue;
        }
        uint32_t sv[16], dv[16];
        tmem_ld16(tmem_St + lane_off + st * 64 + part * 16, sv);
        tmem_ld16(tmem_dPt + lane_off + st * 64 + part * 16, dv);
        float lq[16], dq_[16];
        {
          const float4* ls = reinterpret_cast<const float4*>(smem + L::oStat + rs * 512) + part * 4;
          const float4* ds4 = reinterpret_cast<const float4*>(smem + L::oStat + rs * 512 + 256) + part * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 a = ls[i], c = ds4[i];
            lq[4 * i] = a.x; lq[4 * i + 1] = a.y; lq[4 * i + 2] = a.z; lq[4 * i + 3] = a.w;
            dq_[4 * i] = c.x; dq_[4 * i + 1] = c.y; dq_[4 * i + 2] = c.z; dq_[4 * i + 3] = c.w;
          }
        }
        tmem_ld_wait();
        uint32_t pw[8], dw[8];
        // no masking needed for THIS WARP's 32 kv rows x 16 q columns (the tile-level test sent the lower half of the kv tile down the
        // masked path on the second diagonal q tile although all of it is visible)
        const bool full_tile = (qc0 + 15 < it.len) && (it.kv0 + wq * 32 + 31 < it.len_kv) && (!kCausal || it.kv0 + wq * 32 + 31 <= qc0);
        if (full_tile) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float p0 = exp2f(__uint_as_float(sv[2 * c]) * scale_log2 - lq[2 * c]);
            const float p1 = exp2f(__uint_as_float(sv[2 * c + 1]) * scale_log2 - lq[2 * c + 1]);
            pw[c] = pack_bf16(p0, p1);
            dw[c] = pack_bf16(p0 * (__uint_as_float(dv[2 * c]) - dq_[2 * c]), p1 * (__uint_as_float(dv[2 * c + 1]) - dq_[2 * c + 1]));
          }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float pe[2], de[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int cc = 2 * c + e;
              const int qi = qc0 + cc;
              const bool ok = (qi < it.len) && (kv_row < it.len_kv) && (!kCausal || kv_row <= qi);
              pe[e] = ok ? exp2f(__uint_as_float(sv[cc]) * scale_log2 - lq[cc]) : 0.f;
              de[e] = ok ? pe[e] * (__uint_as_float(dv[cc]) - dq_[cc]) : 0.f;
            }
            pw[c] = pack_bf16(pe[0], pe[1]);
            dw[c] = pack_bf16(de[0], de[1]);
          }
        }
        tmem_st8(tmem_St + lane_off + st * 64 + part * 16, pw);
        tmem_st8(tmem_dPt + lane_off + st * 64 + part * 16, dw);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pds_full[st]);
      }
      // epilogue: job `part` = one 64-column chunk of dV or dK, staged in the ring stages of the item's last two q tiles
      const uint32_t gl = g0 + it.n_it - 1;
      mbar_wait(&acc_done[gl & 1], (gl >> 1) & 1, 46);
      tc_fence_after();
      if (my_store) {
        smem_count_wait_acquire(stored_cnt, 4u * kJobs * ic);   // every warp's bulk stores of the earlier items have read their staging
        const uint32_t sA = gl & 3, sB = (gl + 3) & 3;
        uint8_t* stage;
        if constexpr (NCH == 2) {   // four 16 KB chunks: {Q part, dO part} of stage sA, then of stage sB
          stage = smem + ((part & 1) ? L::oDO : L::oQ) + ((part >> 1) ? sB : sA) * L::kQ + wq * 4096;
        } else {                    // two 16 KB chunks, each = the 8 KB Q part + 8 KB dO part of one stage
          stage = smem + ((wq >> 1) ? L::oDO : L::oQ) + (part ? sB : sA) * L::kQ + (wq & 1) * 4096;
        }
        store_acc_rows64((acc == 0 ? tmem_dV : tmem_dK) + lane_off + ch * 64, stage, acc == 0 ? 1.f : scale, acc == 0 ? &tdv : &tdk,
                         it.h * D + ch * 64, it.kv0 + wq * 32, it.b, lane);
        if (lane == 0) smem_count_add_release(stored_cnt);
      }
      tc_fence_before();
      g0 += it.n_it;
      ++ic;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kBwdRowWarps + 1) tmem_dealloc<1>(tmem_base, 512);
}

template <int D, bool kCausal>
__global__ void __launch_bounds__(kBwdTsThreads, 1)
attn_bwd_dq_persist_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                           const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                           const __grid_constant__ CUtensorMap tdq, const float* __restrict__ lse2, const float* __restrict__ delta,
                           const int* __restrict__ seqlens, BwdOut go, int B, int S, int Skv, int S_pad, int nh, float scale,
                           float scale_log2, int* __restrict__ item_counter) {
  using L = BwdQTsSmem<D>;
  constexpr int NCH = L::NCH;
  constexpr int R4 = kBwdRing;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::oBar);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;                 // the item's last S / dP tile-GEMMs have retired: Q / dO may be refilled
  uint64_t* kv_full = bars + 2;                 // [4]  ring over ALL kv tiles this CTA processes
  uint64_t* kv_empty = kv_full + R4;            // [4]
  uint64_t* sdp_full = kv_empty + R4;           // [2]
  uint64_t* ds_full = sdp_full + 2;             // [2]
  uint64_t* acc_done = ds_full + 2;             // [2]
  uint64_t* it_full = acc_done + 2;             // [2]
  uint64_t* it_empty = it_full + 2;             // [2]
  uint32_t* item_ring = reinterpret_cast<uint32_t*>(it_empty + 2);
  uint32_t* tmem_ptr = item_ring + 2;
  uint32_t* stored_cnt = tmem_ptr + 1;          // storing warps whose bulk store has read its staging (4 * NCH per non-empty item)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = (S + 127) / 128, n_hb = nh * B, n_items = nq * n_hb;
  const int win_heads = max(1, min(n_hb, (2 * static_cast<int>(gridDim.x) + nq - 1) / nq));

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) { printf("attn_bwd_dq_persist: smem misaligned\n"); __trap(); }
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < R4; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sdp_full[i], 1); mbar_init(&acc_done[i], 1);
      mbar_init(&ds_full[i], kBwdRowWarps);
      mbar_init(&it_full[i], 1);
      mbar_init(&it_empty[i], kBwdRowWarps + 1);
    }
    *stored_cnt = 0;
    fence_mbar_init();
  }
  if (warp == kBwdRowWarps + 1) { tmem_alloc<1>(tmem_ptr, 512); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 128, tmem_dQ = tmem_base + 256;

  struct Item { int q0, h, b, len, len_kv, n_kv; };
  auto item = [&](int w) {
    const ItemId id = decode_item(w, nq, n_hb, win_heads, kCausal);   // causal: the last q tile sees every kv tile -> heaviest first
    Item it;
    it.q0 = id.tile * 128; it.h = id.hb % nh; it.b = id.hb / nh;
    it.len = seqlens ? min(seqlens[it.b], S) : S;
    it.len_kv = seqlens ? it.len : Skv;
    const int kv_end = kCausal ? min(it.len_kv, it.q0 + 128) : it.len_kv;
    it.n_kv = (it.q0 < it.len) ? (kv_end + 63) / 64 : 0;
    return it;
  };
  auto next_item = [&](uint32_t n, bool one_lane_arrives) {
    mbar_wait(&it_full[n & 1], (n >> 1) & 1, 59);
    const int w = static_cast<int>(item_ring[n & 1]);
    if (one_lane_arrives) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&it_empty[n & 1]);
    } else {
      mbar_arrive(&it_empty[n & 1]);
    }
    return w;
  };

  if (warp == kBwdRowWarps) { if (elect_one_sync()) {
    // ---------------- TMA producer + item scheduler ----------------
    tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); tma_prefetch_desc(&tdo);
    uint32_t kg = 0, ic = 0, hold = 0;
    for (uint32_t n = 0;; ++n) {
      int w = atomicAdd(item_counter, 1);
      if (w > n_items) w = n_items;
      mbar_wait(&it_empty[n & 1], ((n >> 1) & 1) ^ 1, 59);
      item_ring[n & 1] = static_cast<uint32_t>(w);
      mbar_arrive(&it_full[n & 1]);
      if (w >= n_items) break;
      const Item it = item(w);
      if (it.n_kv == 0) continue;
      mbar_wait(q_empty, (ic & 1) ^ 1, 50);
      mbar_arrive_expect_tx(q_full, 2 * 128 * D * 2);
      for (int c = 0; c < NCH; ++c) {
        tma_load_3d(smem + L::oQ + c * 16384, &tq, q_full, it.h * D + c * 64, it.q0, it.b);
        tma_load_3d(smem + L::oDO + c * 16384, &tdo, q_full, it.h * D + c * 64, it.q0, it.b);
      }
      for (int j = 0; j < it.n_kv; ++j) {
        const uint32_t rs = kg & 3;
        if ((hold >> rs) & 1u) { smem_count_wait_acquire(stored_cnt, 4u * NCH * ic); hold = 0; }
        mbar_wait(&kv_empty[rs], ((kg >> 2) & 1) ^ 1, 50);
        mbar_arrive_expect_tx(&kv_full[rs], 2 * 64 * D * 2);
        for (int c = 0; c < NCH; ++c) {
          tma_load_3d(smem + L::oK + rs * L::kKV + c * 8192, &tk, &kv_full[rs], it.h * D + c * 64, j * 64, it.b);
          tma_load_3d(smem + L::oV + rs * L::kKV + c * 8192, &tv, &kv_full[rs], it.h * D + c * 64, j * 64, it.b);
        }
        ++kg;
      }
      hold |= 1u << ((kg + 3) & 3);   // the stage of the item's last kv tile becomes its dQ staging area
      ++ic;
    }
  } } else if (warp == kBwdRowWarps + 1) { if (elect_one_sync()) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
    constexpr uint32_t idesc_acc = make_idesc_bf16(128, D, false, true);
    const uint32_t sQ = smem_u32(smem + L::oQ), sDO = smem_u32(smem + L::oDO), sK = smem_u32(smem + L::oK),
                   sV = smem_u32(smem + L::oV);
    uint32_t g0 = 0, ic = 0;
    for (uint32_t n = 0;; ++n) {
      const int w = next_item(n, false);
      if (w >= n_items) break;
      const Item it = item(w);
      if (it.n_kv == 0) continue;
      auto issue_s = [&](int j) {
        const uint32_t g = g0 + j, st = g & 1, rs = g & 3;
        mbar_wait(&kv_full[rs], (g >> 2) & 1, 51);
        tc_fence_after();
        mma_tile<D / 16>(tmem_S + st * 64, sQ, false, 16384, sK + rs * L::kKV, false, 8192, idesc_s, false);
        mma_tile<D / 16>(tmem_dP + st * 64, sDO, false, 16384, sV + rs * L::kKV, false, 8192, idesc_s, false);
        umma_commit(&sdp_full[st]);
        if (j == it.n_kv - 1) umma_commit(q_empty);
      };
      mbar_wait(q_full, ic & 1, 52);
      issue_s(0);
      for (int j = 0; j < it.n_kv; ++j) {
        if (j + 1 < it.n_kv) issue_s(j + 1);
        const uint32_t g = g0 + j, st = g & 1, rs = g & 3;
        mbar_wait(&ds_full[st], (g >> 1) & 1, 53);
        tc_fence_after();
        mma_tile_ts<4>(tmem_dQ, tmem_dP + st * 64, 16, sK + rs * L::kKV, true, 8192, idesc_acc, j > 0);   // dQ (+)= dS K_j, dS from TMEM
        umma_commit(&kv_empty[rs]);
        umma_commit(&acc_done[st]);
      }
      g0 += it.n_kv;
      ++ic;
    }
  } } else if (warp < kBwdRowWarps) {
    const int wq = warp & 3, part = warp >> 2;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const bool my_store = part < NCH;
    uint32_t g0 = 0, ic = 0;
    for (uint32_t n = 0;; ++n) {
      const int w = next_item(n, true);
      if (w >= n_items) break;
      const Item it = item(w);
      const int q_row = it.q0 + row;
      if (it.n_kv == 0) {   // padding rows of a right-padded batch: zeros
        if (my_store && q_row < S) zero_row64(go.dq + (static_cast<size_t>(it.b) * S + q_row) * go.ld_dq + it.h * D + part * 64);
        continue;
      }
      const size_t sidx = (static_cast<size_t>(it.b) * nh + it.h) * S_pad + min(q_row, S_pad - 1);
      const float my_lse = lse2[sidx], my_delta = delta[sidx];
      const bool row_ok = q_row < it.len;
      for (int j = 0; j < it.n_kv; ++j) {
        const uint32_t g = g0 + j, st = g & 1;
        const int kc0 = j * 64 + part * 16;
        mbar_wait(&sdp_full[st], (g >> 1) & 1, 54);
        tc_fence_after();
        // causal diagonal per warp: all 16 kv columns lie beyond the last q row of this warp -> dS = 0 without loads or exponentials
        if (kCausal && kc0 > it.q0 + wq * 32 + 31) {
          uint32_t z[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) z[c] = 0u;
          tmem_st8(tmem_dP + lane_off + st * 64 + part * 16, z);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ds_full[st]);
          continue;
        }
        uint32_t sv[16], dv[16];
        tmem_ld16(tmem_S + lane_off + st * 64 + part * 16, sv);
        tmem_ld16(tmem_dP + lane_off + st * 64 + part * 16, dv);
        tmem_ld_wait();
        uint32_t dw[8];
        const bool full_tile = (it.q0 + wq * 32 + 31 < it.len) && (kc0 + 15 < it.len_kv) && (!kCausal || kc0 + 15 <= it.q0 + wq * 32);   // per warp
        if (full_tile) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float p0 = exp2f(__uint_as_float(sv[2 * c]) * scale_log2 - my_lse);
            const float p1 = exp2f(__uint_as_float(sv[2 * c + 1]) * scale_log2 - my_lse);
            dw[c] = pack_bf16(p0 * (__uint_as_float(dv[2 * c]) - my_delta), p1 * (__uint_as_float(dv[2 * c + 1]) - my_delta));
          }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float de[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int cc = 2 * c + e;
              const int kvi = kc0 + cc;
              const bool ok = row_ok && (kvi < it.len_kv) && (!kCausal || kvi <= q_row);
              const float pe = ok ? exp2f(__uint_as_float(sv[cc]) * scale_log2 - my_lse) : 0.f;
              de[e] = ok ? pe * (__uint_as_float(dv[cc]) - my_delta) : 0.f;
            }
            dw[c] = pack_bf16(de[0], de[1]);
          }
        }
        tmem_st8(tmem_dP + lane_off + st * 64 + part * 16, dw);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ds_full[st]);
      }
      // epilogue: chunk `part` of dQ, staged in the ring stage of the item's last kv tile
      const uint32_t gl = g0 + it.n_kv - 1;
      mbar_wait(&acc_done[gl & 1], (gl >> 1) & 1, 56);
      tc_fence_after();
      if (my_store) {
        smem_count_wait_acquire(stored_cnt, 4u * NCH * ic);
        const uint32_t sA = gl & 3;
        uint8_t* stage;
        if constexpr (NCH == 2) stage = smem + (part ? L::oV : L::oK) + sA * L::kKV + wq * 4096;         // 16 KB chunk = one K or V stage
        else stage = smem + ((wq >> 1) ? L::oV : L::oK) + sA * L::kKV + (wq & 1) * 4096;                 // 16 KB chunk = 8 KB K + 8 KB V stage
        store_acc_rows64(tmem_dQ + lane_off + part * 64, stage, scale, &tdq, it.h * D + part * 64, it.q0 + wq * 32, it.b, lane);
        if (lane == 0) smem_count_add_release(stored_cnt);
      }
      tc_fence_before();
      g0 += it.n_kv;
      ++ic;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kBwdRowWarps + 1) tmem_dealloc<1>(tmem_base, 512);
}

// ================================================================================================ host
template <typename K>
static int set_smem(K kern, int bytes) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess ? 0 : DLLM_ERR_LAUNCH;
}

// DLLM_ATTN_LEGACY=1 selects the round-1 data paths (P / dS staged through shared memory) for same-box A/B runs
static bool attn_legacy() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DLLM_ATTN_LEGACY");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

template <int D, bool C, bool PT>
static int launch_fwd_pt(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, float* lse,
                         const int* seqlens, int B, int S, int Skv, int nh, float scale_log2, cudaStream_t st, const uint8_t* kv_mask,
                         int mask_ld) {
  auto kern = attn_fwd_kernel<D, C, PT>;
  static bool once = false;
  if (!once) {
    if (set_smem(kern, FwdSmem<D, PT>::kBytes)) return DLLM_ERR_LAUNCH;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    once = true;
  }
  dim3 grid((S + 127) / 128, nh, B);
  kern<<<grid, kAttnThreads, FwdSmem<D, PT>::kBytes, st>>>(tq, tk, tv, to, lse, seqlens, S, Skv, nh, scale_log2, kv_mask, mask_ld);
  return cudaGetLastError() == cudaSuccess ? 0 : DLLM_ERR_LAUNCH;
}
// DLLM_ATTN_NONPERSIST=1: one CTA per (q tile, head, batch) item (the r02 kernel before persistence) for same-box A/B runs
static bool attn_nonpersist() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DLLM_ATTN_NONPERSIST");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
template <int D, bool C>
static int launch_fwd_persist(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, void* out,
                              long ld_o, float* lse,
                              const int* seqlens, int B, int S, int Skv, int nh, float scale_log2, cudaStream_t st,
                              const uint8_t* kv_mask, int mask_ld) {
  auto kern = attn_fwd_persist_kernel<D, C>;
  static bool once = false;
  if (!once) {
    if (set_smem(kern, FwdSmem<D, true>::kBytes)) return DLLM_ERR_LAUNCH;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    once = true;
  }
  const long items = static_cast<long>((S + 127) / 128) * nh * B;
  const long slots = 2L * num_sms();
  const unsigned grid = static_cast<unsigned>(items < slots ? items : slots);
  int* counter = tile_counter_slot(st);
  if (!counter) return DLLM_ERR_LAUNCH;
  kern<<<grid, kAttnThreads, FwdSmem<D, true>::kBytes, st>>>(tq, tk, tv, to, static_cast<bf16*>(out), ld_o, lse, seqlens, B, S, Skv, nh,
                                                              scale_log2, kv_mask, mask_ld, counter);
  return cudaGetLastError() == cudaSuccess ? 0 : DLLM_ERR_LAUNCH;
}
template <int D, bool C>
static int launch_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, void* out, long ld_o,
                      float* lse, const int* seqlens, int B, int S, int Skv, int nh, float scale_log2, cudaStream_t st,
                      const uint8_t* kv_mask = nullptr, int mask_ld = 0) {
  if (attn_legacy()) return launch_fwd_pt<D, C, false>(tq, tk, tv, to, lse, seqlens, B, S, Skv, nh, scale_log2, st, kv_mask, mask_ld);
  if (attn_nonpersist()) return launch_fwd_pt<D, C, true>(tq, tk, tv, to, lse, seqlens, B, S, Skv, nh, scale_log2, st, kv_mask, mask_ld);
  return launch_fwd_persist<D, C>(tq, tk, tv, to, out, ld_o, lse, seqlens, B, S, Skv, nh, scale_log2, st, kv_mask, mask_ld);
}

int attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, const int* seqlens, int B, int S, int nh,
             int d, long ld_qkv, long ld_o, int causal, float scale, cudaStream_t st) {
  return attn_fwd_ex(q, k, v, out, lse, seqlens, B, S, S, nh, d, ld_qkv, ld_qkv, ld_o, causal, scale, st);
}

// q: [B, S, nh, d] (token stride ld_q); k, v: [B, Skv, nh, d] (token stride ld_kv).  Skv != S = cross-attention.
int attn_fwd_ex(const void* q, const void* k, const void* v, void* out, float* lse, const int* seqlens, int B, int S, int Skv,
                int nh, int d, long ld_q, long ld_kv, long ld_o, int causal, float scale, cudaStream_t st) {
  return attn_fwd_cache(q, k, v, out, lse, seqlens, B, S, Skv, Skv, nh, d, ld_q, ld_kv, ld_o, causal, scale, st);
}

// kv_rows = rows allocated per batch entry in the k/v buffers (>= Skv): lets a preallocated kv-cache [B, max_len, nh*d] be read in
// place with Skv valid rows.  causal with Skv > S uses the bottom-right aligned mask (decode / continuation).
int attn_fwd_cache(const void* q, const void* k, const void* v, void* out, float* lse, const int* seqlens, int B, int S, int Skv,
                   int kv_rows, int nh, int d, long ld_q, long ld_kv, long ld_o, int causal, float scale, cudaStream_t st) {
  return attn_fwd_cache_mask(q, k, v, out, lse, seqlens, nullptr, 0, B, S, Skv, kv_rows, nh, d, ld_q, ld_kv, ld_o, causal, scale, st);
}

// + kv_mask [B, mask_ld] bytes (0 = padded key): a padded prompt batch inside the kv-cache.  mask_ld must cover whole 64-key tiles
// (mask_ld % 16 == 0, mask_ld >= round_up(Skv, 64)) because the kernel fetches the mask 64 bytes at a time.
int attn_fwd_cache_mask(const void* q, const void* k, const void* v, void* out, float* lse, const int* seqlens, const void* kv_mask,
                        int mask_ld, int B, int S, int Skv, int kv_rows, int nh, int d, long ld_q, long ld_kv, long ld_o, int causal,
                        float scale, cudaStream_t st) {
  if (B <= 0 || S <= 0 || Skv <= 0 || nh <= 0 || kv_rows < Skv) return DLLM_ERR_SHAPE;
  if (kv_mask && ((mask_ld & 15) || mask_ld < (Skv + 63) / 64 * 64 || (reinterpret_cast<uintptr_t>(kv_mask) & 15))) return DLLM_ERR_ALIGN;
  const uint8_t* km = static_cast<const uint8_t*>(kv_mask);
  if (d != 128 && d != 64) return DLLM_ERR_UNSUPPORTED;
  if (S != Skv && seqlens) return DLLM_ERR_UNSUPPORTED;
  if (causal && Skv < S) return DLLM_ERR_SHAPE;
  CUtensorMap tq, tk, tv, to;
  int rc;
  if ((rc = make_tmap_bsc(&tq, q, B, S, nh * d, ld_q, 128))) return rc;
  if ((rc = make_tmap_bsc(&tk, k, B, kv_rows, nh * d, ld_kv, 64))) return rc;
  if ((rc = make_tmap_bsc(&tv, v, B, kv_rows, nh * d, ld_kv, 64))) return rc;
  if ((rc = make_tmap_bsc(&to, out, B, S, nh * d, ld_o, 32))) return rc;
  const float sl2 = scale * kLog2e;
  if (d == 128) return causal ? launch_fwd<128, true>(tq, tk, tv, to, out, ld_o, lse, seqlens, B, S, Skv, nh, sl2, st, km, mask_ld)
                              : launch_fwd<128, false>(tq, tk, tv, to, out, ld_o, lse, seqlens, B, S, Skv, nh, sl2, st, km, mask_ld);
  return causal ? launch_fwd<64, true>(tq, tk, tv, to, out, ld_o, lse, seqlens, B, S, Skv, nh, sl2, st, km, mask_ld)
                : launch_fwd<64, false>(tq, tk, tv, to, out, ld_o, lse, seqlens, B, S, Skv, nh, sl2, st, km, mask_ld);
}

static inline int s_pad(int S) { return (S + 63) / 64 * 64; }
size_t attn_bwd_workspace(int B, int S, int nh, int) { return static_cast<size_t>(B) * s_pad(S) * nh * 2 * sizeof(float); }

template <int D, bool C>
static int launch_bwd(const CUtensorMap& tq64, const CUtensorMap& tq128, const CUtensorMap& tk64, const CUtensorMap& tk128,
                      const CUtensorMap& tv64, const CUtensorMap& tv128, const CUtensorMap& tdo64,
                      const CUtensorMap& tdo128, const CUtensorMap& tdq, const CUtensorMap& tdk, const CUtensorMap& tdv,
                      const bf16* dout, const bf16* out, const float* lse, float* delta, float* lse2, const int* seqlens,
                      int B, int S, int Skv, int nh, long ld_o, float scale, cudaStream_t st, const BwdOut& go) {
  auto k1 = attn_bwd_dkdv_kernel<D, C>;
  auto k2 = attn_bwd_dq_kernel<D, C>;
  auto k1t = attn_bwd_dkdv_ts_kernel<D, C>;
  auto k2t = attn_bwd_dq_ts_kernel<D, C>;
  auto k1p = attn_bwd_dkdv_persist_kernel<D, C>;
  auto k2p = attn_bwd_dq_persist_kernel<D, C>;
  static bool once = false;
  if (!once) {
    if (set_smem(k1, BwdKVSmem<D>::kBytes) || set_smem(k2, BwdQSmem<D>::kBytes)) return DLLM_ERR_LAUNCH;
    if (set_smem(k1t, BwdKVTsSmem<D>::kBytes) || set_smem(k2t, BwdQTsSmem<D>::kBytes)) return DLLM_ERR_LAUNCH;
    if (set_smem(k1p, BwdKVTsSmem<D>::kBytes) || set_smem(k2p, BwdQTsSmem<D>::kBytes)) return DLLM_ERR_LAUNCH;
    once = true;
  }
  const int Sp = s_pad(S);
  if (nh > 384) return DLLM_ERR_SHAPE;   // prep kernel parks nh x 32 floats in (default-limit) shared memory
  attn_bwd_prep_kernel<D><<<static_cast<unsigned>(static_cast<long>(B) * (Sp / 32)), 256, static_cast<size_t>(nh) * 32 * sizeof(float), st>>>(
      dout, out, lse, delta, lse2, B, S, Sp, nh, ld_o);
  dim3 grid_kv((Skv + 127) / 128, nh, B), grid_q((S + 127) / 128, nh, B);
  if (attn_legacy()) {
    k1<<<grid_kv, kBwdThreads, BwdKVSmem<D>::kBytes, st>>>(tq64, tk128, tv128, tdo64, tdk, tdv, lse2, delta, seqlens, S, Skv, Sp, nh,
                                                            scale, scale * kLog2e);
    k2<<<grid_q, kBwdThreads, BwdQSmem<D>::kBytes, st>>>(tq128, tk64, tv64, tdo128, tdq, lse2, delta, seqlens, S, Skv, Sp, nh, scale,
                                                          scale * kLog2e);
  } else if (!attn_nonpersist()) {
    const long items_kv = static_cast<long>(grid_kv.x) * nh * B, items_q = static_cast<long>(grid_q.x) * nh * B;
    const long sms = num_sms();
    int* c1 = tile_counter_slot(st);
    int* c2 = tile_counter_slot(st);
    if (!c1 || !c2) return DLLM_ERR_LAUNCH;
    k1p<<<static_cast<unsigned>(items_kv < sms ? items_kv : sms), kBwdTsThreads, BwdKVTsSmem<D>::kBytes, st>>>(
        tq64, tk128, tv128, tdo64, tdk, tdv, lse2, delta, seqlens, go, B, S, Skv, Sp, nh, scale, scale * kLog2e, c1);
    k2p<<<static_cast<unsigned>(items_q < sms ? items_q : sms), kBwdTsThreads, BwdQTsSmem<D>::kBytes, st>>>(
        tq128, tk64, tv64, tdo128, tdq, lse2, delta, seqlens, go, B, S, Skv, Sp, nh, scale, scale * kLog2e, c2);
  } else {
    k1t<<<grid_kv, kBwdTsThreads, BwdKVTsSmem<D>::kBytes, st>>>(tq64, tk128, tv128, tdo64, tdk, tdv, lse2, delta, seqlens, S, Skv, Sp,
                                                                 nh, scale, scale * kLog2e);
    k2t<<<grid_q, kBwdTsThreads, BwdQTsSmem<D>::kBytes, st>>>(tq128, tk64, tv64, tdo128, tdq, lse2, delta, seqlens, S, Skv, Sp, nh,
                                                               scale, scale * kLog2e);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : DLLM_ERR_LAUNCH;
}

int attn_bwd(const void* dout, const void* q, const void* k, const void* v, const void* out, const float* lse, void* dq,
             void* dk, void* dv, const int* seqlens, void* workspace, size_t workspace_bytes, int B, int S, int nh, int d,
             long ld_qkv, long ld_o, long ld_dqkv, int causal, float scale, cudaStream_t st) {
  return attn_bwd_ex(dout, q, k, v, out, lse, dq, dk, dv, seqlens, workspace, workspace_bytes, B, S, S, nh, d, ld_qkv, ld_qkv, ld_o,
                     ld_dqkv, ld_dqkv, causal, scale, st);
}

int attn_bwd_ex(const void* dout, const void* q, const void* k, const void* v, const void* out, const float* lse, void* dq,
                void* dk, void* dv, const int* seqlens, void* workspace, size_t workspace_bytes, int B, int S, int Skv, int nh,
                int d, long ld_q, long ld_kv, long ld_o, long ld_dq, long ld_dkv, int causal, float scale, cudaStream_t st) {
  if (B <= 0 || S <= 0 || Skv <= 0 || nh <= 0) return DLLM_ERR_SHAPE;
  if (d != 128 && d != 64) return DLLM_ERR_UNSUPPORTED;
  if (S != Skv && (causal || seqlens)) return DLLM_ERR_UNSUPPORTED;
  if (workspace_bytes < attn_bwd_workspace(B, S, nh, d)) return DLLM_ERR_SHAPE;
  float* delta = static_cast<float*>(workspace);
  float* lse2 = delta + static_cast<size_t>(B) * s_pad(S) * nh;
  CUtensorMap tq64, tq128, tk64, tk128, tv64, tv128, tdo64, tdo128, tdq, tdk, tdv;
  int rc;
  const int C = nh * d;
  if ((rc = make_tmap_bsc(&tq64, q, B, S, C, ld_q, 64))) return rc;
  if ((rc = make_tmap_bsc(&tq128, q, B, S, C, ld_q, 128))) return rc;
  if ((rc = make_tmap_bsc(&tk64, k, B, Skv, C, ld_kv, 64))) return rc;
  if ((rc = make_tmap_bsc(&tk128, k, B, Skv, C, ld_kv, 128))) return rc;
  if ((rc = make_tmap_bsc(&tv64, v, B, Skv, C, ld_kv, 64))) return rc;
  if ((rc = make_tmap_bsc(&tv128, v, B, Skv, C, ld_kv, 128))) return rc;
  if ((rc = make_tmap_bsc(&tdo64, dout, B, S, C, ld_o, 64))) return rc;
  if ((rc = make_tmap_bsc(&tdo128, dout, B, S, C, ld_o, 128))) return rc;
  if ((rc = make_tmap_bsc(&tdq, dq, B, S, C, ld_dq, 32))) return rc;
  if ((rc = make_tmap_bsc(&tdk, dk, B, Skv, C, ld_dkv, 32))) return rc;
  if ((rc = make_tmap_bsc(&tdv, dv, B, Skv, C, ld_dkv, 32))) return rc;
  const BwdOut go{static_cast<bf16*>(dq), static_cast<bf16*>(dk), static_cast<bf16*>(dv), ld_dq, ld_dkv};
#define DLLM_BWD(DD, CC)                                                                                             \
  return launch_bwd<DD, CC>(tq64, tq128, tk64, tk128, tv64, tv128, tdo64, tdo128, tdq, tdk, tdv, (const bf16*)dout,     \
                            (const bf16*)out, lse, delta, lse2, seqlens, B, S, Skv, nh, ld_o, scale, st, go)
  if (d == 128) { if (causal) DLLM_BWD(128, true); else DLLM_BWD(128, false); }
  if (causal) DLLM_BWD(64, true); else DLLM_BWD(64, false);
#undef DLLM_BWD
}

// host-side mirror of decode_item (tests / tooling): item w of a grid of `grid` CTAs -> (tile, head*batch index).  Kept as a separate
// host function with the same arithmetic so that the device translation unit is exactly the one validated on hardware.
void attn_item_order(int w, int ntiles, int n_hb, int grid, int descending, int* tile, int* hb, int* win_heads_out) {
  int win_heads = (2 * grid + ntiles - 1) / ntiles;
  if (win_heads > n_hb) win_heads = n_hb;
  if (win_heads < 1) win_heads = 1;
  const int wsz = win_heads * ntiles, n_full = n_hb / win_heads, full_items = n_full * wsz;
  int win, idx, R;
  if (w < full_items) { win = w / wsz; idx = w - win * wsz; R = win_heads; }
  else { win = n_full; idx = w - full_items; R = n_hb - n_full * win_heads; }
  const int tpos = idx / R;
  *tile = descending ? ntiles - 1 - tpos : tpos;
  *hb = win * win_heads + idx - tpos * R;
  if (win_heads_out) *win_heads_out = win_heads;
}

}  // namespace dllm

#ifdef DLLM_ATTN_TRACE
extern "C" int dllm_attn_trace_read(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, dllm::g_attn_trace, sizeof(long long) * n);
}
extern "C" int dllm_attn_cta_log_read(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, dllm::g_attn_cta_log, sizeof(long long) * n);
}
#endif

/* libdreamllm_sm100.so — C ABI of the B200-native DreamLLM hot path.
 *
 * The reference (RunpeiDong/DreamLLM) is pure Python and has no FFI of its own; these entry points
 * are what a ctypes / cffi binding inside `omni/models/dreamllm` calls in place of the torch ops the reference
 * issues.  Each function cites the reference call site it replaces
 * (paths relative to /root/reference/omni/models/dreamllm/).
 *
 * Conventions: plain device pointers + sizes, no torch types; bf16 unless noted; row-major; `ld*` = row stride in
 * ELEMENTS; every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates nothing, never
 * synchronises; returns 0 or a negative DLLM_ERR_* code.  Thread-safe for distinct streams.
 */
#ifndef DREAMLLM_SM100_H
#define DREAMLLM_SM100_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define DLLM_OK 0
#define DLLM_ERR_SHAPE (-1)
#define DLLM_ERR_ALIGN (-2)
#define DLLM_ERR_DRIVER (-3)
#define DLLM_ERR_TMAP (-4)
#define DLLM_ERR_LAUNCH (-5)
#define DLLM_ERR_UNSUPPORTED (-6)

int dllm_version(void);
const char* dllm_error_string(int code);
/* Leave `n` SMs out of the persistent GEMM grids (for an overlapped NCCL all-reduce; data-parallel training only). */
int dllm_set_reserved_sms(int n);
int dllm_get_reserved_sms(void);

/* Dense contraction on tcgen05 tensor cores: C[M,N] = op(A) * op(B), fp32 accumulate in TMEM.
 *   a_mn = 0: A is [M,K] row-major; a_mn = 1: A is stored [K,M] row-major (i.e. A^T, used for wgrad)
 *   b_mn = 0: B is [N,K] row-major (an nn.Linear weight); b_mn = 1: B is stored [K,N] row-major
 *   out_fp32: C dtype (0 = bf16, 1 = fp32).  cta_pair: -1 auto, 0 = one CTA per tile, 1 = cta_group::2 pairs.
 * Replaces F.linear / nn.Linear and their autograd dgrad/wgrad:
 *   modeling_dreamllm.py:336-338, :395 (q/k/v/o_proj), :237 (gate/up/down_proj), :1452 (lm_head),
 *   omni/models/projector/mlp_projector.py:11-50 (projectors). */
int dllm_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, long lda, long ldb, long ldc,
                   int a_mn, int b_mn, int out_fp32, int cta_pair, void* stream);

/* DreamLLMRMSNorm.forward (modeling_dreamllm.py:86-91), optionally fused with the preceding residual add
 * (:638, :644): if `add` != NULL, x_out = bf16(x + add) is written and normalised.  rstd[T] (fp32) is saved for bwd. */
int dllm_rmsnorm_fwd(const void* x, const void* add, const void* weight, void* x_out, void* y, float* rstd, int T,
                     int H, float eps, void* stream);
size_t dllm_rmsnorm_bwd_workspace_bytes(int T, int H);
/* dx = d(norm)/dx (+ dres if != NULL); dweight (=|+=) column sums (bf16). */
int dllm_rmsnorm_bwd(const void* dy, const void* x, const void* weight, const float* rstd, const void* dres, void* dx,
                     void* dweight, int dweight_accumulate, void* workspace, size_t workspace_bytes, int T, int H,
                     void* stream);

/* apply_rotary_pos_emb (modeling_dreamllm.py:184-209) in place on `heads_total` heads laid side by side in a
 * [T, ld] buffer (q and k blocks of the fused qkv projection); cos/sin tables [max_pos, head_dim] bf16 (:126-127),
 * pos[T] int32.  mode = +1 forward, -1 backward. */
int dllm_rope_inplace(void* buf, const void* cos_table, const void* sin_table, const int* pos, long ld, int T,
                      int heads_total, int head_dim, int mode, void* stream);

/* DreamLLMMLP activation (modeling_dreamllm.py:237): act = silu(gate) * up on the fused [T, 2I] gate|up buffer. */
int dllm_swiglu_fwd(const void* gate_up, void* act, long ld_gate_up, int T, int I, void* stream);
int dllm_swiglu_bwd(const void* dact, const void* gate_up, void* dgate_up, long ld_gate_up, int T, int I, void* stream);

int dllm_add_bf16(const void* a, const void* b, void* out, long n, void* stream);

/* Shifted masked-mean cross entropy (modeling_dreamllm.py:1453-1470). labels[T] int64, already shifted, -100 =
 * ignore. loss[1] fp32.  If write_grad, logits are overwritten in place by dloss * d(loss)/d(logits) (bf16).
 * workspace: (T + 2) floats. */
int dllm_cross_entropy(void* logits, const long long* labels, float* loss, float dloss, void* workspace, long ld,
                       int T, int V, int write_grad, void* stream);

/* embed_tokens lookup (modeling_dreamllm.py:1066-1067) and its deterministic gradient (sorted segment sums). */
int dllm_embedding_fwd(const long long* ids, const void* weight, void* out, int T, int H, void* stream);
int dllm_embedding_bwd(const long long* sorted_ids, const long long* order, const void* dy, void* dweight, int T,
                       int H, int accumulate, void* stream);

/* Causal / full flash attention on tcgen05 (replaces flash_attn_func / flash_attn_varlen_func,
 * modeling_dreamllm.py:500-551, and the eager path :357-379).
 * q,k,v: [B, S, nh, d] views with row stride ld_qkv (elements) between consecutive tokens — the fused qkv GEMM
 * output is consumed directly, no transposes.  out: [B, S, nh*d] (ld_o).  lse: [B, nh, S] fp32.
 * seqlens: int32[B] valid (right-padded) lengths or NULL.  Rows >= seqlens[b] are written as zeros (pad_input). */
int dllm_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, const int* seqlens, int B,
                  int S, int nh, int d, long ld_qkv, long ld_o, int causal, float scale, void* stream);
size_t dllm_attn_bwd_workspace_bytes(int B, int S, int nh, int d);
int dllm_attn_bwd(const void* dout, const void* q, const void* k, const void* v, const void* out, const float* lse,
                  void* dq, void* dk, void* dv, const int* seqlens, void* workspace, size_t workspace_bytes, int B,
                  int S, int nh, int d, long ld_qkv, long ld_o, long ld_dqkv, int causal, float scale, void* stream);

/* GEMM with fused epilogue: out = act(bf16(acc + bias[col])) (+ residual[row, col]); act 0 none, 1 quick_gelu
 * (CLIP MLP, transformers activations.QuickGELU), 2 gelu-erf (MLPProjector, mlp_projector.py:30-50), 3 silu.
 * Replaces nn.Linear(bias=True) + activation + residual add in the CLIP tower / projectors / UNet transformer blocks. */
int dllm_gemm_bf16_ex(const void* A, const void* B, void* C, int M, int N, int K, long lda, long ldb, long ldc,
                      int a_mn, int b_mn, int out_fp32, int cta_pair, const void* bias, const void* residual, long ldr,
                      int act, void* stream);

/* nn.LayerNorm forward (CLIP layer_norm1/2 + pre_layrnorm, modeling_plugins.py:321 -> transformers CLIPVisionModel). */
int dllm_layernorm_fwd(const void* x, const void* weight, const void* bias, void* y, int T, int H, float eps, void* stream);

/* CLIP patch embedding (Conv2d k=s=patch, no bias) as unfold + GEMM, and CLS/position-embedding assembly. */
int dllm_clip_patchify(const void* images_nchw, void* out, int N, int R, int patch, int Kpad, void* stream);
int dllm_clip_assemble(const void* patches, const void* cls, const void* pos, void* out, int N, int P, int C, void* stream);

/* Index-driven row copies: the embedding splice of DreamLLMModel.forward (modeling_dreamllm.py:1082-1141) and the
 * dream-query conditioning gather (:1401-1418).  mode 0: dst[dst_idx[r]] = src[src_idx[r]]; mode 1: += (unique dst). */
int dllm_copy_rows(void* dst, const int* dst_idx, const void* src, const int* src_idx, int R, int H, int mode, void* stream);
int dllm_segment_sum_rows(void* dst, const void* src, const int* seg, const int* rows, int Q, int H, void* stream);
int dllm_zero_rows(void* dst, const int* idx, int R, int H, void* stream);

/* ---- Stable-Diffusion-2.1 UNet denoising step (reference: StableDiffusionHead.pipeline, modeling_plugins.py:809-833;
 * arithmetic = diffusers 0.24 UNet2DConditionModel / DDPMScheduler / DDIM, restated in oracle/unet_oracle.py). NHWC bf16. ---- */

/* attention with separate kv length (UNet cross-attention on the dream-query conditioning: Skv = 64 / 77). */
int dllm_attn_fwd_ex(const void* q, const void* k, const void* v, void* out, float* lse, const int* seqlens, int B, int Sq,
                     int Skv, int nh, int d, long ld_q, long ld_kv, long ld_o, int causal, float scale, void* stream);
/* kv-cache attention (reference :344-355, :444-449 concat past k/v): k/v live in a preallocated [B, kv_rows, nh*d] cache with Skv valid
 * rows; causal uses the bottom-right aligned mask (query i is at absolute position Skv - Sq + i). */
int dllm_attn_fwd_cache(const void* q, const void* k, const void* v, void* out, float* lse, int B, int Sq, int Skv, int kv_rows, int nh,
                        int d, long ld_q, long ld_kv, long ld_o, int causal, float scale, void* stream);
/* same, for a PADDED prompt batch held in the cache: kv_mask [B, mask_ld] bytes, 0 = pad key (the 2-D `attention_mask` HF `generate` keeps
 * extending, modeling_dreamllm.py:1511-1547; the reference's flash path drops those tokens with `_upad_input`, :553-583).  Pad QUERY rows
 * whose keys are all masked come back as zeros, as `pad_input` re-inserts them (:545).  mask_ld % 16 == 0, >= round_up(Skv, 64). */
int dllm_attn_fwd_cache_mask(const void* q, const void* k, const void* v, void* out, float* lse, const void* kv_mask, int mask_ld, int B,
                             int Sq, int Skv, int kv_rows, int nh, int d, long ld_q, long ld_kv, long ld_o, int causal, float scale,
                             void* stream);
/* ResnetBlock2D / Upsample2D conv: implicit-GEMM 3x3 stride-1 pad-1 on tcgen05 (4-D TMA im2col, zero-fill padding).
 * y = conv(x, w) + bias[c] + rowbias[n, c] (+ residual);  w is [Cout, 3, 3, Cin]. */
int dllm_conv3x3_nhwc(const void* x, const void* w, void* y, int N, int H, int W, int Cin, int Cout, const void* bias,
                      const void* rowbias, const void* residual, void* stream);
size_t dllm_groupnorm_workspace_bytes(int N, int HW, int G);
int dllm_groupnorm_nhwc(const void* x, const void* w, const void* b, void* y, void* workspace, size_t ws_bytes, int N, int HW,
                        int C, int G, float eps, int silu, void* stream);
int dllm_geglu(const void* in, void* out, int T, int I, void* stream);
int dllm_upsample2x_nhwc(const void* x, void* y, int N, int H, int W, int C, void* stream);
/* pad = 1: UNet Downsample2D; pad = 0: VAE downsample (F.pad (0,1,0,1) + padding-0 conv) */
int dllm_im2col_s2_nhwc(const void* x, void* out, int N, int H, int W, int C, int pad, void* stream);
int dllm_copy_cols(const void* src, void* dst, long rows, int Cs, int Cd, int col0, void* stream);
/* latents x_nchw [Bsrc,Cin,H,W] fp32 -> y_nhwc [B,H,W,Cout]; image n reads latents[n % Bsrc] (CFG duplication, plugins:811) */
int dllm_conv_in(const float* x_nchw, const void* w, const void* bias, void* y_nhwc, int B, int Bsrc, int Cin, int H, int W, int Cout,
                 void* stream);
int dllm_conv_out(const void* x_nhwc, const void* w, const void* bias, float* y_nchw, int B, int C, int H, int W, int Cout, void* stream);
/* Split-K forms for small-M shapes (stage-1 steps run the UNet on 4 samples: an 8x8 plane is 256 GEMM rows and would occupy 5 of 74 CTA
 * pairs).  `*_workspace_bytes` returns 0 when the shape is not worth splitting; otherwise pass that many bytes of 16-byte aligned scratch:
 * the GEMM writes fp32 K-slice partials there and a reduce kernel applies bias / row-group bias / activation / residual with the fused
 * epilogue's rounding points.  C = act(bf16(A B^T + bias)) (+ residual) for b_mn = 0 (nn.Linear forward), C = A B for b_mn = 1 (dgrad). */
size_t dllm_gemm_splitk_workspace_bytes(int M, int N, int K);
int dllm_gemm_bf16_ws(const void* A, const void* B, void* C, int M, int N, int K, long lda, long ldb, long ldc, int b_mn, const void* bias,
                      const void* residual, long ldr, int act, void* ws, size_t ws_bytes, void* stream);
size_t dllm_conv3x3_splitk_workspace_bytes(int N, int H, int W, int Cin, int Cout);
int dllm_conv3x3_nhwc_ws(const void* x, const void* w, void* y, int N, int H, int W, int Cin, int Cout, const void* bias,
                         const void* rowbias, const void* residual, void* ws, size_t ws_bytes, void* stream);
/* UNet FeedForward-in projection with GEGLU in the GEMM epilogue (diffusers `GEGLU.forward`: h, gate = proj(x).chunk(2); h * gelu(gate)):
 * out[M, N/2] = h * gelu(gate) where [h | gate] = A @ W^T + b and Wp / bias_p hold W / b with rows permuted to [64 h rows | 64 gate rows]
 * per 128-row group; the [M, N] projection is never written.  N % 128 == 0. */
int dllm_gemm_bf16_geglu(const void* A, const void* Wp, const void* bias_p, void* out, int M, int N, int K, long lda, long ldb, long ldc,
                         void* stream);
/* tensor-core forms of the two tiny-channel convolutions (conv_in: Conv2d(4|3, C, 3, pad 1) of UNet2DConditionModel / AutoencoderKL;
 * conv_out: Conv2d(C, 4|8|3, 3, pad 1)):
 *   cols [B*H*W, 64] bf16 = im2col(x_nchw fp32), k = (c*3 + r)*3 + s zero-padded to 64  ->  dllm_gemm_bf16_ex(cols, Wk [Cout, 64], bias)
 *   y8 [N*H*W, 8] bf16 = dllm_conv3x3_nhwc(x, W8 [8, 9*C])                                ->  out fp32 NCHW [N, Cout, H*W] (first Cout channels) */
int dllm_im2col_in(const float* x_nchw, void* cols, int B, int Bsrc, int Cin, int H, int W, void* stream);
int dllm_nhwc_to_nchw_f32(const void* y_nhwc8, float* out_nchw, int N, int HW, int Cp, int Cout, void* stream);
/* timestep = timesteps[*step] (device-side schedule so one captured CUDA graph serves every step) */
int dllm_timestep_embedding(const int* timesteps, const int* step, void* out, int B, int dim, void* stream);
int dllm_timestep_embedding_batch(const int* t_per_sample, void* out, int B, int dim, void* stream);
/* fused CFG combine + DDIM (mode 0) / DDPM (mode 1) update; coef[step] = {sqrt(a_t), sqrt(1-a_t), c_x0, c_eps|c_xt, sigma};
 * advances *step. */
int dllm_sampler_step(const float* eps, float* latents, const float* noise, const float* coef, int* step, float guidance,
                      int use_cfg, int mode, long n, void* stream);

/* ---- input-gradient (dgrad-only) backward of the frozen UNet + diffusion loss plumbing: StableDiffusionHead.forward,
 * modeling_plugins.py:493-577 (add_noise :536, unet(...) :556, MSE :559). ---- */
int dllm_attn_bwd_ex(const void* dout, const void* q, const void* k, const void* v, const void* out, const float* lse, void* dq,
                     void* dk, void* dv, const int* seqlens, void* workspace, size_t workspace_bytes, int B, int Sq, int Skv, int nh,
                     int d, long ld_q, long ld_kv, long ld_o, long ld_dq, long ld_dkv, int causal, float scale, void* stream);
/* Host-only: the order in which the persistent attention kernels (flash attention replacing flash_attn_func / flash_attn_varlen_func,
 * modeling_dreamllm.py:532-551) hand out their (tile, head x batch) work items — item w of a launch with `grid` resident CTAs.  Windows of
 * win_heads = ceil(2 * grid / ntiles) consecutive (head, batch) pairs; inside a window heaviest tiles first (descending != 0: highest
 * tile index first — forward / dQ under a causal mask; 0: lowest first — dK/dV).  For tests and tooling; no device work. */
void dllm_attn_item_order(int w, int ntiles, int n_hb, int grid, int descending, int* tile, int* hb, int* win_heads);
int dllm_groupnorm_stats(const void* x, float* stats, void* workspace, size_t ws_bytes, int N, int HW, int C, int G, float eps, void* stream);
int dllm_groupnorm_apply(const void* x, const void* w, const void* b, const float* stats, void* y, int N, int HW, int C, int G, int silu,
                         void* stream);
int dllm_groupnorm_bwd_nhwc(const void* dy, const void* x, const void* w, const void* b, const float* stats, const void* dres, void* dx,
                            void* workspace, size_t ws_bytes, int N, int HW, int C, int G, int silu, void* stream);
int dllm_layernorm_bwd(const void* dy, const void* x, const void* w, const void* dres, void* dx, int T, int H, float eps, void* stream);
int dllm_geglu_bwd(const void* dout, const void* in, void* din, int T, int I, void* stream);
int dllm_upsample2x_bwd_nhwc(const void* dy, void* dx, int N, int H, int W, int C, void* stream);
int dllm_col2im_s2_nhwc(const void* dcols, void* dx, int N, int H, int W, int C, void* stream);
int dllm_copy_cols2(const void* src, void* dst, long rows, int Cs, int Cd, int scol0, int dcol0, int ncols, void* stream);
int dllm_conv_out_bwd(const float* dy_nchw, const void* w, void* dx_nhwc, int B, int C, int H, int W, int Cout, void* stream);
int dllm_add_noise(const float* x0, const float* noise, const int* t, const float* alphas_cumprod, float* out, int B, long per_sample,
                   void* stream);
int dllm_mse_fwd_bwd(const float* pred, const float* target, float* loss, float* dpred, long n, void* stream);
/* min-SNR weighted MSE (modeling_plugins.py:561-572 with `_compute_snr` :468-491): w_b = min(snr(t_b), gamma) / snr(t_b) */
int dllm_mse_minsnr_fwd_bwd(const float* pred, const float* target, const int* t, const float* alphas_cumprod, float snr_gamma,
                            float* loss, float* dpred, int B, long per_sample, void* stream);

/* ---- VAE encoder helpers (AutoencoderKL.encode(...).latent_dist.sample() * scaling_factor, modeling_plugins.py:511-512) ---- */
int dllm_softmax_rows(void* x, long rows, int cols, float scale, void* stream);
int dllm_vae_sample(const float* h, const void* wq, const void* bq, const float* z, float* out, int B, int L, long plane, float scaling,
                    void* stream);

/* ---- data-parallel optimizer shard ----
 * Replaces torch.optim.AdamW (`optim="adamw_torch"`, projects/dreamllm/configs/stage1/base.py:85, stage2/base.py:95) applied to this
 * rank's shard of a flat bf16 gradient bucket under FSDP shard_grad_op (stage2/base.py:91-94), with the gradient clipping of
 * omni/train/trainer.py:800-807 folded in: the step scales grads by min(1, max_grad_norm / (sqrt(*grad_sumsq) + 1e-6)) when
 * grad_sumsq != NULL and max_grad_norm > 0 (grad_sumsq = global sum of squares, device pointer, read at kernel time: no host sync).
 * n % 8 == 0, 16-byte aligned pointers.  bf16_state = 0: master / exp_avg / exp_avg_sq are fp32 shards, `param` (bf16) is written from
 * the updated master.  bf16_state = 1: `param`, exp_avg, exp_avg_sq are bf16 and every ATen op of the reference optimizer's
 * single-tensor path rounds to bf16 (bit-for-bit the arithmetic the reference runs on its bf16-loaded model); master is ignored.
 * `step` is the 1-based step count used for the bias corrections.  Hyper-parameters are doubles: Python computes 1 - beta1, lr / (1 - beta1^t)
 * ... in double before ATen narrows them to the fp32 opmath type, and the kernel's scalars are derived the same way. */
int dllm_adamw_step(const void* grad, void* master, void* exp_avg, void* exp_avg_sq, void* param, long n, int bf16_state, double lr,
                    double beta1, double beta2, double eps, double weight_decay, int step, const float* grad_sumsq,
                    double max_grad_norm,
                    void* stream);
/* sum of squares of a bf16 vector in fp32 (deterministic two-stage reduction); *out = (accumulate ? *out : 0) + sum(x^2) */
size_t dllm_sumsq_workspace_bytes(void);
int dllm_sumsq_bf16(const void* x, long n, float* out, int accumulate, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif

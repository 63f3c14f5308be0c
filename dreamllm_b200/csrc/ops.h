// Internal C++ declarations of the non-GEMM ops (definitions: elementwise.cu, attn_sm100.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace dllm {
int rmsnorm_fwd(const void* x, const void* add, const void* w, void* x_out, void* y, float* rstd, int T, int H, float eps,
                cudaStream_t s);
size_t rmsnorm_bwd_workspace(int T, int H);
int rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, const void* dres, void* dx, void* dw,
                int dw_accumulate, void* workspace, size_t workspace_bytes, int T, int H, cudaStream_t s);
int rope_inplace(void* buf, const void* cos_t, const void* sin_t, const int* pos, long ld, int T, int heads_total,
                 int head_dim, int mode, cudaStream_t s);
int swiglu_fwd(const void* gu, void* act, long ld_gu, int T, int I, cudaStream_t s);
int swiglu_bwd(const void* dact, const void* gu, void* dgu, long ld_gu, int T, int I, cudaStream_t s);
int add_bf16(const void* a, const void* b, void* o, long n, cudaStream_t s);
int cross_entropy(void* logits, const long long* labels, float* loss, float dloss, void* workspace, long ld, int T, int V,
                  int write_grad, cudaStream_t s);
int embedding_fwd(const long long* ids, const void* W, void* out, int T, int H, cudaStream_t s);
int embedding_bwd(const long long* sorted_ids, const long long* order, const void* dy, void* dW, int T, int H,
                  int accumulate, cudaStream_t s);
int layernorm_fwd(const void* x, const void* w, const void* b, void* y, int T, int H, float eps, cudaStream_t s);
int clip_patchify(const void* img, void* out, int N, int R, int patch, int Kpad, cudaStream_t s);
int clip_assemble(const void* patches, const void* cls, const void* pos, void* out, int N, int P, int C, cudaStream_t s);
int copy_rows(void* dst, const int* dst_idx, const void* src, const int* src_idx, int R, int H, int mode, cudaStream_t s);
int segment_sum_rows(void* dst, const void* src, const int* seg, const int* rows, int Q, int H, cudaStream_t s);
int zero_rows(void* dst, const int* idx, int R, int H, cudaStream_t s);
int attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, const int* seqlens, int B, int S, int nh,
             int d, long ld_qkv, long ld_o, int causal, float scale, cudaStream_t s);
int attn_fwd_ex(const void* q, const void* k, const void* v, void* out, float* lse, const int* seqlens, int B, int S, int Skv,
                int nh, int d, long ld_q, long ld_kv, long ld_o, int causal, float scale, cudaStream_t s);
int attn_fwd_cache(const void* q, const void* k, const void* v, void* out, float* lse, const int* seqlens, int B, int S, int Skv,
                   int kv_rows, int nh, int d, long ld_q, long ld_kv, long ld_o, int causal, float scale, cudaStream_t s);
int attn_fwd_cache_mask(const void* q, const void* k, const void* v, void* out, float* lse, const int* seqlens, const void* kv_mask,
                        int mask_ld, int B, int S, int Skv, int kv_rows, int nh, int d, long ld_q, long ld_kv, long ld_o, int causal,
                        float scale, cudaStream_t st);
size_t groupnorm_workspace(int N, int HW, int G);
int groupnorm_nhwc(const void* x, const void* w, const void* b, void* y, void* workspace, size_t ws_bytes, int N, int HW, int C,
                   int G, float eps, int silu, cudaStream_t s);
int geglu(const void* in, void* out, int T, int I, cudaStream_t s);
int upsample2x_nhwc(const void* x, void* y, int N, int H, int W, int C, cudaStream_t s);
int im2col_s2_nhwc(const void* x, void* out, int N, int H, int W, int C, int pad, cudaStream_t s);
int copy_cols(const void* src, void* dst, long rows, int Cs, int Cd, int col0, cudaStream_t s);
int conv_in_nchw_to_nhwc(const float* x, const void* w, const void* bias, void* y, int B, int Bsrc, int Cin, int H, int W,
                         int Cout, cudaStream_t s);
int conv_out_nhwc_to_nchw(const void* x, const void* w, const void* bias, float* y, int B, int C, int H, int W, int Cout,
                          cudaStream_t s);
int im2col_in(const float* x, void* cols, int B, int Bsrc, int Cin, int H, int W, cudaStream_t s);
int nhwc_to_nchw_f32(const void* y, float* out, int N, int HW, int Cp, int Cout, cudaStream_t s);
int timestep_embedding(const int* timesteps, const int* step, void* out, int B, int dim, cudaStream_t s);
int timestep_embedding_batch(const int* t, void* out, int B, int dim, cudaStream_t s);
int sampler_step(const float* eps, float* latents, const float* noise, const float* coef, int* step, float guidance, int use_cfg,
                 int mode, long n, cudaStream_t s);
size_t attn_bwd_workspace(int B, int S, int nh, int d);
int attn_bwd(const void* dout, const void* q, const void* k, const void* v, const void* out, const float* lse, void* dq,
             void* dk, void* dv, const int* seqlens, void* workspace, size_t workspace_bytes, int B, int S, int nh, int d,
             long ld_qkv, long ld_o, long ld_dqkv, int causal, float scale, cudaStream_t s);
int attn_bwd_ex(const void* dout, const void* q, const void* k, const void* v, const void* out, const float* lse, void* dq,
                void* dk, void* dv, const int* seqlens, void* workspace, size_t workspace_bytes, int B, int S, int Skv, int nh,
                int d, long ld_q, long ld_kv, long ld_o, long ld_dq, long ld_dkv, int causal, float scale, cudaStream_t s);
int groupnorm_bwd_nhwc(const void* dy, const void* x, const void* w, const void* b, const float* stats, const void* dres, void* dx,
                       void* workspace, size_t ws_bytes, int N, int HW, int C, int G, int silu, cudaStream_t s);
void attn_item_order(int w, int ntiles, int n_hb, int grid, int descending, int* tile, int* hb, int* win_heads_out);
int groupnorm_stats(const void* x, float* stats, void* workspace, size_t ws_bytes, int N, int HW, int C, int G, float eps, cudaStream_t s);
int groupnorm_apply(const void* x, const void* w, const void* b, const float* stats, void* y, int N, int HW, int C, int G, int silu,
                    cudaStream_t s);
int layernorm_bwd(const void* dy, const void* x, const void* w, const void* dres, void* dx, int T, int H, float eps, cudaStream_t s);
int geglu_bwd(const void* dout, const void* in, void* din, int T, int I, cudaStream_t s);
int upsample2x_bwd_nhwc(const void* dy, void* dx, int N, int H, int W, int C, cudaStream_t s);
int col2im_s2_nhwc(const void* dcols, void* dx, int N, int H, int W, int C, cudaStream_t s);
int copy_cols2(const void* src, void* dst, long rows, int Cs, int Cd, int scol0, int dcol0, int ncols, cudaStream_t s);
int conv_out_bwd(const float* dy, const void* w, void* dx, int B, int C, int H, int W, int Cout, cudaStream_t s);
int add_noise(const float* x0, const float* noise, const int* t, const float* ac, float* out, int B, long per_sample, cudaStream_t s);
int mse_fwd_bwd(const float* pred, const float* target, float* loss, float* dpred, long n, cudaStream_t s);
int mse_minsnr_fwd_bwd(const float* pred, const float* target, const int* t, const float* ac, float gamma, float* loss, float* dpred,
                       int B, long per_sample, cudaStream_t s);
int softmax_rows(void* x, long rows, int cols, float scale, cudaStream_t s);
int vae_sample(const float* h, const void* wq, const void* bq, const float* z, float* out, int B, int L, long plane, float scaling,
               cudaStream_t s);
// optim.cu — sharded AdamW + grad-norm
int adamw_step(const void* grad, void* master, void* mom, void* var, void* param, long n, int bf16_state, double lr, double beta1,
               double beta2, double eps, double weight_decay, int step, const float* sumsq, double max_norm, cudaStream_t s);
size_t sumsq_workspace();
int sumsq_bf16(const void* x, long n, float* out, int accumulate, void* workspace, size_t ws_bytes, cudaStream_t s);
}  // namespace dllm

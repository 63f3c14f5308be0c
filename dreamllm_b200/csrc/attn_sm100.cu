// placeholder until the tcgen05 attention kernels land (next commit) — keeps the C ABI complete
#include "gemm_sm100.h"
#include "ops.h"
namespace dllm {
int attn_fwd(const void*, const void*, const void*, void*, float*, const int*, int, int, int, int, long, long, int, float,
             cudaStream_t) { return DLLM_ERR_UNSUPPORTED; }
size_t attn_bwd_workspace(int, int, int, int) { return 0; }
int attn_bwd(const void*, const void*, const void*, const void*, const void*, const float*, void*, void*, void*, const int*,
             void*, size_t, int, int, int, int, long, long, long, int, float, cudaStream_t) { return DLLM_ERR_UNSUPPORTED; }
}

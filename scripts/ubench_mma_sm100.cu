// tcgen05.mma dispatch cost of the tile-GEMM shapes the attention kernels issue: clocks per UMMA_K = 16 instruction,
// measured from the first issue to the retirement of the last (tcgen05.commit -> mbarrier), one issuing thread per CTA, 1 or 2 CTAs per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/ubench_mma scripts/ubench_mma_sm100.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../dreamllm_b200/csrc/common.cuh"

using namespace dllm;

struct Variant {
  const char* name;
  int N;        // MMA N
  int ksteps;   // UMMA_K steps per tile-GEMM
  int a_tmem;   // A operand from tensor memory (TS form)
  int b_mn;     // B operand MN-major
};

__global__ void __launch_bounds__(128, 2) mma_kernel(long long* out, int N, int ksteps, int a_tmem, int b_mn, int rounds) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc<1>(&tptr, 256); tmem_relinquish<1>(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tptr;
  if (threadIdx.x == 0) {
    const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + 32768);
    const uint32_t idesc = make_idesc_bf16(128, N, false, b_mn != 0);
    // warm-up round
    for (int ks = 0; ks < ksteps; ++ks) {
      if (a_tmem) umma_ts(tbase, tbase + 128 + ks * 8, op_desc(sB, b_mn, b_mn ? static_cast<uint32_t>(ksteps) * 2048u : static_cast<uint32_t>(N) * 128u, ks), idesc, ks > 0);
      else umma_ss<1>(tbase, op_desc(sA, false, 16384, ks), op_desc(sB, b_mn, b_mn ? static_cast<uint32_t>(ksteps) * 2048u : static_cast<uint32_t>(N) * 128u, ks), idesc, ks > 0);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0, 1);
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r)
      for (int ks = 0; ks < ksteps; ++ks) {
        if (a_tmem) umma_ts(tbase, tbase + 128 + ks * 8, op_desc(sB, b_mn, b_mn ? static_cast<uint32_t>(ksteps) * 2048u : static_cast<uint32_t>(N) * 128u, ks), idesc, 1u);
        else umma_ss<1>(tbase, op_desc(sA, false, 16384, ks), op_desc(sB, b_mn, b_mn ? static_cast<uint32_t>(ksteps) * 2048u : static_cast<uint32_t>(N) * 128u, ks), idesc, 1u);
      }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 1, 2);
    const long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tbase, 256);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* out;
  cudaMalloc(&out, 16 * 1024);
  const int smem_bytes = 32768 + 65536 + 1024;
  cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const Variant vs[] = {
      {"SS_N64_k8   (S = Q K^T, 64 keys, d=128)", 64, 8, 0, 0},
      {"SS_N128_k8  (S = Q K^T, 128 keys, d=128)", 128, 8, 0, 0},
      {"SS_N256_k8", 256, 8, 0, 0},
      {"SS_N64_k4   (S, 64 keys, d=64)", 64, 4, 0, 0},
      {"SS_N128_k4  (S, 128 keys, d=64)", 128, 4, 0, 0},
      {"TS_N128_k4_Bmn (O += P V, d=128, P in TMEM)", 128, 4, 1, 1},
      {"SS_N128_k4_Bmn (O += P V, d=128, P in smem)", 128, 4, 0, 1},
      {"TS_N64_k4_Bmn  (O += P V, d=64)", 64, 4, 1, 1},
      {"TS_N64_k8   (S with Q in TMEM, 64 keys)", 64, 8, 1, 0},
      {"TS_N128_k8  (S with Q in TMEM, 128 keys)", 128, 8, 1, 0},
      {"TS_N128_k8_Bmn (O += P V over 128 keys, d=128)", 128, 8, 1, 1},
  };
  printf("{\"sms\": %d, \"unit\": \"clk per UMMA_K=16 instruction (M=128), [issue loop only, until retired]\"", sms);
  const int rounds = 64;
  for (const Variant& v : vs)
    for (int ctas = 1; ctas <= 2; ++ctas) {
      mma_kernel<<<sms * ctas, 128, smem_bytes>>>(out, v.N, v.ksteps, v.a_tmem, v.b_mn, rounds);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[4];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      const double n = double(rounds) * v.ksteps;
      printf(",\n \"%s | %d CTA/SM\": [%.1f, %.1f]%s", v.name, ctas, h[0] / n, h[1] / n, e == cudaSuccess ? "" : " /*ERR*/");
      if (e != cudaSuccess) { printf(", \"cuda_error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    }
  printf("}\n");
  return 0;
}

// Micro-benchmarks of the three per-SM resources the attention softmax warps contend for: TMEM read bandwidth
// (tcgen05.ld 32x32b.x32), MUFU.EX2 issue rate and the fp32 FMA pipe in scalar vs packed (fma.rn.f32x2) form.  Built and run by
// scripts/r02i_gpu.sh:   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_sm100 scripts/ubench_sm100.cu
// Every kernel runs one CTA (or two) per SM on all SMs; the figure printed is per SM per clock, from clock64 on SM 0's CTA(s).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../dreamllm_b200/csrc/common.cuh"

using namespace dllm;

// mode 0: ld + wait per load; mode 1: two loads in flight per wait
template <int kMode>
__global__ void ldtm_kernel(long long* out, int iters, uint32_t* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc<1>(&tptr, 256); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tptr + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t a[32], b[32];
    tmem_ld32(base + ((i & 1) * 64), a);
    if (kMode == 1) tmem_ld32(base + ((i & 1) * 64) + 32, b);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 32; e += 8) acc ^= a[e];
    if (kMode == 1) {
#pragma unroll
      for (int e = 0; e < 32; e += 8) acc ^= b[e];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tptr, 256);
}

__global__ void sttm_kernel(long long* out, int iters) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc<1>(&tptr, 256); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tptr + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t a[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) a[e] = threadIdx.x + e;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    tmem_st32(base + ((i & 1) * 64), a);
    tmem_st_wait();
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tptr, 256);
}

// 16 independent ex2 chains per thread
__global__ void mufu_kernel(long long* out, int iters, float* sink, float seed) {
  float x[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) x[e] = seed * (e + 1) * 1e-3f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int e = 0; e < 16; ++e) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[e]));
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) s += x[e];
  if (s == 1.2345f) sink[0] = s;
}

// kPacked = false: 32 scalar FFMA per iteration; true: 16 fma.rn.f32x2 (same 32 results)
template <bool kPacked>
__global__ void fma_kernel(long long* out, int iters, float* sink, float seed) {
  float x[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) x[e] = seed * (e + 1);
  const float a = 0.999f, b = 1e-3f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if constexpr (!kPacked) {
#pragma unroll
      for (int e = 0; e < 32; ++e) asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x[e]) : "f"(a), "f"(b));
    } else {
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        asm volatile(
            "{\n .reg .b64 v, s, t;\n mov.b64 v, {%0, %1};\n mov.b64 s, {%2, %2};\n mov.b64 t, {%3, %3};\n"
            " fma.rn.ftz.f32x2 v, v, s, t;\n mov.b64 {%0, %1}, v;\n}"
            : "+f"(x[e]), "+f"(x[e + 1]) : "f"(a), "f"(b));
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 32; ++e) s += x[e];
  if (s == 1.2345f) sink[0] = s;
}

// exp2 through MUFU vs the degree-3 polynomial on the FMA pipe (Cody-Waite split: 2^x = 2^floor(x) * p(frac)), accuracy + throughput
__device__ __forceinline__ float exp2_poly(float x) {
  // x <= 0 expected (softmax argument); clamp keeps the exponent field valid
  x = fmaxf(x, -126.f);
  const float fl = floorf(x);
  const float f = x - fl;                         // [0, 1)
  // minimax-ish cubic for 2^f on [0,1): max rel err ~1e-4 (bf16 P needs 4e-3)
  float p = 0.0555041086f;
  p = fmaf(p, f, 0.2402265069f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (static_cast<int>(fl) << 23));
}
__global__ void poly_kernel(long long* out, int iters, float* sink, float seed, float* maxerr) {
  float x[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) x[e] = -seed * (e + 1) * 0.37f - threadIdx.x * 0.01f;
  __syncthreads();
  const long long t0 = clock64();
  float s = 0.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int e = 0; e < 16; ++e) { s += exp2_poly(x[e]); x[e] -= 1e-3f; }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (s == 1.2345f) sink[0] = s;
  if (blockIdx.x == 0) {
    float worst = 0.f;
    for (int k = 0; k < 4096; ++k) {
      const float v = -(threadIdx.x * 4096 + k) * (20.f / (4096.f * blockDim.x));
      const float ref = exp2f(v), got = exp2_poly(v);
      worst = fmaxf(worst, fabsf(got - ref) / ref);
    }
    atomicMax(reinterpret_cast<int*>(maxerr), __float_as_int(worst));
  }
}

static long long first(long long* d) { long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); return h; }

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* out; uint32_t* sink; float* fs; float* merr;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&sink, 64); cudaMalloc(&fs, 64); cudaMalloc(&merr, 4);
  cudaMemset(merr, 0, 4);
  const int iters = 4096;
  printf("{\"sms\": %d", sms);
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int warps : {1, 2, 4, 8}) {
      ldtm_kernel<0><<<sms * ctas, warps * 32>>>(out, iters, sink); cudaDeviceSynchronize();
      double c0 = double(first(out));
      ldtm_kernel<1><<<sms * ctas, warps * 32>>>(out, iters, sink); cudaDeviceSynchronize();
      double c1 = double(first(out));
      // bytes per SM per clk: ctas * warps * 32 lanes * 32 cols * 4 B per load
      printf(",\n \"ldtm_x32_ctas%d_warps%d_B_per_clk_per_sm\": [%.1f, %.1f]", ctas, warps, ctas * warps * 4096.0 * iters / c0,
             ctas * warps * 8192.0 * iters / c1);
    }
  for (int warps : {4, 8}) {
    sttm_kernel<<<sms, warps * 32>>>(out, iters); cudaDeviceSynchronize();
    printf(",\n \"sttm_x32_warps%d_B_per_clk_per_sm\": %.1f", warps, warps * 4096.0 * iters / double(first(out)));
  }
  for (int warps : {4, 8, 16}) {
    mufu_kernel<<<sms, warps * 32>>>(out, iters, fs, 1.f); cudaDeviceSynchronize();
    printf(",\n \"mufu_ex2_warps%d_per_clk_per_sm\": %.2f", warps, warps * 32 * 16.0 * iters / double(first(out)));
  }
  for (int warps : {4, 8, 16}) {
    fma_kernel<false><<<sms, warps * 32>>>(out, iters, fs, 1.f); cudaDeviceSynchronize();
    double a = warps * 32 * 32.0 * iters / double(first(out));
    fma_kernel<true><<<sms, warps * 32>>>(out, iters, fs, 1.f); cudaDeviceSynchronize();
    double b = warps * 32 * 32.0 * iters / double(first(out));
    printf(",\n \"fma_results_per_clk_per_sm_warps%d\": {\"scalar\": %.1f, \"f32x2\": %.1f}", warps, a, b);
  }
  for (int warps : {4, 8}) {
    poly_kernel<<<sms, warps * 32>>>(out, iters / 4, fs, 1.f, merr); cudaDeviceSynchronize();
    printf(",\n \"exp2_poly_warps%d_per_clk_per_sm\": %.2f", warps, warps * 32 * 16.0 * (iters / 4) / double(first(out)));
  }
  float e; cudaMemcpy(&e, merr, 4, cudaMemcpyDeviceToHost);
  printf(",\n \"exp2_poly_max_rel_err\": %.3g", e);
  cudaError_t err = cudaGetLastError();
  printf(",\n \"cuda_error\": \"%s\"}\n", cudaGetErrorString(err));
  return 0;
}

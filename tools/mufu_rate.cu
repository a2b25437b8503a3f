// Micro-benchmark: cycles per warp-level activation element for the epilogue math variants (16 warps/CTA = 4 per SMSP,
// one CTA per SM).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_exp/mufu_rate tools/mufu_rate.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2h2(unsigned x) { unsigned y; asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
constexpr float kC = 0.44361418485641479492f;
template <int V> __device__ __forceinline__ float act(float t) {
  if (V == 0) return ex2a(t);
  if (V == 1) return lg2a(t);
  if (V == 2) { const float e = ex2a(-fabsf(t)); return (fmaxf(t, 0.f) + lg2a(1.f + e)) * kC; }          // production
  if (V == 3) {                                                                                          // 1 MUFU + poly
    const float e = ex2a(-fabsf(t));
    float p = fmaf(e, -0.0783f, 0.3169f); p = fmaf(p, e, -0.6750f); p = fmaf(p, e, 1.4363f);
    return fmaf(p, e, fmaxf(t, 0.f)) * kC;
  }
  return t;
}
template <int V> __global__ void k(float* out, long long* cyc, int iters, float seed) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = seed + threadIdx.x * 1e-3f + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = act<V>(v[i] - 3.f);
  }
  const long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// half2 variant: t pairs -> packed f16x2 result
__global__ void kh(unsigned* out, long long* cyc, int iters, float seed) {
  float v[16];
  for (int i = 0; i < 16; ++i) v[i] = seed + threadIdx.x * 1e-3f + i;
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  const __half2 one = __float2half2_rn(1.f), c1 = __float2half2_rn(1.4363f), c2 = __float2half2_rn(-0.6750f),
                c3 = __float2half2_rn(0.3169f), c4 = __float2half2_rn(-0.0783f), kc = __float2half2_rn(kC), z = __float2half2_rn(0.f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __half2 t = __floats2half2_rn(v[2 * i] - 3.f, v[2 * i + 1] - 3.f);
      __half2 na = __hneg2(__habs2(t));
      unsigned eu = ex2h2(*reinterpret_cast<unsigned*>(&na));
      __half2 e = *reinterpret_cast<__half2*>(&eu);
      __half2 p = __hfma2(e, c4, c3); p = __hfma2(p, e, c2); p = __hfma2(p, e, c1);
      __half2 y = __hmul2(__hfma2(p, e, __hmax2(t, z)), kc);
      acc ^= *reinterpret_cast<unsigned*>(&y);
      const float2 f = __half22float2(y);
      v[2 * i] = f.x + it; v[2 * i + 1] = f.y - it;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  long long h[148];
  const char* names[] = {"ex2 only", "lg2 only", "softplus (ex2+lg2, production)", "softplus (ex2 + deg-4 poly)"};
  for (int v = 0; v < 4; ++v) {
    for (int rep = 0; rep < 2; ++rep) {
      if (v == 0) k<0><<<148, 512>>>(out, cyc, iters, 0.5f);
      if (v == 1) k<1><<<148, 512>>>(out, cyc, iters, 0.5f);
      if (v == 2) k<2><<<148, 512>>>(out, cyc, iters, 0.5f);
      if (v == 3) k<3><<<148, 512>>>(out, cyc, iters, 0.5f);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    // 4 warps per SMSP, 8 elements per iteration each
    printf("%-36s %.2f cycles per warp-element per SMSP\n", names[v], (double)h[0] / iters / 8 / 4);
  }
  for (int rep = 0; rep < 2; ++rep) { kh<<<148, 512>>>((unsigned*)out, cyc, iters, 0.5f); cudaDeviceSynchronize(); }
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-36s %.2f cycles per warp-element per SMSP (16 elements per iteration)\n", "softplus half2 (ex2.f16x2 + poly)", (double)h[0] / iters / 16 / 4);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

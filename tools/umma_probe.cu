// Stand-alone probe for the UMMA shared-memory descriptor / TMEM layout assumptions of mlp_kernel.cuh.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu ; run on a B200.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../mvsdf_b200/csrc/ptx.cuh"
using namespace mvsdf;

constexpr int M = 128, N = 64, K = 16;

__global__ void probe(const __half* A, const __half* B, float* D, int a_lbo, int a_sbo, int b_lbo, int b_sbo, int swap_fields, int b_mn) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s0 = ptx::smem_u32(smem);
  const uint32_t sA = s0, sB = s0 + 16384, sBar = s0 + 16384 + 16384, sT = sBar + 16;
  // place A[r][k], B[n][k] per the K-major no-swizzle formula
  for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
    int r = i / K, k = i % K;
    *reinterpret_cast<__half*>(smem + (r / 8) * a_sbo + (k / 8) * a_lbo + (r % 8) * 16 + (k % 8) * 2) = A[i];
  }
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    int n = i / K, k = i % K;
    if (b_mn) *reinterpret_cast<__half*>(smem + 16384 + (n / 8) * b_sbo + (k / 8) * b_lbo + (k % 8) * 16 + (n % 8) * 2) = B[i];
    else *reinterpret_cast<__half*>(smem + 16384 + (n / 8) * b_sbo + (k / 8) * b_lbo + (n % 8) * 16 + (k % 8) * 2) = B[i];
  }
  if (threadIdx.x == 0) { ptx::mbar_init(sBar, 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(sT, 64); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 16384 + 16384 + 16);
  if (threadIdx.x == 0) {
    uint64_t da = swap_fields ? ptx::smem_desc(sA, a_sbo, a_lbo) : ptx::smem_desc(sA, a_lbo, a_sbo);
    uint64_t db = swap_fields ? ptx::smem_desc(sB, b_sbo, b_lbo) : ptx::smem_desc(sB, b_lbo, b_sbo);
    ptx::umma_f16(tmem, da, db, ptx::idesc_f16_f32(M, N) | (b_mn ? (1u << 16) : 0u), 0u);
    ptx::umma_commit(sBar);
  }
  ptx::mbar_wait(sBar, 0);
  ptx::tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 4) {
    for (int hcol = 0; hcol < 2; ++hcol) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + hcol * 32, v);
      ptx::tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + hcol * 32 + j] = __uint_as_float(v[j]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<__half> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K), ref(M * N), out(M * N);
  srand(1);
  for (int i = 0; i < M * K; ++i) { fA[i] = (float)((rand() % 17) - 8) * 0.125f; hA[i] = __float2half(fA[i]); }
  for (int i = 0; i < N * K; ++i) { fB[i] = (float)((rand() % 13) - 6) * 0.25f; hB[i] = __float2half(fB[i]); }
  for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += fA[r * K + k] * fB[n * K + k]; ref[r * N + n] = s; }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, out.size() * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  struct V { int a_lbo, a_sbo, b_lbo, b_sbo, swap; const char* name; int b_mn; } vs[] = {
      {128, 512, 1040, 128, 0, "design (A lbo128 sbo512, B lbo1040 sbo128)"},
      {128, 512, 1040, 128, 1, "design, descriptor fields swapped"},
      {128, 256, 128, 256, 0, "compact 2-core rows"},
      {2048, 128, 1024, 128, 0, "k-core-major A"},
      {128, 512, 2048, 128, 0, "B MN-major (lbo 2048 = K blocks, sbo 128 = MN blocks)", 1},
      {128, 512, 128, 2048, 0, "B MN-major, roles swapped (lbo 128, sbo 2048)", 1},
  };
  for (auto& v : vs) {
    cudaMemset(dD, 0, out.size() * 4);
    probe<<<1, 128, 40000>>>(dA, dB, dD, v.a_lbo, v.a_sbo, v.b_lbo, v.b_sbo, v.swap, v.b_mn);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", v.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < M * N; ++i) mx = fmax(mx, fabs(out[i] - ref[i]));
    printf("%-48s max|err| = %g   D[0][0..3] = %g %g %g %g (ref %g %g %g %g)  D[77][33]=%g (ref %g)\n", v.name, mx, out[0], out[1], out[2], out[3],
           ref[0], ref[1], ref[2], ref[3], out[77 * N + 33], ref[77 * N + 33]);
  }
  return 0;
}

// Probe for tcgen05.mma.cta_group::2 semantics (M=256 split by rows across the CTA pair, N=128 split by B rows).
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../mvsdf_b200/csrc/ptx.cuh"
using namespace mvsdf;

constexpr int M = 256, N = 128, K = 16;

__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" :: "r"(bar), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) probe(const __half* A, const __half* B, float* D) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t rank = ptx::cluster_ctarank();
  const uint32_t s0 = ptx::smem_u32(smem);
  const uint32_t sA = s0, sB = s0 + 8192, sBar = s0 + 16384, sT = sBar + 16;
  // my half of A: rows [128 rank, +128); my half of B: rows [64 rank, +64)
  for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
    int r = i / K, k = i % K;
    *reinterpret_cast<__half*>(smem + (r / 8) * 256 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2) = A[(rank * 128 + r) * K + k];
  }
  for (int i = threadIdx.x; i < 64 * K; i += blockDim.x) {
    int n = i / K, k = i % K;
    *reinterpret_cast<__half*>(smem + 8192 + (n / 8) * 128 + (k / 8) * 1040 + (n % 8) * 16 + (k % 8) * 2) = B[(rank * 64 + n) * K + k];
  }
  if (threadIdx.x == 0) { ptx::mbar_init(sBar, 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(sT), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 16384 + 16);
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = ptx::idesc_f16_f32(M, N);
    umma_f16_2cta(tmem, ptx::smem_desc(sA, 128, 256), ptx::smem_desc(sB, 1040, 128), idesc, 0u);
    umma_commit_2cta(sBar, 3);
  }
  ptx::mbar_wait(sBar, 0);
  ptx::tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 4) {
    for (int hc = 0; hc < 4; ++hc) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + hc * 32, v);
      ptx::tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[(rank * 128 + warp * 32 + lane) * N + hc * 32 + j] = __uint_as_float(v[j]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128) : "memory");
}

int main() {
  std::vector<__half> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K), ref(M * N), out(M * N);
  srand(2);
  for (int i = 0; i < M * K; ++i) { fA[i] = (float)((rand() % 17) - 8) * 0.125f; hA[i] = __float2half(fA[i]); }
  for (int i = 0; i < N * K; ++i) { fB[i] = (float)((rand() % 13) - 6) * 0.25f; hB[i] = __float2half(fB[i]); }
  for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += fA[r * K + k] * fB[n * K + k]; ref[r * N + n] = s; }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, out.size() * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, out.size() * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 20000);
  probe<<<2, 128, 20000>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("2-CTA UMMA M=256 N=128: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
  double mx = 0; int bad = 0;
  for (int i = 0; i < M * N; ++i) { double d = fabs(out[i] - ref[i]); mx = fmax(mx, d); bad += d > 1e-3; }
  printf("max|err| = %g, mismatches %d of %d ; D[0][0..2]=%g %g %g (ref %g %g %g) D[200][100]=%g (ref %g) D[5][70]=%g (ref %g)\n", mx, bad, M * N,
         out[0], out[1], out[2], ref[0], ref[1], ref[2], out[200 * N + 100], ref[200 * N + 100], out[5 * N + 70], ref[5 * N + 70]);
  return 0;
}

// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, fp16 -> fp32) as a function of N and of the operand pattern,
// one CTA per SM on all SMs (so shared-memory / tensor pipe behaviour is as in the real kernel).
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include "../mvsdf_b200/csrc/ptx.cuh"
using namespace mvsdf;

template <int N>
__global__ void __launch_bounds__(128, 1) rate(long long* out, int iters, int pattern) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s0 = ptx::smem_u32(smem);
  const uint32_t sA = s0, sB = s0 + 65536, sBar = s0 + 65536 + 131072, sT = sBar + 16;
  for (int i = threadIdx.x; i < (65536 + 131072) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(sBar, 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(sT, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 65536 + 131072 + 16);
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::idesc_f16_f32(128, N);
    long long t0 = clock64();
    if (pattern == 3) {
      // loop-invariant descriptors (hi/lo A tiles, hi/lo B buffers): the issue loop is just 3 UTCHMMAs
      const uint64_t da_hi = ptx::smem_desc(sA, 128, 512), da_lo = ptx::smem_desc(sA + 8192, 128, 512);
      const uint64_t db_hi = ptx::smem_desc(sB, 1040, 128), db_lo = ptx::smem_desc(sB + 49152, 1040, 128);
#pragma unroll 4
      for (int it = 0; it < iters; ++it) {
        ptx::umma_f16(tmem, da_hi, db_hi, idesc, 1u);
        ptx::umma_f16(tmem, da_lo, db_hi, idesc, 1u);
        ptx::umma_f16(tmem, da_hi, db_lo, idesc, 1u);
      }
    } else
    for (int it = 0; it < iters; ++it) {
      // pattern 0: one A tile, one B tile, same accumulator;  1: my 3-term pattern (A_hi,B_hi),(A_lo,B_hi),(A_hi,B_lo)
      // walking over 4 stages x 2 k-steps and the B buffer like the real kernel; 2: as 1 but 4 different accumulators
      const int stage = it & 3;
      const uint32_t a_hi = sA + stage * 16384, a_lo = a_hi + 8192;
      const uint32_t boff = (uint32_t)((it & 15) * 2 * 1040);
      const uint64_t da_hi = ptx::smem_desc(a_hi, 128, 512), da_lo = ptx::smem_desc(a_lo, 128, 512);
      const uint64_t db_hi = ptx::smem_desc(sB + boff, 1040, 128), db_lo = ptx::smem_desc(sB + 49152 + boff, 1040, 128);
      if (pattern == 0) {
        ptx::umma_f16(tmem, da_hi, db_hi, idesc, 1u);
        ptx::umma_f16(tmem, da_hi, db_hi, idesc, 1u);
        ptx::umma_f16(tmem, da_hi, db_hi, idesc, 1u);
      } else if (pattern == 1) {
        ptx::umma_f16(tmem, da_hi, db_hi, idesc, 1u);
        ptx::umma_f16(tmem, da_lo, db_hi, idesc, 1u);
        ptx::umma_f16(tmem, da_hi, db_lo, idesc, 1u);
      } else {
        const uint32_t d = tmem + (uint32_t)((it & 3) * N) % 512;
        ptx::umma_f16(d, da_hi, db_hi, idesc, 1u);
        ptx::umma_f16(d, da_lo, db_hi, idesc, 1u);
        ptx::umma_f16(d, da_hi, db_lo, idesc, 1u);
      }
    }
    ptx::umma_commit(sBar);
    ptx::mbar_wait(sBar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 512);
}

template <int N>
void run(long long* d, int pattern, const char* name) {
  const int iters = 4096;
  const int smem = 65536 + 131072 + 64;
  cudaFuncSetAttribute(rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate<N><<<148, 128, smem>>>(d, iters, pattern);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d %-28s : %s  %.1f cycles per UMMA (ideal %d)\n", N, name, cudaGetErrorString(e), (double)h / (3.0 * iters), N / 2);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  run<64>(d, 0, "same A,B");
  run<64>(d, 1, "3-term hi/lo, 1 accumulator");
  run<64>(d, 2, "3-term hi/lo, 4 accumulators");
  run<64>(d, 3, "3-term, invariant descriptors");
  run<32>(d, 3, "3-term, invariant descriptors");
  run<128>(d, 3, "3-term, invariant descriptors");
  run<256>(d, 3, "3-term, invariant descriptors");
  run<128>(d, 0, "same A,B");
  run<128>(d, 1, "3-term hi/lo (B 128 rows)");
  run<256>(d, 0, "same A,B");
  return 0;
}

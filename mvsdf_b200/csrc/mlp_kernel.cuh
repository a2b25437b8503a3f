// Fused MLP tile core for sm_100a: one persistent CTA per SM pushes tiles of 64 columns
// (64 points, or 16 points x [value, d/dx, d/dy, d/dz]) through every layer of a weight-norm
// MLP without the activations ever leaving the SM.
//
//   D[feature (TMEM lane), column] (+)= W_tile[feature, k] * X[column, k]        tcgen05.mma, M=128 N=64 K=16
//
//   * weights: streamed from L2 as pre-tiled fp16 hi/lo pairs by 1-D bulk async copies (TMA engine)
//     through a 4-stage mbarrier ring (warp 8);
//   * activations: resident in shared memory as fp16 hi/lo pairs, MN-major un-swizzled core matrices: a thread
//     of the epilogue owns one feature and consecutive columns, so it writes 16-byte vectors;
//   * 3 UMMAs per K step (hi*hi + lo*hi + hi*lo) give ~22-bit operands with fp32 accumulation in
//     TMEM -- the reference's fp32 SGEMM accuracy (SURVEY.md fact 0.9 rules out plain TF32/BF16);
//   * epilogue (warps 0-15): as soon as the UMMAs of one 128-row output tile have committed (one mbarrier
//     per tile) its accumulators go TMEM -> registers -> bias + softplus(beta=100)/ReLU -> fp16 hi/lo and
//     are parked in registers while the tensor core works on the next tiles; once the layer's last UMMA
//     has retired the parked values overwrite the activation buffer in place (it is the next layer's B
//     operand).  Forward-mode tangents ride along as extra columns (value+gradient mode) and are scaled
//     by sigmoid(100 z).
//
// Replaces, per call, the reference's embedder + 9 SGEMMs + softplus kernels
// (model/embedder.py:35, model/implicit_differentiable_renderer.py:77-94) and, in value+gradient mode,
// the autograd pass of ImplicitNetwork.gradient (:96-107); with NET_RENDER it is
// RenderingNetwork.forward (:145-167).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "netplan.h"
#include "ptx.cuh"

namespace mvsdf {

constexpr int kEpiWarps = 16;              // 4 TMEM lane quarters x 4 column groups of 16
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kMlpThreads = kEpiThreads + 64;   // + producer warp + MMA warp
constexpr int kPeCores = 8;                     // PE tile: K padded to 64
constexpr int kPeTileBytes = kPeCores * kBCoreStride;
constexpr int kScratchStride = 40;
constexpr int kMaxTiles = 4;                   // 128-row output tiles per layer (width <= 512)              // floats per point in the prologue scratch

struct MlpArgs {
  const uint8_t* packed;
  const float* bias;
  int n_run;                 // layers to run: hidden layers + one head
  int skip_layer;
  int skip_rows_begin;
  int pe_dim;
  int k_cores_max;
  int head;                  // HeadKind
  int feat_size;
  int debug;                 // bit0 skip weight copies, bit1 skip UMMAs, bit2 skip epilogue math (bottleneck isolation only)
  int feat_stride;           // row stride (floats) of `feats`; lets the render pass read full[:, 2:] in place
  int lp;                    // screening precision (pair2 kernel only): W_hi X_hi^T, see mlp_pair2_kernel.cuh
  int fuse_head;             // pair2 kernel, exact SDF-only head: the 1-row head is a dot product in the last hidden layer's epilogue
  long long n;               // number of points (ignored when n_ptr != nullptr)
  const int* n_ptr;          // optional device-side count
  const float* x;            // [n,3]
  const float* view;         // render: [n,3]
  const float* normals;      // render: [n,3]
  const float* feats;        // render: [n,feat_size]
  float* out_sdf;            // [n]            (optional with HEAD_FULL)
  float* out_full;           // [n, 2+feat]    (HEAD_FULL)
  float* out_grad;           // [n,3]          (value+gradient mode)
  float* out_rgb;            // [n,3]          (render)
  int* status;               // packed blob's status words (netplan.h kStatus*)
  // training forward (single-CTA kernel only): the input operand of every layer is also written to global memory for the
  // native backward (mlp_bwd_kernel.cuh).  save_off[l] = byte offset of the image of layer l's input inside `save`;
  // layout per layer: [tile][16-column slice][8-feature block][hi 2x128 B | lo 2x128 B] (save_addr below).
  uint8_t* save;
  long long save_off[kMaxLayers];
  unsigned long long* trace; // debug only: clock64 timeline of pair 0 (tools/diag_trace.py); nullptr in production
  LayerPlan L[10];
};

constexpr int kHeadPartBytes = 8 * kTileN * 4;     // fused SDF head (pair2 kernel): partial sums [rows' CTA][lane quarter][my 64 columns]
__host__ __device__ inline size_t mlp_smem_bytes(int k_cores_max) {
  return (size_t)kStages * kStageBytes + (size_t)k_cores_max * kBCoreStride + kPeTileBytes + 256 + kHeadPartBytes;
}

// Range monitor: activations are stored as fp16 hi/lo of kActScale * x, so |x| >= 1023 overflows to inf; the split then
// yields NaN (inf - inf), which poisons every accumulator of that column and surfaces at the head.  Checking the values
// the head writes therefore catches any overflow on the way at the cost of one predicate per OUTPUT element.
// (The ReLU of the rendering net is the exception -- fmaxf(NaN, 0) = 0 swallows the NaN -- so its pre-activations are
// checked too, and so is anything beyond the fp16 range on its way into the split: limit = 65504 / kActScale.)
__device__ __forceinline__ float checked(float v, int* status, float limit = 3.0e38f) {
  if (!(fabsf(v) <= limit) && status) atomicAdd(status + kStatusNonFinite, 1);
  return v;
}
constexpr float kActLimit = 65504.0f / kActScale;

// byte offset of (column n, feature k) inside an activation operand buffer
__device__ __forceinline__ uint32_t xoff(int n, int k) {
  return (uint32_t)((k >> 3) * kBCoreStride + (n >> 3) * 128 + (k & 7) * 16 + (n & 7) * 2);
}

// Saved-operand image ("K-sliced"): the same fp16 hi/lo pairs as the shared-memory operand, but grouped so that a 16-column
// slice of ALL features is contiguous -- read along the columns it is a K-major UMMA operand (8 rows x 16 B core matrices,
// LBO = 128, SBO = 512), which is what dW = dZ * H^T (contraction over the points) needs.  Byte offset of the 16-byte hi
// vector of (feature f, columns [8 cb, 8 cb + 8)) inside the image of one 64-column tile with `kc` feature blocks; the lo
// vector sits 256 B behind it.
__device__ __forceinline__ size_t save_addr(int kc, int f, int cb) {
  return (size_t)(cb >> 1) * (size_t)kc * 512 + (size_t)(f >> 3) * 512 + (size_t)(cb & 1) * 128 + (size_t)(f & 7) * 16;
}

__device__ __forceinline__ void store_split(uint32_t hi_addr, uint32_t lo_addr, float v) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  ptx::st_shared_u16(hi_addr, __half_as_ushort(h));
  ptx::st_shared_u16(lo_addr, __half_as_ushort(l));
}

// two activated values -> packed fp16 (hi, hi) and (lo, lo) registers: x = hi + lo to ~22 bits
__device__ __forceinline__ void pack_split(float y0, float y1, uint32_t& hi2, uint32_t& lo2) {
  const __half2 h = __floats2half2_rn(y0, y1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(y0 - hf.x, y1 - hf.y);
  hi2 = *reinterpret_cast<const uint32_t*>(&h);
  lo2 = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ float select_gt(float a, float thr, float if_true, float if_false) {
  // branch-free select; inline PTX so that ptxas cannot turn it into a divergent branch around the MUFU ops
  float r;
  asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\tselp.f32 %0, %3, %4, p;\n\t}" : "=f"(r) : "f"(a), "f"(thr), "f"(if_true), "f"(if_false));
  return r;
}

// kActScale * softplus(z; beta=100, threshold=20), as torch computes it: z if 100 z > 20 else log1p(exp(100 z))/100
// (nn.Softplus(beta=100), implicit_differentiable_renderer.py:75).  Straight-line code: ex2/lg2 on the SFU,
// overflow to +inf above the threshold is discarded by the select.
__device__ __forceinline__ float softplus100_scaled(float z) {
  const float tl = z * (100.0f * 1.4426950408889634f);                 // 100 z log2(e)
  const float e = ptx::ex2_approx(tl);
  const float y = ptx::lg2_approx(1.0f + e) * (kActScale * 0.6931471805599453f * 0.01f);
  return select_gt(tl, 20.0f * 1.4426950408889634f, z * kActScale, y);
}
// same, also returning d softplus / dz = sigmoid(100 z)
__device__ __forceinline__ float softplus100_scaled_grad(float z, float& sig) {
  const float tl = z * (100.0f * 1.4426950408889634f);
  const float e = ptx::ex2_approx(tl);
  const float ope = 1.0f + e;
  const float y = ptx::lg2_approx(ope) * (kActScale * 0.6931471805599453f * 0.01f);
  sig = select_gt(tl, 20.0f * 1.4426950408889634f, 1.0f, e * ptx::rcp_approx(ope));
  return select_gt(tl, 20.0f * 1.4426950408889634f, z * kActScale, y);
}

// ---- leaner epilogue math (pair2 kernel) ------------------------------------------------------------------
// softplus in the log2 domain without a select:  with t = 100 z log2(e),
//   kActScale * softplus(z; beta=100) = C * (max(t,0) + lg2(1 + 2^-|t|)),   C = kActScale ln(2)/100.
// Never overflows; for 100 z > 20 (torch's linear branch) the lg2 term is < 3e-9 and the result equals z to fp32
// rounding, so the threshold of nn.Softplus needs no special case.
constexpr float kSpT = 100.0f * 1.4426950408889634f;                  // t = kSpT * z
constexpr float kSpC = kActScale * 0.6931471805599453f * 0.01f;
// lg2(1 + e), e in (0, 1].  Default: the MUFU unit.  -DMVSDF_SOFTPLUS_POLY: e * P6(e) on the FMA pipe (minimax fit, max error
// 3.1e-7, i.e. 2e-9 on the activation -- below its fp32 rounding -- and exactly 0 at e = 0).  Round-2 experiment: the idea was
// that the epilogue stage E_0, which the tensor pipe waits for at every layer, is SFU-bound (two MUFU ops per element at
// 16 lanes / clock / SM).  Measured on the 378 880-point probe: 3.25 ms against 3.18 ms with MUFU, cfg2 step 615.8 ms against
// 606.9 ms -- the six extra FMAs cost more issue slots (shared with the UMMA issuer warp) than the MUFU op they replace.
// Rejected; kept for A/B.
#ifdef MVSDF_SOFTPLUS_POLY
__device__ __forceinline__ float lg2_1p(float e) {
  float p = fmaf(e, 0.0155299071f, -0.0795576957f);
  p = fmaf(p, e, 0.194294275f);
  p = fmaf(p, e, -0.325901943f);
  p = fmaf(p, e, 0.473553401f);
  p = fmaf(p, e, -0.720585467f);
  p = fmaf(p, e, 1.44266783f);
  return p * e;
}
#else
__device__ __forceinline__ float lg2_1p(float e) { return ptx::lg2_approx(1.0f + e); }
#endif
__device__ __forceinline__ float softplus_t_scaled(float t) {
  const float e = ptx::ex2_approx(-fabsf(t));
  return (fmaxf(t, 0.0f) + lg2_1p(e)) * kSpC;
}
// same, also returning d softplus / dz = sigmoid(100 z) = 1/(1+2^-t)
__device__ __forceinline__ float softplus_t_scaled_grad(float t, float& sig) {
  const float e = ptx::ex2_approx(-fabsf(t));
  const float r = ptx::rcp_approx(1.0f + e);
  sig = t >= 0.0f ? r : e * r;
  return (fmaxf(t, 0.0f) + lg2_1p(e)) * kSpC;
}
// hi/lo split with the residual computed by mixed-precision FMAs (y - float(h) = h * (-1) + y, SASS: FHFMA)
__device__ __forceinline__ void pack_split_fh(float y0, float y1, uint32_t& hi2, uint32_t& lo2) {
  float l0, l1;
  asm("{\n\t.reg .b16 hl, hh, m1;\n\t"
      "cvt.rn.f16x2.f32 %0, %4, %3;\n\t"
      "mov.b32 {hl, hh}, %0;\n\t"
      "mov.b16 m1, 0xBC00;\n\t"
      "fma.rn.f32.f16 %1, hl, m1, %3;\n\t"
      "fma.rn.f32.f16 %2, hh, m1, %4;\n\t}"
      : "=r"(hi2), "=f"(l0), "=f"(l1)
      : "f"(y0), "f"(y1));
  asm("cvt.rn.f16x2.f32 %0, %2, %1;" : "=r"(lo2) : "f"(l0), "f"(l1));
}

// screening-precision softplus: lg2(1 + e), e in (0, 1], as e * P4(e) (minimax, max error 1.6e-5 -- below the fp16 rounding
// of the stored activation) instead of the second MUFU op: the SFU (8 cycles per warp instruction per SM sub-partition)
// bounds the production epilogue at 16 cycles per element, this form is issue-bound at ~9.5 (tools/mufu_rate.cu)
__device__ __forceinline__ float softplus_t_scaled_screen(float t) {
  const float e = ptx::ex2_approx(-fabsf(t));
  float p = fmaf(e, 0.044938717f, -0.19310565f);
  p = fmaf(p, e, 0.41525051f);
  p = fmaf(p, e, -0.70899308f);
  p = fmaf(p, e, 1.4419080f);
  return fmaf(p, e, fmaxf(t, 0.0f)) * kSpC;
}

// fp16 hi parts only (screening precision)
__device__ __forceinline__ uint32_t pack_hi(float y0, float y1) {
  uint32_t h;
  asm("cvt.rn.f16x2.f32 %0, %2, %1;" : "=r"(h) : "f"(y0), "f"(y1));
  return h;
}

template <int KIND, int MODE, int CL>
__global__ void __launch_bounds__(kMlpThreads, 1) mlp_tile_kernel(const MlpArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t xbytes = (uint32_t)a.k_cores_max * kBCoreStride;
  const uint32_t s_stage = ptx::smem_u32(smem);
  const uint32_t s_xhi = s_stage + kStages * kStageBytes;   // activation operand [hi | lo] interleaved per K core
  const uint32_t s_xlo = s_xhi + kBLoOffset;
  const uint32_t s_pehi = s_xhi + xbytes;                    // positional-encoding operand, same structure
  const uint32_t s_pelo = s_pehi + kBLoOffset;
  const uint32_t s_bar = s_pehi + kPeTileBytes;              // 16-byte aligned by construction
  const uint32_t bar_full = s_bar;                       // kStages x 8 B
  const uint32_t bar_empty = s_bar + 8 * kStages;
  const uint32_t bar_acc = s_bar + 16 * kStages;         // kMaxTiles x 8 B, MMA -> epilogue: tile m accumulators complete
  const uint32_t bar_act = bar_acc + 8 * kMaxTiles;      // epilogue -> MMA: next B operand ready, TMEM drained
  const uint32_t s_tmem = bar_act + 8;
  uint8_t* const g_scratch = (KIND == NET_SDF) ? (smem + kStages * kStageBytes)
                                               : (smem + kStages * kStageBytes + xbytes);
  float* const scratch = reinterpret_cast<float*>(g_scratch);

  // (warp-shuffled so that the compiler can treat the trip counts below as warp-uniform)
  const int n_dev_count = a.n_ptr ? __shfl_sync(0xffffffffu, *a.n_ptr, 0) : 0;
  const long long n_pts = a.n_ptr ? (long long)n_dev_count : a.n;
  constexpr int kPtsPerTile = (MODE == 0) ? kTileN : kTileN / 4;
  const long long n_tiles = (n_pts + kPtsPerTile - 1) / kPtsPerTile;
  // A cluster of CL CTAs walks the tile list together (CTA rank r takes tile g*CL + r) so that all of them
  // consume the same weight stream: every stage is fetched from L2 once per cluster and multicast.
  const uint32_t crank = CL > 1 ? ptx::cluster_ctarank() : 0u;
  const long long n_groups = (n_tiles + CL - 1) / CL;
  const long long group0 = blockIdx.x / CL;
  const long long group_stride = gridDim.x / CL;
  constexpr uint16_t kClusterMask = (uint16_t)((1u << CL) - 1u);
  if (n_tiles == 0) return;      // empty request list: nothing to set up (uniform over the grid)

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, CL);
    }
    for (int m = 0; m < kMaxTiles; ++m) ptx::mbar_init(bar_acc + 8 * m, 1);
    ptx::mbar_init(bar_act, kEpiWarps);
    ptx::fence_mbar_init();
  }
  if (warp == kEpiWarps + 1) {
    ptx::tmem_alloc(s_tmem, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync();      // barrier inits visible cluster-wide before any remote arrive / multicast
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(smem + (s_tmem - s_stage)), 0);

  if (warp == kEpiWarps) {
    // ------------------------------------------------------------------ weight producer
    uint32_t it = 0;
    constexpr uint32_t kSlice = kStageBytes / CL;
    for (long long g = group0; g < n_groups; g += group_stride) {
      for (int l = 0; l < a.n_run; ++l) {
        const LayerPlan& lp = a.L[l];
        const uint8_t* src = a.packed + lp.w_off;
        const int n_stage = lp.m_tiles * lp.k_chunks;
        for (int i = 0; i < n_stage; ++i, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1);     // every CTA of the cluster has drained this stage
          if (lane == 0) {
            if (a.debug & 1) {
              ptx::mbar_arrive(bar_full + 8 * s);
            } else {
              ptx::mbar_arrive_expect_tx(bar_full + 8 * s, kStageBytes);
              if (CL == 1)
                ptx::bulk_g2s(s_stage + s * kStageBytes, src + (size_t)i * kStageBytes, kStageBytes, bar_full + 8 * s);
              else
                ptx::bulk_g2s_multicast(s_stage + s * kStageBytes + crank * kSlice,
                                        src + (size_t)i * kStageBytes + crank * kSlice, kSlice, bar_full + 8 * s,
                                        kClusterMask);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ------------------------------------------------------------------ UMMA issuer
    // Per 16-wide K step:  D[:, 0:128] += W_hi * [X_hi ; X_lo]^T   (N = 128: full tensor rate)
    //                      D[:, 0:64]  += W_lo * X_hi^T            (N = 64)
    // the epilogue adds the two 64-column halves, giving W_hi X_hi + W_lo X_hi + W_hi X_lo.
    constexpr uint32_t idesc128 = ptx::idesc_f16_f32_bmn(kTileM, 2 * kTileN);
    constexpr uint32_t idesc64 = ptx::idesc_f16_f32_bmn(kTileM, kTileN);
    const bool leader = ptx::elect_one();
    uint32_t it = 0, act_ctr = 0;
    for (long long g = group0; g < n_groups; g += group_stride) {
      for (int l = 0; l < a.n_run; ++l) {
        const LayerPlan& lp = a.L[l];
        ptx::mbar_wait(bar_act, act_ctr & 1);
        ++act_ctr;
        ptx::tc_fence_after();
        const uint32_t b_base = lp.b_from_pe ? s_pehi : s_xhi;
        for (int m = 0; m < lp.m_tiles; ++m) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(m * 2 * kTileN);
          for (int kc = 0; kc < lp.k_chunks; ++kc, ++it) {
            const uint32_t s = it % kStages, ph = (it / kStages) & 1;
            ptx::mbar_wait(bar_full + 8 * s, ph);
            ptx::tc_fence_after();
            const uint32_t a_hi = s_stage + s * kStageBytes;
            const uint32_t a_lo = a_hi + kTileBytes;
#pragma unroll
            for (int ks = 0; ks < kChunkK / 16; ++ks) {
              const uint64_t da_hi = ptx::smem_desc(a_hi + ks * 256, 128, 512);
              const uint64_t da_lo = ptx::smem_desc(a_lo + ks * 256, 128, 512);
              const uint64_t db = ptx::smem_desc(b_base + (uint32_t)((kc * (kChunkK / 8) + ks * 2) * kBCoreStride),
                                                 kBCoreStride, 128);
              if (leader && !(a.debug & 2)) {
                ptx::umma_f16(d_tmem, da_hi, db, idesc128, (kc | ks) != 0 ? 1u : 0u);
                ptx::umma_f16(d_tmem, da_lo, db, idesc64, 1u);
              }
            }
            // frees the stage (in every CTA of the cluster) once these UMMAs have read it
            if (leader) {
              if (CL == 1) ptx::umma_commit(bar_empty + 8 * s);
              else ptx::umma_commit_multicast(bar_empty + 8 * s, kClusterMask);
            }
            __syncwarp();
          }
          // this tile's accumulators are final: the epilogue may start on it while the next tile is computed
          if (leader) ptx::umma_commit(bar_acc + 8 * m);
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ prologue + epilogue warps
    const int q = warp & 3;              // TMEM lane quarter this warp may read
    const int cg = warp >> 2;            // column group: columns 16*cg .. 16*cg+15
    const int row = q * 32 + lane;       // row inside a 128-row output tile
    const int t = threadIdx.x;           // 0..kEpiThreads-1
    constexpr float kInvScale = 1.0f / (kWeightScale * kActScale);
    uint32_t acc_ctr[kMaxTiles] = {0, 0, 0, 0};

    for (long long g = group0; g < n_groups; g += group_stride) {
      const long long tile = g * CL + crank;            // tiles past the end run on zero points, outputs are guarded
      const long long p0 = tile * kPtsPerTile;

      // ---------------- prologue: build the first layer's B operand
      if (KIND == NET_SDF) {
        // phase A: positional encoding (model/embedder.py:5-50) of 64 (16) points, one (point, coordinate) per thread
        if (t < 3 * kPtsPerTile) {
          const int pt = t / 3, c = t - 3 * pt;
          const long long gp = p0 + pt;
          const float xc = gp < n_pts ? __ldg(a.x + gp * 3 + c) : 0.0f;
          float* pe = scratch + pt * kScratchStride;
          pe[c] = xc;
          float* dpe = scratch + (kPtsPerTile + pt) * kScratchStride;
          if (MODE == 1) dpe[c] = 1.0f;
          float f = 1.0f;
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            float sn, cs;
            sincosf(xc * f, &sn, &cs);
            pe[3 + 6 * i + c] = sn;
            pe[6 + 6 * i + c] = cs;
            if (MODE == 1) {
              dpe[3 + 6 * i + c] = f * cs;
              dpe[6 + 6 * i + c] = -f * sn;
            }
            f *= 2.0f;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        // phase B: scale, split to fp16 hi/lo, scatter into the PE operand tile (64 columns x K=64)
        for (int cidx = t; cidx < kTileN * 64; cidx += kEpiThreads) {
          const int col = cidx >> 6, k = cidx & 63;
          float v = 0.0f;
          if (k < a.pe_dim) {
            if (MODE == 0) {
              v = scratch[col * kScratchStride + k];
            } else {
              const int pt = col >> 2, j = col & 3;
              const int coord = k < 3 ? k : (k - 3) % 3;
              v = j == 0 ? scratch[pt * kScratchStride + k]
                         : (coord == j - 1 ? scratch[(kPtsPerTile + pt) * kScratchStride + k] : 0.0f);
            }
          }
          const uint32_t o = xoff(col, k);
          store_split(s_pehi + o, s_pelo + o, v * kActScale);
          if (a.save) {
            const __half h = __float2half_rn(v * kActScale);
            const __half lo = __float2half_rn(v * kActScale - __half2float(h));
            uint8_t* g = a.save + a.save_off[0] + (size_t)tile * (kPeCores * kBCoreStride) + save_addr(kPeCores, k, col >> 3) + (col & 7) * 2;
            *reinterpret_cast<__half*>(g) = h;
            *reinterpret_cast<__half*>(g + 256) = lo;
          }
        }
      } else {
        // render input [points(3), PE4(view)(27), normals(3), features(F)]  (implicit_differentiable_renderer.py:150)
        if (t < kTileN) {
          const long long gp = p0 + t;
          float* pe = scratch + t * kScratchStride;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float vc = gp < n_pts ? __ldg(a.view + gp * 3 + c) : 0.0f;
            pe[c] = vc;
            float f = 1.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float sn, cs;
              sincosf(vc * f, &sn, &cs);
              pe[3 + 6 * i + c] = sn;
              pe[6 + 6 * i + c] = cs;
              f *= 2.0f;
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        const int kpad = a.L[0].k_chunks * kChunkK;
        const int F = a.feat_size;
        for (int cidx = t; cidx < kTileN * kpad; cidx += kEpiThreads) {
          const int col = cidx / kpad, k = cidx - col * kpad;
          const long long gp = p0 + col;
          float v = 0.0f;
          if (gp < n_pts) {
            if (k < 3) v = __ldg(a.x + gp * 3 + k);
            else if (k < 30) v = scratch[col * kScratchStride + (k - 3)];
            else if (k < 33) v = __ldg(a.normals + gp * 3 + (k - 30));
            else if (k < 33 + F) v = __ldg(a.feats + gp * a.feat_stride + (k - 33));
          }
          const uint32_t o = xoff(col, k);
          store_split(s_xhi + o, s_xlo + o, v * kActScale);
          if (a.save) {
            const int kc0 = kpad >> 3;
            const __half h = __float2half_rn(v * kActScale);
            const __half lo = __float2half_rn(v * kActScale - __half2float(h));
            uint8_t* g = a.save + a.save_off[0] + (size_t)tile * ((size_t)kc0 * kBCoreStride) + save_addr(kc0, k, col >> 3) + (col & 7) * 2;
            *reinterpret_cast<__half*>(g) = h;
            *reinterpret_cast<__half*>(g + 256) = lo;
          }
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_act);

      // ---------------- layers
      for (int l = 0; l < a.n_run; ++l) {
        const LayerPlan& lp = a.L[l];
        const bool last = (l == a.n_run - 1);
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 16);
        const bool skip_src = KIND == NET_SDF && l == a.skip_layer - 1;
        float bias_r[kMaxTiles];          // fetched before the waits so the global-load latency hides behind the UMMAs
#pragma unroll
        for (int m = 0; m < kMaxTiles; ++m) bias_r[m] = m < lp.m_tiles ? __ldg(a.bias + lp.bias_off + m * kTileM + row) : 0.f;
        // activated outputs of the layer, fp16 hi/lo, two columns per register; parked until the layer's UMMAs retire
        uint32_t phi[kMaxTiles][8], plo[kMaxTiles][8];
#pragma unroll
        for (int m = 0; m < kMaxTiles; ++m) {
          if (m < lp.m_tiles) {
            ptx::mbar_wait(bar_acc + 8 * m, acc_ctr[m] & 1);
            ++acc_ctr[m];
            ptx::tc_fence_after();
            uint32_t v[16], v2[16];
            ptx::tmem_ld_32x16(t_row + (uint32_t)(m * 2 * kTileN), v);
            ptx::tmem_ld_32x16(t_row + (uint32_t)(m * 2 * kTileN + kTileN), v2);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
            const int f = m * kTileM + row;
            const float bias = bias_r[m];
            if (!last) {
              if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float z0 = fmaf(__uint_as_float(v[2 * i]), kInvScale, bias);
                  const float z1 = fmaf(__uint_as_float(v[2 * i + 1]), kInvScale, bias);
                  if (KIND == NET_RENDER) {       // fmaxf(NaN, 0) = 0 would swallow an overflow of the previous layer
                    checked(z0, a.status, kActLimit);
                    checked(z1, a.status, kActLimit);
                  }
                  const float y0 = (KIND == NET_SDF) ? softplus100_scaled(z0) : fmaxf(z0, 0.0f) * kActScale;
                  const float y1 = (KIND == NET_SDF) ? softplus100_scaled(z1) : fmaxf(z1, 0.0f) * kActScale;
                  pack_split(y0, y1, phi[m][i], plo[m][i]);
                }
              } else {
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                  const float z = fmaf(__uint_as_float(v[4 * gq]), kInvScale, bias);
                  float sg;
                  const float y = softplus100_scaled_grad(z, sg);
                  const float ts = sg * (kInvScale * kActScale);
                  pack_split(y, __uint_as_float(v[4 * gq + 1]) * ts, phi[m][2 * gq], plo[m][2 * gq]);
                  pack_split(__uint_as_float(v[4 * gq + 2]) * ts, __uint_as_float(v[4 * gq + 3]) * ts, phi[m][2 * gq + 1],
                             plo[m][2 * gq + 1]);
                }
              }
            } else {
              // ---------------- head: write results to global memory
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int col = cg * 16 + j;
                const float acc = __uint_as_float(v[j]) * kInvScale;
                if (KIND == NET_RENDER) {
                  const long long gp = p0 + col;
                  if (row < 3 && gp < n_pts) a.out_rgb[gp * 3 + row] = tanhf(checked(acc + bias, a.status));
                } else {
                  const long long gp = (MODE == 0) ? p0 + col : p0 + (col >> 2);
                  const int jj = (MODE == 0) ? 0 : (col & 3);
                  if (gp < n_pts) {
                    if (a.head == HEAD_SDF_ONLY) {
                      if (row == 0) {
                        if (jj == 0) a.out_sdf[gp] = checked(acc + bias, a.status);
                        else a.out_grad[gp * 3 + jj - 1] = checked(acc, a.status);
                      }
                    } else {
                      const int F = a.feat_size;
                      if (jj == 0) {
                        if (f < F) a.out_full[gp * (F + 2) + 2 + f] = acc + bias;
                        else if (f < F + 2) {
                          a.out_full[gp * (F + 2) + (f - F)] = checked(acc + bias, a.status);
                          if (f == F && a.out_sdf) a.out_sdf[gp] = acc + bias;
                        }
                      } else if (f == F) {
                        a.out_grad[gp * 3 + jj - 1] = checked(acc, a.status);
                      }
                    }
                  }
                }
              }
            }
          }
        }
        if (!last) {
          // every UMMA of this layer has retired (commits complete in order): overwrite the activation buffer in place
#pragma unroll
          for (int m = 0; m < kMaxTiles; ++m) {
            if (m < lp.m_tiles) {
              const int f = m * kTileM + row;
              const uint32_t o0 = xoff(cg * 16, f);          // my 16 columns = two 16-byte vectors (8 columns each)
              const int kc_next = a.L[l + 1].k_chunks * (kChunkK / 8);
              uint8_t* gsave = (a.save && f < kc_next * 8)
                                   ? a.save + a.save_off[l + 1] + (size_t)tile * ((size_t)kc_next * kBCoreStride) + save_addr(kc_next, f, cg * 2)
                                   : nullptr;
              if (skip_src && f >= a.skip_rows_begin) {
                // rows that hold the skip connection: copy PE (already scaled & split) instead of softplus
                const int k = f - a.skip_rows_begin;
                const uint32_t s0 = xoff(cg * 16, k);
                const bool real = k < a.pe_dim;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  uint4 vh = make_uint4(0, 0, 0, 0), vl = make_uint4(0, 0, 0, 0);
                  if (real) {
                    vh = ptx::ld_shared_v4(s_pehi + s0 + j * 128);
                    vl = ptx::ld_shared_v4(s_pelo + s0 + j * 128);
                  }
                  ptx::st_shared_v4(s_xhi + o0 + j * 128, vh.x, vh.y, vh.z, vh.w);
                  ptx::st_shared_v4(s_xlo + o0 + j * 128, vl.x, vl.y, vl.z, vl.w);
                  if (gsave) {
                    *reinterpret_cast<uint4*>(gsave + j * 128) = vh;
                    *reinterpret_cast<uint4*>(gsave + 256 + j * 128) = vl;
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  ptx::st_shared_v4(s_xhi + o0 + j * 128, phi[m][4 * j], phi[m][4 * j + 1], phi[m][4 * j + 2], phi[m][4 * j + 3]);
                  ptx::st_shared_v4(s_xlo + o0 + j * 128, plo[m][4 * j], plo[m][4 * j + 1], plo[m][4 * j + 2], plo[m][4 * j + 3]);
                  if (gsave) {
                    *reinterpret_cast<uint4*>(gsave + j * 128) = make_uint4(phi[m][4 * j], phi[m][4 * j + 1], phi[m][4 * j + 2], phi[m][4 * j + 3]);
                    *reinterpret_cast<uint4*>(gsave + 256 + j * 128) = make_uint4(plo[m][4 * j], plo[m][4 * j + 1], plo[m][4 * j + 2], plo[m][4 * j + 3]);
                  }
                }
              }
            }
          }
          ptx::tc_fence_before();
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_act);
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync();      // no CTA may exit while peers can still multicast into it
  if (warp == kEpiWarps + 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace mvsdf

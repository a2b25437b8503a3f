// CTA-pair variant of the fused MLP tile core: two SMs (a cluster of 2) share every UMMA through
// tcgen05 cta_group::2.  A pair processes 128 columns per tile (64 per CTA); every weight tile is 256 output rows,
// each CTA streams only its own 128-row half, so the per-SM weight stream and the per-SM shared-memory operand
// traffic are half those of the single-CTA kernel for the same columns.
//
//   D[256 features (128 TMEM lanes per CTA), 128 columns]:
//       D_a += W_hi [X_hi]^T      D_b += W_hi [X_lo]^T      D_a += W_lo [X_hi]^T          (3 UMMAs, M=256 N=128 K=16)
//   X_hi / X_lo rows (columns of the tile) 0-63 come from CTA 0's shared memory, 64-127 from CTA 1's.
//
// The epilogue of CTA r owns features [256 mp + 128 r, +128) for all 128 columns: the half that belongs to the peer's
// columns is written straight into the peer's shared memory (st.shared::cluster).  One leader thread (CTA 0) issues
// all UMMAs; CTA 1's otherwise idle MMA warp relays "my weight stage has landed" to the leader.
// Same numerics, same packed weights and same epilogue math as mlp_kernel.cuh.
#pragma once
#include "mlp_kernel.cuh"

namespace mvsdf {

constexpr int kPairTiles = 2;          // 256-row tiles per layer (width <= 512)
constexpr int kPairCols = 2 * kTileN;  // 128 columns per pair tile

template <int KIND, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1) mlp_pair_kernel(const MlpArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = ptx::cluster_ctarank();
  const uint32_t xbytes = (uint32_t)a.k_cores_max * kBCoreStride;
  const uint32_t s_stage = ptx::smem_u32(smem);
  const uint32_t s_xhi = s_stage + kStages * kStageBytes;
  const uint32_t s_xlo = s_xhi + kBLoOffset;
  const uint32_t s_pehi = s_xhi + xbytes;
  const uint32_t s_pelo = s_pehi + kBLoOffset;
  const uint32_t s_bar = s_pehi + kPeTileBytes;
  const uint32_t bar_full = s_bar;                        // kStages: my weight half has landed
  const uint32_t bar_full2 = s_bar + 8 * kStages;         // kStages: (leader only) the peer's half has landed
  const uint32_t bar_empty = s_bar + 16 * kStages;        // kStages: the UMMAs reading the stage retired (both CTAs)
  const uint32_t bar_acc = s_bar + 24 * kStages;          // kPairTiles: accumulators of pair tile mp complete
  const uint32_t bar_act = bar_acc + 8 * kPairTiles;      // (leader only) next B operand ready in BOTH CTAs, TMEM drained
  const uint32_t s_tmem = bar_act + 8;
  uint8_t* const g_scratch = (KIND == NET_SDF) ? (smem + kStages * kStageBytes) : (smem + kStages * kStageBytes + xbytes);
  float* const scratch = reinterpret_cast<float*>(g_scratch);

  const int n_dev_count = a.n_ptr ? __shfl_sync(0xffffffffu, *a.n_ptr, 0) : 0;
  const long long n_pts = a.n_ptr ? (long long)n_dev_count : a.n;
  constexpr int kPtsPerCta = (MODE == 0) ? kTileN : kTileN / 4;       // points per CTA per tile
  const long long n_tiles = (n_pts + 2 * kPtsPerCta - 1) / (2 * kPtsPerCta);
  const long long pair0 = blockIdx.x >> 1;
  const long long pair_stride = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_full2 + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int m = 0; m < kPairTiles; ++m) ptx::mbar_init(bar_acc + 8 * m, 1);
    ptx::mbar_init(bar_act, 2 * kEpiWarps);
    ptx::fence_mbar_init();
  }
  if (warp == kEpiWarps + 1) {
    ptx::tmem_alloc_2cta(s_tmem, kTmemCols);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(smem + (s_tmem - s_stage)), 0);

  if (warp == kEpiWarps) {
    // ------------------------------------------------------------------ weight producer: my 128-row half of every 256-row tile
    uint32_t it = 0;
    for (long long g = pair0; g < n_tiles; g += pair_stride) {
      for (int l = 0; l < a.n_run; ++l) {
        const LayerPlan& lp = a.L[l];
        const int n_pair_tiles = (lp.m_tiles + 1) >> 1;
        for (int mp = 0; mp < n_pair_tiles; ++mp) {
          const int m = 2 * mp + (int)crank;
          const bool have = m < lp.m_tiles;
          const uint8_t* src = a.packed + lp.w_off + (size_t)m * lp.k_chunks * kStageBytes;
          for (int kc = 0; kc < lp.k_chunks; ++kc, ++it) {
            const uint32_t s = it % kStages, ph = (it / kStages) & 1;
            ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1);
            if (lane == 0) {
              if (have && !(a.debug & 1)) {
                ptx::mbar_arrive_expect_tx(bar_full + 8 * s, kStageBytes);
                ptx::bulk_g2s(s_stage + s * kStageBytes, src + (size_t)kc * kStageBytes, kStageBytes, bar_full + 8 * s);
              } else {
                ptx::mbar_arrive(bar_full + 8 * s);     // odd tile count: this half multiplies stale data into rows nobody reads
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    if (crank == 0) {
      // ------------------------------------------------------------------ UMMA issuer (leader CTA)
      constexpr uint32_t idesc = ptx::idesc_f16_f32_bmn(2 * kTileM, kPairCols);
      const bool leader = ptx::elect_one();
      uint32_t it = 0, act_ctr = 0;
      for (long long g = pair0; g < n_tiles; g += pair_stride) {
        for (int l = 0; l < a.n_run; ++l) {
          const LayerPlan& lp = a.L[l];
          const int n_pair_tiles = (lp.m_tiles + 1) >> 1;
          ptx::mbar_wait(bar_act, act_ctr & 1);
          ++act_ctr;
          ptx::tc_fence_after();
          const uint32_t b_base = lp.b_from_pe ? s_pehi : s_xhi;
          for (int mp = 0; mp < n_pair_tiles; ++mp) {
            const uint32_t d_a = tmem_base + (uint32_t)(mp * 2 * kPairCols);
            const uint32_t d_b = d_a + kPairCols;
            for (int kc = 0; kc < lp.k_chunks; ++kc, ++it) {
              const uint32_t s = it % kStages, ph = (it / kStages) & 1;
              ptx::mbar_wait(bar_full + 8 * s, ph);
              ptx::mbar_wait(bar_full2 + 8 * s, ph);
              ptx::tc_fence_after();
              const uint32_t a_hi = s_stage + s * kStageBytes;
              const uint32_t a_lo = a_hi + kTileBytes;
#pragma unroll
              for (int ks = 0; ks < kChunkK / 16; ++ks) {
                const uint64_t da_hi = ptx::smem_desc(a_hi + ks * 256, 128, 512);
                const uint64_t da_lo = ptx::smem_desc(a_lo + ks * 256, 128, 512);
                const uint32_t boff = (uint32_t)((kc * (kChunkK / 8) + ks * 2) * kBCoreStride);
                const uint64_t db_hi = ptx::smem_desc(b_base + boff, kBCoreStride, 128);
                const uint64_t db_lo = ptx::smem_desc(b_base + kBLoOffset + boff, kBCoreStride, 128);
                if (leader && !(a.debug & 2)) {
                  const uint32_t acc = (kc | ks) != 0 ? 1u : 0u;
                  ptx::umma_f16_2cta(d_a, da_hi, db_hi, idesc, acc);
                  ptx::umma_f16_2cta(d_b, da_hi, db_lo, idesc, acc);
                  ptx::umma_f16_2cta(d_a, da_lo, db_hi, idesc, 1u);
                }
              }
              if (leader) ptx::umma_commit_2cta(bar_empty + 8 * s, 3);
              __syncwarp();
            }
            if (leader) ptx::umma_commit_2cta(bar_acc + 8 * mp, 3);
            __syncwarp();
          }
        }
      }
    } else {
      // ------------------------------------------------------------------ peer relay: tell the leader my stage has landed
      const uint32_t remote_full2 = ptx::mapa(bar_full2, 0);
      uint32_t it = 0;
      for (long long g = pair0; g < n_tiles; g += pair_stride) {
        for (int l = 0; l < a.n_run; ++l) {
          const LayerPlan& lp = a.L[l];
          const int n_stage = ((lp.m_tiles + 1) >> 1) * lp.k_chunks;
          for (int i = 0; i < n_stage; ++i, ++it) {
            const uint32_t s = it % kStages, ph = (it / kStages) & 1;
            ptx::mbar_wait(bar_full + 8 * s, ph);
            if (lane == 0) ptx::mbar_arrive_remote_relaxed(remote_full2 + 8 * s);
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ prologue + epilogue warps
    const int q = warp & 3;              // TMEM lane quarter
    const int cg = warp >> 2;            // 16-column group: this warp produces columns [16cg,+16) of MY 64 columns
                                         // (chunk 0, local stores) and of the PEER's 64 columns (chunk 1, DSMEM stores),
                                         // so that every warp carries the same share of remote traffic
    const int row = q * 32 + lane;       // row inside my 128-row half
    const int t = threadIdx.x;
    const int lcol0 = cg * 16;           // first column inside the destination CTA's 64
    // base addresses of the peer's activation buffer in the cluster window
    const uint32_t dst_xhi = ptx::mapa(s_xhi, crank ^ 1u);
    const uint32_t dst_xlo = dst_xhi + kBLoOffset;
    const uint32_t remote_act = ptx::mapa(bar_act, 0);
    constexpr float kInvScale = 1.0f / (kWeightScale * kActScale);
    uint32_t acc_ctr[kPairTiles] = {0, 0};

    for (long long g = pair0; g < n_tiles; g += pair_stride) {
      const long long p0 = g * (2 * kPtsPerCta) + (long long)crank * kPtsPerCta;      // my first point
      const long long p0_pair = g * (2 * kPtsPerCta);

      // ---------------- prologue: first layer's B operand for MY 64 columns (identical to the single-CTA kernel)
      if (KIND == NET_SDF) {
        if (t < 3 * kPtsPerCta) {
          const int pt = t / 3, c = t - 3 * pt;
          const long long gp = p0 + pt;
          const float xc = gp < n_pts ? __ldg(a.x + gp * 3 + c) : 0.0f;
          float* pe = scratch + pt * kScratchStride;
          pe[c] = xc;
          float* dpe = scratch + (kPtsPerCta + pt) * kScratchStride;
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
                         : (coord == j - 1 ? scratch[(kPtsPerCta + pt) * kScratchStride + k] : 0.0f);
            }
          }
          const uint32_t o = xoff(col, k);
          store_split(s_pehi + o, s_pelo + o, v * kActScale);
        }
      } else {
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
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_all();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(remote_act);      // release.cluster: publishes the warp's (remote) writes

      // ---------------- layers
      for (int l = 0; l < a.n_run; ++l) {
        const LayerPlan& lp = a.L[l];
        const bool last = (l == a.n_run - 1);
        const int n_pair_tiles = (lp.m_tiles + 1) >> 1;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool skip_src = KIND == NET_SDF && l == a.skip_layer - 1;
        float bias_r[kPairTiles];
#pragma unroll
        for (int mp = 0; mp < kPairTiles; ++mp) {
          const int m = 2 * mp + (int)crank;
          bias_r[mp] = m < lp.m_tiles ? __ldg(a.bias + lp.bias_off + m * kTileM + row) : 0.f;
        }
        uint32_t phi[kPairTiles][16], plo[kPairTiles][16];
#pragma unroll
        for (int mp = 0; mp < kPairTiles; ++mp) {
          if (mp < n_pair_tiles) {
            ptx::mbar_wait(bar_acc + 8 * mp, acc_ctr[mp] & 1);
            ++acc_ctr[mp];
            ptx::tc_fence_after();
            const int m = 2 * mp + (int)crank;
            const int f = m * kTileM + row;                 // feature (output row) this thread owns
            const bool have = m < lp.m_tiles;
            const float bias = bias_r[mp];
#pragma unroll
            for (int hcol = 0; hcol < 2; ++hcol) {           // two 16-column halves of my 32 columns
              // chunk 0: my own columns, chunk 1: the peer's columns (columns 0-63 of the tile are CTA 0's)
              const uint32_t dest_h = hcol == 0 ? crank : (crank ^ 1u);
              const uint32_t tcol = (uint32_t)(mp * 2 * kPairCols) + dest_h * kTileN + (uint32_t)lcol0;
              uint32_t v[16], v2[16];
              ptx::tmem_ld_32x16(t_row + tcol, v);
              ptx::tmem_ld_32x16(t_row + tcol + kPairCols, v2);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
              if (!last) {
                if (MODE == 0) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float z0 = fmaf(__uint_as_float(v[2 * i]), kInvScale, bias);
                    const float z1 = fmaf(__uint_as_float(v[2 * i + 1]), kInvScale, bias);
                    const float y0 = (KIND == NET_SDF) ? softplus100_scaled(z0) : fmaxf(z0, 0.0f) * kActScale;
                    const float y1 = (KIND == NET_SDF) ? softplus100_scaled(z1) : fmaxf(z1, 0.0f) * kActScale;
                    pack_split(y0, y1, phi[mp][hcol * 8 + i], plo[mp][hcol * 8 + i]);
                  }
                } else {
#pragma unroll
                  for (int gq = 0; gq < 4; ++gq) {
                    const float z = fmaf(__uint_as_float(v[4 * gq]), kInvScale, bias);
                    float sg;
                    const float y = softplus100_scaled_grad(z, sg);
                    const float ts = sg * (kInvScale * kActScale);
                    pack_split(y, __uint_as_float(v[4 * gq + 1]) * ts, phi[mp][hcol * 8 + 2 * gq], plo[mp][hcol * 8 + 2 * gq]);
                    pack_split(__uint_as_float(v[4 * gq + 2]) * ts, __uint_as_float(v[4 * gq + 3]) * ts,
                               phi[mp][hcol * 8 + 2 * gq + 1], plo[mp][hcol * 8 + 2 * gq + 1]);
                  }
                }
              } else if (have) {
                // ---------------- head: write results to global memory
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int ccta = (int)dest_h, lc = lcol0 + j;   // owner CTA of the column and its index there
                  const float acc = __uint_as_float(v[j]) * kInvScale;
                  if (KIND == NET_RENDER) {
                    const long long gp = p0_pair + (long long)ccta * kPtsPerCta + lc;
                    if (row < 3 && m == 0 && gp < n_pts) a.out_rgb[gp * 3 + row] = tanhf(acc + bias);
                  } else {
                    const long long gp = p0_pair + (long long)ccta * kPtsPerCta + ((MODE == 0) ? lc : (lc >> 2));
                    const int jj = (MODE == 0) ? 0 : (lc & 3);
                    if (gp < n_pts) {
                      if (a.head == HEAD_SDF_ONLY) {
                        if (f == 0) {
                          if (jj == 0) a.out_sdf[gp] = acc + bias;
                          else a.out_grad[gp * 3 + jj - 1] = acc;
                        }
                      } else {
                        const int F = a.feat_size;
                        if (jj == 0) {
                          if (f < F) a.out_full[gp * (F + 2) + 2 + f] = acc + bias;
                          else if (f < F + 2) {
                            a.out_full[gp * (F + 2) + (f - F)] = acc + bias;
                            if (f == F && a.out_sdf) a.out_sdf[gp] = acc + bias;
                          }
                        } else if (f == F) {
                          a.out_grad[gp * 3 + jj - 1] = acc;
                        }
                      }
                    }
                  }
                }
              }
            }
          }
        }
        if (!last) {
          // all UMMAs of the layer have retired in both CTAs: overwrite the activation operands in place,
          // my rows for the peer's columns go straight into the peer's shared memory
#pragma unroll
          for (int mp = 0; mp < kPairTiles; ++mp) {
            if (mp < n_pair_tiles) {
              const int m = 2 * mp + (int)crank;
              const int f = m * kTileM + row;
              const bool write = m < lp.m_tiles && !(skip_src && f >= a.skip_rows_begin);
              if (write) {
                const uint32_t o0 = xoff(lcol0, f);          // 16 local + 16 remote columns = 2 + 2 16-byte vectors (hi, lo each)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  ptx::st_shared_v4(s_xhi + o0 + j * 128, phi[mp][4 * j], phi[mp][4 * j + 1], phi[mp][4 * j + 2], phi[mp][4 * j + 3]);
                  ptx::st_shared_v4(s_xlo + o0 + j * 128, plo[mp][4 * j], plo[mp][4 * j + 1], plo[mp][4 * j + 2], plo[mp][4 * j + 3]);
                  ptx::st_cluster_v4(dst_xhi + o0 + j * 128, phi[mp][8 + 4 * j], phi[mp][8 + 4 * j + 1], phi[mp][8 + 4 * j + 2], phi[mp][8 + 4 * j + 3]);
                  ptx::st_cluster_v4(dst_xlo + o0 + j * 128, plo[mp][8 + 4 * j], plo[mp][8 + 4 * j + 1], plo[mp][8 + 4 * j + 2], plo[mp][8 + 4 * j + 3]);
                }
              }
            }
          }
          if (skip_src) {
            // skip connection: features [skip_rows_begin, +pe_dim) of MY 64 columns are the positional encoding
            // (already scaled and split in my PE tile) -- each CTA fills them for its own columns
            for (int cidx = t; cidx < a.pe_dim * 16; cidx += kEpiThreads) {
              const int k = cidx >> 4, blk = cidx & 15;           // 16 column blocks of 16 bytes per feature: 8 hi + 8 lo
              const int kd = a.skip_rows_begin + k;
              const uint4 v = ptx::ld_shared_v4(s_pehi + (uint32_t)((k >> 3) * kBCoreStride + (k & 7) * 16 + blk * 128));
              ptx::st_shared_v4(s_xhi + (uint32_t)((kd >> 3) * kBCoreStride + (kd & 7) * 16 + blk * 128), v.x, v.y, v.z, v.w);
            }
          }
          ptx::tc_fence_before();
          ptx::fence_proxy_async_all();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(remote_act);
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();      // no CTA may exit (or free TMEM) while its peer can still write into it
  if (warp == kEpiWarps + 1) ptx::tmem_dealloc_2cta(tmem_base, kTmemCols);
}

}  // namespace mvsdf

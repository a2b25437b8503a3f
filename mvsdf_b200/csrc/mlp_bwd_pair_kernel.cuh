// CTA-pair version of the reverse sweep (mlp_bwd_kernel.cuh): the split-K pipelined cta_group::2 tile core of
// mlp_pair2_kernel.cuh -- M = 256 UMMAs shared by two SMs, each CTA streaming its 128-row half of every (transposed) weight
// tile, 128 columns per pair tile, DSMEM exchange of the produced operand with st.async -- with the backward epilogue:
//   D = W_l^T [dZ_l | dS_l]  ->  sp' / sp'' (or ReLU') from the saved layer input H_l  ->  [dZ_{l-1} | dS_{l-1}]  (next B operand,
//   also dumped K-sliced for the dW GEMM),  bias-gradient partial sums,  PE rows chained into d/dx at layer 0.
// Why: the single-CTA sweep is bound by its weight ring (8.6 MB of W^T per 64-column tile through 64 KiB: 26 B/clock/SM, tensor pipe
// 27 % active, profiles/r02/ncu_bwd_sweep_r2_summary.txt); a pair streams half as much per SM for twice the columns.
// Barriers, roles, UMMA order and byte counts are those of mlp_pair2_kernel (LP = 0); only the data the epilogue warps
// read and write differs.  The gradient reaching the skip-connection PE rows is produced in CTA 1 (rows 384-511 of the layer
// input) and consumed in CTA 0 (rows 0-38 at layer 0): it travels through a small global scratch (fp32, L2), fenced.
#pragma once
#include "mlp_bwd_kernel.cuh"
#include "mlp_pair2_kernel.cuh"

namespace mvsdf {

constexpr int kBwdStashRows = 40;
constexpr size_t kBwdStashBytesPerPair = (size_t)kBwdStashRows * kP2Cols * sizeof(float);       // 20 KiB

template <int KIND, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kP2Threads, 1) mlp_bwd_sweep_pair_kernel(const BwdArgs a, float* stash_g) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = ptx::cluster_ctarank();
  const uint32_t xbytes = (uint32_t)a.k_cores_max * kBCoreStride;
  const uint32_t s_stage = ptx::smem_u32(smem);
  const uint32_t s_xhi = s_stage + kStages * kStageBytes;
  const uint32_t s_xlo = s_xhi + kBLoOffset;
  const uint32_t s_bar = s_xhi + xbytes + kPeTileBytes;
  const uint32_t bar_full = s_bar;
  const uint32_t bar_empty = s_bar + 8 * kStages;
  const uint32_t bar_acc = s_bar + 16 * kStages;
  const uint32_t bar_lx = bar_acc + 8 * kP2Tiles;
  const uint32_t bar_px = bar_lx + 16;
  const uint32_t bar_d1 = bar_px + 16;
  const uint32_t s_tmem = bar_d1 + 8;

  const long long n_pts = a.n;
  constexpr int kPtsPerCta = (MODE == 0) ? kTileN : kTileN / 4;
  const long long n_tiles = (n_pts + 2 * kPtsPerCta - 1) / (2 * kPtsPerCta);          // pair tiles
  const long long pair0 = blockIdx.x >> 1;
  const long long pair_stride = gridDim.x >> 1;
  if (n_tiles == 0) return;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, crank == 0 ? 2 : 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int m = 0; m < kP2Tiles; ++m) ptx::mbar_init(bar_acc + 8 * m, 1);
    ptx::mbar_init(bar_lx, kEpiWarps);
    ptx::mbar_init(bar_lx + 8, kEpiWarps);
    ptx::mbar_init(bar_px, 1);
    ptx::mbar_init(bar_px + 8, 1);
    ptx::mbar_init(bar_d1, 2 * kEpiWarps);
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
    // ------------------------------------------------------------------ weight producer: my 128-row half of W^T, in UMMA order
    uint32_t it = 0;
    for (long long g = pair0; g < n_tiles; g += pair_stride) {
      for (int l = 0; l < a.n_run; ++l) {
        const LayerPlan& lp = a.Lt[l];
        for (int part = 0; part < 4; ++part) {
          int mp, k0, k1;
          bool run;
          p2_part(lp, part, mp, k0, k1, run);
          if (!run) continue;
          const int m = 2 * mp + (int)crank;
          const bool have = m < lp.m_tiles;
          const uint8_t* src = a.packed_t + lp.w_off + (size_t)m * lp.k_chunks * kStageBytes;
          for (int kc = k0; kc < k1; ++kc, ++it) {
            const uint32_t s = it % kStages, ph = (it / kStages) & 1;
            ptx::mbar_wait_sleep(bar_empty + 8 * s, ph ^ 1);
            if (lane == 0) {
              if (have) {
                ptx::mbar_arrive_expect_tx(bar_full + 8 * s, kStageBytes);
                ptx::bulk_g2s(s_stage + s * kStageBytes, src + (size_t)kc * kStageBytes, kStageBytes, bar_full + 8 * s);
              } else {
                ptx::mbar_arrive(bar_full + 8 * s);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (crank == 0 && warp == kEpiWarps + 1) {
    // ------------------------------------------------------------------ UMMA issuer (leader CTA): as in mlp_pair2_kernel
    constexpr uint32_t idesc = ptx::idesc_f16_f32_bmn(2 * kTileM, kP2Cols);
    const bool leader = ptx::elect_one();
    const uint32_t issue = leader ? 1u : 0u;
    uint32_t it0 = 0, x_ctr = 0, d1_uses = 0;
    bool ready = false;
    const uint32_t dlo_a = ((s_stage & 0x3FFFFu) >> 4) | ((128u >> 4) << 16);
    const uint32_t dlo_x = ((s_xhi & 0x3FFFFu) >> 4) | (((uint32_t)kBCoreStride >> 4) << 16);
    constexpr uint32_t dhi_a = (512u >> 4) | (1u << 14);
    constexpr uint32_t dhi_b = (128u >> 4) | (1u << 14);
    for (long long g = pair0; g < n_tiles; g += pair_stride) {
      for (int l = 0; l < a.n_run; ++l) {
        const LayerPlan& lp = a.Lt[l];
        for (int part = 0; part < 4; ++part) {
          int mp, k0, k1;
          bool run;
          p2_part(lp, part, mp, k0, k1, run);
          if (part == 0) {
            ptx::mbar_wait(bar_lx, x_ctr & 1);
            ptx::mbar_wait(bar_px, x_ctr & 1);
            ptx::fence_proxy_async_smem();
            ptx::tc_fence_after();
          } else if (part == 2) {
            ptx::mbar_wait(bar_lx + 8, x_ctr & 1);
            ptx::mbar_wait(bar_px + 8, x_ctr & 1);
            ptx::fence_proxy_async_smem();
            ptx::tc_fence_after();
            ++x_ctr;
          }
          if (!run) continue;
          if (part == 1) {
            if (d1_uses > 0) {
              ptx::mbar_wait(bar_d1, (d1_uses - 1) & 1);
              ptx::tc_fence_after();
            }
            ++d1_uses;
          }
          const uint32_t d_a = tmem_base + (uint32_t)(mp * 2 * kP2Cols);
          const uint32_t d_b = d_a + kP2Cols;
          const int n_st = k1 - k0;
          for (int j = 0; j < n_st; ++j) {
            const int kc = k0 + j;
            const uint32_t it = it0 + (uint32_t)j;
            const uint32_t st = it % kStages, ph = (it / kStages) & 1;
            if (!ready) ptx::mbar_wait(bar_full + 8 * st, ph);
            ptx::tc_fence_after();
            const uint32_t itn = it + 1;
            ready = ptx::mbar_test_wait(bar_full + 8 * (itn % kStages), (itn / kStages) & 1);
            const uint32_t a_off = dlo_a + st * (kStageBytes >> 4);
            const uint32_t b_off = dlo_x + (uint32_t)kc * ((kChunkK / 8) * kBCoreStride >> 4);
#pragma unroll
            for (int ks = 0; ks < kChunkK / 16; ++ks) {
              const uint64_t da_hi = ptx::desc_from_words(a_off + ks * (256 >> 4), dhi_a);
              const uint64_t da_lo = ptx::desc_from_words(a_off + ((kTileBytes + ks * 256) >> 4), dhi_a);
              const uint64_t db_hi = ptx::desc_from_words(b_off + ks * (2 * kBCoreStride >> 4), dhi_b);
              const uint64_t db_lo = ptx::desc_from_words(b_off + ((kBLoOffset + ks * 2 * kBCoreStride) >> 4), dhi_b);
              ptx::umma3_f16_2cta(d_a, d_b, da_hi, da_lo, db_hi, db_lo, idesc, (kc | ks) != 0 ? 1u : 0u, issue);
            }
            if (leader) ptx::umma_commit_2cta(bar_empty + 8 * st, 3);
            __syncwarp();
          }
          if (k1 == lp.k_chunks) {
            if (leader) ptx::umma_commit_2cta(bar_acc + 8 * mp, 3);
            __syncwarp();
          }
          it0 += (uint32_t)n_st;
        }
      }
    }
  } else if (crank == 1 && warp == kEpiWarps + 1) {
    // ------------------------------------------------------------------ CTA 1, warp 17: tell the leader my stage has landed
    const uint32_t remote_full = ptx::mapa(bar_full, 0);
    uint32_t it = 0;
    for (long long g = pair0; g < n_tiles; g += pair_stride) {
      for (int l = 0; l < a.n_run; ++l) {
        const LayerPlan& lp = a.Lt[l];
        int n_stage = 0;
        for (int part = 0; part < 4; ++part) {
          int mp, k0, k1;
          bool run;
          p2_part(lp, part, mp, k0, k1, run);
          if (run) n_stage += k1 - k0;
        }
        for (int i = 0; i < n_stage; ++i, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          ptx::mbar_wait(bar_full + 8 * s, ph);
          if (lane == 0) ptx::mbar_arrive_remote_relaxed(remote_full + 8 * s);
          __syncwarp();
        }
      }
    }
  } else if (crank == 1 && warp == kEpiWarps + 2) {
    // ------------------------------------------------------------------ CTA 1, warp 18: my bar_lx[h] -> the issuer's bar_px[h]
    const uint32_t remote_px = ptx::mapa(bar_px, 0);
    uint32_t x_ctr = 0;
    for (long long g = pair0; g < n_tiles; g += pair_stride) {
      for (int l = 0; l < a.n_run; ++l) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          ptx::mbar_wait(bar_lx + 8 * h, x_ctr & 1);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(remote_px + 8 * h);
          __syncwarp();
        }
        ++x_ctr;
      }
    }
  } else if (warp < kEpiWarps) {
    // ------------------------------------------------------------------ prologue + epilogue warps
    const int q = warp & 3;
    const int cg = warp >> 2;
    const int row = q * 32 + lane;
    const int t = threadIdx.x;
    const int lcol0 = cg * 16;
    const uint32_t dst_xhi = ptx::mapa(s_xhi, crank ^ 1u);
    const uint32_t dst_xlo = dst_xhi + kBLoOffset;
    const uint32_t dst_lx = ptx::mapa(bar_lx, crank ^ 1u);
    const uint32_t remote_d1 = ptx::mapa(bar_d1, 0);
    const float S = __ldg(a.gscale), invS = __ldg(a.gscale + 1);
    constexpr float kInvW = 1.0f / kWeightScale;
    const int F = a.feat_size;
    uint32_t acc_ctr[kP2Tiles] = {0, 0};
    float db_acc[kMaxLayers][kP2Tiles];
#pragma unroll
    for (int i = 0; i < kMaxLayers; ++i)
#pragma unroll
      for (int m = 0; m < kP2Tiles; ++m) db_acc[i][m] = 0.0f;
    float* const stash = stash_g + (size_t)pair0 * (kBwdStashRows * kP2Cols);       // [k][128 columns of the pair tile]

    for (long long g = pair0; g < n_tiles; g += pair_stride) {
      const long long tile_me = 2 * g + crank;                  // my 64-column tile
      const long long p0 = tile_me * kPtsPerCta;

      // ---------------- prologue: upstream gradients of MY 64 columns -> first B operand (scaled by S), also dumped for dW
      {
        const int fl = a.fwd_layer[0];
        const int kc = a.dz_kc[fl];
        const int kpad = kc * 8;
        uint8_t* gimg = a.dz ? a.dz + a.dz_off[fl] + (size_t)tile_me * ((size_t)kc * kBCoreStride) : nullptr;
        for (int item = t; item < 8 * kpad; item += kEpiThreads) {
          const int cb = item / kpad, k = item - cb * kpad;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = 0.0f;
          if (KIND == NET_SDF) {
            if (k < F + 2) {
              const int src = k < F ? k + 2 : k - F;
#pragma unroll
              for (int pp = 0; pp < 2; ++pp) {
                const long long gp = p0 + cb * 2 + pp;
                if (gp < n_pts) {
                  if (a.g_full) v[4 * pp] = __ldg(a.g_full + gp * (F + 2) + src);
                  if (k == F && a.g_grad) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) v[4 * pp + 1 + j] = __ldg(a.g_grad + gp * 3 + j);
                  }
                }
              }
            }
          } else if (k < 3) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const long long gp = p0 + cb * 8 + j;
              if (gp < n_pts) {
                const float y = __ldg(a.rgb + gp * 3 + k);
                v[j] = __ldg(a.g_full + gp * 3 + k) * (1.0f - y * y);
              }
            }
          }
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) pack_split(v[2 * i] * S, v[2 * i + 1] * S, hi[i], lo[i]);
          const uint32_t o = xoff(cb * 8, k);
          ptx::st_shared_v4(s_xhi + o, hi[0], hi[1], hi[2], hi[3]);
          ptx::st_shared_v4(s_xlo + o, lo[0], lo[1], lo[2], lo[3]);
          if (gimg) {
            uint8_t* gd = gimg + save_addr(kc, k, cb);
            *reinterpret_cast<uint4*>(gd) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(gd + 256) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(bar_lx);
        ptx::mbar_arrive(bar_lx + 8);
      }

      // ---------------- backward steps
      for (int l = 0; l < a.n_run; ++l) {
        const LayerPlan& lp = a.Lt[l];
        const int fl = a.fwd_layer[l];
        const bool last = (l == a.n_run - 1);
        const bool skip_here = KIND == NET_SDF && fl == a.skip_layer;
        const int n_pair_tiles = (lp.m_tiles + 1) >> 1;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        const int kc_in = last ? 0 : a.save_kc[fl];
        const int kc_out = last ? 0 : a.dz_kc[fl - 1];
#pragma unroll
        for (int mp = 0; mp < kP2Tiles; ++mp) {
          if (mp < n_pair_tiles) {
            const int m = 2 * mp + (int)crank;
            const int f = m * kTileM + row;                   // input feature of forward layer fl that this thread owns
            const bool have = m < lp.m_tiles;
            const bool have_h = have && !last && f < kc_in * 8;
            ptx::mbar_wait_sleep(bar_acc + 8 * mp, acc_ctr[mp] & 1);
            ++acc_ctr[mp];
            ptx::tc_fence_after();
            uint32_t va[2][16], vb[2][16];
#pragma unroll
            for (int hcol = 0; hcol < 2; ++hcol) {
              const uint32_t dest_h = hcol == 0 ? (crank ^ 1u) : crank;
              const uint32_t tcol = (uint32_t)(mp * 2 * kP2Cols) + dest_h * kTileN + (uint32_t)lcol0;
              ptx::tmem_ld_32x16(t_row + tcol, va[hcol]);
              ptx::tmem_ld_32x16(t_row + tcol + kP2Cols, vb[hcol]);
            }
            ptx::tmem_ld_wait();
            if (mp == 1) {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive_remote_relaxed(remote_d1);
            }
            const bool pe_row = KIND == NET_SDF && !last && skip_here && f >= a.skip_rows_begin;
            const uint32_t o0 = xoff(lcol0, f);
#pragma unroll
            for (int hcol = 0; hcol < 2; ++hcol) {
              const uint32_t dest_h = hcol == 0 ? (crank ^ 1u) : crank;
              const long long tile_d = 2 * g + dest_h;               // the 64-column tile these 16 columns belong to
              float d[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) d[j] = (__uint_as_float(va[hcol][j]) + __uint_as_float(vb[hcol][j])) * kInvW;
              if (!last) {
                uint32_t phi[8], plo[8];
                if (pe_row) {
                  const int k = f - a.skip_rows_begin;
                  if (k < a.pe_dim) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) __stcg(stash + (size_t)k * kP2Cols + dest_h * kTileN + lcol0 + j, d[j]);
                  }
#pragma unroll
                  for (int i = 0; i < 8; ++i) phi[i] = plo[i] = 0u;
                } else {
                  float o[16];
#pragma unroll
                  for (int j = 0; j < 16; ++j) o[j] = 0.0f;
                  if (have_h) {
                    const uint8_t* hp = a.save + a.save_off[fl] + (size_t)tile_d * ((size_t)kc_in * kBCoreStride) + save_addr(kc_in, f, cg * 2);
                    const uint4 sh0 = __ldg(reinterpret_cast<const uint4*>(hp));
                    const uint4 sh1 = __ldg(reinterpret_cast<const uint4*>(hp + 128));
                    const uint4 sl0 = __ldg(reinterpret_cast<const uint4*>(hp + 256));
                    const uint4 sl1 = __ldg(reinterpret_cast<const uint4*>(hp + 384));
                    if (KIND == NET_SDF) {
#pragma unroll
                      for (int gq = 0; gq < 4; ++gq) {
                        const uint4& hh = gq < 2 ? sh0 : sh1;
                        const uint4& ll = gq < 2 ? sl0 : sl1;
                        const int c0 = (gq & 1) * 4;
                        const float h = saved_val(hh, ll, c0);
                        const float t0 = saved_val(hh, ll, c0 + 1), t1 = saved_val(hh, ll, c0 + 2), t2 = saved_val(hh, ll, c0 + 3);
                        const float em = expm1f(-100.0f * h);
                        const float s1 = -em, one_m = 1.0f + em;
                        const float dh = d[4 * gq], dt0 = d[4 * gq + 1], dt1 = d[4 * gq + 2], dt2 = d[4 * gq + 3];
                        o[4 * gq] = s1 * dh + 100.0f * one_m * (t0 * dt0 + t1 * dt1 + t2 * dt2);
                        o[4 * gq + 1] = s1 * dt0;
                        o[4 * gq + 2] = s1 * dt1;
                        o[4 * gq + 3] = s1 * dt2;
                      }
                      db_acc[l][mp] += o[0] + o[4] + o[8] + o[12];
                    } else {
                      float bsum = 0.0f;
#pragma unroll
                      for (int j = 0; j < 16; ++j) {
                        const uint4& hh = j < 8 ? sh0 : sh1;
                        const uint4& ll = j < 8 ? sl0 : sl1;
                        const float h = saved_val(hh, ll, j & 7);
                        o[j] = h > 0.0f ? d[j] : 0.0f;
                        bsum += o[j];
                      }
                      db_acc[l][mp] += bsum;
                    }
                  }
#pragma unroll
                  for (int i = 0; i < 8; ++i) pack_split(o[2 * i], o[2 * i + 1], phi[i], plo[i]);
                }
                if (have) {
                  if (a.dz && f < kc_out * 8) {
                    uint8_t* gd = a.dz + a.dz_off[fl - 1] + (size_t)tile_d * ((size_t)kc_out * kBCoreStride) + save_addr(kc_out, f, cg * 2);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                      *reinterpret_cast<uint4*>(gd + j * 128) = make_uint4(phi[4 * j], phi[4 * j + 1], phi[4 * j + 2], phi[4 * j + 3]);
                      *reinterpret_cast<uint4*>(gd + 256 + j * 128) = make_uint4(plo[4 * j], plo[4 * j + 1], plo[4 * j + 2], plo[4 * j + 3]);
                    }
                  }
                  if (hcol == 1) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                      ptx::st_shared_v4(s_xhi + o0 + j * 128, phi[4 * j], phi[4 * j + 1], phi[4 * j + 2], phi[4 * j + 3]);
                      ptx::st_shared_v4(s_xlo + o0 + j * 128, plo[4 * j], plo[4 * j + 1], plo[4 * j + 2], plo[4 * j + 3]);
                    }
                  } else {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                      ptx::st_async_v4(dst_xhi + o0 + j * 128, phi[4 * j], phi[4 * j + 1], phi[4 * j + 2], phi[4 * j + 3], dst_lx + 8 * mp);
                      ptx::st_async_v4(dst_xlo + o0 + j * 128, plo[4 * j], plo[4 * j + 1], plo[4 * j + 2], plo[4 * j + 3], dst_lx + 8 * mp);
                    }
                  }
                }
              } else if (have) {
                // ---------------- input of the first layer (rows of CTA 0 / CTA 1 as they fall), for the columns of tile_d
                const long long pd0 = tile_d * kPtsPerCta;
                if (KIND == NET_SDF) {
                  if (f < a.pe_dim && a.dx) {
                    const int k = f;
                    const int coord = k < 3 ? k : (k - 3) % 3;
                    const int fi = k < 3 ? 0 : (k - 3) / 6;
                    const bool is_cos = k >= 3 && ((k - 3) % 6) >= 3;
                    const float fr = (float)(1 << fi);
                    const bool has_skip = a.skip_layer >= 0;
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq) {
                      const long long gp = pd0 + cg * 4 + gq;
                      if (gp < n_pts) {
                        const int c0 = (int)dest_h * kTileN + lcol0 + 4 * gq;
                        const float gv = d[4 * gq] + (has_skip ? __ldcg(stash + (size_t)k * kP2Cols + c0) : 0.0f);
                        const float gt = d[4 * gq + 1 + coord] + (has_skip ? __ldcg(stash + (size_t)k * kP2Cols + c0 + 1 + coord) : 0.0f);
                        const float xc = __ldg(a.x + gp * 3 + coord);
                        float p1, p2;
                        if (k < 3) {
                          p1 = 1.0f;
                          p2 = 0.0f;
                        } else {
                          float sn, cs;
                          sincosf(xc * fr, &sn, &cs);
                          p1 = is_cos ? -fr * sn : fr * cs;
                          p2 = is_cos ? -fr * fr * cs : -fr * fr * sn;
                        }
                        atomicAdd(a.dx + gp * 3 + coord, (gv * p1 + gt * p2) * invS);
                      }
                    }
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const long long gp = pd0 + lcol0 + j;
                    if (gp < n_pts) {
                      const float val = d[j] * invS;
                      if (f < 3) {
                        if (a.d_points) a.d_points[gp * 3 + f] = val;
                      } else if (f < 30) {
                        if (a.dx) {
                          const int k = f - 3;
                          const int coord = k < 3 ? k : (k - 3) % 3;
                          float p1 = 1.0f;
                          if (k >= 3) {
                            const float fr = (float)(1 << ((k - 3) / 6));
                            float sn, cs;
                            sincosf(__ldg(a.x + gp * 3 + coord) * fr, &sn, &cs);
                            p1 = ((k - 3) % 6) >= 3 ? -fr * sn : fr * cs;
                          }
                          atomicAdd(a.dx + gp * 3 + coord, val * p1);
                        }
                      } else if (f < 33) {
                        if (a.d_normals) a.d_normals[gp * 3 + (f - 30)] = val;
                      } else if (f < 33 + F) {
                        if (a.d_feats) a.d_feats[gp * F + (f - 33)] = val;
                      }
                    }
                  }
                }
              }
            }
            if (!last) {
              if (pe_row) __threadfence();                     // the stash is read by the other CTA of the pair (layer 0)
              ptx::tc_fence_before();
              ptx::fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                if (warp == 0) {
                  const int pm = 2 * mp + (int)(crank ^ 1u);       // the peer's output tile: 256 B per row it writes for my columns
                  const int rows = pm < lp.m_tiles ? kTileM : 0;
                  if (rows > 0) ptx::mbar_arrive_expect_tx(bar_lx + 8 * mp, (uint32_t)rows * (kTileN * 4));
                  else ptx::mbar_arrive(bar_lx + 8 * mp);
                } else {
                  ptx::mbar_arrive(bar_lx + 8 * mp);
                }
                if (n_pair_tiles == 1) ptx::mbar_arrive(bar_lx + 8);
              }
            }
          }
        }
      }
    }
    if (a.db) {
      for (int l = 0; l + 1 < a.n_run; ++l) {
        const int fl = a.fwd_layer[l];
#pragma unroll
        for (int mp = 0; mp < kP2Tiles; ++mp) {
          const int m = 2 * mp + (int)crank;
          const int f = m * kTileM + row;
          if (m < a.Lt[l].m_tiles && f < a.Lt[l + 1].in_dim && db_acc[l][mp] != 0.0f) atomicAdd(a.db + a.db_off[fl - 1] + f, db_acc[l][mp] * invS);
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == kEpiWarps + 1) ptx::tmem_dealloc_2cta(tmem_base, kTmemCols);
}

}  // namespace mvsdf

// Native backward of the two MLPs: the reverse sweep through ImplicitNetwork.forward +
// .gradient (code/model/implicit_differentiable_renderer.py:77-107; the reference differentiates autograd's own backward,
// create_graph=True, :104) and through RenderingNetwork.forward (:145-167), on the same tcgen05 tile core as the forward.
// The chain executed here is written out and pinned against autograd in oracle/backward_spec.py.
//
// Two kernels:
//
//  mlp_bwd_sweep_kernel<KIND, MODE>  -- per 64-column tile (16 points x [value, d/dx, d/dy, d/dz] for the SDF net, 64 points
//    for the rendering net) walks the layers in reverse:   D = W_l^T [dZ_l | dS_l]   (tcgen05.mma, transposed weight blob
//    streamed through the same 4-stage bulk-copy ring, fp16 hi/lo split operands, 3 products per MAC, fp32 accumulate in TMEM),
//    epilogue:   dZ_{l-1} = sp'(Z_{l-1}) dH_l + sp''(Z_{l-1}) sum_j S_{l-1,j} dT_{l,j},   dS_{l-1,j} = sp'(Z_{l-1}) dT_{l,j}
//    with everything that depends on the forward pass recovered from the SAVED layer input H_l = [sp(Z), sp'(Z) S_j]
//    (mlp_kernel.cuh, MlpArgs::save):  sp'(z) = 1 - exp(-100 sp(z)),  sp''(z) S_j = 100 (1 - sp'(z)) T_j.
//    The new operand overwrites the activation buffer in place (next layer's B operand) and is also written to global
//    memory in the K-sliced layout (save_addr) for the dW kernel; bias gradients are column sums (red.global.add); at the
//    skip layer and at layer 0 the rows that belong to the positional encoding are chained through PE' and PE'' into dx.
//
//  mlp_bwd_dw_kernel -- dW_l = [dZ_l | dS_l] [H_l | T_l]^T, a GEMM whose contraction runs over the POINTS: both saved
//    images are K-major operands when read along the columns.  One CTA owns 128 output rows x all (<= 512) input features
//    of one layer (the whole TMEM: 128 lanes x 512 fp32 columns) for a range of column tiles; K = 16 columns per pipeline
//    stage (one contiguous 8 KiB + one <= 32 KiB bulk copy), split-K partial sums are added with red.global.add.f32.
//
// Gradients are carried scaled by a power of two (BwdArgs::gscale, chosen on the device from max |upstream|) so that they
// sit in the middle of the fp16 range, exactly like AMP loss scaling; the backward is linear, the scale is exact.
#pragma once
#include "mlp_kernel.cuh"

namespace mvsdf {

constexpr int kBwdStashFloats = 40 * kTileN;       // gradient reaching the skip-connection PE rows (fp32), lives in the PE tile

struct BwdArgs {
  const uint8_t* packed_t;     // transposed weight blob (mvsdf_pack_weights_t)
  int n_run;                   // backward steps: head^T, hidden layers in reverse
  int skip_layer;              // forward index of the layer whose input is cat([h, PE]) / sqrt 2 (-1: none)
  int skip_rows_begin;
  int pe_dim;
  int k_cores_max;
  int feat_size;
  long long n;
  const float* gscale;         // [2] device: {S, 1/S}
  // forward state
  const float* x;              // [n,3] SDF: points (PE derivatives for dx); render: view directions (only with dx)
  const uint8_t* save;         // saved layer inputs of the forward pass
  long long save_off[kMaxLayers];    // by FORWARD layer index
  int save_kc[kMaxLayers];           // feature blocks of that image
  // upstream gradients
  const float* g_full;         // SDF: [n, 2+F] or nullptr;   render: g_rgb [n,3]
  const float* g_grad;         // SDF: [n,3] or nullptr
  const float* rgb;            // render: forward output [n,3] (tanh'), else nullptr
  // outputs
  uint8_t* dz;                 // dumped [dZ_l | dS_l] images, K-sliced, by FORWARD layer index
  long long dz_off[kMaxLayers];
  int dz_kc[kMaxLayers];
  float* db;                   // bias gradients (scaled by S), plan coordinates: db + db_off[l] + row
  int db_off[kMaxLayers];
  float* dx;                   // SDF: dL/dx [n,3]; render: dL/d view [n,3]; accumulated with atomics (pre-zeroed), or nullptr
  float* d_points;             // render: [n,3]
  float* d_normals;            // render: [n,3]
  float* d_feats;              // render: [n, feat_size]
  int* status;
  LayerPlan Lt[kMaxLayers];    // transposed plans in backward order (Lt[0] = head^T)
  int fwd_layer[kMaxLayers];   // forward layer index of backward step i
};

__device__ __forceinline__ float half2_lo_f(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v & 0xffffu))); }
__device__ __forceinline__ float half2_hi_f(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v >> 16))); }

// value of column c (0..7) of a pair of saved (hi, lo) 16-byte vectors, un-scaled
__device__ __forceinline__ float saved_val(const uint4& hi, const uint4& lo, int c) {
  const uint32_t wh = c < 2 ? hi.x : (c < 4 ? hi.y : (c < 6 ? hi.z : hi.w));
  const uint32_t wl = c < 2 ? lo.x : (c < 4 ? lo.y : (c < 6 ? lo.z : lo.w));
  const float h = (c & 1) ? half2_hi_f(wh) : half2_lo_f(wh);
  const float l = (c & 1) ? half2_hi_f(wl) : half2_lo_f(wl);
  return (h + l) * (1.0f / kActScale);
}

template <int KIND, int MODE>
__global__ void __launch_bounds__(kMlpThreads, 1) mlp_bwd_sweep_kernel(const BwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t xbytes = (uint32_t)a.k_cores_max * kBCoreStride;
  const uint32_t s_stage = ptx::smem_u32(smem);
  const uint32_t s_xhi = s_stage + kStages * kStageBytes;
  const uint32_t s_xlo = s_xhi + kBLoOffset;
  const uint32_t s_pe = s_xhi + xbytes;                     // stash area (the forward's PE tile)
  const uint32_t s_bar = s_pe + kPeTileBytes;
  const uint32_t bar_full = s_bar;
  const uint32_t bar_empty = s_bar + 8 * kStages;
  const uint32_t bar_acc = s_bar + 16 * kStages;
  const uint32_t bar_act = bar_acc + 8 * kMaxTiles;
  const uint32_t s_tmem = bar_act + 8;
  float* const stash = reinterpret_cast<float*>(smem + kStages * kStageBytes + xbytes);

  const long long n_pts = a.n;
  constexpr int kPtsPerTile = (MODE == 0) ? kTileN : kTileN / 4;
  const long long n_tiles = (n_pts + kPtsPerTile - 1) / kPtsPerTile;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
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
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(smem + (s_tmem - s_stage)), 0);

  if (warp == kEpiWarps) {
    // ------------------------------------------------------------------ weight producer (transposed blob)
    uint32_t it = 0;
    for (long long g = blockIdx.x; g < n_tiles; g += gridDim.x) {
      for (int l = 0; l < a.n_run; ++l) {
        const LayerPlan& lp = a.Lt[l];
        const uint8_t* src = a.packed_t + lp.w_off;
        const int n_stage = lp.m_tiles * lp.k_chunks;
        for (int i = 0; i < n_stage; ++i, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1);
          if (lane == 0) {
            ptx::mbar_arrive_expect_tx(bar_full + 8 * s, kStageBytes);
            ptx::bulk_g2s(s_stage + s * kStageBytes, src + (size_t)i * kStageBytes, kStageBytes, bar_full + 8 * s);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ------------------------------------------------------------------ UMMA issuer (identical to the forward kernel)
    constexpr uint32_t idesc128 = ptx::idesc_f16_f32_bmn(kTileM, 2 * kTileN);
    constexpr uint32_t idesc64 = ptx::idesc_f16_f32_bmn(kTileM, kTileN);
    const bool leader = ptx::elect_one();
    uint32_t it = 0, act_ctr = 0;
    for (long long g = blockIdx.x; g < n_tiles; g += gridDim.x) {
      for (int l = 0; l < a.n_run; ++l) {
        const LayerPlan& lp = a.Lt[l];
        ptx::mbar_wait(bar_act, act_ctr & 1);
        ++act_ctr;
        ptx::tc_fence_after();
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
              const uint64_t db = ptx::smem_desc(s_xhi + (uint32_t)((kc * (kChunkK / 8) + ks * 2) * kBCoreStride), kBCoreStride, 128);
              if (leader) {
                ptx::umma_f16(d_tmem, da_hi, db, idesc128, (kc | ks) != 0 ? 1u : 0u);
                ptx::umma_f16(d_tmem, da_lo, db, idesc64, 1u);
              }
            }
            if (leader) ptx::umma_commit(bar_empty + 8 * s);
            __syncwarp();
          }
          if (leader) ptx::umma_commit(bar_acc + 8 * m);
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ prologue + epilogue warps
    const int q = warp & 3;
    const int cg = warp >> 2;
    const int row = q * 32 + lane;
    const int t = threadIdx.x;
    const float S = __ldg(a.gscale), invS = __ldg(a.gscale + 1);
    constexpr float kInvW = 1.0f / kWeightScale;
    uint32_t acc_ctr[kMaxTiles] = {0, 0, 0, 0};
    const int F = a.feat_size;
    // bias gradients: this thread's column sums, accumulated over every tile the CTA processes and flushed with one atomic
    // per (layer, row) at the end (an atomic per tile and row serialised 43 M red operations on 4.6 k addresses)
    float db_acc[kMaxLayers][kMaxTiles];
#pragma unroll
    for (int i = 0; i < kMaxLayers; ++i)
#pragma unroll
      for (int m = 0; m < kMaxTiles; ++m) db_acc[i][m] = 0.0f;

    for (long long g = blockIdx.x; g < n_tiles; g += gridDim.x) {
      const long long tile = g;
      const long long p0 = tile * kPtsPerTile;

      // ---------------- prologue: upstream gradients -> first B operand [dZ_L | dS_L] (scaled by S), also dumped for dW
      {
        const int fl = a.fwd_layer[0];
        const int kc = a.dz_kc[fl];
        const int kpad = kc * 8;
        uint8_t* gimg = a.dz ? a.dz + a.dz_off[fl] + (size_t)tile * ((size_t)kc * kBCoreStride) : nullptr;   // nullptr: dx-only sweep
        // one work item = (feature k, block of 8 columns): 16-byte vectors for the shared-memory operand and for the dump
        for (int item = t; item < 8 * kpad; item += kEpiThreads) {
          const int cb = item / kpad, k = item - cb * kpad;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = 0.0f;
          if (KIND == NET_SDF) {
            if (k < F + 2) {
              // head rows are stored features-first: k < F -> full[:, 2 + k], k = F -> sdf, k = F + 1 -> indicator
              const int src = k < F ? k + 2 : k - F;
#pragma unroll
              for (int pp = 0; pp < 2; ++pp) {           // 8 columns = 2 points x [value, d/dx, d/dy, d/dz]
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
                const float y = __ldg(a.rgb + gp * 3 + k);          // d tanh = 1 - y^2
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
      if (lane == 0) ptx::mbar_arrive(bar_act);

      // ---------------- backward steps
      for (int l = 0; l < a.n_run; ++l) {
        const LayerPlan& lp = a.Lt[l];
        const int fl = a.fwd_layer[l];                 // D = gradient w.r.t. the INPUT of forward layer fl
        const bool last = (l == a.n_run - 1);          // fl == 0: the input is the positional encoding / the render input
        const bool skip_here = KIND == NET_SDF && fl == a.skip_layer;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 16);
        uint32_t phi[kMaxTiles][8], plo[kMaxTiles][8];
        const int kc_in = last ? 0 : a.save_kc[fl];
        const uint8_t* himg = last ? nullptr : a.save + a.save_off[fl] + (size_t)tile * ((size_t)kc_in * kBCoreStride);
#pragma unroll
        for (int m = 0; m < kMaxTiles; ++m) {
          if (m < lp.m_tiles) {
            const int f = m * kTileM + row;            // input feature of forward layer fl
            // saved H_fl[f, my 16 columns]: fetched before the wait so that the global-load latency hides behind the UMMAs
            uint4 sh0 = make_uint4(0, 0, 0, 0), sh1 = sh0, sl0 = sh0, sl1 = sh0;
            const bool have_h = !last && f < kc_in * 8;
            if (have_h) {
              const uint8_t* hp = himg + save_addr(kc_in, f, cg * 2);
              sh0 = __ldg(reinterpret_cast<const uint4*>(hp));
              sh1 = __ldg(reinterpret_cast<const uint4*>(hp + 128));
              sl0 = __ldg(reinterpret_cast<const uint4*>(hp + 256));
              sl1 = __ldg(reinterpret_cast<const uint4*>(hp + 384));
            }
            ptx::mbar_wait(bar_acc + 8 * m, acc_ctr[m] & 1);
            ++acc_ctr[m];
            ptx::tc_fence_after();
            uint32_t v[16], v2[16];
            ptx::tmem_ld_32x16(t_row + (uint32_t)(m * 2 * kTileN), v);
            ptx::tmem_ld_32x16(t_row + (uint32_t)(m * 2 * kTileN + kTileN), v2);
            ptx::tmem_ld_wait();
            float d[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) d[j] = (__uint_as_float(v[j]) + __uint_as_float(v2[j])) * kInvW;     // still x S
            const bool pe_row = KIND == NET_SDF && ((skip_here && f >= a.skip_rows_begin) || last);
            if (!last && !pe_row) {
              float o[16];
              if (KIND == NET_SDF) {
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                  // columns 4 gq .. 4 gq + 3 = [value, d/dx, d/dy, d/dz] of one point
                  const uint4& hh = gq < 2 ? sh0 : sh1;
                  const uint4& ll = gq < 2 ? sl0 : sl1;
                  const int c0 = (gq & 1) * 4;
                  const float h = saved_val(hh, ll, c0);
                  const float t0 = saved_val(hh, ll, c0 + 1), t1 = saved_val(hh, ll, c0 + 2), t2 = saved_val(hh, ll, c0 + 3);
                  const float em = expm1f(-100.0f * h);          // sp'(z) = 1 - exp(-100 sp(z)) = -em;  1 - sp'(z) = 1 + em
                  const float s1 = -em, one_m = 1.0f + em;
                  const float dh = d[4 * gq], dt0 = d[4 * gq + 1], dt1 = d[4 * gq + 2], dt2 = d[4 * gq + 3];
                  o[4 * gq] = s1 * dh + 100.0f * one_m * (t0 * dt0 + t1 * dt1 + t2 * dt2);
                  o[4 * gq + 1] = s1 * dt0;
                  o[4 * gq + 2] = s1 * dt1;
                  o[4 * gq + 3] = s1 * dt2;
                }
                if (MODE == 1 && have_h) {
                  const float bsum = o[0] + o[4] + o[8] + o[12];
                  db_acc[l][m] += bsum;
                }
              } else {
                float bsum = 0.0f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const uint4& hh = j < 8 ? sh0 : sh1;
                  const uint4& ll = j < 8 ? sl0 : sl1;
                  const float h = saved_val(hh, ll, j & 7);
                  o[j] = h > 0.0f ? d[j] : 0.0f;                 // ReLU'
                  bsum += o[j];
                }
                if (have_h) db_acc[l][m] += bsum;
              }
              if (!have_h) {
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = 0.0f;
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) pack_split(o[2 * i], o[2 * i + 1], phi[m][i], plo[m][i]);
            } else if (pe_row && !last) {
              // skip connection: these rows of layer fl's input are the positional encoding -- keep their gradient (fp32)
              // for the PE chain at the end, and feed zeros to the rows of the next operand they occupy
              const int k = f - a.skip_rows_begin;
              if (k < a.pe_dim) {
#pragma unroll
                for (int j = 0; j < 16; ++j) stash[k * kTileN + cg * 16 + j] = d[j];
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) phi[m][i] = plo[m][i] = 0u;
            } else {
              // ---------------- input of the first layer
              if (KIND == NET_SDF) {
                if (f < a.pe_dim && a.dx) {
                  const int k = f;
                  const int coord = k < 3 ? k : (k - 3) % 3;
                  const int fi = k < 3 ? 0 : (k - 3) / 6;
                  const bool is_cos = k >= 3 && ((k - 3) % 6) >= 3;
                  const float fr = (float)(1 << fi);
#pragma unroll
                  for (int gq = 0; gq < 4; ++gq) {
                    const long long gp = p0 + cg * 4 + gq;
                    if (gp < n_pts) {
                      const int c0 = cg * 16 + 4 * gq;
                      const bool has_skip = a.skip_layer >= 0;
                      const float gv = d[4 * gq] + (has_skip ? stash[k * kTileN + c0] : 0.0f);
                      const float gt = d[4 * gq + 1 + coord] + (has_skip ? stash[k * kTileN + c0 + 1 + coord] : 0.0f);
                      const float xc = __ldg(a.x + gp * 3 + coord);
                      float p1, p2;                          // PE' and PE'' of this feature w.r.t. its coordinate
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
                // render input rows: [points(3), PE4(view)(27), normals(3), features(F)]; the view direction is a constant
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const long long gp = p0 + cg * 16 + j;
                  if (gp < n_pts) {
                    const float val = d[j] * invS;
                    if (f < 3) {
                      if (a.d_points) a.d_points[gp * 3 + f] = val;
                    } else if (f < 30) {
                      // PE4(view) rows: chained through PE' into the view direction (only needed when the camera poses are
                      // trained, rend_util.py:49-57: the view direction then depends on the pose parameters)
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
                    } else if (f >= 30 && f < 33) {
                      if (a.d_normals) a.d_normals[gp * 3 + (f - 30)] = val;
                    } else if (f >= 33 && f < 33 + F) {
                      if (a.d_feats) a.d_feats[gp * F + (f - 33)] = val;
                    }
                  }
                }
              }
            }
          }
        }
        if (!last) {
          // every UMMA of this step has retired: the new operand [dZ_{fl-1} | dS_{fl-1}] overwrites the buffer in place
          const int kc_out = a.dz_kc[fl - 1];
          uint8_t* gimg = a.dz ? a.dz + a.dz_off[fl - 1] + (size_t)tile * ((size_t)kc_out * kBCoreStride) : nullptr;
#pragma unroll
          for (int m = 0; m < kMaxTiles; ++m) {
            if (m < lp.m_tiles) {
              const int f = m * kTileM + row;
              const uint32_t o0 = xoff(cg * 16, f);
              uint8_t* gd = (gimg && f < kc_out * 8) ? gimg + save_addr(kc_out, f, cg * 2) : nullptr;
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                ptx::st_shared_v4(s_xhi + o0 + j * 128, phi[m][4 * j], phi[m][4 * j + 1], phi[m][4 * j + 2], phi[m][4 * j + 3]);
                ptx::st_shared_v4(s_xlo + o0 + j * 128, plo[m][4 * j], plo[m][4 * j + 1], plo[m][4 * j + 2], plo[m][4 * j + 3]);
                if (gd) {
                  *reinterpret_cast<uint4*>(gd + j * 128) = make_uint4(phi[m][4 * j], phi[m][4 * j + 1], phi[m][4 * j + 2], phi[m][4 * j + 3]);
                  *reinterpret_cast<uint4*>(gd + 256 + j * 128) = make_uint4(plo[m][4 * j], plo[m][4 * j + 1], plo[m][4 * j + 2], plo[m][4 * j + 3]);
                }
              }
            }
          }
          ptx::tc_fence_before();
          ptx::fence_proxy_async_smem();
          if (skip_here) asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");     // stash complete before anyone reads it
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_act);
        } else {
          // the tile is done; the next tile's prologue rewrites the operand buffer and the stash: wait for every warp
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        }
      }
    }
    if (a.db) {
      for (int l = 0; l + 1 < a.n_run; ++l) {
        const int fl = a.fwd_layer[l];
#pragma unroll
        for (int m = 0; m < kMaxTiles; ++m) {
          const int f = m * kTileM + row;
          if (m < a.Lt[l].m_tiles && f < a.Lt[l + 1].in_dim && db_acc[l][m] != 0.0f) atomicAdd(a.db + a.db_off[fl - 1] + f, db_acc[l][m] * invS);
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------------------------
// dW GEMM over the points
// ------------------------------------------------------------------------------------------------------------------
constexpr int kDwStages = 5;
constexpr int kDwABytes = 16 * 512;                  // 128 output rows x 16 columns, hi + lo
constexpr int kDwBBytesMax = 64 * 512;               // <= 512 input features x 16 columns, hi + lo
constexpr int kDwStageBytes = kDwABytes + kDwBBytesMax;
constexpr int kDwThreads = 6 * 32;                   // 4 epilogue warps (128 TMEM lanes) + producer + issuer

struct DwLayer {
  long long a_off;      // image of [dZ | dS] of this forward layer (K-sliced), inside `dz`
  long long b_off;      // image of the saved layer input, inside `save`
  int kc_a, kc_b;       // feature blocks per image (out_pad / 8, in_pad / 8)
  int m_tiles;          // 128-row output tiles
  int n_in;             // input features, padded (multiple of 16, <= 512)
  long long w_off;      // float offset of dW_l [m_tiles * 128, n_in] inside `dw`
};

struct DwArgs {
  const uint8_t* dz;
  const uint8_t* save;
  float* dw;
  const float* gscale;  // {S, 1/S}
  long long n_tiles;    // 64-column tiles
  int n_layers;
  int splits;           // split-K factor (ranges of column tiles)
  int item_begin[kMaxLayers + 1];   // work items are (layer, m tile, split): prefix sums over layers of m_tiles * splits
  DwLayer L[kMaxLayers];
};

__host__ __device__ inline size_t dw_smem_bytes() { return (size_t)kDwStages * kDwStageBytes + 256; }

__global__ void __launch_bounds__(kDwThreads, 1) mlp_bwd_dw_kernel(const DwArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t s_stage = ptx::smem_u32(smem);
  const uint32_t s_bar = s_stage + kDwStages * kDwStageBytes;
  const uint32_t bar_full = s_bar;
  const uint32_t bar_empty = s_bar + 8 * kDwStages;
  const uint32_t bar_done = s_bar + 16 * kDwStages;
  const uint32_t s_tmem = bar_done + 8;

  // which work item
  int layer = 0;
  while (layer + 1 < a.n_layers && (int)blockIdx.x >= a.item_begin[layer + 1]) ++layer;
  const DwLayer& L = a.L[layer];
  const int local = (int)blockIdx.x - a.item_begin[layer];
  const int m = local / a.splits, split = local - m * a.splits;
  const long long t_begin = a.n_tiles * split / a.splits, t_end = a.n_tiles * (split + 1) / a.splits;
  const long long n_steps = (t_end - t_begin) * 4;            // 16-column slices
  const uint32_t b_bytes = (uint32_t)L.kc_b * 512u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kDwStages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    ptx::mbar_init(bar_done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 5) {
    ptx::tmem_alloc(s_tmem, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(smem + (s_tmem - s_stage)), 0);

  if (warp == 4) {
    // ------------------------------------------------------------------ producer: one A block + one B block per 16 columns
    const uint8_t* a_img = a.dz + L.a_off;
    const uint8_t* b_img = a.save + L.b_off;
    const size_t a_tile = (size_t)L.kc_a * kBCoreStride, b_tile = (size_t)L.kc_b * kBCoreStride;
    for (long long i = 0; i < n_steps; ++i) {
      const uint32_t s = (uint32_t)(i % kDwStages), ph = (uint32_t)((i / kDwStages) & 1);
      ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1);
      if (lane == 0) {
        const long long tile = t_begin + (i >> 2);
        const int sl = (int)(i & 3);
        ptx::mbar_arrive_expect_tx(bar_full + 8 * s, kDwABytes + b_bytes);
        ptx::bulk_g2s(s_stage + s * kDwStageBytes, a_img + tile * a_tile + (size_t)sl * L.kc_a * 512 + (size_t)m * kDwABytes, kDwABytes,
                      bar_full + 8 * s);
        ptx::bulk_g2s(s_stage + s * kDwStageBytes + kDwABytes, b_img + tile * b_tile + (size_t)sl * L.kc_b * 512, b_bytes,
                      bar_full + 8 * s);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ issuer: D[128, n_in] += A[128, 16] B[n_in, 16]^T, 3 products
    const bool leader = ptx::elect_one();
    const int n_half = (L.n_in + 255) / 256;
    for (long long i = 0; i < n_steps; ++i) {
      const uint32_t s = (uint32_t)(i % kDwStages), ph = (uint32_t)((i / kDwStages) & 1);
      ptx::mbar_wait(bar_full + 8 * s, ph);
      ptx::tc_fence_after();
      const uint32_t a_hi = s_stage + s * kDwStageBytes;
      const uint32_t b_hi = a_hi + kDwABytes;
      const uint64_t da_hi = ptx::smem_desc(a_hi, 128, 512);
      const uint64_t da_lo = ptx::smem_desc(a_hi + 256, 128, 512);
      if (leader) {
        for (int h = 0; h < n_half; ++h) {
          const int n_this = min(256, L.n_in - 256 * h);
          const uint32_t idesc = ptx::idesc_f16_f32(kTileM, n_this);
          const uint64_t db_hi = ptx::smem_desc(b_hi + (uint32_t)h * 32u * 512u, 128, 512);
          const uint64_t db_lo = ptx::smem_desc(b_hi + (uint32_t)h * 32u * 512u + 256, 128, 512);
          const uint32_t d = tmem_base + (uint32_t)(h * 256);
          ptx::umma_f16(d, da_hi, db_hi, idesc, i != 0 ? 1u : 0u);
          ptx::umma_f16(d, da_lo, db_hi, idesc, 1u);
          ptx::umma_f16(d, da_hi, db_lo, idesc, 1u);
        }
        ptx::umma_commit(bar_empty + 8 * s);
      }
      __syncwarp();
    }
    if (leader) ptx::umma_commit(bar_done);
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> red.global.add into dW
    if (n_steps > 0) {
      ptx::mbar_wait(bar_done, 0);
      ptx::tc_fence_after();
      const float scale = __ldg(a.gscale + 1) * (1.0f / kActScale);          // operands were (S dZ) and (64 H)
      float* out = a.dw + L.w_off + (size_t)(m * kTileM + warp * 32 + lane) * L.n_in;
      const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int c0 = 0; c0 < L.n_in; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld_32x16(t_row + (uint32_t)c0, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) atomicAdd(out + c0 + j, __uint_as_float(v[j]) * scale);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace mvsdf

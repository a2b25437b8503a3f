// CTA-pair MLP tile core, split-K pipelined ("pair2"): same data path as mlp_pair_kernel.cuh (cta_group::2 UMMAs,
// M=256 x N=128 per pair tile, each CTA streams its 128-row half of every weight tile, epilogue exchange through
// DSMEM) but the layer-to-layer dependency is tracked per HALF of the K dimension, so the tensor cores no longer
// idle while the epilogue of a layer's last output tile runs.
//
// A 512-wide layer is a 2x2 block product   D_mp = W[mp][0] X[0] + W[mp][1] X[1]   (mp = 256-row pair tile,
// X[h] = input features [256h, 256h+256)).  Epilogue stage E_mp turns D_mp into X'[mp] of the next layer.
// UMMA order per layer:   (mp0,X0) (mp1,X0) (mp0,X1) | commit D_0 | (mp1,X1) | commit D_1
//   * E_0 starts after 3/4 of the layer's UMMAs, and X[0] has been fully consumed by then, so E_0 overwrites
//     X[0] in place as it goes (no register parking);
//   * the next layer's (mp0,X'0) (mp1,X'0) only need E_0 and "D_1 drained to registers" (bar_d1);
//     they run while E_1 produces X'[1];
//   * the half of every output tile that belongs to the peer's columns travels as st.async (STAS) with
//     complete_tx on a barrier in the DESTINATION CTA (bar_lx[h], which also counts the 16 local epilogue warps),
//     so no epilogue warp ever waits for a DSMEM round trip (measured: fence.proxy.async + release.cluster
//     arrives after st.shared::cluster cost ~30 % of the epilogue warps' time).  CTA 1's warp 18 forwards
//     "bar_lx[h] complete" to the issuer in CTA 0 (bar_px[h]).
// Numerics, packed weights and epilogue math are those of mlp_kernel.cuh.
#pragma once
#include "mlp_kernel.cuh"

namespace mvsdf {

constexpr int kP2Tiles = 2;                 // 256-row pair tiles per layer (width <= 512)
constexpr int kP2Cols = 2 * kTileN;         // 128 columns per pair tile
constexpr int kP2SplitChunks = 256 / kChunkK;   // K chunks that belong to X[0]
constexpr int kP2Threads = kMlpThreads + 32;    // + the X-ready relay warp (used in CTA 1)

// part p of a layer: 0 = (mp0, X0), 1 = (mp1, X0), 2 = (mp0, X1), 3 = (mp1, X1)
__device__ __forceinline__ void p2_part(const LayerPlan& lp, int part, int& mp, int& k0, int& k1, bool& run) {
  const int n_pair_tiles = (lp.m_tiles + 1) >> 1;
  const int ks = lp.k_chunks < kP2SplitChunks ? lp.k_chunks : kP2SplitChunks;
  mp = part & 1;
  k0 = part < 2 ? 0 : ks;
  k1 = part < 2 ? ks : lp.k_chunks;
  run = mp < n_pair_tiles && k0 < k1;
}

// debug timeline: slot = role * 16384 + index (role 0 issuer, 1 epilogue warp 0 of CTA 0, 2 epilogue warp 0 of CTA 1,
// 3 producer of CTA 0)
// bottleneck-isolation flags (MlpArgs::debug) exist in the diagnostic build only: in production they are the constant 0,
// which removes a constant-bank load + branch from every epilogue element pair and every issuer stage
#ifdef MVSDF_TRACE
#define P2_DEBUG (a.debug)
#else
#define P2_DEBUG 0
#endif
#ifdef MVSDF_TRACE
#define P2_TRACE(role, index)                                                                         \
  do {                                                                                                \
    if (a.trace && blockIdx.x < 2 && (index) < 16384) a.trace[(role) * 16384 + (index)] = clock64(); \
  } while (0)
#else
#define P2_TRACE(role, index) do { } while (0)
#endif

// LP = 1: single-product "screening" precision -- only W_hi X_hi^T (fp16 operands, ~1e-3 absolute SDF error).  The pair
// tile is 256 columns wide (128 points per CTA): the second 64 columns of a CTA live where the exact kernel keeps the lo
// parts -- in the activation operand (feature-block slots 8-15, so one N=256 UMMA descriptor walk covers both) and in TMEM
// (the D_b columns) -- so every buffer, barrier and byte count of the exact kernel is reused unchanged: one N=256 UMMA
// per K step instead of three N=128 ones (2/3 of the tensor work for twice the points), the hi half of each weight
// stage only.  Used by the tracer to decide which of the 100 samples per ray need the exact evaluation at all
// (csrc/tracer.cu, prefilter); never for an output.
// SAVE = 1 (training forward): the operand of every layer is also written to global memory in the K-sliced layout of
// mlp_kernel.cuh (save_addr) for the native backward; a 128-column pair tile g is the 64-column tiles 2g (CTA 0) and 2g + 1.
template <int KIND, int MODE, int LP = 0, int SAVE = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kP2Threads, 1) mlp_pair2_kernel(const MlpArgs a) {
  // K chunks per weight-ring stage.  LP: two hi tiles (K = 64) per stage -- at one N=256 UMMA per K step a 32-wide
  // stage is 256 tensor cycles, less than the issuer thread's per-stage latency (~350 cycles)
  constexpr int kChunksPerStage = LP ? 2 : 1;
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
  const uint32_t bar_full = s_bar;                        // kStages: my weight half has landed (leader: AND the peer's half)
  const uint32_t bar_empty = s_bar + 8 * kStages;         // kStages: the UMMAs reading the stage retired (both CTAs)
  const uint32_t bar_acc = s_bar + 16 * kStages;          // kP2Tiles: accumulators of pair tile mp complete (both CTAs)
  const uint32_t bar_lx = bar_acc + 8 * kP2Tiles;         // 2: half h of MY B operand is complete: 16 local warps + the peer's st.async bytes
  const uint32_t bar_px = bar_lx + 16;                    // 2 (leader only): the peer's bar_lx[h] completed
  const uint32_t bar_d1 = bar_px + 16;                    // (leader only) accumulator 1 drained to registers in BOTH CTAs
  const uint32_t s_tmem = bar_d1 + 8;
  const uint32_t bar_head = s_tmem + 8;                   // fused head: 16 local warps + the peer's partial sums (st.async bytes)
  const uint32_t s_part = s_bar + 256;                    // fused head: float [8 = rows' CTA * 4 + lane quarter][64 columns of mine]
  // Fused head (exact SDF-only evaluations): the head layer has ONE output row, but as an UMMA it costs two full parts
  // (M = 256 rows, K = 512: 6.6 % of a tile's tensor time).  Instead the epilogue of the last hidden layer multiplies its
  // fp32 softplus outputs with the head weight of its feature, reduces over the 32 features of the warp with a shuffle
  // butterfly (the same tree for every column, so a point's value does not depend on where it sits), adds the two 256-row
  // stages, and leaves one partial sum per (rows' CTA, lane quarter, column); thread c of the column's own CTA adds the
  // eight partials in a fixed order.  The activations never pass through the fp16 split on this path.
  constexpr bool kCanFuse = KIND == NET_SDF && MODE == 0 && LP == 0 && SAVE == 0;
  const bool fuse = kCanFuse && a.fuse_head != 0;
  const int n_mma = fuse ? a.n_run - 1 : a.n_run;         // layers that run on the tensor cores
  uint8_t* const g_scratch = (KIND == NET_SDF) ? (smem + kStages * kStageBytes) : (smem + kStages * kStageBytes + xbytes);
  float* const scratch = reinterpret_cast<float*>(g_scratch);

  const int n_dev_count = a.n_ptr ? __shfl_sync(0xffffffffu, *a.n_ptr, 0) : 0;
  const long long n_pts = a.n_ptr ? (long long)n_dev_count : a.n;
  constexpr int kPtsPerCta = LP ? 2 * kTileN : ((MODE == 0) ? kTileN : kTileN / 4);       // points per CTA per tile
  const long long n_tiles = (n_pts + 2 * kPtsPerCta - 1) / (2 * kPtsPerCta);
  const long long pair0 = blockIdx.x >> 1;
  const long long pair_stride = gridDim.x >> 1;
  // an empty request list (most back-off passes of the sphere tracer) costs a launch and nothing else: every CTA of every
  // cluster leaves before barriers, TMEM or the cluster handshake are touched
  if (n_tiles == 0) return;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, crank == 0 ? 2 : 1);   // leader: own producer + the peer's relay
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int m = 0; m < kP2Tiles; ++m) ptx::mbar_init(bar_acc + 8 * m, 1);
    ptx::mbar_init(bar_lx, kEpiWarps);
    ptx::mbar_init(bar_lx + 8, kEpiWarps);
    ptx::mbar_init(bar_px, 1);
    ptx::mbar_init(bar_px + 8, 1);
    ptx::mbar_init(bar_d1, 2 * kEpiWarps);
    ptx::mbar_init(bar_head, kEpiWarps);
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
    // ------------------------------------------------------------------ weight producer: my 128-row half, in UMMA order
    uint32_t it = 0;
    for (long long g = pair0; g < n_tiles; g += pair_stride) {
      for (int l = 0; l < n_mma; ++l) {
        const LayerPlan& lp = a.L[l];
        for (int part = 0; part < 4; ++part) {
          int mp, k0, k1;
          bool run;
          p2_part(lp, part, mp, k0, k1, run);
          if (!run) continue;
          const int m = 2 * mp + (int)crank;
          const bool have = m < lp.m_tiles;
          const uint8_t* src = a.packed + lp.w_off + (size_t)m * lp.k_chunks * kStageBytes;
          for (int kc = k0; kc < k1; kc += kChunksPerStage, ++it) {
            const uint32_t s = it % kStages, ph = (it / kStages) & 1;
            if (P2_DEBUG & 16) ptx::mbar_wait(bar_empty + 8 * s, ph ^ 1);
            else ptx::mbar_wait_sleep(bar_empty + 8 * s, ph ^ 1);
            if (lane == 0) {
              if (have && !(P2_DEBUG & 1)) {
                if (LP) {       // the hi tiles of two consecutive K chunks share a stage (the second where lo would be)
                  const int n_sub = min(2, k1 - kc);
                  ptx::mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)n_sub * kTileBytes);
                  for (int sub = 0; sub < n_sub; ++sub)
                    ptx::bulk_g2s(s_stage + s * kStageBytes + sub * kTileBytes, src + (size_t)(kc + sub) * kStageBytes, kTileBytes,
                                  bar_full + 8 * s);
                } else {
                  ptx::mbar_arrive_expect_tx(bar_full + 8 * s, kStageBytes);
                  ptx::bulk_g2s(s_stage + s * kStageBytes, src + (size_t)kc * kStageBytes, kStageBytes, bar_full + 8 * s);
                }
              } else {
                ptx::mbar_arrive(bar_full + 8 * s);     // odd tile count: this half multiplies stale data into rows nobody reads
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (crank == 0 && warp == kEpiWarps + 1) {
    // ------------------------------------------------------------------ UMMA issuer (leader CTA)
    // Per 16-wide K step three UMMAs (M = 256 over the pair, N = 128):
    //   D_a (+)= W_hi X_hi^T     D_b (+)= W_hi X_lo^T     D_a += W_lo X_hi^T
    // (Merging the first two into one N = 256 UMMA over [X_hi ; X_lo] reads W_hi once and was measured 1 % faster, but
    //  the third product then lands on b for CTA 0's points and on a for CTA 1's: the fp32 rounding of a point would
    //  depend on which CTA processes it, and the outputs would no longer be bit-identical under re-sharding.)
    constexpr uint32_t idesc = ptx::idesc_f16_f32_bmn(2 * kTileM, LP ? 2 * kP2Cols : kP2Cols);
    const bool leader = ptx::elect_one();
    const uint32_t issue = (leader && !(P2_DEBUG & 2)) ? 1u : 0u;
    uint32_t it0 = 0, x_ctr = 0, d1_uses = 0;
    bool ready = false;      // the barrier of my next stage was already seen complete
    // shared-memory descriptors (ptx::smem_desc) as (low word, high word): only the 14-bit address field varies
    const uint32_t dlo_a = ((s_stage & 0x3FFFFu) >> 4) | ((128u >> 4) << 16);
    const uint32_t dlo_x = ((s_xhi & 0x3FFFFu) >> 4) | (((uint32_t)kBCoreStride >> 4) << 16);
    const uint32_t dlo_pe = ((s_pehi & 0x3FFFFu) >> 4) | (((uint32_t)kBCoreStride >> 4) << 16);
    constexpr uint32_t dhi_a = (512u >> 4) | (1u << 14);
    constexpr uint32_t dhi_b = (128u >> 4) | (1u << 14);
    int tr = 0;
    for (long long g = pair0; g < n_tiles; g += pair_stride) {
      for (int l = 0; l < n_mma; ++l, tr += 8) {
        const LayerPlan& lp = a.L[l];
        if (lane == 0) P2_TRACE(0, tr);
        for (int part = 0; part < 4; ++part) {
          int mp, k0, k1;
          bool run;
          p2_part(lp, part, mp, k0, k1, run);
          if (lane == 0 && (part == 1 || part == 3)) P2_TRACE(0, tr + 2 + part);     // 3: part 1 reached, 5: part 3 reached
          {
            if (part == 0) {          // X[0] of this layer complete in both CTAs (also: D_0 drained by the previous E_0)
              ptx::mbar_wait(bar_lx, x_ctr & 1);
              ptx::mbar_wait(bar_px, x_ctr & 1);
              ptx::fence_proxy_async_smem();     // observer-side fence for the bytes the peer delivered with st.async
              ptx::tc_fence_after();
              if (lane == 0) P2_TRACE(0, tr + 1);     // X0 wait done
            } else if (part == 2) {   // X[1] complete
              if (lane == 0) P2_TRACE(0, tr + 2);     // part 2 reached (parts 0,1 issued)
              ptx::mbar_wait(bar_lx + 8, x_ctr & 1);
              ptx::mbar_wait(bar_px + 8, x_ctr & 1);
              ptx::fence_proxy_async_smem();
              ptx::tc_fence_after();
              ++x_ctr;
              if (lane == 0) P2_TRACE(0, tr + 4);     // X1 wait done
            }
          }
          if (!run) continue;
          {
            if (part == 1) {          // first write into D_1 this layer: its previous contents must have been read
              if (d1_uses > 0) {
                ptx::mbar_wait(bar_d1, (d1_uses - 1) & 1);
                ptx::tc_fence_after();
              }
              ++d1_uses;
              if (lane == 0) P2_TRACE(0, tr + 6);     // d1 wait done
            }
          }
          const uint32_t d_a = tmem_base + (uint32_t)(mp * 2 * kP2Cols);
          const uint32_t d_b = d_a + kP2Cols;
          const uint32_t dlo_b = lp.b_from_pe ? dlo_pe : dlo_x;
          const int n_st = (k1 - k0 + kChunksPerStage - 1) / kChunksPerStage;
          for (int j = 0; j < n_st; ++j) {
            const int kc = k0 + j * kChunksPerStage;
            const uint32_t it = it0 + (uint32_t)j;
            const uint32_t st = it % kStages, ph = (it / kStages) & 1;
            if (!ready) ptx::mbar_wait(bar_full + 8 * st, ph);
            ptx::tc_fence_after();
            // query my next stage's barrier now: the round trip of the query hides behind the UMMA issue below
            const uint32_t itn = it + 1;
            ready = ptx::mbar_test_wait(bar_full + 8 * (itn % kStages), (itn / kStages) & 1);
            const uint32_t a_off = dlo_a + st * (kStageBytes >> 4);
            const uint32_t b_off = dlo_b + (uint32_t)kc * ((kChunkK / 8) * kBCoreStride >> 4);
            if (LP) {
              const int n_sub = min(2, k1 - kc);
#pragma unroll
              for (int sk = 0; sk < 2 * (kChunkK / 16); ++sk) {      // sk = 2 sub + ks: K steps of 16 over up to two chunks
                if (sk < n_sub * (kChunkK / 16)) {
                  const uint64_t da_hi = ptx::desc_from_words(a_off + (((sk >> 1) * kTileBytes + (sk & 1) * 256) >> 4), dhi_a);
                  const uint64_t db_hi = ptx::desc_from_words(b_off + sk * (2 * kBCoreStride >> 4), dhi_b);
                  ptx::umma1_f16_2cta(d_a, da_hi, db_hi, idesc, (kc | sk) != 0 ? 1u : 0u, issue);
                }
              }
            } else {
#pragma unroll
              for (int ks = 0; ks < kChunkK / 16; ++ks) {
                const uint64_t da_hi = ptx::desc_from_words(a_off + ks * (256 >> 4), dhi_a);
                const uint64_t da_lo = ptx::desc_from_words(a_off + ((kTileBytes + ks * 256) >> 4), dhi_a);
                const uint64_t db_hi = ptx::desc_from_words(b_off + ks * (2 * kBCoreStride >> 4), dhi_b);
                const uint64_t db_lo = ptx::desc_from_words(b_off + ((kBLoOffset + ks * 2 * kBCoreStride) >> 4), dhi_b);
                ptx::umma3_f16_2cta(d_a, d_b, da_hi, da_lo, db_hi, db_lo, idesc, (kc | ks) != 0 ? 1u : 0u, issue);
              }
            }
            if (leader) ptx::umma_commit_2cta(bar_empty + 8 * st, 3);
            __syncwarp();
          }
          if (k1 == lp.k_chunks) {      // this part completes D_mp
            if (leader) ptx::umma_commit_2cta(bar_acc + 8 * mp, 3);
            __syncwarp();
          }
          it0 += (uint32_t)n_st;
        }
        if (lane == 0) P2_TRACE(0, tr + 7);         // layer fully issued
      }
    }
  } else if (crank == 1 && warp == kEpiWarps + 1) {
    // ------------------------------------------------------------------ CTA 1, warp 17: tell the leader my stage has landed
    const uint32_t remote_full = ptx::mapa(bar_full, 0);
    uint32_t it = 0;
    for (long long g = pair0; g < n_tiles; g += pair_stride) {
      for (int l = 0; l < n_mma; ++l) {
        const LayerPlan& lp = a.L[l];
        int n_stage = 0;
        for (int part = 0; part < 4; ++part) {
          int mp, k0, k1;
          bool run;
          p2_part(lp, part, mp, k0, k1, run);
          if (run) n_stage += (k1 - k0 + kChunksPerStage - 1) / kChunksPerStage;
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
      for (int l = 0; l < n_mma; ++l) {
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
    const int q = warp & 3;              // TMEM lane quarter
    const int cg = warp >> 2;            // 16-column group: this warp produces columns [16cg,+16) of MY 64 columns
                                         // (local stores) and of the PEER's 64 columns (DSMEM stores)
    const int row = q * 32 + lane;       // row inside my 128-row half
    const int t = threadIdx.x;
    const int lcol0 = cg * 16;           // first column inside the destination CTA's 64
    const uint32_t dst_xhi = ptx::mapa(s_xhi, crank ^ 1u);   // the peer's activation buffer in the cluster window
    const uint32_t dst_xlo = dst_xhi + kBLoOffset;
    const uint32_t dst_lx = ptx::mapa(bar_lx, crank ^ 1u);   // the barrier that counts the bytes I send to the peer
    const uint32_t remote_d1 = ptx::mapa(bar_d1, 0);
    constexpr float kInvScale = 1.0f / (kWeightScale * kActScale);
    uint32_t acc_ctr[kP2Tiles] = {0, 0};
    uint32_t head_ctr = 0;
    int tr = 0;
    const bool tracer = threadIdx.x == 0;

    for (long long g = pair0; g < n_tiles; g += pair_stride) {
      const long long p0 = g * (2 * kPtsPerCta) + (long long)crank * kPtsPerCta;      // my first point
      if (tracer) P2_TRACE(1 + crank, tr);            // prologue start
      tr += 8;
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
        for (int cidx = t; cidx < (LP ? 2 * kTileN : kTileN) * 64; cidx += kEpiThreads) {
          const int col = cidx >> 6, k = cidx & 63;        // LP: columns 64-127 fall into slots 8-15 of the feature block
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
          if (LP) ptx::st_shared_u16(s_pehi + o, __half_as_ushort(__float2half_rn(v * kActScale)));
          else store_split(s_pehi + o, s_pelo + o, v * kActScale);
          if (SAVE) {
            const __half h = __float2half_rn(v * kActScale);
            const __half lo = __float2half_rn(v * kActScale - __half2float(h));
            uint8_t* gsv = a.save + a.save_off[0] + (size_t)(2 * g + crank) * (kPeCores * kBCoreStride) + save_addr(kPeCores, k, col >> 3) +
                           (col & 7) * 2;
            *reinterpret_cast<__half*>(gsv) = h;
            *reinterpret_cast<__half*>(gsv + 256) = lo;
          }
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
          if (SAVE) {
            const int kc0 = kpad >> 3;
            const __half h = __float2half_rn(v * kActScale);
            const __half lo = __float2half_rn(v * kActScale - __half2float(h));
            uint8_t* gsv = a.save + a.save_off[0] + (size_t)(2 * g + crank) * ((size_t)kc0 * kBCoreStride) + save_addr(kc0, k, col >> 3) +
                           (col & 7) * 2;
            *reinterpret_cast<__half*>(gsv) = h;
            *reinterpret_cast<__half*>(gsv + 256) = lo;
          }
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {                 // the whole first-layer operand (my columns, local writes only) is ready: both K halves
        ptx::mbar_arrive(bar_lx);
        ptx::mbar_arrive(bar_lx + 8);
      }

      if (tracer) P2_TRACE(1 + crank, tr - 7);        // prologue done
      // ---------------- layers
      for (int l = 0; l < n_mma; ++l, tr += 8) {
        const LayerPlan& lp = a.L[l];
        const bool last = (l == n_mma - 1);
        const int n_pair_tiles = (lp.m_tiles + 1) >> 1;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool skip_src = KIND == NET_SDF && l == a.skip_layer - 1;
        float head_acc = 0.0f;            // fused head: lane L = partial sum of column (L & 15) of the peer (L < 16) / my half
        float bias_r[kP2Tiles];
#pragma unroll
        for (int mp = 0; mp < kP2Tiles; ++mp) {
          const int m = 2 * mp + (int)crank;
          bias_r[mp] = m < lp.m_tiles ? __ldg(a.bias + lp.bias_off + m * kTileM + row) : 0.f;
        }
#pragma unroll
        for (int mp = 0; mp < kP2Tiles; ++mp) {
          if (mp < n_pair_tiles) {
            if (tracer) P2_TRACE(1 + crank, tr + 4 * mp);          // waiting for D_mp
            if (P2_DEBUG & 16) ptx::mbar_wait(bar_acc + 8 * mp, acc_ctr[mp] & 1);
            else ptx::mbar_wait_sleep(bar_acc + 8 * mp, acc_ctr[mp] & 1);
            ++acc_ctr[mp];
            ptx::tc_fence_after();
            if (tracer) P2_TRACE(1 + crank, tr + 4 * mp + 1);      // D_mp ready
            const int m = 2 * mp + (int)crank;
            const int f = m * kTileM + row;                 // feature (output row) this thread owns
            const bool have = m < lp.m_tiles;
            const float bias = bias_r[mp];
            // drain my 16 own + 16 peer columns of both accumulators; columns 0-63 of the pair tile are CTA 0's
            uint32_t va[2][16], vb[2][16];
#pragma unroll
            for (int hcol = 0; hcol < 2; ++hcol) {
              const uint32_t dest_h = hcol == 0 ? (crank ^ 1u) : crank;     // peer columns first: their st.async overlaps my own half
              const uint32_t tcol = (uint32_t)(mp * 2 * kP2Cols) + dest_h * kTileN + (uint32_t)lcol0;
              if (LP) {       // N = 256: columns [128 dest_h, +128) are dest_h's; vb = its second 64-column block
                ptx::tmem_ld_32x16(t_row + (uint32_t)(mp * 2 * kP2Cols) + dest_h * (2 * kTileN) + (uint32_t)lcol0, va[hcol]);
                ptx::tmem_ld_32x16(t_row + (uint32_t)(mp * 2 * kP2Cols) + dest_h * (2 * kTileN) + kTileN + (uint32_t)lcol0, vb[hcol]);
              } else {
                ptx::tmem_ld_32x16(t_row + tcol, va[hcol]);
                ptx::tmem_ld_32x16(t_row + tcol + kP2Cols, vb[hcol]);
              }
            }
            ptx::tmem_ld_wait();
            if (tracer) P2_TRACE(1 + crank, tr + 4 * mp + 2);      // drained
            if (mp == 1) {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive_remote_relaxed(remote_d1);
            }
            const bool write = have && !(skip_src && f >= a.skip_rows_begin);
            const uint32_t o0 = xoff(lcol0, f);
            float head_w = 0.0f;
            if (kCanFuse && fuse && last && have) {
              // row 0 of the head layer's packed tile (hi + lo = kWeightScale * w_head[f]): K chunk f / 32, core f % 32 / 8
              const LayerPlan& hl = a.L[a.n_run - 1];
              const uint8_t* wt = a.packed + hl.w_off + (size_t)(f >> 5) * kStageBytes + ((f & 31) >> 3) * 128 + (f & 7) * 2;
              head_w = __half2float(*reinterpret_cast<const __half*>(wt)) + __half2float(*reinterpret_cast<const __half*>(wt + kTileBytes));
            }
#pragma unroll
            for (int hcol = 0; hcol < 2; ++hcol) {
              const uint32_t dest_h = hcol == 0 ? (crank ^ 1u) : crank;     // peer columns first: their st.async overlaps my own half
              if (kCanFuse && fuse && last) {
                // this thread's 16 products  w_head[f] * softplus(z[f, column])  (scaled by kWeightScale * kActScale) ...
                constexpr float kK = kInvScale * kSpT;
                const float bias_k = bias * kSpT;
                float hprod[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float t0 = fmaf(__uint_as_float(va[hcol][j]), kK, fmaf(__uint_as_float(vb[hcol][j]), kK, bias_k));
                  hprod[j] = have ? head_w * softplus_t_scaled(t0) : 0.0f;
                }
                // ... summed over the warp's 32 features: lanes L and L ^ 16 first add their 16 values, then every step halves
                // the values a lane keeps (those whose index has the lane's bit), so lane L ends with the total of column L & 15
#pragma unroll
                for (int j = 0; j < 16; ++j) hprod[j] += __shfl_xor_sync(0xffffffffu, hprod[j], 16);
#pragma unroll
                for (int sft = 8, nv = 16; sft >= 1; sft >>= 1, nv >>= 1) {
                  const bool upper = (lane & sft) != 0;
#pragma unroll
                  for (int i = 0; i < nv / 2; ++i) {
                    const float send = upper ? hprod[i] : hprod[i + nv / 2];
                    const float keep = upper ? hprod[i + nv / 2] : hprod[i];
                    hprod[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
                  }
                }
                if ((lane >= 16) == (hcol == 1)) head_acc += hprod[0];     // lanes 0-15 carry the peer's columns, 16-31 mine
                continue;
              }
              if (!last) {
                uint32_t phi[8], plo[8];
                if (MODE == 0) {
                  // pre-activation straight in the softplus / ReLU domain: two FMAs fold the accumulator sum,
                  // the un-scaling and the bias
                  constexpr float kK = (KIND == NET_SDF) ? kInvScale * kSpT : kInvScale * kActScale;
                  const float bias_k = bias * ((KIND == NET_SDF) ? kSpT : kActScale);
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    if (LP) {       // two independent column blocks; "plo" carries the hi parts of columns 64-127
                      const float ta0 = fmaf(__uint_as_float(va[hcol][2 * i]), kK, bias_k);
                      const float ta1 = fmaf(__uint_as_float(va[hcol][2 * i + 1]), kK, bias_k);
                      const float tb0 = fmaf(__uint_as_float(vb[hcol][2 * i]), kK, bias_k);
                      const float tb1 = fmaf(__uint_as_float(vb[hcol][2 * i + 1]), kK, bias_k);
                      if (P2_DEBUG & 8) {      // diagnostic build only: ReLU instead of softplus (epilogue issue load / 3)
                        phi[i] = pack_hi(fmaxf(ta0, 0.f), fmaxf(ta1, 0.f));
                        plo[i] = pack_hi(fmaxf(tb0, 0.f), fmaxf(tb1, 0.f));
                        continue;
                      }
                      phi[i] = pack_hi(softplus_t_scaled_screen(ta0), softplus_t_scaled_screen(ta1));
                      plo[i] = pack_hi(softplus_t_scaled_screen(tb0), softplus_t_scaled_screen(tb1));
                      continue;
                    }
                    const float t0 = fmaf(__uint_as_float(va[hcol][2 * i]), kK, fmaf(__uint_as_float(vb[hcol][2 * i]), kK, bias_k));
                    const float t1 = fmaf(__uint_as_float(va[hcol][2 * i + 1]), kK, fmaf(__uint_as_float(vb[hcol][2 * i + 1]), kK, bias_k));
                    if (KIND == NET_RENDER) {     // fmaxf(NaN, 0) = 0 would swallow an overflow; t = kActScale * z here
                      checked(t0, a.status, 65504.0f);
                      checked(t1, a.status, 65504.0f);
                    }
                    const float y0 = (KIND == NET_SDF && !(P2_DEBUG & 8)) ? softplus_t_scaled(t0) : fmaxf(t0, 0.0f);
                    const float y1 = (KIND == NET_SDF && !(P2_DEBUG & 8)) ? softplus_t_scaled(t1) : fmaxf(t1, 0.0f);
                    pack_split_fh(y0, y1, phi[i], plo[i]);
                  }
                } else {
                  constexpr float kK = kInvScale * kSpT;
                  const float bias_k = bias * kSpT;
#pragma unroll
                  for (int gq = 0; gq < 4; ++gq) {
                    const float t = fmaf(__uint_as_float(va[hcol][4 * gq]), kK, fmaf(__uint_as_float(vb[hcol][4 * gq]), kK, bias_k));
                    float sg;
                    const float y = softplus_t_scaled_grad(t, sg);
                    const float ts = sg * (kInvScale * kActScale);
                    const float d0 = __uint_as_float(va[hcol][4 * gq + 1]) + __uint_as_float(vb[hcol][4 * gq + 1]);
                    const float d1 = __uint_as_float(va[hcol][4 * gq + 2]) + __uint_as_float(vb[hcol][4 * gq + 2]);
                    const float d2 = __uint_as_float(va[hcol][4 * gq + 3]) + __uint_as_float(vb[hcol][4 * gq + 3]);
                    pack_split_fh(y, d0 * ts, phi[2 * gq], plo[2 * gq]);
                    pack_split_fh(d1 * ts, d2 * ts, phi[2 * gq + 1], plo[2 * gq + 1]);
                  }
                }
                if (SAVE && write) {
                  const int kc_next = a.L[l + 1].k_chunks * (kChunkK / 8);
                  if (f < kc_next * 8) {
                    uint8_t* gsv = a.save + a.save_off[l + 1] + (size_t)(2 * g + dest_h) * ((size_t)kc_next * kBCoreStride) +
                                   save_addr(kc_next, f, cg * 2);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                      *reinterpret_cast<uint4*>(gsv + j * 128) = make_uint4(phi[4 * j], phi[4 * j + 1], phi[4 * j + 2], phi[4 * j + 3]);
                      *reinterpret_cast<uint4*>(gsv + 256 + j * 128) = make_uint4(plo[4 * j], plo[4 * j + 1], plo[4 * j + 2], plo[4 * j + 3]);
                    }
                  }
                }
                // every UMMA that reads X[mp] of the CURRENT layer retired before D_mp was committed: write in place
                if (write) {
                  if (hcol == 1) {
#pragma unroll
                    for (int j = 0; j < 2 && !(P2_DEBUG & 32); ++j) {      // bit 5 (diagnostic build): skip the local stores
                      ptx::st_shared_v4(s_xhi + o0 + j * 128, phi[4 * j], phi[4 * j + 1], phi[4 * j + 2], phi[4 * j + 3]);
                      ptx::st_shared_v4(s_xlo + o0 + j * 128, plo[4 * j], plo[4 * j + 1], plo[4 * j + 2], plo[4 * j + 3]);
                    }
                  } else if (!(P2_DEBUG & 4)) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                      ptx::st_async_v4(dst_xhi + o0 + j * 128, phi[4 * j], phi[4 * j + 1], phi[4 * j + 2], phi[4 * j + 3], dst_lx + 8 * mp);
                      ptx::st_async_v4(dst_xlo + o0 + j * 128, plo[4 * j], plo[4 * j + 1], plo[4 * j + 2], plo[4 * j + 3], dst_lx + 8 * mp);
                    }
                  }
                }
              } else if (have) {
                // ---------------- head: write results to global memory
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int ccta = (int)dest_h, lc = lcol0 + j;   // owner CTA of the column and its index there
                  if (LP) {       // SDF-only head, two column blocks
                    const long long gp = p0_pair + (long long)ccta * kPtsPerCta + lc;
                    if (f == 0) {
                      if (gp < n_pts) a.out_sdf[gp] = checked(__uint_as_float(va[hcol][j]) * kInvScale + bias, a.status);
                      if (gp + kTileN < n_pts) a.out_sdf[gp + kTileN] = checked(__uint_as_float(vb[hcol][j]) * kInvScale + bias, a.status);
                    }
                    continue;
                  }
                  const float acc = (__uint_as_float(va[hcol][j]) + __uint_as_float(vb[hcol][j])) * kInvScale;
                  if (KIND == NET_RENDER) {
                    const long long gp = p0_pair + (long long)ccta * kPtsPerCta + lc;
                    if (row < 3 && m == 0 && gp < n_pts) a.out_rgb[gp * 3 + row] = tanhf(checked(acc + bias, a.status));
                  } else {
                    const long long gp = p0_pair + (long long)ccta * kPtsPerCta + ((MODE == 0) ? lc : (lc >> 2));
                    const int jj = (MODE == 0) ? 0 : (lc & 3);
                    if (gp < n_pts) {
                      if (a.head == HEAD_SDF_ONLY) {
                        if (f == 0) {
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
            if (!last) {
              if (skip_src && mp == n_pair_tiles - 1) {
                // skip connection: features [skip_rows_begin, +pe_dim) of MY 64 columns are the positional encoding
                // (already scaled and split in my PE tile) -- each CTA fills them for its own columns
                for (int cidx = t; cidx < a.pe_dim * 16; cidx += kEpiThreads) {
                  const int k = cidx >> 4, blk = cidx & 15;           // 16 column blocks of 16 bytes per feature: 8 hi + 8 lo
                  const int kd = a.skip_rows_begin + k;
                  const uint4 pv = ptx::ld_shared_v4(s_pehi + (uint32_t)((k >> 3) * kBCoreStride + (k & 7) * 16 + blk * 128));
                  ptx::st_shared_v4(s_xhi + (uint32_t)((kd >> 3) * kBCoreStride + (kd & 7) * 16 + blk * 128), pv.x, pv.y, pv.z, pv.w);
                  if (SAVE) {      // blocks 0-7: hi parts of column blocks 0-7, blocks 8-15: their lo parts
                    const int kc_next = a.L[l + 1].k_chunks * (kChunkK / 8);
                    uint8_t* gsv = a.save + a.save_off[l + 1] + (size_t)(2 * g + crank) * ((size_t)kc_next * kBCoreStride) +
                                   save_addr(kc_next, kd, blk & 7) + (blk >> 3) * 256;
                    *reinterpret_cast<uint4*>(gsv) = pv;
                  }
                }
              }
              // my rows of X'[mp] for my own columns are written (the peer's rows arrive as st.async bytes on the
              // same barrier): publish to the async proxy and arrive; warp 0 also arms the expected byte count
              ptx::tc_fence_before();
              ptx::fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                if (warp == 0) {
                  const int pm = 2 * mp + (int)(crank ^ 1u);       // the peer's output tile: 256 B per row it writes
                  int rows = pm < lp.m_tiles ? kTileM : 0;
                  if (skip_src) rows = min(rows, max(a.skip_rows_begin - pm * kTileM, 0));
                  if (rows > 0 && !(P2_DEBUG & 4)) ptx::mbar_arrive_expect_tx(bar_lx + 8 * mp, (uint32_t)rows * (kTileN * 4));
                  else ptx::mbar_arrive(bar_lx + 8 * mp);
                } else {
                  ptx::mbar_arrive(bar_lx + 8 * mp);
                }
                if (n_pair_tiles == 1) ptx::mbar_arrive(bar_lx + 8);
              }
              if (tracer) P2_TRACE(1 + crank, tr + 4 * mp + 3);    // E_mp done (arrived)
            }
          }
        }
        if (kCanFuse && fuse && last) {
          // partial sums of this warp's 32 rows (both 256-row stages): my half's columns stay here, the peer's travel;
          // slot = CTA that owns the ROWS (not local / remote), so both CTAs add the eight partials in the same order
          const uint32_t slot = (crank * 4u + (uint32_t)q) * (kTileN * 4) + (uint32_t)(lcol0 + (lane & 15)) * 4;
          if (lane >= 16) ptx::st_shared_f32(s_part + slot, head_acc);
          else ptx::st_async_b32(ptx::mapa(s_part, crank ^ 1u) + slot, __float_as_uint(head_acc), ptx::mapa(bar_head, crank ^ 1u));
          __syncwarp();
          if (lane == 0) {
            if (warp == 0) ptx::mbar_arrive_expect_tx(bar_head, kEpiWarps * 16 * 4);     // the peer's 16 warps x 16 columns x 4 B
            else ptx::mbar_arrive(bar_head);
          }
          if (t < kTileN) {
            ptx::mbar_wait(bar_head, head_ctr & 1);
            float sdf = __ldg(a.bias + a.L[a.n_run - 1].bias_off);
            float acc = 0.0f;
#pragma unroll
            for (int sidx = 0; sidx < 8; ++sidx) acc += ptx::ld_shared_f32(s_part + (uint32_t)(sidx * kTileN + t) * 4);
            sdf = fmaf(acc, kInvScale, sdf);
            const long long gp = p0 + t;
            if (gp < n_pts) a.out_sdf[gp] = checked(sdf, a.status);
          }
          ++head_ctr;
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

// C ABI, part 3: surface shading and the multi-view feature-consistency / rgb losses
// (rows a11-a16 of SURVEY.md section 8).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "../../include/mvsdf_b200.h"
#include "internal.h"

namespace mvsdf {

constexpr int kBlk = 256;
constexpr int kScanItems = 2048;       // rays per block of the hit compaction
constexpr int kHitChunk = 1 << 18;     // hit points shaded per pass (bounds the [chunk, 2+F] feature buffer)

// ------------------------------------------------------------------ order-preserving compaction of the hit rays
// points[surface_mask] / ray_dirs[surface_mask] (implicit_differentiable_renderer.py:207-213, :296) without nonzero():
// per-block counts -> single-block exclusive scan -> scatter.
__global__ void hit_count_kernel(const uint8_t* __restrict__ mask, int R, int* __restrict__ block_counts) {
  __shared__ int sh[kBlk / 32];
  const int base = blockIdx.x * kScanItems;
  int cnt = 0;
  for (int i = threadIdx.x; i < kScanItems; i += kBlk) {
    const int r = base + i;
    cnt += (r < R && mask[r]) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kBlk / 32; ++w) t += sh[w];
    block_counts[blockIdx.x] = t;
  }
}

// single block: exclusive scan of block counts; also the per-image offsets (rays are image-major) and chunk counts
__global__ void hit_scan_kernel(int* __restrict__ block_counts, int n_blocks, int n_images, int n_pixels,
                                const uint8_t* __restrict__ mask, int* __restrict__ img_offsets,
                                int* __restrict__ chunk_counts, int n_chunks, int* __restrict__ miss_count) {
  __shared__ int carry;
  __shared__ int sh[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n_blocks ? block_counts[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n_blocks) block_counts[i] = carry + sh[threadIdx.x] - v;   // exclusive
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  // per-image offsets: offset of the first ray of image b = exclusive count at ray b*N
  for (int b = threadIdx.x; b <= n_images; b += blockDim.x) {
    if (b == n_images) {
      img_offsets[b] = carry;
    } else {
      const long long r0 = (long long)b * n_pixels;
      const int blk = (int)(r0 / kScanItems);
      int cnt = block_counts[blk];
      for (long long r = (long long)blk * kScanItems; r < r0; ++r) cnt += mask[r] ? 1 : 0;
      img_offsets[b] = cnt;
    }
  }
  for (int c = threadIdx.x; c < n_chunks; c += blockDim.x)
    chunk_counts[c] = min(max(carry - c * kHitChunk, 0), kHitChunk);
  if (threadIdx.x == 0 && miss_count) *miss_count = n_images * n_pixels - carry;
}

__global__ void hit_scatter_kernel(const uint8_t* __restrict__ mask, int R, int n_pixels, const int* __restrict__ block_offsets,
                                   const float* __restrict__ points, const float* __restrict__ dirs,
                                   int* __restrict__ hit_index, float* __restrict__ pts_hit, float* __restrict__ view_hit,
                                   int* __restrict__ miss_index, float* __restrict__ pts_miss) {
  // one warp handles 32 consecutive rays at a time so that the order is preserved with ballots
  __shared__ int warp_base[kBlk / 32];
  __shared__ int sh_cnt[kBlk / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int base = blockIdx.x * kScanItems;
  const int per_warp = kScanItems / (kBlk / 32);   // 256 consecutive rays per warp
  int cnt = 0;
  for (int i = lane; i < per_warp; i += 32) {
    const int r = base + warp * per_warp + i;
    cnt += (r < R && mask[r]) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) sh_cnt[warp] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = block_offsets[blockIdx.x];
    for (int w = 0; w < kBlk / 32; ++w) {
      warp_base[w] = t;
      t += sh_cnt[w];
    }
  }
  __syncthreads();
  int out = warp_base[warp];
  for (int i = 0; i < per_warp; i += 32) {
    const int r = base + warp * per_warp + i + lane;
    const bool hit = r < R && mask[r];
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const int pos = out + __popc(bal & ((1u << lane) - 1));
      hit_index[pos] = r;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        pts_hit[3 * (size_t)pos + k] = points[3 * (size_t)r + k];
        view_hit[3 * (size_t)pos + k] = -dirs[3 * (size_t)r + k];
      }
    } else if (r < R && miss_index) {
      const int pos = r - (out + __popc(bal & ((1u << lane) - 1)));      // rays before r minus surface rays before r
      miss_index[pos] = r;
#pragma unroll
      for (int k = 0; k < 3; ++k) pts_miss[3 * (size_t)pos + k] = points[3 * (size_t)r + k];
    }
    out += __popc(bal);
  }
}

__global__ void fill_ones_kernel(float* __restrict__ p, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 1.0f;
}

// rgb_values[surface_mask] = rgb (:301-304); also keeps (sdf, indicator) of the hit points
__global__ void shade_scatter_kernel(const int* __restrict__ chunk_count, int begin, const int* __restrict__ hit_index,
                                     const float* __restrict__ rgb_hit, const float* __restrict__ full, int full_stride,
                                     float* __restrict__ rgb_values, float* __restrict__ surf_head, float* __restrict__ sdf_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *chunk_count) return;
  const int pos = begin + i;
  const int r = hit_index[pos];
  if (sdf_out) sdf_out[r] = full[(size_t)i * full_stride];      // sdf_output of a surface ray: the value column of this pass
#pragma unroll
  for (int k = 0; k < 3; ++k) rgb_values[3 * (size_t)r + k] = rgb_hit[3 * (size_t)pos + k];
  if (surf_head) {
    surf_head[2 * (size_t)pos] = full[(size_t)i * full_stride];
    surf_head[2 * (size_t)pos + 1] = full[(size_t)i * full_stride + 1];
  }
}

__global__ void miss_scatter_kernel(const int* __restrict__ miss_count, const int* __restrict__ miss_index,
                                    const float* __restrict__ val, float* __restrict__ sdf_out) {
  const int n = *miss_count;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) sdf_out[miss_index[i]] = val[i];
}

// ------------------------------------------------------------------ feature maps: NCHW -> channels-last
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int C, int hw, float* __restrict__ dst) {
  // block: 32 pixels x C(<=32) channels through shared memory so both sides are coalesced
  __shared__ float tile[32][33];
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int c = ty; c < C; c += 8) {
    const int p = p0 + tx;
    tile[c][tx] = p < hw ? src[((size_t)n * C + c) * hw + p] : 0.f;
  }
  __syncthreads();
  for (int pp = ty; pp < 32; pp += 8) {
    const int p = p0 + pp;
    if (p < hw && tx < C) dst[((size_t)n * hw + p) * C + tx] = tile[tx][pp];
  }
}

// ------------------------------------------------------------------ a15: multi-view feature consistency
// IDRLoss.get_feat_loss_corr (code/model/loss.py:115-165) with the helpers of code/utils/my_utils.py:98-165.
// One warp per surface point, lane = feature channel: every bilinear tap of the channels-last map is one
// coalesced 128-byte load (the reference's NCHW grid_sample touches 32 sectors per tap).
struct FeatArgs {
  const float* pts;        // [M,3] packed by image
  const int* offsets;      // [B+1]
  const float* cams;       // [B,V,2,4,4]
  const float* maps;       // [B,V,h,w,32]
  const float* size;       // [1]
  const float* center;     // [3]
  double* partial;         // [B,2] (sum, count)
  int B, V, h, w;
  const int* map_index;    // optional [B,V]: map (b, v) is maps[map_index[b*V+v]] (scene feature store); nullptr: b*V+v
};

__device__ __forceinline__ float dot4(const float* m, float x, float y, float z, float w) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], z)), __fmul_rn(m[3], w));
}
__device__ __forceinline__ float dot3(const float* m, float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], z));
}

__device__ __forceinline__ float sample_bilinear(const float* __restrict__ map, int h, int w, float gx, float gy, int lane) {
  // F.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=False)
  const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)w), 1.f), 0.5f);
  const float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)h), 1.f), 0.5f);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = __fsub_rn(ix, x0f), wx0 = __fsub_rn(x0f + 1.f, ix);
  const float wy1 = __fsub_rn(iy, y0f), wy0 = __fsub_rn(y0f + 1.f, iy);
  float acc = 0.f;
  const bool xin0 = x0 >= 0 && x0 < w, xin1 = x1 >= 0 && x1 < w, yin0 = y0 >= 0 && y0 < h, yin1 = y1 >= 0 && y1 < h;
  if (xin0 && yin0) acc = __fadd_rn(acc, __fmul_rn(__ldg(map + ((size_t)y0 * w + x0) * 32 + lane), __fmul_rn(wx0, wy0)));
  if (xin1 && yin0) acc = __fadd_rn(acc, __fmul_rn(__ldg(map + ((size_t)y0 * w + x1) * 32 + lane), __fmul_rn(wx1, wy0)));
  if (xin0 && yin1) acc = __fadd_rn(acc, __fmul_rn(__ldg(map + ((size_t)y1 * w + x0) * 32 + lane), __fmul_rn(wx0, wy1)));
  if (xin1 && yin1) acc = __fadd_rn(acc, __fmul_rn(__ldg(map + ((size_t)y1 * w + x1) * 32 + lane), __fmul_rn(wx1, wy1)));
  return acc;
}

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void feat_loss_kernel(FeatArgs a) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const int M = a.offsets[a.B];
  const float size = a.size[0];
  const float cx = a.center[0], cy = a.center[1], cz = a.center[2];
  int img = 0;
  int cur_img = -1;
  double acc = 0.0;
  for (int p = warp_global; p < M; p += n_warps) {
    while (img + 1 < a.B && p >= a.offsets[img + 1]) ++img;
    if (img != cur_img) {
      if (cur_img >= 0 && lane == 0 && acc != 0.0) atomicAdd(a.partial + 2 * cur_img, acc);
      acc = 0.0;
      cur_img = img;
    }
    // pts_world = pts / 2 * size + center  (loss.py:132)
    const float X = __fadd_rn(__fmul_rn(__fmul_rn(a.pts[3 * (size_t)p], 0.5f), size), cx);
    const float Y = __fadd_rn(__fmul_rn(__fmul_rn(a.pts[3 * (size_t)p + 1], 0.5f), size), cy);
    const float Z = __fadd_rn(__fmul_rn(__fmul_rn(a.pts[3 * (size_t)p + 2], 0.5f), size), cz);
    float f0 = 0.f, n0 = 0.f;
    bool in0 = false;
    for (int v = 0; v < a.V; ++v) {
      const float* cam = a.cams + ((size_t)img * a.V + v) * 32;
      // idx_world2cam (my_utils.py:98-102)
      float c0 = dot4(cam, X, Y, Z, 1.f), c1 = dot4(cam + 4, X, Y, Z, 1.f), c2 = dot4(cam + 8, X, Y, Z, 1.f),
            c3 = dot4(cam + 12, X, Y, Z, 1.f);
      const float d0 = __fadd_rn(c3, 1e-9f);
      c0 = __fdiv_rn(c0, d0);
      c1 = __fdiv_rn(c1, d0);
      c2 = __fdiv_rn(c2, d0);
      c3 = __fdiv_rn(c3, d0);
      // idx_cam2img (my_utils.py:105-110)
      const float d1 = __fadd_rn(c3, 1e-9f);
      const float x = __fdiv_rn(c0, d1), y = __fdiv_rn(c1, d1), z = __fdiv_rn(c2, d1);
      const float* K = cam + 16;
      float u = dot3(K, x, y, z), vv = dot3(K + 4, x, y, z), ww = dot3(K + 8, x, y, z);
      const float d2 = __fadd_rn(ww, 1e-9f);
      u = __fdiv_rn(u, d2);
      vv = __fdiv_rn(vv, d2);
      // grid/2, normalize_for_grid_sample (my_utils.py:152-156), get_in_range (:159-165)
      float gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fmul_rn(u, 0.5f), (float)a.w), 2.f), 1.f);
      float gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fmul_rn(vv, 0.5f), (float)a.h), 2.f), 1.f);
      gx = fminf(fmaxf(gx, -1.1f), 1.1f);
      gy = fminf(fmaxf(gy, -1.1f), 1.1f);
      const bool in = gx <= 1.f && gx >= -1.f && gy <= 1.f && gy >= -1.f;
      const size_t mi = a.map_index ? (size_t)a.map_index[img * a.V + v] : (size_t)img * a.V + v;
      const float* map = a.maps + mi * (size_t)a.h * a.w * 32;
      const float f = sample_bilinear(map, a.h, a.w, gx, gy, lane);
      const float nrm = sqrtf(warp_sum(f * f));
      if (v == 0) {
        f0 = f;
        n0 = nrm;
        in0 = in;
      } else {
        const float dt = warp_sum(f0 * f);
        const float corr = __fdiv_rn(__fdiv_rn(dt, fmaxf(n0, 1e-9f)), fmaxf(nrm, 1e-9f));
        const float l = fabsf(__fsub_rn(1.f, corr));
        if (in0 && in && l < 0.5f) acc += (double)l;
      }
    }
  }
  if (cur_img >= 0 && lane == 0 && acc != 0.0) atomicAdd(a.partial + 2 * cur_img, acc);
}

// ---- backward of the feature-consistency term w.r.t. the surface points (row f1, first native piece) -----------------
// d loss / d pts for loss = mean_i( sum_{v>=1, kept} |1 - corr_v| / count_i ): the feature maps are constants, so the
// whole chain is  point -> (world2cam, cam2img: three eps-guarded divisions) -> pixel -> bilinear taps -> cosine
// similarity.  One warp per surface point, lane = channel; the 2x3 Jacobian d(ix, iy)/d pts of a view is carried in
// forward mode (identical in every lane), the channel sums are warp reductions.  Gradient flows through BOTH the
// reference-view sample and the source-view sample, and only where the forward kept the term (in range in both views --
// which also means the +-1.1 clamp is inactive -- and |1 - corr| < 0.5, loss.py:143-153).
struct ViewSample {
  float f, fx, fy;      // sampled channel value and its derivatives w.r.t. the pixel coordinates (ix, iy)
  float J[2][3];        // d(ix, iy) / d pts
  bool in;
};

__device__ __forceinline__ ViewSample project_and_sample(const float* __restrict__ cam, const float* __restrict__ map, int h, int w,
                                                         float X, float Y, float Z, float half_size, int lane) {
  ViewSample r;
  float c[4], dc[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    c[i] = dot4(cam + 4 * i, X, Y, Z, 1.f);
#pragma unroll
    for (int k = 0; k < 3; ++k) dc[i][k] = cam[4 * i + k] * half_size;       // d(world)/d(pts) = size / 2
  }
  const float d0 = __fadd_rn(c[3], 1e-9f);
  float cp[4], dcp[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    cp[i] = __fdiv_rn(c[i], d0);
#pragma unroll
    for (int k = 0; k < 3; ++k) dcp[i][k] = (dc[i][k] - cp[i] * dc[3][k]) / d0;
  }
  const float d1 = __fadd_rn(cp[3], 1e-9f);
  float x[3], dx[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    x[i] = __fdiv_rn(cp[i], d1);
#pragma unroll
    for (int k = 0; k < 3; ++k) dx[i][k] = (dcp[i][k] - x[i] * dcp[3][k]) / d1;
  }
  const float* K = cam + 16;
  float U[3], dU[3][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    U[j] = dot3(K + 4 * j, x[0], x[1], x[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) dU[j][k] = K[4 * j] * dx[0][k] + K[4 * j + 1] * dx[1][k] + K[4 * j + 2] * dx[2][k];
  }
  const float d2 = __fadd_rn(U[2], 1e-9f);
  const float u = __fdiv_rn(U[0], d2), vv = __fdiv_rn(U[1], d2);
  float gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fmul_rn(u, 0.5f), (float)w), 2.f), 1.f);
  float gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fmul_rn(vv, 0.5f), (float)h), 2.f), 1.f);
  gx = fminf(fmaxf(gx, -1.1f), 1.1f);
  gy = fminf(fmaxf(gy, -1.1f), 1.1f);
  r.in = gx <= 1.f && gx >= -1.f && gy <= 1.f && gy >= -1.f;
  // ix = ((gx + 1) w - 1) / 2 = u / 2 - 1/2  =>  d ix = d u / 2 (clamp inactive wherever the term is kept)
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    r.J[0][k] = 0.5f * (dU[0][k] - u * dU[2][k]) / d2;
    r.J[1][k] = 0.5f * (dU[1][k] - vv * dU[2][k]) / d2;
  }
  const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)w), 1.f), 0.5f);
  const float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)h), 1.f), 0.5f);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = ix - x0f, wx0 = x0f + 1.f - ix, wy1 = iy - y0f, wy0 = y0f + 1.f - iy;
  const bool xin0 = x0 >= 0 && x0 < w, xin1 = x1 >= 0 && x1 < w, yin0 = y0 >= 0 && y0 < h, yin1 = y1 >= 0 && y1 < h;
  // zeros padding: an out-of-bounds tap contributes 0 to the value and to both derivatives (F.grid_sample backward)
  const float t00 = (xin0 && yin0) ? __ldg(map + ((size_t)y0 * w + x0) * 32 + lane) : 0.f;
  const float t01 = (xin1 && yin0) ? __ldg(map + ((size_t)y0 * w + x1) * 32 + lane) : 0.f;
  const float t10 = (xin0 && yin1) ? __ldg(map + ((size_t)y1 * w + x0) * 32 + lane) : 0.f;
  const float t11 = (xin1 && yin1) ? __ldg(map + ((size_t)y1 * w + x1) * 32 + lane) : 0.f;
  r.f = t00 * wx0 * wy0 + t01 * wx1 * wy0 + t10 * wx0 * wy1 + t11 * wx1 * wy1;
  r.fx = (t01 - t00) * wy0 + (t11 - t10) * wy1;
  r.fy = (t10 - t00) * wx0 + (t11 - t01) * wx1;
  return r;
}

struct FeatBwdArgs {
  const float* pts;
  const int* offsets;
  const float* cams;
  const float* maps;
  const float* size;
  const float* center;
  const double* partial;   // [B,2]: the (global) counts (V-1) m_i are the denominators
  const float* upstream;   // [1] d L / d loss
  float* grad;             // [M,3]
  int B, V, h, w;
  const int* map_index;    // see FeatArgs
};

__global__ void feat_loss_bwd_kernel(FeatBwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const int M = a.offsets[a.B];
  const float size = a.size[0];
  const float cx = a.center[0], cy = a.center[1], cz = a.center[2];
  const float g_up = a.upstream[0];
  int img = 0;
  for (int p = warp_global; p < M; p += n_warps) {
    while (img + 1 < a.B && p >= a.offsets[img + 1]) ++img;
    const double cnt = a.partial[2 * img + 1];
    const float scale = cnt > 0.0 ? g_up / ((float)cnt * (float)a.B) : 0.f;
    const float X = __fadd_rn(__fmul_rn(__fmul_rn(a.pts[3 * (size_t)p], 0.5f), size), cx);
    const float Y = __fadd_rn(__fmul_rn(__fmul_rn(a.pts[3 * (size_t)p + 1], 0.5f), size), cy);
    const float Z = __fadd_rn(__fmul_rn(__fmul_rn(a.pts[3 * (size_t)p + 2], 0.5f), size), cz);
    const size_t map_stride = (size_t)a.h * a.w * 32;
    const size_t mi0 = a.map_index ? (size_t)a.map_index[img * a.V] : (size_t)img * a.V;
    const ViewSample s0 = project_and_sample(a.cams + ((size_t)img * a.V) * 32, a.maps + mi0 * map_stride, a.h, a.w,
                                             X, Y, Z, 0.5f * size, lane);
    const float n0 = sqrtf(warp_sum(s0.f * s0.f));
    const float N0 = fmaxf(n0, 1e-9f);
    float g[3] = {0.f, 0.f, 0.f};
    for (int v = 1; v < a.V; ++v) {
      const size_t miv = a.map_index ? (size_t)a.map_index[img * a.V + v] : (size_t)img * a.V + v;
      const ViewSample sv = project_and_sample(a.cams + ((size_t)img * a.V + v) * 32, a.maps + miv * map_stride,
                                               a.h, a.w, X, Y, Z, 0.5f * size, lane);
      const float nv = sqrtf(warp_sum(sv.f * sv.f));
      const float Nv = fmaxf(nv, 1e-9f);
      const float dt = warp_sum(s0.f * sv.f);
      const float corr = __fdiv_rn(__fdiv_rn(dt, N0), Nv);
      const float one_m = __fsub_rn(1.f, corr);
      const float l = fabsf(one_m);
      if (!(s0.in && sv.in && l < 0.5f)) continue;          // warp-uniform
      const float sgn = one_m > 0.f ? -1.f : (one_m < 0.f ? 1.f : 0.f);      // d|1-c|/dc
      const float inv = 1.f / (N0 * Nv);
      // d corr / d f0[c], d corr / d fv[c]  (the norm clamps have zero derivative when active)
      const float a0 = sv.f * inv - (n0 > 1e-9f ? corr * s0.f / (n0 * n0) : 0.f);
      const float av = s0.f * inv - (nv > 1e-9f ? corr * sv.f / (nv * nv) : 0.f);
      const float s0x = warp_sum(a0 * s0.fx), s0y = warp_sum(a0 * s0.fy);
      const float svx = warp_sum(av * sv.fx), svy = warp_sum(av * sv.fy);
#pragma unroll
      for (int k = 0; k < 3; ++k)
        g[k] += sgn * (s0.J[0][k] * s0x + s0.J[1][k] * s0y + sv.J[0][k] * svx + sv.J[1][k] * svy);
    }
    if (lane < 3) a.grad[3 * (size_t)p + lane] = scale * (lane == 0 ? g[0] : (lane == 1 ? g[1] : g[2]));
  }
}

__global__ void feat_counts_kernel(const int* __restrict__ offsets, int B, int V, double* __restrict__ partial) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) partial[2 * b + 1] = (double)(offsets[b + 1] - offsets[b]) * (double)(V - 1);
}

// loss = mean over images of sum_i / count_i (0 for images without hits)  (loss.py:155-163)
__global__ void feat_finalize_kernel(const double* __restrict__ partial, int B, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float total = 0.f;
    for (int b = 0; b < B; ++b) {
      const double c = partial[2 * b + 1];
      total += c > 0.0 ? (float)(partial[2 * b] / c) : 0.f;
    }
    *out = total / (float)B;
  }
}

// ------------------------------------------------------------------ f2: depth-carving loss (loss.py:37-63, my_utils.py:269-331)
// One thread per eikonal point: project into every MVS depth view (idx_world2cam, idx_cam2img, normalize_for_grid_sample,
// get_in_range), nearest-neighbour depth lookup (grid_sample mode='nearest', zeros padding, align_corners=False), the
// inside / outside vote of carving_t2 and its RunningTopK(k=1) aggregates (running min of the positive gaps, running max of
// the negative ones, +-1e30/v sentinels), then the weights of get_depth_loss.  Writes the per-point L1 target (-dist_r)
// and weight (far * near * in_range) for the backward pass and accumulates sum(weight * |f - target|).
struct DepthArgs {
  const float* pts;       // [E, stride] normalised object frame (eikonal_points_hom rows)
  int pts_stride;
  const float* f;         // [E] eikonal_output
  const float* depths;    // [V,h,w]
  const float* cams;      // [V,2,4,4]
  const float* size;
  const float* center;
  int E, V, h, w;
  float out_thresh_perc, far_thresh, far_att, near_thresh, near_att;
  float* target;          // [E]
  float* weight;          // [E]
  double* partial;        // [2] = (sum, E)
};

__global__ void depth_carve_kernel(DepthArgs a) {
  __shared__ double sh[kBlk / 32];
  const float size = a.size[0];
  const float cx = a.center[0], cy = a.center[1], cz = a.center[2];
  const float big = 1e30f / (float)a.V;
  double acc = 0.0;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < a.E; e += gridDim.x * blockDim.x) {
    const float* p = a.pts + (size_t)e * a.pts_stride;
    const float X = __fadd_rn(__fmul_rn(__fmul_rn(p[0], 0.5f), size), cx);
    const float Y = __fadd_rn(__fmul_rn(__fmul_rn(p[1], 0.5f), size), cy);
    const float Z = __fadd_rn(__fmul_rn(__fmul_rn(p[2], 0.5f), size), cz);
    int n_valid = 0, n_inside = 0;
    float pos = big, neg = -big;
    for (int v = 0; v < a.V; ++v) {
      const float* cam = a.cams + (size_t)v * 32;
      float c0 = dot4(cam, X, Y, Z, 1.f), c1 = dot4(cam + 4, X, Y, Z, 1.f), c2 = dot4(cam + 8, X, Y, Z, 1.f),
            c3 = dot4(cam + 12, X, Y, Z, 1.f);
      const float d0 = __fadd_rn(c3, 1e-9f);
      c0 = __fdiv_rn(c0, d0);
      c1 = __fdiv_rn(c1, d0);
      c2 = __fdiv_rn(c2, d0);
      c3 = __fdiv_rn(c3, d0);
      const float point_depth = c2;
      const float d1 = __fadd_rn(c3, 1e-9f);
      const float x = __fdiv_rn(c0, d1), y = __fdiv_rn(c1, d1), z = __fdiv_rn(c2, d1);
      const float* K = cam + 16;
      float u = dot3(K, x, y, z), vv = dot3(K + 4, x, y, z), ww = dot3(K + 8, x, y, z);
      const float d2 = __fadd_rn(ww, 1e-9f);
      u = __fdiv_rn(u, d2);
      vv = __fdiv_rn(vv, d2);
      float gx = __fsub_rn(__fmul_rn(__fdiv_rn(u, (float)a.w), 2.f), 1.f);
      float gy = __fsub_rn(__fmul_rn(__fdiv_rn(vv, (float)a.h), 2.f), 1.f);
      gx = fminf(fmaxf(gx, -1.1f), 1.1f);
      gy = fminf(fmaxf(gy, -1.1f), 1.1f);
      const bool in = gx <= 1.f && gx >= -1.f && gy <= 1.f && gy >= -1.f;
      // grid_sampler_2d, nearest: unnormalise with align_corners=False, round half to even
      const float fx = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)a.w), 1.f), 0.5f);
      const float fy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)a.h), 1.f), 0.5f);
      const int ix = (int)nearbyintf(fx), iy = (int)nearbyintf(fy);
      float gd = 0.f;
      if (ix >= 0 && ix < a.w && iy >= 0 && iy < a.h) gd = __ldg(a.depths + ((size_t)v * a.h + iy) * a.w + ix);
      const bool valid = gd > 0.f && in;
      const bool inside = valid && point_depth > __fmul_rn(gd, 0.99f);
      const bool outside = valid && !inside;
      const float dist = valid ? __fsub_rn(point_depth, gd) : 0.f;
      n_valid += valid ? 1 : 0;
      n_inside += inside ? 1 : 0;
      pos = fminf(pos, inside ? dist : big);
      neg = fmaxf(neg, outside ? dist : -big);
    }
    if (!(fabsf(pos) < big * 0.99f)) pos = big;
    if (!(fabsf(neg) < big * 0.99f)) neg = -big;
    const float outside_perc = __fdiv_rn((float)(n_valid - n_inside), __fadd_rn((float)n_valid, 1e-9f));
    const bool scene_valid = n_valid > 0;
    const bool scene_outside = scene_valid && outside_perc > a.out_thresh_perc;
    const bool scene_inside = scene_valid && !scene_outside;
    const float ave = __fadd_rn(scene_inside ? pos : 0.f, scene_outside ? neg : 0.f);
    float dist_r = __fadd_rn(__fmul_rn(__fdiv_rn(ave, size), 2.f), scene_valid ? 0.f : -1.25f);
    dist_r = fminf(fmaxf(dist_r, -1.25f), 1.25f);
    const float far_w = fabsf(dist_r) > a.far_thresh ? a.far_att : 1.f;
    const float near_w = fabsf(dist_r) < a.near_thresh ? a.near_att : 1.f;
    const float wgt = scene_valid ? __fmul_rn(far_w, near_w) : 0.f;
    const float tgt = -dist_r;
    a.target[e] = tgt;
    a.weight[e] = wgt;
    acc += (double)__fmul_rn(fabsf(__fsub_rn(a.f[e], tgt)), wgt);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kBlk / 32; ++w) t += sh[w];
    if (t != 0.0) atomicAdd(a.partial, t);
  }
}

// ------------------------------------------------------------------ a16: rgb L1 over hit pixels / all pixels (loss.py:21-28)
__global__ void rgb_l1_kernel(const float* __restrict__ rgb, const float* __restrict__ gt, const uint8_t* __restrict__ mask,
                              long long R, double* __restrict__ partial) {
  __shared__ double sh[kBlk / 32];
  double acc = 0.0;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (long long)gridDim.x * blockDim.x) {
    if (mask[r]) {
#pragma unroll
      for (int k = 0; k < 3; ++k) acc += (double)fabsf(__fsub_rn(rgb[3 * r + k], gt[3 * r + k]));
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kBlk / 32; ++w) t += sh[w];
    if (t != 0.0) atomicAdd(partial, t);
  }
}

__global__ void rgb_finalize_kernel(const double* __restrict__ partial, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = partial[1] > 0.0 ? (float)(partial[0] / partial[1]) : 0.f;
}
__global__ void set_double_kernel(double* p, double v) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *p = v;
}

struct ShadeLayout {
  size_t off_blocks, off_chunks, off_view, off_full, off_rgb, off_miss_pts, off_miss_idx, off_miss_val, off_miss_cnt, total;
  int n_blocks, n_chunks;
};
static ShadeLayout shade_layout(int64_t R, int feat) {
  ShadeLayout l{};
  size_t off = 0;
  auto take = [&](size_t b) {
    size_t o = off;
    off += (b + 255) / 256 * 256;
    return o;
  };
  l.n_blocks = (int)((R + kScanItems - 1) / kScanItems);
  l.n_chunks = (int)((R + kHitChunk - 1) / kHitChunk);
  l.off_blocks = take((size_t)l.n_blocks * 4);
  l.off_chunks = take((size_t)l.n_chunks * 4);
  l.off_view = take((size_t)R * 12);
  l.off_full = take((size_t)std::min<int64_t>(R, kHitChunk) * (feat + 2) * 4);
  l.off_rgb = take((size_t)R * 12);
  l.off_miss_pts = take((size_t)R * 12);      // sdf_output of the rays that are not surface rays: compacted request list
  l.off_miss_idx = take((size_t)R * 4);
  l.off_miss_val = take((size_t)R * 4);
  l.off_miss_cnt = take(4);
  l.total = off;
  return l;
}

}  // namespace mvsdf

using namespace mvsdf;

// ------------------------------------------------------------------ depth-surface points (phase 0 of training)
// One thread per depth pixel: get_pixel_grids -> idx_img2cam -> idx_cam2world (utils/my_utils.py:71-95) and the
// normalisation (x - center) / size * 2 of implicit_differentiable_renderer.py:236-239.  The two matrix inverses are
// supplied by the caller (torch.inverse of 3x3 / 4x4 blocks, like the reference); products are summed in index order
// without FMA contraction, like the separate torch ops.
__global__ void depth_backproject_kernel(const float* __restrict__ depths, const float* __restrict__ k_inv,
                                         const float* __restrict__ e_inv, int n_maps, int h, int w,
                                         const float* __restrict__ center, const float* __restrict__ size,
                                         float* __restrict__ pts, uint8_t* __restrict__ valid) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n_maps * h * w;
  if (idx >= total) return;
  const int x = (int)(idx % w);
  const int y = (int)((idx / w) % h);
  const int n = (int)(idx / ((long long)w * h));
  const float d = depths[idx];
  const float px = (float)x + 0.5f, py = (float)y + 0.5f;
  const float* Ki = k_inv + n * 9;
  float c[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    c[r] = __fadd_rn(__fadd_rn(__fmul_rn(Ki[r * 3 + 0], px), __fmul_rn(Ki[r * 3 + 1], py)), Ki[r * 3 + 2]);
  const float cz = __fadd_rn(c[2], 1e-9f);
  float ch[4];
#pragma unroll
  for (int r = 0; r < 3; ++r) ch[r] = __fmul_rn(__fdiv_rn(c[r], cz), d);
  ch[3] = 1.0f;
  const float* Ei = e_inv + n * 16;
  float wv[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float acc = __fmul_rn(Ei[r * 4 + 0], ch[0]);
    acc = __fadd_rn(acc, __fmul_rn(Ei[r * 4 + 1], ch[1]));
    acc = __fadd_rn(acc, __fmul_rn(Ei[r * 4 + 2], ch[2]));
    acc = __fadd_rn(acc, __fmul_rn(Ei[r * 4 + 3], ch[3]));
    wv[r] = acc;
  }
  const float ww = __fadd_rn(wv[3], 1e-9f);
  const float sz = size[0];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float p = __fdiv_rn(wv[r], ww);
    pts[idx * 3 + r] = __fmul_rn(__fdiv_rn(__fsub_rn(p, center[r]), sz), 2.0f);
  }
  valid[idx] = d > 0.0f ? 1 : 0;
}

extern "C" {

int mvsdf_depth_backproject(const float* depths, const float* k_inv, const float* e_inv, int n_maps, int h, int w,
                            const float* center, const float* size, float* out_pts, uint8_t* out_valid, void* stream) {
  if (sm_count() <= 0) return fail(MVSDF_ERR_CUDA, "no CUDA device (the product path has no CPU fallback)");
  if (n_maps < 0 || h <= 0 || w <= 0) return fail(MVSDF_ERR_INVALID, "bad depth-map shape");
  const long long total = (long long)n_maps * h * w;
  if (total == 0) return MVSDF_OK;
  if (!depths || !k_inv || !e_inv || !center || !size || !out_pts || !out_valid) return fail(MVSDF_ERR_INVALID, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  note_launch();
  depth_backproject_kernel<<<(unsigned)((total + kBlk - 1) / kBlk), kBlk, 0, st>>>(depths, k_inv, e_inv, n_maps, h, w, center,
                                                                               size, out_pts, out_valid);
  return check_cuda(cudaGetLastError(), "launch depth_backproject_kernel");
}

size_t mvsdf_shade_workspace_bytes(int64_t n_rays, int feature_size) { return shade_layout(n_rays, feature_size).total + 256; }

int mvsdf_shade_rays(const mvsdf_net* sdf_net, const void* sdf_packed, const mvsdf_net* render_net,
                     const void* render_packed, const float* ray_dirs, const float* points, const uint8_t* surface_mask,
                     int n_images, int n_pixels, int feature_size, size_t workspace_bytes, void* workspace,
                     float* out_sdf, float* out_rgb_values, float* out_surf_pts, float* out_normals, float* out_surf_head,
                     int32_t* out_hit_index, int32_t* out_hit_offsets, void* stream) {
  if (!sdf_net || !sdf_packed || !render_net || !render_packed || !ray_dirs || !points || !surface_mask || !workspace ||
      !out_rgb_values || !out_surf_pts || !out_normals || !out_hit_index || !out_hit_offsets)
    return fail(MVSDF_ERR_INVALID, "mvsdf_shade_rays: null argument");
  const int64_t R = (int64_t)n_images * n_pixels;
  if (R <= 0 || R > (1ll << 30)) return fail(MVSDF_ERR_INVALID, "mvsdf_shade_rays: bad ray count");
  if (workspace_bytes < mvsdf_shade_workspace_bytes(R, feature_size))
    return fail(MVSDF_ERR_WORKSPACE, "mvsdf_shade_rays: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const ShadeLayout l = shade_layout(R, feature_size);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int* block_counts = reinterpret_cast<int*>(ws + l.off_blocks);
  int* chunk_counts = reinterpret_cast<int*>(ws + l.off_chunks);
  float* view_hit = reinterpret_cast<float*>(ws + l.off_view);
  float* full = reinterpret_cast<float*>(ws + l.off_full);
  float* rgb_hit = reinterpret_cast<float*>(ws + l.off_rgb);
  int rc;
  // sdf_output = implicit_network(points)[:, :1] for every ray (:202-203); only column 0 is ever read.  For a surface ray it
  // is the value column of the value + normal + feature pass below (same point, same network: the reference evaluates it
  // twice, :202 and :325), so the SDF-only launch covers the other rays only (43 % of them at cfg2).
  int* miss_idx = reinterpret_cast<int*>(ws + l.off_miss_idx);
  float* miss_pts = reinterpret_cast<float*>(ws + l.off_miss_pts);
  float* miss_val = reinterpret_cast<float*>(ws + l.off_miss_val);
  int* miss_cnt = reinterpret_cast<int*>(ws + l.off_miss_cnt);
  note_launch(); hit_count_kernel<<<l.n_blocks, kBlk, 0, st>>>(surface_mask, (int)R, block_counts);
  note_launch(); hit_scan_kernel<<<1, 1024, 0, st>>>(block_counts, l.n_blocks, n_images, n_pixels, surface_mask, out_hit_offsets,
                                      chunk_counts, l.n_chunks, miss_cnt);
  note_launch(); hit_scatter_kernel<<<l.n_blocks, kBlk, 0, st>>>(surface_mask, (int)R, n_pixels, block_counts, points, ray_dirs,
                                                  out_hit_index, out_surf_pts, view_hit, out_sdf ? miss_idx : nullptr, miss_pts);
  if (out_sdf) {
    if ((rc = mlp_sdf(sdf_net, sdf_packed, miss_pts, 0, miss_cnt, MVSDF_HEAD_SDF_ONLY, miss_val, nullptr, nullptr, false, st)))
      return rc;
    note_launch(); miss_scatter_kernel<<<sm_count() * 4, kBlk, 0, st>>>(miss_cnt, miss_idx, miss_val, out_sdf);
  }
  note_launch(); fill_ones_kernel<<<(int)((3 * R + kBlk - 1) / kBlk), kBlk, 0, st>>>(out_rgb_values, 3 * R);
  const int stride = feature_size + 2;
  for (int c = 0; c < l.n_chunks; ++c) {
    const size_t begin = (size_t)c * kHitChunk;
    // get_rbg_value (:324-338): features + un-normalised normals at the surface points, then the light-field MLP
    if ((rc = mlp_sdf(sdf_net, sdf_packed, out_surf_pts + 3 * begin, 0, chunk_counts + c, MVSDF_HEAD_FULL, nullptr, full,
                      out_normals + 3 * begin, true, st)))
      return rc;
    if ((rc = mlp_render(render_net, render_packed, out_surf_pts + 3 * begin, view_hit + 3 * begin,
                         out_normals + 3 * begin, full + 2, stride, 0, chunk_counts + c, rgb_hit + 3 * begin, st)))
      return rc;
    const int64_t cap = std::min<int64_t>(R - (int64_t)begin, kHitChunk);
    note_launch(); shade_scatter_kernel<<<(int)((cap + kBlk - 1) / kBlk), kBlk, 0, st>>>(chunk_counts + c, (int)begin, out_hit_index,
                                                                          rgb_hit, full, stride, out_rgb_values,
                                                                          out_surf_head, out_sdf);
  }
  return check_cuda(cudaGetLastError(), "mvsdf_shade_rays launches");
}

int mvsdf_feat_nchw_to_nhwc(const float* src, int n, int channels, int h, int w, float* dst, void* stream) {
  if (!src || !dst || n <= 0 || channels <= 0 || channels > 32 || h <= 0 || w <= 0)
    return fail(MVSDF_ERR_INVALID, "mvsdf_feat_nchw_to_nhwc: bad argument (channels must be <= 32)");
  const int hw = h * w;
  dim3 grid((hw + 31) / 32, n);
  note_launch(); nchw_to_nhwc_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, channels, hw, dst);
  return check_cuda(cudaGetLastError(), "nchw_to_nhwc launch");
}

int mvsdf_feat_loss_partials(const float* surf_pts, const int32_t* hit_offsets, const float* cams, const float* maps_nhwc,
                             int n_images, int n_views, int h, int w, int channels, const float* size,
                             const float* center, double* partials, void* stream) {
  return mvsdf_feat_loss_partials_indexed(surf_pts, hit_offsets, cams, maps_nhwc, nullptr, n_images, n_views, h, w, channels, size,
                                          center, partials, stream);
}

int mvsdf_feat_loss_partials_indexed(const float* surf_pts, const int32_t* hit_offsets, const float* cams,
                                     const float* maps_nhwc, const int32_t* map_index, int n_images, int n_views, int h, int w,
                                     int channels, const float* size, const float* center, double* partials, void* stream) {
  if (!surf_pts || !hit_offsets || !cams || !maps_nhwc || !size || !center || !partials)
    return fail(MVSDF_ERR_INVALID, "mvsdf_feat_loss_partials: null argument");
  if (channels != 32) return fail(MVSDF_ERR_INVALID, "mvsdf_feat_loss_partials: 32 feature channels expected (FeatExt, my_utils.py:697-708)");
  if (n_images <= 0 || n_views < 2 || h <= 0 || w <= 0) return fail(MVSDF_ERR_INVALID, "mvsdf_feat_loss_partials: bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = check_cuda(cudaMemsetAsync(partials, 0, sizeof(double) * 2 * n_images, st), "memset partials");
  if (rc) return rc;
  FeatArgs a{surf_pts, hit_offsets, cams, maps_nhwc, size, center, partials, n_images, n_views, h, w, map_index};
  const int sms = sm_count();
  if (sms <= 0) return fail(MVSDF_ERR_CUDA, "no CUDA device (the product path has no CPU fallback)");
  note_launch(); feat_counts_kernel<<<(n_images + 63) / 64, 64, 0, st>>>(hit_offsets, n_images, n_views, partials);
  note_launch(); feat_loss_kernel<<<sms * 8, 256, 0, st>>>(a);
  return check_cuda(cudaGetLastError(), "feat_loss launch");
}

int mvsdf_feat_loss_backward(const float* surf_pts, const int32_t* hit_offsets, const float* cams, const float* maps_nhwc,
                             int n_images, int n_views, int h, int w, int channels, const float* size, const float* center,
                             const double* partials, const float* upstream_grad, float* out_grad_pts, void* stream) {
  return mvsdf_feat_loss_backward_indexed(surf_pts, hit_offsets, cams, maps_nhwc, nullptr, n_images, n_views, h, w, channels, size,
                                          center, partials, upstream_grad, out_grad_pts, stream);
}

int mvsdf_feat_loss_backward_indexed(const float* surf_pts, const int32_t* hit_offsets, const float* cams,
                                     const float* maps_nhwc, const int32_t* map_index, int n_images, int n_views, int h, int w,
                                     int channels, const float* size, const float* center, const double* partials,
                                     const float* upstream_grad, float* out_grad_pts, void* stream) {
  if (!surf_pts || !hit_offsets || !cams || !maps_nhwc || !size || !center || !partials || !upstream_grad || !out_grad_pts)
    return fail(MVSDF_ERR_INVALID, "mvsdf_feat_loss_backward: null argument");
  if (channels != 32) return fail(MVSDF_ERR_INVALID, "mvsdf_feat_loss_backward: 32 feature channels expected");
  if (n_images <= 0 || n_views < 2 || h <= 0 || w <= 0) return fail(MVSDF_ERR_INVALID, "mvsdf_feat_loss_backward: bad sizes");
  const int sms = sm_count();
  if (sms <= 0) return fail(MVSDF_ERR_CUDA, "no CUDA device (the product path has no CPU fallback)");
  FeatBwdArgs a{surf_pts, hit_offsets, cams, maps_nhwc, size, center, partials, upstream_grad, out_grad_pts,
                n_images, n_views, h, w, map_index};
  note_launch(); feat_loss_bwd_kernel<<<sms * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return check_cuda(cudaGetLastError(), "feat_loss_bwd launch");
}

int mvsdf_feat_loss_finalize(const double* partials, int n_images, float* out_loss, void* stream) {
  if (!partials || !out_loss || n_images <= 0) return fail(MVSDF_ERR_INVALID, "mvsdf_feat_loss_finalize: bad argument");
  note_launch(); feat_finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(partials, n_images, out_loss);
  return check_cuda(cudaGetLastError(), "feat_finalize launch");
}

int mvsdf_depth_loss_partials(const float* eik_points, int point_stride, const float* eik_output, int64_t n_points,
                              const float* depths, const float* depth_cams, int n_views, int h, int w, const float* size,
                              const float* center, float out_thresh_perc, float far_thresh, float far_att, float near_thresh,
                              float near_att, float* out_target, float* out_weight, double* partials, void* stream) {
  if (!eik_points || !eik_output || !depths || !depth_cams || !size || !center || !out_target || !out_weight || !partials ||
      n_points <= 0 || n_views <= 0 || h <= 0 || w <= 0 || point_stride < 3)
    return fail(MVSDF_ERR_INVALID, "mvsdf_depth_loss_partials: bad argument");
  const int sms = sm_count();
  if (sms <= 0) return fail(MVSDF_ERR_CUDA, "no CUDA device (the product path has no CPU fallback)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = check_cuda(cudaMemsetAsync(partials, 0, sizeof(double) * 2, st), "memset partials");
  if (rc) return rc;
  note_launch(); set_double_kernel<<<1, 32, 0, st>>>(partials + 1, (double)n_points);
  DepthArgs a{eik_points, point_stride, eik_output, depths, depth_cams, size, center, (int)n_points, n_views, h, w,
              out_thresh_perc, far_thresh, far_att, near_thresh, near_att, out_target, out_weight, partials};
  const int grid = (int)std::min<int64_t>((n_points + kBlk - 1) / kBlk, (int64_t)sms * 8);
  note_launch(); depth_carve_kernel<<<grid, kBlk, 0, st>>>(a);
  return check_cuda(cudaGetLastError(), "depth_carve launch");
}

int mvsdf_rgb_l1_partials(const float* rgb_values, const float* rgb_gt, const uint8_t* mask, int64_t n_rays,
                          double* partials, void* stream) {
  if (!rgb_values || !rgb_gt || !mask || !partials || n_rays <= 0)
    return fail(MVSDF_ERR_INVALID, "mvsdf_rgb_l1_partials: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = check_cuda(cudaMemsetAsync(partials, 0, sizeof(double) * 2, st), "memset partials");
  if (rc) return rc;
  const int sms = sm_count();
  if (sms <= 0) return fail(MVSDF_ERR_CUDA, "no CUDA device (the product path has no CPU fallback)");
  note_launch(); set_double_kernel<<<1, 32, 0, st>>>(partials + 1, (double)n_rays);
  const int grid = (int)std::min<int64_t>((n_rays + kBlk - 1) / kBlk, (int64_t)sms * 8);
  note_launch(); rgb_l1_kernel<<<grid, kBlk, 0, st>>>(rgb_values, rgb_gt, mask, n_rays, partials);
  return check_cuda(cudaGetLastError(), "rgb_l1 launch");
}

int mvsdf_rgb_l1_finalize(const double* partials, float* out_loss, void* stream) {
  if (!partials || !out_loss) return fail(MVSDF_ERR_INVALID, "mvsdf_rgb_l1_finalize: bad argument");
  note_launch(); rgb_finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(partials, out_loss);
  return check_cuda(cudaGetLastError(), "rgb_finalize launch");
}

}  // extern "C"

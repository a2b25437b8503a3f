// C ABI, part 5: the feature extractor that fills the scene feature store.
//
// FeatExt (code/utils/my_utils.py:693-708; UNet :595-690, BasicBlock :531-576) in eval mode -- a 5x5/2 stem, a three-level
// residual encoder (32 / 64 / 128 channels at 1/2, 1/4, 1/8 resolution), two transposed-convolution decoder levels with
// skip concatenation, and 3x3 heads -- runs once per scene over all images (datasets/scene_dataset.py:138-149).  Here:
//   * activations are channels-last fp32 throughout, so the finest output IS the [image, h, w, 32] layout the feature-warp
//     kernel gathers from (one 128-byte line per bilinear tap); no NCHW map is ever materialised;
//   * BatchNorm (eval) is folded into the convolution weights / bias once, on the device (fe_fold_kernel);
//   * one tiled direct-convolution kernel (shared-memory input patch + weight slab, register accumulators, fused bias +
//     residual + ReLU, optional second input = the skip concatenation) and one gather-form transposed convolution.
// A one-off per scene (about 0.2 TFLOP per 1200x1600 image): fp32 CUDA-core math, bandwidth / FMA bound, not on the
// per-step path -- the point of the native version is the output format and having no cuDNN / PyTorch dependency.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstring>

#include "../../include/mvsdf_b200.h"
#include "internal.h"

namespace mvsdf {

// ---- network description: the order in which mvsdf_featext_pack expects the source tensors ---------------------------
struct FeConv {
  int cin, cout, k, stride, pad;
  int bn;          // followed by BatchNorm (4 more source tensors: weight, bias, running_mean, running_var)
  int transposed;  // ConvTranspose2d(k=3, s=2, p=1, output_padding=1); weight layout [cin, cout, k, k]
};
// index: 0 stem | 1-5 enc0.block0 (conv1, conv2, down) + enc0.block1 (conv1, conv2) | 6-10 enc1 | 11-15 enc2 |
//        16 deconv, 17 post-concat conv, 18-19 block | 20 deconv, 21 post-concat, 22-23 block | 24-26 final_conv_1..3
constexpr int kFeConvs = 27;
static const FeConv kFe[kFeConvs] = {
    {3, 16, 5, 2, 2, 1, 0},
    {16, 32, 3, 1, 1, 1, 0}, {32, 32, 3, 1, 1, 1, 0}, {16, 32, 1, 1, 0, 1, 0}, {32, 32, 3, 1, 1, 1, 0}, {32, 32, 3, 1, 1, 1, 0},
    {32, 64, 3, 2, 1, 1, 0}, {64, 64, 3, 1, 1, 1, 0}, {32, 64, 1, 2, 0, 1, 0}, {64, 64, 3, 1, 1, 1, 0}, {64, 64, 3, 1, 1, 1, 0},
    {64, 128, 3, 2, 1, 1, 0}, {128, 128, 3, 1, 1, 1, 0}, {64, 128, 1, 2, 0, 1, 0}, {128, 128, 3, 1, 1, 1, 0}, {128, 128, 3, 1, 1, 1, 0},
    {128, 64, 3, 2, 1, 0, 1}, {128, 64, 3, 1, 1, 0, 0}, {64, 64, 3, 1, 1, 1, 0}, {64, 64, 3, 1, 1, 1, 0},
    {64, 32, 3, 2, 1, 0, 1}, {64, 32, 3, 1, 1, 0, 0}, {32, 32, 3, 1, 1, 1, 0}, {32, 32, 3, 1, 1, 1, 0},
    {128, 32, 3, 1, 1, 0, 0}, {64, 32, 3, 1, 1, 0, 0}, {32, 32, 3, 1, 1, 0, 0},
};

static size_t fe_w_off(int i) {       // float offset of conv i inside the packed blob: [k*k][cin][cout] weights, then [cout] bias
  size_t o = 0;
  for (int j = 0; j < i; ++j) o += (size_t)kFe[j].k * kFe[j].k * kFe[j].cin * kFe[j].cout + kFe[j].cout;
  return o;
}

// packed[tap][ci][co] = w * gamma / sqrt(var + eps),  bias[co] = beta - mean * gamma / sqrt(var + eps)
__global__ void fe_fold_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps, int cin, int cout, int k,
                               int transposed, float* __restrict__ packed) {
  const int total = k * k * cin * cout;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < cout) {
    float b = 0.f;
    if (gamma) b = beta[idx] - mean[idx] * gamma[idx] / sqrtf(var[idx] + eps);
    packed[total + idx] = b;
  }
  if (idx >= total) return;
  const int co = idx % cout, ci = (idx / cout) % cin, tap = idx / (cout * cin);
  const int ky = tap / k, kx = tap % k;
  const float src = transposed ? w[(((size_t)ci * cout + co) * k + ky) * k + kx] : w[(((size_t)co * cin + ci) * k + ky) * k + kx];
  const float s = gamma ? gamma[co] / sqrtf(var[co] + eps) : 1.0f;
  packed[idx] = src * s;
}

__global__ void fe_nchw_to_nhwc3_kernel(const float* __restrict__ src, long long hw, int n, float* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw * n) return;
  const long long img = i / hw, p = i - img * hw;
#pragma unroll
  for (int c = 0; c < 3; ++c) dst[i * 3 + c] = src[(img * 3 + c) * hw + p];
}

struct FeConvArgs {
  const float* in_a;     // [n, hin, win, ca]
  const float* in_b;     // optional second input (channel concatenation), [n, hin, win, cin - ca]
  const float* w;        // [k*k][cin][cout]
  const float* bias;     // [cout]
  const float* res;      // optional residual [n, hout, wout, cout]
  float* out;            // [n, hout, wout, cout]
  int hin, win, hout, wout, ca, cin, cout, stride, pad, relu;
};

constexpr int kFeTile = 8;        // 8 x 8 output pixels per block
constexpr int kFeCi = 8;          // input channels per shared-memory slab

// 256 threads: thread = (pixel of the tile, quarter of the output channels); CPT = cout / 4 accumulators per thread
template <int K, int CPT>
__global__ void __launch_bounds__(256) fe_conv_kernel(FeConvArgs a) {
  extern __shared__ float fe_smem[];
  const int S = a.stride;
  const int T = (kFeTile - 1) * S + K;                  // input patch edge
  float* s_in = fe_smem;                                // [T*T][kFeCi]
  float* s_w = fe_smem + T * T * kFeCi;                 // [K*K][kFeCi][cout]
  const int tiles_x = (a.wout + kFeTile - 1) / kFeTile;
  const int ty0 = (blockIdx.x / tiles_x) * kFeTile, tx0 = (blockIdx.x % tiles_x) * kFeTile;
  const int n = blockIdx.y;
  const int pix = threadIdx.x & 63, cgp = threadIdx.x >> 6;
  const int py = pix >> 3, px = pix & 7;
  const int cout = a.cout;
  const int cb = a.cin - a.ca;
  float acc[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) acc[j] = 0.f;
  const int iy0 = ty0 * S - a.pad, ix0 = tx0 * S - a.pad;
  for (int c0 = 0; c0 < a.cin; c0 += kFeCi) {
    for (int idx = threadIdx.x; idx < T * T * kFeCi; idx += 256) {
      const int ci = idx % kFeCi, pos = idx / kFeCi;
      const int y = iy0 + pos / T, x = ix0 + pos % T;
      const int c = c0 + ci;
      float v = 0.f;
      if (y >= 0 && y < a.hin && x >= 0 && x < a.win && c < a.cin) {
        const size_t p = ((size_t)n * a.hin + y) * a.win + x;
        v = c < a.ca ? a.in_a[p * a.ca + c] : a.in_b[p * cb + (c - a.ca)];
      }
      s_in[idx] = v;
    }
    for (int idx = threadIdx.x; idx < K * K * kFeCi * cout; idx += 256) {
      const int co = idx % cout, ci = (idx / cout) % kFeCi, tap = idx / (cout * kFeCi);
      const int c = c0 + ci;
      s_w[idx] = c < a.cin ? a.w[((size_t)tap * a.cin + c) * cout + co] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const float* ip = s_in + ((py * S + ky) * T + px * S + kx) * kFeCi;
        const float* wp = s_w + (ky * K + kx) * kFeCi * cout + cgp * CPT;
#pragma unroll
        for (int ci = 0; ci < kFeCi; ++ci) {
          const float v = ip[ci];
#pragma unroll
          for (int j = 0; j < CPT; ++j) acc[j] = fmaf(v, wp[ci * cout + j], acc[j]);
        }
      }
    }
    __syncthreads();
  }
  const int oy = ty0 + py, ox = tx0 + px;
  if (oy < a.hout && ox < a.wout) {
    const size_t o = (((size_t)n * a.hout + oy) * a.wout + ox) * cout + cgp * CPT;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      float v = acc[j] + a.bias[cgp * CPT + j];
      if (a.res) v += a.res[o + j];
      if (a.relu) v = fmaxf(v, 0.f);
      a.out[o + j] = v;
    }
  }
}

// ConvTranspose2d(k=3, s=2, p=1, output_padding=1), gather form: out[oy, ox] collects in[(oy + 1 - ky) / 2, (ox + 1 - kx) / 2]
// for the taps whose parity matches.  One warp-row of threads = the output channels of one pixel (coalesced weight reads,
// broadcast input reads).
__global__ void fe_deconv_kernel(FeConvArgs a) {
  const int co = threadIdx.x % a.cout;
  const int pix_per_block = blockDim.x / a.cout;
  const long long p = (long long)blockIdx.x * pix_per_block + threadIdx.x / a.cout;
  const int n = blockIdx.y;
  if (p >= (long long)a.hout * a.wout) return;
  const int oy = (int)(p / a.wout), ox = (int)(p % a.wout);
  float acc = a.bias[co];
  for (int ky = 0; ky < 3; ++ky) {
    const int ty = oy + 1 - ky;
    if (ty < 0 || (ty & 1)) continue;
    const int iy = ty >> 1;
    if (iy >= a.hin) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int tx = ox + 1 - kx;
      if (tx < 0 || (tx & 1)) continue;
      const int ix = tx >> 1;
      if (ix >= a.win) continue;
      const float* ip = a.in_a + (((size_t)n * a.hin + iy) * a.win + ix) * a.cin;
      const float* wp = a.w + (size_t)(ky * 3 + kx) * a.cin * a.cout + co;
      for (int ci = 0; ci < a.cin; ++ci) acc = fmaf(ip[ci], wp[(size_t)ci * a.cout], acc);
    }
  }
  a.out[(((size_t)n * a.hout + oy) * a.wout + ox) * a.cout + co] = acc;
}

template <int K, int CPT>
static int launch_conv_t(const FeConvArgs& a, int n, cudaStream_t st) {
  const int T = (kFeTile - 1) * a.stride + K;
  const size_t smem = ((size_t)T * T * kFeCi + (size_t)K * K * kFeCi * a.cout) * sizeof(float);
  auto kern = fe_conv_kernel<K, CPT>;
  if (smem > 48 * 1024) {
    int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(fe_conv)");
    if (rc) return rc;
  }
  const int tiles = ((a.wout + kFeTile - 1) / kFeTile) * ((a.hout + kFeTile - 1) / kFeTile);
  note_launch();
  kern<<<dim3(tiles, n), 256, smem, st>>>(a);
  return check_cuda(cudaGetLastError(), "launch fe_conv_kernel");
}

static int launch_conv(const FeConvArgs& a, int k, int n, cudaStream_t st) {
  const int cpt = a.cout / 4;
  if (k == 5 && cpt == 4) return launch_conv_t<5, 4>(a, n, st);
  if (k == 3 && cpt == 8) return launch_conv_t<3, 8>(a, n, st);
  if (k == 3 && cpt == 16) return launch_conv_t<3, 16>(a, n, st);
  if (k == 3 && cpt == 32) return launch_conv_t<3, 32>(a, n, st);
  if (k == 1 && cpt == 8) return launch_conv_t<1, 8>(a, n, st);
  if (k == 1 && cpt == 16) return launch_conv_t<1, 16>(a, n, st);
  if (k == 1 && cpt == 32) return launch_conv_t<1, 32>(a, n, st);
  return fail(MVSDF_ERR_INVALID, "featext: no convolution instantiation for k=%d cout=%d", k, a.cout);
}

}  // namespace mvsdf

using namespace mvsdf;

extern "C" {

int mvsdf_featext_num_convs(void) { return kFeConvs; }

size_t mvsdf_featext_packed_floats(void) { return fe_w_off(kFeConvs); }

int mvsdf_featext_pack(const float* const* conv_weight_host, const float* const* bn_weight_host, const float* const* bn_bias_host,
                       const float* const* bn_mean_host, const float* const* bn_var_host, float bn_eps, float* packed, void* stream) {
  if (!conv_weight_host || !bn_weight_host || !bn_bias_host || !bn_mean_host || !bn_var_host || !packed)
    return fail(MVSDF_ERR_INVALID, "mvsdf_featext_pack: null argument");
  if (sm_count() <= 0) return fail(MVSDF_ERR_CUDA, "no CUDA device (the product path has no CPU fallback)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < kFeConvs; ++i) {
    const FeConv& c = kFe[i];
    if (!conv_weight_host[i] || (c.bn && (!bn_weight_host[i] || !bn_bias_host[i] || !bn_mean_host[i] || !bn_var_host[i])))
      return fail(MVSDF_ERR_INVALID, "mvsdf_featext_pack: missing tensor for convolution %d", i);
    const int total = c.k * c.k * c.cin * c.cout;
    note_launch();
    fe_fold_kernel<<<(std::max(total, c.cout) + 255) / 256, 256, 0, st>>>(conv_weight_host[i], c.bn ? bn_weight_host[i] : nullptr,
                                                                          c.bn ? bn_bias_host[i] : nullptr, c.bn ? bn_mean_host[i] : nullptr,
                                                                          c.bn ? bn_var_host[i] : nullptr, bn_eps, c.cin, c.cout, c.k,
                                                                          c.transposed, packed + fe_w_off(i));
  }
  return check_cuda(cudaGetLastError(), "featext pack launch");
}

// activations (floats per image): image nhwc 3, stem 16, enc0 32 x3 scratch, enc1 64 x3, enc2 128 x3, dec ...
static size_t fe_ws_floats(int n, int H, int W) {
  const size_t h2 = H / 2, w2 = W / 2, h4 = (h2 + 1) / 2, w4 = (w2 + 1) / 2, h8 = (h4 + 1) / 2, w8 = (w4 + 1) / 2;
  size_t per = (size_t)H * W * 3 + h2 * w2 * 16 + 4 * h2 * w2 * 32 + 4 * h4 * w4 * 64 + 4 * h8 * w8 * 128;
  return per * n + 1024;
}

size_t mvsdf_featext_workspace_bytes(int n_images, int height, int width) {
  if (n_images <= 0 || height <= 0 || width <= 0) return 0;
  return fe_ws_floats(n_images, height, width) * sizeof(float);
}

int mvsdf_featext_forward(const float* packed, const float* images_nchw, int n_images, int height, int width, size_t workspace_bytes,
                          void* workspace, float* out_eighth_nhwc, float* out_quarter_nhwc, float* out_half_nhwc, void* stream) {
  if (!packed || !images_nchw || !workspace || !out_half_nhwc) return fail(MVSDF_ERR_INVALID, "mvsdf_featext_forward: null argument");
  if (n_images <= 0 || height < 16 || width < 16 || (height % 8) || (width % 8))
    return fail(MVSDF_ERR_INVALID, "mvsdf_featext_forward: image sides must be multiples of 8 (the decoder doubles the 1/8 map twice)");
  if (workspace_bytes < mvsdf_featext_workspace_bytes(n_images, height, width))
    return fail(MVSDF_ERR_WORKSPACE, "mvsdf_featext_forward: workspace too small");
  if (sm_count() <= 0) return fail(MVSDF_ERR_CUDA, "no CUDA device (the product path has no CPU fallback)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = n_images, H = height, W = width;
  const int h2 = H / 2, w2 = W / 2, h4 = h2 / 2, w4 = w2 / 2, h8 = h4 / 2, w8 = w4 / 2;
  float* ws = static_cast<float*>(workspace);
  auto take = [&](size_t floats) {
    float* p = ws;
    ws += floats;
    return p;
  };
  float* img = take((size_t)n * H * W * 3);
  float* stem = take((size_t)n * h2 * w2 * 16);
  float* a2[4];
  float* a4[4];
  float* a8[4];
  for (int i = 0; i < 4; ++i) a2[i] = take((size_t)n * h2 * w2 * 32);
  for (int i = 0; i < 4; ++i) a4[i] = take((size_t)n * h4 * w4 * 64);
  for (int i = 0; i < 4; ++i) a8[i] = take((size_t)n * h8 * w8 * 128);
  int rc;
  note_launch();
  fe_nchw_to_nhwc3_kernel<<<(unsigned)(((long long)n * H * W + 255) / 256), 256, 0, st>>>(images_nchw, (long long)H * W, n, img);
  auto conv = [&](int i, const float* in_a, int ca, const float* in_b, int hin, int win, const float* res, int relu, float* out) {
    const FeConv& c = kFe[i];
    FeConvArgs a{};
    a.in_a = in_a;
    a.in_b = in_b;
    a.w = packed + fe_w_off(i);
    a.bias = a.w + (size_t)c.k * c.k * c.cin * c.cout;
    a.res = res;
    a.out = out;
    a.hin = hin;
    a.win = win;
    a.ca = ca;
    a.cin = c.cin;
    a.cout = c.cout;
    a.stride = c.stride;
    a.pad = c.pad;
    a.relu = relu;
    if (c.transposed) {
      a.hout = hin * 2;
      a.wout = win * 2;
      const int ppb = 256 / c.cout;
      note_launch();
      fe_deconv_kernel<<<dim3((unsigned)(((long long)a.hout * a.wout + ppb - 1) / ppb), n), ppb * c.cout, 0, st>>>(a);
      return check_cuda(cudaGetLastError(), "launch fe_deconv_kernel");
    }
    a.hout = (hin + 2 * c.pad - c.k) / c.stride + 1;
    a.wout = (win + 2 * c.pad - c.k) / c.stride + 1;
    return launch_conv(a, c.k, n, st);
  };
  // BasicBlock (my_utils.py:558-576): relu(bn2(conv2(relu(bn1(conv1(x))))) + residual); residual = x or bn(conv1x1(x))
  auto block = [&](int i_conv1, int i_conv2, int i_down, const float* x, int cx, int hin, int win, float* t1, float* t2, float* out,
                   int hout, int wout) {
    int e = conv(i_conv1, x, cx, nullptr, hin, win, nullptr, 1, t1);
    if (e) return e;
    const float* res = x;
    if (i_down >= 0) {
      if ((e = conv(i_down, x, cx, nullptr, hin, win, nullptr, 0, t2))) return e;
      res = t2;
    }
    return conv(i_conv2, t1, kFe[i_conv1].cout, nullptr, hout, wout, res, 1, out);
  };
  // stem
  if ((rc = conv(0, img, 3, nullptr, H, W, nullptr, 1, stem))) return rc;
  // encoder
  if ((rc = block(1, 2, 3, stem, 16, h2, w2, a2[0], a2[1], a2[2], h2, w2))) return rc;
  if ((rc = block(4, 5, -1, a2[2], 32, h2, w2, a2[0], nullptr, a2[3], h2, w2))) return rc;          // e0 = a2[3]
  if ((rc = block(6, 7, 8, a2[3], 32, h2, w2, a4[0], a4[1], a4[2], h4, w4))) return rc;
  if ((rc = block(9, 10, -1, a4[2], 64, h4, w4, a4[0], nullptr, a4[3], h4, w4))) return rc;          // e1 = a4[3]
  if ((rc = block(11, 12, 13, a4[3], 64, h4, w4, a8[0], a8[1], a8[2], h8, w8))) return rc;
  if ((rc = block(14, 15, -1, a8[2], 128, h8, w8, a8[0], nullptr, a8[3], h8, w8))) return rc;        // e2 = a8[3]
  if (out_eighth_nhwc && (rc = conv(24, a8[3], 128, nullptr, h8, w8, nullptr, 0, out_eighth_nhwc))) return rc;
  // decoder level 1/4: deconv(e2) ++ e1 -> conv -> block
  if ((rc = conv(16, a8[3], 128, nullptr, h8, w8, nullptr, 0, a4[0]))) return rc;
  if ((rc = conv(17, a4[0], 64, a4[3], h4, w4, nullptr, 0, a4[1]))) return rc;
  if ((rc = block(18, 19, -1, a4[1], 64, h4, w4, a4[0], nullptr, a4[2], h4, w4))) return rc;         // d1 = a4[2]
  if (out_quarter_nhwc && (rc = conv(25, a4[2], 64, nullptr, h4, w4, nullptr, 0, out_quarter_nhwc))) return rc;
  // decoder level 1/2: deconv(d1) ++ e0 -> conv -> block
  if ((rc = conv(20, a4[2], 64, nullptr, h4, w4, nullptr, 0, a2[0]))) return rc;
  if ((rc = conv(21, a2[0], 32, a2[3], h2, w2, nullptr, 0, a2[1]))) return rc;
  if ((rc = block(22, 23, -1, a2[1], 32, h2, w2, a2[0], nullptr, a2[2], h2, w2))) return rc;         // d2 = a2[2]
  return conv(26, a2[2], 32, nullptr, h2, w2, nullptr, 0, out_half_nhwc);
}

}  // extern "C"

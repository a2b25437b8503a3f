// C ABI, part 4: the training step's native backward -- transposed weight packer, saving
// forward, reverse sweep + dW GEMM launches, weight-norm chain, fused Adam with gradient-norm clipping.
// Replaces what the reference gets from eager autograd: loss.backward() through ImplicitNetwork / RenderingNetwork
// (code/training/idr_train.py:287, model/implicit_differentiable_renderer.py:96-107 create_graph=True),
// torch.nn.utils.clip_grad_norm_ (:292) and torch.optim.Adam.step (:300).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "../../include/mvsdf_b200.h"
#include "internal.h"
#include "mlp_bwd_kernel.cuh"
#include "mlp_bwd_pair_kernel.cuh"

namespace mvsdf {

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

struct TrainPlan {
  int n_run;
  LayerPlan fwd[kMaxLayers];
  LayerPlan Lt[kMaxLayers];
  int fwd_layer[kMaxLayers];
  int save_kc[kMaxLayers];
  int dz_kc[kMaxLayers];
  int db_off[kMaxLayers];
  int db_total;
  long long dw_off[kMaxLayers];
  long long dw_total;
  int k_cores_max;
  long long packed_t_bytes;
  long long t_scale_off;
};

static void make_train_plan(const NetPlan& p, TrainPlan& t) {
  memset(&t, 0, sizeof(t));
  t.n_run = p.n_hidden + 1;
  for (int l = 0; l < p.n_hidden; ++l) t.fwd[l] = p.L[l];
  t.fwd[p.n_hidden] = p.L[p.head_index[HEAD_FULL]];
  int db = 0;
  long long dw = 0;
  for (int l = 0; l < t.n_run; ++l) {
    t.save_kc[l] = t.fwd[l].k_chunks * (kChunkK / 8);
    t.dz_kc[l] = t.fwd[l].m_tiles * (kTileM / 8);
    t.db_off[l] = db;
    db += t.fwd[l].m_tiles * kTileM;
    t.dw_off[l] = dw;
    dw += (long long)t.fwd[l].m_tiles * kTileM * t.save_kc[l] * 8;
  }
  t.db_total = db;
  t.dw_total = dw;
  long long off = 0;
  t.k_cores_max = t.dz_kc[t.n_run - 1];
  for (int i = 0; i < t.n_run; ++i) {
    const int fl = t.n_run - 1 - i;
    LayerPlan& L = t.Lt[i];
    L = t.fwd[fl];
    L.in_dim = t.fwd[fl].out_dim;                       // K of the transposed product: the forward layer's output rows
    L.out_dim = t.fwd[fl].in_dim;                       // rows of D: the forward layer's input features
    L.k_chunks = cdiv(L.in_dim, kChunkK);
    L.m_tiles = cdiv(t.save_kc[fl] * 8, kTileM);
    L.w_off = off;
    off += (long long)L.m_tiles * L.k_chunks * kStageBytes;
    t.fwd_layer[i] = fl;
    t.k_cores_max = std::max(t.k_cores_max, std::max(L.k_chunks * (kChunkK / 8), L.m_tiles * (kTileM / 8)));
  }
  off = (off + 255) / 256 * 256;
  t.t_scale_off = off;
  int sc = 0;
  for (int l = 0; l < t.n_run; ++l) sc += t.fwd[l].out_dim;
  off += (long long)sc * 4;
  t.packed_t_bytes = (off + 255) / 256 * 256;
}

// g / ||v||_2 per source row (torch._weight_norm, dim = 0)
__global__ void t_row_scale_kernel(const float* __restrict__ v, const float* __restrict__ g, int rows, int cols,
                                   float* __restrict__ scale) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) {
    const float x = v[(size_t)row * cols + c];
    s = fmaf(x, x, s);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) scale[row] = g ? g[row] / sqrtf(s) : 1.0f;
}

// source row of head-ordered row k (the inverse of pack_layer_kernel's row map)
__host__ __device__ inline int src_row_of(int row_map, int k, int out_rows, int feat) {
  if (row_map == 0) return k < out_rows ? k : -1;
  if (row_map == 1) return k == 0 ? 0 : -1;
  return k < feat ? k + 2 : (k < feat + 2 ? k - feat : -1);
}

// transposed tiles: dest row = input feature i of the forward layer, K index = (head-ordered) output row
__global__ void pack_layer_t_kernel(const float* __restrict__ v, const float* __restrict__ scale, LayerPlan lt, int src_in_dim,
                                    int src_out_dim, int feat, uint8_t* __restrict__ packed_t) {
  const int cores = lt.k_chunks * (kChunkK / 8);
  const int rows = lt.m_tiles * kTileM;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cores) return;
  const int dr = idx / cores, kcore = idx - dr * cores;
  __align__(16) __half hi[8];
  __align__(16) __half lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = kcore * 8 + e;
    const int sr = src_row_of(lt.row_map, k, src_out_dim, feat);
    float w = 0.f;
    if (sr >= 0 && dr < src_in_dim) w = v[(size_t)sr * src_in_dim + dr] * scale[sr] * lt.col_scale * kWeightScale;
    hi[e] = __float2half_rn(w);
    lo[e] = __float2half_rn(w - __half2float(hi[e]));
  }
  const int m = dr / kTileM, r = dr - m * kTileM;
  const int kc = kcore / (kChunkK / 8), kin = kcore - kc * (kChunkK / 8);
  uint8_t* tile = packed_t + lt.w_off + (size_t)(m * lt.k_chunks + kc) * kStageBytes;
  const int off = (r >> 3) * 512 + kin * 128 + (r & 7) * 16;
  *reinterpret_cast<uint4*>(tile + off) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(tile + kTileBytes + off) = *reinterpret_cast<const uint4*>(lo);
}

// ---- gradient scale: S = 2^-ceil(log2(max |upstream|)), so that S * max lies in (1/2, 1]
__global__ void absmax_kernel(const float* __restrict__ a, long long n, unsigned* __restrict__ out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = fabsf(a[i]);
    m = (v > m || !(v == v)) ? v : m;          // NaN wins so that it is seen
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float other = __shfl_xor_sync(0xffffffffu, m, o);
    m = (other > m || !(other == other)) ? other : m;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));      // non-negative floats order like their bit patterns
}
__global__ void gscale_kernel(const unsigned* __restrict__ maxbits, float* __restrict__ gscale, int* __restrict__ status) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float m = __uint_as_float(*maxbits);
  float S = 1.0f;
  if (!(m <= 3.0e38f)) {
    if (status) atomicAdd(status + kStatusNonFinite, 1);      // non-finite upstream gradient
  } else if (m > 0.f) {
    int e;
    frexpf(m, &e);                 // m = f * 2^e, f in [0.5, 1)
    S = ldexpf(1.0f, -e);
  }
  gscale[0] = S;
  gscale[1] = 1.0f / S;
}

// bias gradient of the head: column sums of the upstream gradient (head row order); tanh' for the rendering net
__global__ void head_db_kernel(const float* __restrict__ g, const float* __restrict__ rgb, long long n, int stride, int row_map,
                               int out_rows, int feat, float* __restrict__ db) {
  __shared__ float sh[8];
  const int k = blockIdx.x;                      // head-ordered row
  const int src = src_row_of(row_map, k, out_rows, feat);
  if (src < 0) return;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    float v = g[i * stride + src];
    if (rgb) {
      const float y = rgb[i * stride + src];
      v *= 1.0f - y * y;
    }
    acc += v;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    db[k] = t;
  }
}

// weight-norm chain (W = g v / ||v||, :70-71): one warp per source row
//   dg = <dW, v^>,   dv = g / ||v|| (dW - <dW, v^> v^),   plain Linear (g == nullptr): dv = dW
__global__ void weight_grads_kernel(const float* __restrict__ dw, const float* __restrict__ db, int dw_stride, int row_map, int feat,
                                    float col_scale, const float* __restrict__ v, const float* __restrict__ g, int rows, int cols,
                                    float* __restrict__ dv, float* __restrict__ dg, float* __restrict__ dbias, int* __restrict__ status) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  // plan row of source row r
  const int dr = row_map == 0 ? r : (r < 2 ? feat + r : r - 2);
  const float* dwr = dw + (size_t)dr * dw_stride;
  const float* vr = v + (size_t)r * cols;
  float nn = 0.f, dot = 0.f;
  for (int c = lane; c < cols; c += 32) {
    const float x = vr[c];
    nn = fmaf(x, x, nn);
    dot = fmaf(dwr[c] * col_scale, x, dot);
  }
  for (int o = 16; o > 0; o >>= 1) {
    nn += __shfl_xor_sync(0xffffffffu, nn, o);
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
  }
  const float nv = sqrtf(nn);
  if (g) {
    const float dgr = dot / nv;
    const float s = g[r] / nv;
    for (int c = lane; c < cols; c += 32) dv[(size_t)r * cols + c] = s * (dwr[c] * col_scale - dgr * vr[c] / nv);
    if (lane == 0) dg[r] = dgr;
  } else {
    for (int c = lane; c < cols; c += 32) dv[(size_t)r * cols + c] = dwr[c] * col_scale;
  }
  if (lane == 0) {
    dbias[r] = db[dr];
    if (!(fabsf(dot) <= 3.0e38f) && status) atomicAdd(status + kStatusNonFinite, 1);
  }
}

// ---- fused Adam (torch.optim.Adam defaults: no weight decay, no amsgrad) with clip_grad_norm_ folded in
constexpr int kAdamMaxTensors = 64;
struct AdamArgs {
  int n;
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  float* m[kAdamMaxTensors];
  float* v[kAdamMaxTensors];
  long long size[kAdamMaxTensors];
  float lr, b1, b2, eps, bc1, bc2, max_norm;
  double* sumsq;          // device scalar: sum of squared gradients (all tensors)
  float* out_norm;        // optional device scalar: the gradient norm before clipping
};

__global__ void grad_sumsq_kernel(AdamArgs a) {
  __shared__ double sh[8];
  const int ti = blockIdx.y;
  const float* g = a.g[ti];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.size[ti]; i += (long long)gridDim.x * blockDim.x) {
    const float x = g[i];
    acc += (double)x * (double)x;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    if (t != 0.0) atomicAdd(a.sumsq, t);
  }
}

__global__ void adam_kernel(AdamArgs a) {
  const int ti = blockIdx.y;
  float coef = 1.0f;
  const float norm = (float)sqrt(*a.sumsq);
  if (a.max_norm > 0.f) coef = fminf(a.max_norm / (norm + 1e-6f), 1.0f);       // torch.nn.utils.clip_grad_norm_
  if (a.out_norm && ti == 0 && blockIdx.x == 0 && threadIdx.x == 0) *a.out_norm = norm;
  float* p = a.p[ti];
  const float* g = a.g[ti];
  float* m = a.m[ti];
  float* v = a.v[ti];
  const float step = a.lr / a.bc1;
  const float inv_sqrt_bc2 = rsqrtf(a.bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.size[ti]; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = a.b1 * m[i] + (1.0f - a.b1) * gi;
    const float vi = a.b2 * v[i] + (1.0f - a.b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step * mi / (sqrtf(vi) * inv_sqrt_bc2 + a.eps);
  }
}

static long long tiles_for(long long n, int with_grad) {
  const int per = with_grad ? kTileN / 4 : kTileN;
  return (n + per - 1) / per;
}

// buffers are laid out for an EVEN number of 64-column tiles: the CTA-pair forward writes whole 128-column pair tiles
static long long alloc_tiles(long long n_tiles) { return (n_tiles + 1) / 2 * 2; }

static void save_offsets(const TrainPlan& t, long long n_tiles, long long* off, long long* total) {
  n_tiles = alloc_tiles(n_tiles);
  long long o = 0;
  for (int l = 0; l < t.n_run; ++l) {
    off[l] = o;
    o += (long long)t.save_kc[l] * kBCoreStride * n_tiles;
  }
  for (int l = t.n_run; l < kMaxLayers; ++l) off[l] = o;
  *total = o;
}

static void dz_offsets(const TrainPlan& t, long long n_tiles, long long base, long long* off, long long* total) {
  n_tiles = alloc_tiles(n_tiles);
  long long o = base;
  for (int l = 0; l < t.n_run; ++l) {
    off[l] = o;
    o += (long long)t.dz_kc[l] * kBCoreStride * n_tiles;
  }
  *total = o;
}

constexpr long long kWsScale = 256;        // gscale[2] floats at 0, max bits at 16
// then the CTA-pair sweep's scratch for the gradient of the skip-connection PE rows (20 KiB per pair, up to 128 pairs)
constexpr long long kWsStash = 128 * (long long)kBwdStashBytesPerPair;
constexpr long long kWsHeader = kWsScale + kWsStash;

template <int KIND, int MODE>
static int run_backward(const mvsdf_net* net, const void* packed_t, const float* x, long long n, const void* save, const float* g_full,
                        const float* g_grad, const float* rgb, size_t ws_bytes, void* ws, float* out_dx, float* d_points,
                        float* d_normals, float* d_feats, float* out_dw, float* out_db, cudaStream_t st) {
  if (sm_count() <= 0) return fail(MVSDF_ERR_CUDA, "no CUDA device (the product path has no CPU fallback)");
  const NetPlan& p = net->plan;
  TrainPlan t;
  make_train_plan(p, t);
  const long long n_tiles = tiles_for(n, MODE);
  long long dz_total = 0;
  BwdArgs a;
  memset(&a, 0, sizeof(a));
  dz_offsets(t, n_tiles, kWsHeader, a.dz_off, &dz_total);
  const bool with_dw = out_dw != nullptr;       // false: dx-only sweep (no gradient dumps, no dW GEMM, no bias gradients)
  if ((long long)ws_bytes < (with_dw ? dz_total : kWsHeader))
    return fail(MVSDF_ERR_WORKSPACE, "backward workspace too small (%zu < %lld)", ws_bytes, with_dw ? dz_total : kWsHeader);
  uint8_t* w8 = static_cast<uint8_t*>(ws);
  float* gscale = reinterpret_cast<float*>(w8);
  unsigned* maxbits = reinterpret_cast<unsigned*>(w8 + 16);
  int rc = check_cuda(cudaMemsetAsync(w8, 0, kWsScale, st), "memset backward header");
  if (rc) return rc;
  if (with_dw && (rc = check_cuda(cudaMemsetAsync(out_dw, 0, (size_t)t.dw_total * 4, st), "memset dW"))) return rc;
  if (with_dw && (rc = check_cuda(cudaMemsetAsync(out_db, 0, (size_t)t.db_total * 4, st), "memset db"))) return rc;
  if (out_dx && (rc = check_cuda(cudaMemsetAsync(out_dx, 0, (size_t)n * 12, st), "memset dx"))) return rc;
  const int sms = sm_count();
  // gradient scale from the upstream magnitudes
  const long long n_full = KIND == NET_SDF ? n * (p.feat_size + 2) : n * 3;
  if (g_full) {
    note_launch();
    absmax_kernel<<<std::min<long long>((n_full + 255) / 256, sms * 8), 256, 0, st>>>(g_full, n_full, maxbits);
  }
  if (g_grad) {
    note_launch();
    absmax_kernel<<<std::min<long long>((n * 3 + 255) / 256, sms * 8), 256, 0, st>>>(g_grad, n * 3, maxbits);
  }
  note_launch();
  gscale_kernel<<<1, 32, 0, st>>>(maxbits, gscale, nullptr);

  a.packed_t = static_cast<const uint8_t*>(packed_t);
  a.n_run = t.n_run;
  a.skip_layer = p.skip_layer;
  a.skip_rows_begin = p.skip_rows_begin;
  a.pe_dim = p.pe_dim;
  a.k_cores_max = t.k_cores_max;
  a.feat_size = p.feat_size;
  a.n = n;
  a.gscale = gscale;
  a.x = x;
  a.save = static_cast<const uint8_t*>(save);
  long long save_total = 0;
  save_offsets(t, n_tiles, a.save_off, &save_total);
  for (int l = 0; l < t.n_run; ++l) {
    a.save_kc[l] = t.save_kc[l];
    a.dz_kc[l] = t.dz_kc[l];
    a.db_off[l] = t.db_off[l];
    a.Lt[l] = t.Lt[l];
    a.fwd_layer[l] = t.fwd_layer[l];
  }
  a.g_full = g_full;
  a.g_grad = g_grad;
  a.rgb = rgb;
  a.dz = with_dw ? w8 : nullptr;
  a.db = with_dw ? out_db : nullptr;
  a.dx = out_dx;
  a.d_points = d_points;
  a.d_normals = d_normals;
  a.d_feats = d_feats;
  a.status = nullptr;

  // head bias gradient
  {
    const LayerPlan& H = t.fwd[t.n_run - 1];
    const int rows = H.m_tiles * kTileM;
    if (g_full && with_dw) {
      note_launch();
      head_db_kernel<<<rows, 256, 0, st>>>(g_full, rgb, n, KIND == NET_SDF ? p.feat_size + 2 : 3, H.row_map, H.out_dim, p.feat_size,
                                          out_db + t.db_off[t.n_run - 1]);
    }
  }

  const size_t smem = mlp_smem_bytes(t.k_cores_max);
  // batches that fill every SM twice run the CTA-pair sweep (half the weight stream per SM); MVSDF_PAIR_SWEEP=0 keeps the single-CTA one
  const char* pair_e = getenv("MVSDF_PAIR_SWEEP");       // read per call: the A/B test toggles it
  const bool pair_env = !(pair_e && pair_e[0] == '0');
  const int max_pairs = std::min(sms / 2, 128);
  if (pair_env && n_tiles >= 2LL * sms) {
    auto kern = mlp_bwd_sweep_pair_kernel<KIND, MODE>;
    if ((rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                         "cudaFuncSetAttribute(mlp_bwd_sweep_pair_kernel)")))
      return rc;
    const int n_pairs = (int)std::min<long long>((n_tiles + 1) / 2, max_pairs);
    note_launch();
    void* pe = prof_begin_ext(5, st);
    kern<<<2 * n_pairs, kP2Threads, smem, st>>>(a, reinterpret_cast<float*>(w8 + kWsScale));
    prof_end_ext(pe, st);
    if ((rc = check_cuda(cudaGetLastError(), "launch mlp_bwd_sweep_pair_kernel"))) return rc;
  } else {
    auto kern = mlp_bwd_sweep_kernel<KIND, MODE>;
    if ((rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                         "cudaFuncSetAttribute(mlp_bwd_sweep_kernel)")))
      return rc;
    const int grid = (int)std::min<long long>(n_tiles, sms);
    note_launch();
    void* pe = prof_begin_ext(5, st);
    kern<<<grid, kMlpThreads, smem, st>>>(a);
    prof_end_ext(pe, st);
    if ((rc = check_cuda(cudaGetLastError(), "launch mlp_bwd_sweep_kernel"))) return rc;
  }

  if (!with_dw) return MVSDF_OK;
  // dW over the points
  DwArgs d;
  memset(&d, 0, sizeof(d));
  d.dz = w8;
  d.save = static_cast<const uint8_t*>(save);
  d.dw = out_dw;
  d.gscale = gscale;
  d.n_tiles = n_tiles;
  d.n_layers = t.n_run;
  int total_m = 0;
  for (int l = 0; l < t.n_run; ++l) total_m += t.fwd[l].m_tiles;
  d.splits = std::max(1, std::min<int>((int)std::min<long long>(n_tiles, 64), sms / total_m));      // one wave of CTAs
  int items = 0;
  for (int l = 0; l < t.n_run; ++l) {
    d.item_begin[l] = items;
    items += t.fwd[l].m_tiles * d.splits;
    d.L[l].a_off = a.dz_off[l];
    d.L[l].b_off = a.save_off[l];
    d.L[l].kc_a = t.dz_kc[l];
    d.L[l].kc_b = t.save_kc[l];
    d.L[l].m_tiles = t.fwd[l].m_tiles;
    d.L[l].n_in = t.save_kc[l] * 8;
    d.L[l].w_off = t.dw_off[l];
  }
  d.item_begin[t.n_run] = items;
  if ((rc = check_cuda(cudaFuncSetAttribute(mlp_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dw_smem_bytes()),
                       "cudaFuncSetAttribute(mlp_bwd_dw_kernel)")))
    return rc;
  note_launch();
  void* pe = prof_begin_ext(6, st);
  mlp_bwd_dw_kernel<<<items, kDwThreads, dw_smem_bytes(), st>>>(d);
  prof_end_ext(pe, st);
  return check_cuda(cudaGetLastError(), "launch mlp_bwd_dw_kernel");
}

}  // namespace mvsdf

using namespace mvsdf;

extern "C" {

size_t mvsdf_train_packed_t_bytes(const mvsdf_net* net) {
  if (!net) return 0;
  TrainPlan t;
  make_train_plan(net->plan, t);
  return (size_t)t.packed_t_bytes;
}

size_t mvsdf_train_save_bytes(const mvsdf_net* net, int64_t n, int with_grad) {
  if (!net || n < 0) return 0;
  TrainPlan t;
  make_train_plan(net->plan, t);
  long long off[kMaxLayers], total = 0;
  save_offsets(t, tiles_for(n, with_grad), off, &total);
  return (size_t)total + 256;
}

size_t mvsdf_train_workspace_bytes(const mvsdf_net* net, int64_t n, int with_grad) {
  if (!net || n < 0) return 0;
  TrainPlan t;
  make_train_plan(net->plan, t);
  long long off[kMaxLayers], total = 0;
  dz_offsets(t, tiles_for(n, with_grad), kWsHeader, off, &total);
  return (size_t)total + 256;
}

size_t mvsdf_train_dw_floats(const mvsdf_net* net) {
  if (!net) return 0;
  TrainPlan t;
  make_train_plan(net->plan, t);
  return (size_t)t.dw_total;
}

size_t mvsdf_train_db_floats(const mvsdf_net* net) {
  if (!net) return 0;
  TrainPlan t;
  make_train_plan(net->plan, t);
  return (size_t)t.db_total;
}

int mvsdf_pack_weights_t(const mvsdf_net* net, const float* const* weight_v_host, const float* const* weight_g_host, void* packed_t,
                         void* stream) {
  if (!net || !weight_v_host || !packed_t) return fail(MVSDF_ERR_INVALID, "mvsdf_pack_weights_t: null argument");
  if (sm_count() <= 0) return fail(MVSDF_ERR_CUDA, "no CUDA device (the product path has no CPU fallback)");
  const NetPlan& p = net->plan;
  TrainPlan t;
  make_train_plan(p, t);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* blob = static_cast<uint8_t*>(packed_t);
  float* scale = reinterpret_cast<float*>(blob + t.t_scale_off);
  int sc = 0;
  for (int i = 0; i < t.n_run; ++i) {
    const int fl = t.fwd_layer[i];
    const LayerPlan& F = t.fwd[fl];
    const int s = F.src_layer;
    if (!weight_v_host[s]) return fail(MVSDF_ERR_INVALID, "mvsdf_pack_weights_t: null weight pointer for layer %d", s);
    const float* g = weight_g_host ? weight_g_host[s] : nullptr;
    note_launch();
    t_row_scale_kernel<<<cdiv(F.out_dim, 8), 256, 0, st>>>(weight_v_host[s], g, F.out_dim, F.in_dim, scale + sc);
    const LayerPlan& Lt = t.Lt[i];
    const int total = Lt.m_tiles * kTileM * Lt.k_chunks * (kChunkK / 8);
    note_launch();
    pack_layer_t_kernel<<<cdiv(total, 256), 256, 0, st>>>(weight_v_host[s], scale + sc, Lt, F.in_dim, F.out_dim, p.feat_size, blob);
    sc += F.out_dim;
  }
  return check_cuda(cudaGetLastError(), "pack_weights_t launch");
}

int mvsdf_sdf_forward_train(const mvsdf_net* net, const void* packed, const float* x, int64_t n, size_t save_bytes, void* save,
                            float* out_full, float* out_grad, void* stream) {
  if (!net || net->plan.kind != NET_SDF) return fail(MVSDF_ERR_INVALID, "expected an SDF net plan");
  if (n == 0) return MVSDF_OK;
  if (!save || !out_full || !out_grad) return fail(MVSDF_ERR_INVALID, "mvsdf_sdf_forward_train: null argument");
  if (save_bytes < mvsdf_train_save_bytes(net, n, 1)) return fail(MVSDF_ERR_WORKSPACE, "mvsdf_sdf_forward_train: save buffer too small");
  TrainPlan t;
  make_train_plan(net->plan, t);
  long long off[kMaxLayers], total = 0;
  save_offsets(t, tiles_for(n, 1), off, &total);
  return mlp_sdf(net, packed, x, n, nullptr, MVSDF_HEAD_FULL, nullptr, out_full, out_grad, true, static_cast<cudaStream_t>(stream),
                 false, static_cast<uint8_t*>(save), off);
}

int mvsdf_sdf_backward(const mvsdf_net* net, const void* packed_t, const float* x, int64_t n, const void* save, const float* g_full,
                       const float* g_grad, size_t workspace_bytes, void* workspace, float* out_dx, float* out_dw, float* out_db,
                       void* stream) {
  if (!net || net->plan.kind != NET_SDF) return fail(MVSDF_ERR_INVALID, "expected an SDF net plan");
  if (!packed_t || !x || !save || !workspace || (out_dw && !out_db) || (!out_dw && !out_dx) || n <= 0)
    return fail(MVSDF_ERR_INVALID, "mvsdf_sdf_backward: null argument / empty batch");
  return run_backward<NET_SDF, 1>(net, packed_t, x, n, save, g_full, g_grad, nullptr, workspace_bytes, workspace, out_dx, nullptr,
                                  nullptr, nullptr, out_dw, out_db, static_cast<cudaStream_t>(stream));
}

int mvsdf_render_forward_train(const mvsdf_net* net, const void* packed, const float* points, const float* view_dirs,
                               const float* normals, const float* features, int64_t n, size_t save_bytes, void* save, float* out_rgb,
                               void* stream) {
  if (!net || net->plan.kind != NET_RENDER) return fail(MVSDF_ERR_INVALID, "expected a rendering net plan");
  if (n == 0) return MVSDF_OK;
  if (!save || !out_rgb) return fail(MVSDF_ERR_INVALID, "mvsdf_render_forward_train: null argument");
  if (save_bytes < mvsdf_train_save_bytes(net, n, 0)) return fail(MVSDF_ERR_WORKSPACE, "mvsdf_render_forward_train: save buffer too small");
  TrainPlan t;
  make_train_plan(net->plan, t);
  long long off[kMaxLayers], total = 0;
  save_offsets(t, tiles_for(n, 0), off, &total);
  return mlp_render(net, packed, points, view_dirs, normals, features, 0, n, nullptr, out_rgb, static_cast<cudaStream_t>(stream),
                    static_cast<uint8_t*>(save), off);
}

int mvsdf_render_backward(const mvsdf_net* net, const void* packed_t, int64_t n, const void* save, const float* rgb, const float* g_rgb,
                          const float* view_dirs, size_t workspace_bytes, void* workspace, float* d_points, float* d_normals,
                          float* d_feats, float* d_view, float* out_dw, float* out_db, void* stream) {
  if (!net || net->plan.kind != NET_RENDER) return fail(MVSDF_ERR_INVALID, "expected a rendering net plan");
  if (!packed_t || !save || !rgb || !g_rgb || !workspace || !out_dw || !out_db || n <= 0)
    return fail(MVSDF_ERR_INVALID, "mvsdf_render_backward: null argument / empty batch");
  if (d_view && !view_dirs) return fail(MVSDF_ERR_INVALID, "mvsdf_render_backward: d_view needs view_dirs");
  return run_backward<NET_RENDER, 0>(net, packed_t, view_dirs, n, save, g_rgb, nullptr, rgb, workspace_bytes, workspace, d_view, d_points,
                                     d_normals, d_feats, out_dw, out_db, static_cast<cudaStream_t>(stream));
}

int mvsdf_weight_grads(const mvsdf_net* net, const float* dw, const float* db, const float* const* weight_v_host,
                       const float* const* weight_g_host, float* const* out_dv_host, float* const* out_dg_host,
                       float* const* out_dbias_host, void* stream) {
  if (!net || !dw || !db || !weight_v_host || !out_dv_host || !out_dbias_host) return fail(MVSDF_ERR_INVALID, "mvsdf_weight_grads: null argument");
  if (sm_count() <= 0) return fail(MVSDF_ERR_CUDA, "no CUDA device (the product path has no CPU fallback)");
  TrainPlan t;
  make_train_plan(net->plan, t);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int l = 0; l < t.n_run; ++l) {
    const LayerPlan& F = t.fwd[l];
    const int s = F.src_layer;
    const float* g = weight_g_host ? weight_g_host[s] : nullptr;
    if (!weight_v_host[s] || !out_dv_host[s] || !out_dbias_host[s] || (g && (!out_dg_host || !out_dg_host[s])))
      return fail(MVSDF_ERR_INVALID, "mvsdf_weight_grads: null pointer for layer %d", s);
    note_launch();
    weight_grads_kernel<<<cdiv(F.out_dim, 8), 256, 0, st>>>(dw + t.dw_off[l], db + t.db_off[l], t.save_kc[l] * 8, F.row_map,
                                                           net->plan.feat_size, F.col_scale, weight_v_host[s], g, F.out_dim, F.in_dim,
                                                           out_dv_host[s], g ? out_dg_host[s] : nullptr, out_dbias_host[s], nullptr);
  }
  return check_cuda(cudaGetLastError(), "weight_grads launch");
}

int mvsdf_adam_step(int n_tensors, float* const* params_host, const float* const* grads_host, float* const* exp_avg_host,
                    float* const* exp_avg_sq_host, const int64_t* sizes_host, float lr, float beta1, float beta2, float eps, int step,
                    float max_grad_norm, double* scratch_sumsq, float* out_grad_norm, void* stream) {
  if (n_tensors <= 0 || !params_host || !grads_host || !exp_avg_host || !exp_avg_sq_host || !sizes_host || !scratch_sumsq || step < 1)
    return fail(MVSDF_ERR_INVALID, "mvsdf_adam_step: bad argument");
  if (sm_count() <= 0) return fail(MVSDF_ERR_CUDA, "no CUDA device (the product path has no CPU fallback)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = check_cuda(cudaMemsetAsync(scratch_sumsq, 0, sizeof(double), st), "memset sumsq");
  if (rc) return rc;
  const float bc1 = 1.0f - (float)std::pow((double)beta1, step), bc2 = 1.0f - (float)std::pow((double)beta2, step);
  // the squared norm runs over ALL tensors before any of them is updated
  for (int pass = 0; pass < 2; ++pass) {
    for (int begin = 0; begin < n_tensors; begin += kAdamMaxTensors) {
      AdamArgs a;
      memset(&a, 0, sizeof(a));
      a.n = std::min(kAdamMaxTensors, n_tensors - begin);
      long long max_size = 0;
      for (int i = 0; i < a.n; ++i) {
        a.p[i] = params_host[begin + i];
        a.g[i] = grads_host[begin + i];
        a.m[i] = exp_avg_host[begin + i];
        a.v[i] = exp_avg_sq_host[begin + i];
        a.size[i] = sizes_host[begin + i];
        if (!a.p[i] || !a.g[i] || !a.m[i] || !a.v[i] || a.size[i] < 0) return fail(MVSDF_ERR_INVALID, "mvsdf_adam_step: null tensor %d", begin + i);
        max_size = std::max(max_size, a.size[i]);
      }
      a.lr = lr;
      a.b1 = beta1;
      a.b2 = beta2;
      a.eps = eps;
      a.bc1 = bc1;
      a.bc2 = bc2;
      a.max_norm = max_grad_norm;
      a.sumsq = scratch_sumsq;
      a.out_norm = out_grad_norm;
      const int gx = (int)std::max<long long>(1, std::min<long long>((max_size + 1023) / 1024, 64));
      note_launch();
      if (pass == 0) grad_sumsq_kernel<<<dim3(gx, a.n), 256, 0, st>>>(a);
      else adam_kernel<<<dim3(gx, a.n), 256, 0, st>>>(a);
    }
  }
  return check_cuda(cudaGetLastError(), "adam launch");
}

}  // extern "C"

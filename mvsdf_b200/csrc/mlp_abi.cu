// C ABI, part 1: network plans, weight packer, MLP tile launches.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/mvsdf_b200.h"
#include "internal.h"
#include "mlp_pair2_kernel.cuh"

namespace mvsdf {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return MVSDF_OK;
  return fail(MVSDF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 0;
  }
  return n;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

static long long g_launches = 0;
void note_launch() { ++g_launches; }

// Optional per-launch timing of the MLP tile kernels with CUDA events on the launching stream
// (bench.py's roofline leg). kinds: 0 SDF plain / SDF-only head, 1 SDF plain / full head, 2 SDF value+grad, 3 render.
struct ProfEvent { cudaEvent_t e0, e1; int kind; };
static bool g_prof_on = false;
static std::vector<ProfEvent> g_prof_pool;
static size_t g_prof_used = 0;
void* prof_begin_ext(int kind, cudaStream_t st);
void prof_end_ext(void* p, cudaStream_t st);
static ProfEvent* prof_begin(int kind, cudaStream_t st) {
  if (!g_prof_on) return nullptr;
  // a launch that is being captured into a CUDA graph is not timed: an event recorded during capture becomes a graph node
  // and can never be synchronised on (a forward repeated without the prefilter captures a new graph mid-measurement)
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return nullptr;
  if (g_prof_used == g_prof_pool.size()) {
    ProfEvent p{};
    if (cudaEventCreate(&p.e0) != cudaSuccess || cudaEventCreate(&p.e1) != cudaSuccess) return nullptr;
    g_prof_pool.push_back(p);
  }
  ProfEvent* p = &g_prof_pool[g_prof_used++];
  p->kind = kind;
  cudaEventRecord(p->e0, st);
  return p;
}
static void prof_end(ProfEvent* p, cudaStream_t st) {
  if (p) cudaEventRecord(p->e1, st);
}
// for the other translation units (train_abi.cu): the pool may grow between begin and end, so hand out the INDEX
void* prof_begin_ext(int kind, cudaStream_t st) {
  ProfEvent* p = prof_begin(kind, st);
  return p ? reinterpret_cast<void*>((p - g_prof_pool.data()) + 1) : nullptr;
}
void prof_end_ext(void* h, cudaStream_t st) {
  if (h) prof_end(&g_prof_pool[reinterpret_cast<size_t>(h) - 1], st);
}

static void finalize_plan(NetPlan& p) {
  long long off = 0;
  int bias = 0;
  p.k_cores_max = kPeCores;
  for (int i = 0; i < p.n_layers; ++i) {
    LayerPlan& l = p.L[i];
    l.w_off = off;
    off += (long long)l.m_tiles * l.k_chunks * kStageBytes;
    l.bias_off = bias;
    bias += l.m_tiles * kTileM;
    p.k_cores_max = std::max(p.k_cores_max, l.k_chunks * (kChunkK / 8));
  }
  p.bias_area_off = off;
  off += (long long)bias * 4;
  off = (off + 255) / 256 * 256;
  p.scale_area_off = off;
  int sc = 0;
  for (int s = 0; s < p.n_src_layers; ++s) {
    p.scale_off[s] = sc;
    int rows = 0;
    for (int i = 0; i < p.n_layers; ++i)
      if (p.L[i].src_layer == s) rows = std::max(rows, p.L[i].out_dim);
    sc += rows;
  }
  off += (long long)sc * 4;
  off = (off + 255) / 256 * 256;
  p.status_off = off;
  off += kStatusWords * 4;
  p.total_bytes = (off + 255) / 256 * 256;
}

// one warp per source row: scale = g / ||v||_2   (torch._weight_norm, dim=0)
__global__ void row_scale_kernel(const float* __restrict__ v, const float* __restrict__ g, int rows, int cols,
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

// one thread per (padded destination row, 8-wide K core): 16 B of hi and 16 B of lo
__global__ void pack_layer_kernel(const float* __restrict__ v, const float* __restrict__ scale,
                                  const float* __restrict__ bias_src, LayerPlan lp, int feat_size,
                                  uint8_t* __restrict__ packed, float* __restrict__ bias_dst, int* __restrict__ status) {
  const int cores = lp.k_chunks * (kChunkK / 8);
  const int rows = lp.m_tiles * kTileM;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cores) return;
  const int dr = idx / cores, kcore = idx - dr * cores;
  int sr;
  if (lp.row_map == 0) sr = dr < lp.out_dim ? dr : -1;
  else if (lp.row_map == 1) sr = dr == 0 ? 0 : -1;
  else sr = dr < feat_size ? dr + 2 : (dr < feat_size + 2 ? dr - feat_size : -1);
  __align__(16) __half hi[8];
  __align__(16) __half lo[8];
  const float sc = sr >= 0 ? scale[sr] : 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = kcore * 8 + e;
    float w = 0.f;
    if (sr >= 0 && k < lp.in_dim) w = v[(size_t)sr * lp.in_dim + k] * sc * lp.col_scale * kWeightScale;
    if (!(fabsf(w) < 65504.0f)) atomicAdd(status + kStatusPackRange, 1);      // beyond fp16 (or NaN): the hi part would be inf
    hi[e] = __float2half_rn(w);
    lo[e] = __float2half_rn(w - __half2float(hi[e]));
  }
  const int m = dr / kTileM, r = dr - m * kTileM;
  const int kc = kcore / (kChunkK / 8), kin = kcore - kc * (kChunkK / 8);
  uint8_t* tile = packed + lp.w_off + (size_t)(m * lp.k_chunks + kc) * kStageBytes;
  const int off = (r >> 3) * 512 + kin * 128 + (r & 7) * 16;
  *reinterpret_cast<uint4*>(tile + off) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(tile + kTileBytes + off) = *reinterpret_cast<const uint4*>(lo);
  if (kcore == 0) bias_dst[lp.bias_off + dr] = sr >= 0 ? bias_src[sr] : 0.f;
}

static unsigned long long* g_trace = nullptr;

int fill_mlp_args(const NetPlan& p, const void* packed, int head, MlpArgs& a) {
  memset(&a, 0, sizeof(a));
  a.trace = g_trace;
  a.packed = static_cast<const uint8_t*>(packed);
  a.bias = reinterpret_cast<const float*>(a.packed + p.bias_area_off);
  a.status = reinterpret_cast<int*>(const_cast<uint8_t*>(a.packed) + p.status_off);
  a.n_run = p.n_hidden + 1;
  a.skip_layer = p.skip_layer;
  a.skip_rows_begin = p.skip_rows_begin;
  a.pe_dim = p.pe_dim;
  a.k_cores_max = p.k_cores_max;
  a.head = head;
  a.feat_size = p.feat_size;
  for (int l = 0; l < p.n_hidden; ++l) a.L[l] = p.L[l];
  a.L[p.n_hidden] = p.L[p.head_index[head]];
  return MVSDF_OK;
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <int KIND, int MODE, int CL>
static int launch_mlp_cl(const NetPlan& p, MlpArgs& a, long long tiles, bool device_count, cudaStream_t st) {
  const int sms = sm_count();
  const size_t smem = mlp_smem_bytes(p.k_cores_max);
  auto kern = mlp_tile_kernel<KIND, MODE, CL>;
  int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                      "cudaFuncSetAttribute(mlp_tile_kernel)");
  if (rc) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kMlpThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  // persistent grid: as many co-resident clusters as the device can hold (cached per instantiation)
  static int max_clusters = 0;
  if (max_clusters == 0) {
    if (CL == 1) {
      max_clusters = sms;
    } else {
      cfg.gridDim = dim3((sms / CL) * CL);
      if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters <= 0) {
        cudaGetLastError();
        max_clusters = std::max(1, sms / CL - 2);
      }
      max_clusters = std::min(max_clusters, sms / CL);
    }
  }
  const long long groups = (tiles + CL - 1) / CL;
  const int n_clusters = device_count ? max_clusters : (int)std::min<long long>(groups, max_clusters);
  cfg.gridDim = dim3(n_clusters * CL);
  const int kind = KIND == NET_RENDER ? 3 : (MODE == 1 ? 2 : (a.head == HEAD_FULL ? 1 : 0));
  ProfEvent* pe = prof_begin(kind, st);
  note_launch();
  rc = check_cuda(cudaLaunchKernelEx(&cfg, kern, a), "launch mlp_tile_kernel");
  prof_end(pe, st);
  return rc;
}

// CTA-pair kernel (cta_group::2): 74 pairs, each walks 128-column tiles
template <int KIND, int MODE, int VER>
static int launch_mlp_pair(const NetPlan& p, MlpArgs& a, long long pair_tiles, bool device_count, cudaStream_t st) {
  const int sms = sm_count();
  const size_t smem = mlp_smem_bytes(p.k_cores_max);
  // the screening-precision instantiation exists for the plain SDF evaluation only (the tracer's sampler prefilter)
  constexpr bool kCanLp = KIND == NET_SDF && MODE == 0 && VER == 2;
  // the saving instantiation exists for the two training forwards: SDF value+gradient and the rendering net
  constexpr bool kCanSave = (KIND == NET_SDF && MODE == 1) || (KIND == NET_RENDER && MODE == 0);
  // exact SDF-only evaluations: the one-row head runs as a dot product in the last hidden layer's epilogue (MVSDF_FUSE_HEAD=0:
  // as an UMMA layer like every other head; read per call, the A/B test toggles it)
  const char* fuse_e = getenv("MVSDF_FUSE_HEAD");
  a.fuse_head = (VER == 2 && KIND == NET_SDF && MODE == 0 && !a.lp && !a.save && a.head == HEAD_SDF_ONLY && a.n_run >= 2 &&
                 !(fuse_e && fuse_e[0] == '0'))
                    ? 1
                    : 0;
  auto kern = (kCanLp && a.lp) ? mlp_pair2_kernel<KIND, MODE, kCanLp ? 1 : 0>
                               : ((kCanSave && a.save) ? mlp_pair2_kernel<KIND, MODE, 0, kCanSave ? 1 : 0> : mlp_pair2_kernel<KIND, MODE, 0>);
  int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                      "cudaFuncSetAttribute(mlp_pair_kernel)");
  if (rc) return rc;
  const int max_pairs = sms / 2;
  const int n_pairs = device_count ? max_pairs : (int)std::min<long long>(pair_tiles, max_pairs);
  const int kind = KIND == NET_RENDER ? 3 : (MODE == 1 ? 2 : (a.head == HEAD_FULL ? 1 : ((kCanLp && a.lp) ? 4 : 0)));
  ProfEvent* pe = prof_begin(kind, st);
  note_launch();
  kern<<<2 * n_pairs, kP2Threads, smem, st>>>(a);
  prof_end(pe, st);
  return check_cuda(cudaGetLastError(), "launch mlp_pair_kernel");
}

template <int KIND, int MODE>
static int launch_mlp(const NetPlan& p, MlpArgs& a, int64_t n, const int32_t* n_dev, cudaStream_t st) {
  if (sm_count() <= 0) return fail(MVSDF_ERR_CUDA, "no CUDA device (the product path has no CPU fallback)");
  a.n = n;
  a.n_ptr = n_dev;
  static const int cl = env_int("MVSDF_CLUSTER", 1);   // multicast clusters measured no faster (profiles/r01): L2 is not the bound
  static const int dbg = env_int("MVSDF_DEBUG_FLAGS", 0);
  a.debug = dbg;
  const int per_tile = MODE == 0 ? kTileN : kTileN / 4;
  const long long tiles = (n + per_tile - 1) / per_tile;
  if (n_dev == nullptr && tiles == 0) return MVSDF_OK;
  // small host-known batches cannot fill clusters: fall back to single-CTA scheduling (same kernel, CL = 1)
  const bool small = n_dev == nullptr && tiles < 2 * sm_count();
  // training forward (saves the layer inputs): single-CTA kernel for small batches, the CTA-pair kernel otherwise
  constexpr bool kPairSaves = (KIND == NET_SDF && MODE == 1) || (KIND == NET_RENDER && MODE == 0);
  static const int pair_save = env_int("MVSDF_PAIR_SAVE", 1);
  if (a.save && (small || !kPairSaves || !pair_save)) return launch_mlp_cl<KIND, MODE, 1>(p, a, tiles, false, st);
  // 1 = split-K pipelined CTA-pair kernel (default), 0 = single-CTA kernel (A/B runs)
  static const int pair = env_int("MVSDF_PAIR", 1);
  if (pair && !small) return launch_mlp_pair<KIND, MODE, 2>(p, a, (tiles + 1) / 2, n_dev != nullptr, st);
  if (cl >= 4 && !small) return launch_mlp_cl<KIND, MODE, 4>(p, a, tiles, n_dev != nullptr, st);
  if (cl >= 2 && !small) return launch_mlp_cl<KIND, MODE, 2>(p, a, tiles, n_dev != nullptr, st);
  return launch_mlp_cl<KIND, MODE, 1>(p, a, tiles, n_dev != nullptr, st);
}

int mlp_sdf(const mvsdf_net* net, const void* packed, const float* x, int64_t n, const int32_t* n_dev, int head,
            float* out_sdf, float* out_full, float* out_grad, bool with_grad, cudaStream_t st, bool screening, uint8_t* save,
            const long long* save_off) {
  if (!net || net->plan.kind != NET_SDF) return fail(MVSDF_ERR_INVALID, "expected an SDF net plan");
  if (n == 0 && !n_dev) return MVSDF_OK;   // empty batch: nothing to enqueue
  if (!packed || (!x && n > 0) || n < 0) return fail(MVSDF_ERR_INVALID, "null pointer / negative count");
  if (head != HEAD_SDF_ONLY && head != HEAD_FULL) return fail(MVSDF_ERR_INVALID, "bad head %d", head);
  if (head == HEAD_SDF_ONLY && !out_sdf) return fail(MVSDF_ERR_INVALID, "out_sdf is required for the SDF-only head");
  if (head == HEAD_FULL && !out_full) return fail(MVSDF_ERR_INVALID, "out_full is required for the full head");
  if (with_grad && !out_grad) return fail(MVSDF_ERR_INVALID, "out_grad is required");
  MlpArgs a;
  fill_mlp_args(net->plan, packed, head, a);
  a.x = x;
  a.out_sdf = out_sdf;
  a.out_full = out_full;
  a.out_grad = out_grad;
  a.lp = (screening && !with_grad && head == HEAD_SDF_ONLY) ? 1 : 0;
  if (save) {
    if (n_dev) return fail(MVSDF_ERR_INVALID, "the saving forward needs a host-side count");
    a.save = save;
    for (int l = 0; l < kMaxLayers; ++l) a.save_off[l] = save_off[l];
  }
  return with_grad ? launch_mlp<NET_SDF, 1>(net->plan, a, n, n_dev, st) : launch_mlp<NET_SDF, 0>(net->plan, a, n, n_dev, st);
}

int mlp_render(const mvsdf_net* net, const void* packed, const float* pts, const float* view, const float* normals,
               const float* feats, int feat_stride, int64_t n, const int32_t* n_dev, float* rgb, cudaStream_t st, uint8_t* save,
               const long long* save_off) {
  if (!net || net->plan.kind != NET_RENDER) return fail(MVSDF_ERR_INVALID, "expected a rendering net plan");
  if (n == 0 && !n_dev) return MVSDF_OK;
  if (!packed || n < 0 || !rgb || ((!pts || !view || !normals || !feats) && n > 0))
    return fail(MVSDF_ERR_INVALID, "null pointer / negative count");
  MlpArgs a;
  fill_mlp_args(net->plan, packed, HEAD_FULL, a);
  a.x = pts;
  a.view = view;
  a.normals = normals;
  a.feats = feats;
  a.feat_stride = feat_stride > 0 ? feat_stride : net->plan.feat_size;
  a.out_rgb = rgb;
  if (save) {
    if (n_dev) return fail(MVSDF_ERR_INVALID, "the saving forward needs a host-side count");
    a.save = save;
    for (int l = 0; l < kMaxLayers; ++l) a.save_off[l] = save_off[l];
  }
  return launch_mlp<NET_RENDER, 0>(net->plan, a, n, n_dev, st);
}

}  // namespace mvsdf

using namespace mvsdf;

extern "C" {

int mvsdf_abi_version(void) { return 1; }
/* debug only (not in the public header): device buffer of >= 1 MiB that receives a clock64 timeline of CTA pair 0 */
void mvsdf_debug_set_trace(void* buf) { g_trace = static_cast<unsigned long long*>(buf); }
long long mvsdf_launch_count(void) { return g_launches; }
void mvsdf_launch_count_add(long long n) { g_launches += n; }
void mvsdf_profile_enable(int on) {
  g_prof_on = on != 0;
  g_prof_used = 0;
}
int mvsdf_profile_collect(float* ms_by_kind, int* launches_by_kind) {
  if (!ms_by_kind || !launches_by_kind) return fail(MVSDF_ERR_INVALID, "mvsdf_profile_collect: null argument");
  for (int k = 0; k < MVSDF_PROFILE_KINDS; ++k) {
    ms_by_kind[k] = 0.f;
    launches_by_kind[k] = 0;
  }
  for (size_t i = 0; i < g_prof_used; ++i) {
    ProfEvent& p = g_prof_pool[i];
    int rc = check_cuda(cudaEventSynchronize(p.e1), "profile event sync");
    if (rc) return rc;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.e0, p.e1);
    ms_by_kind[p.kind] += ms;
    launches_by_kind[p.kind] += 1;
  }
  g_prof_used = 0;
  return MVSDF_OK;
}
const char* mvsdf_last_error(void) { return g_err; }

mvsdf_net* mvsdf_sdf_net_create(int width, int n_hidden, int skip_layer, int n_freqs, int feature_size) {
  const int d0 = 3 + 6 * n_freqs;
  // widths are whole 128-row output tiles: a partial tile would write activation rows beyond the operand buffer
  if (width < 128 || width > 512 || width % 128 != 0 || n_hidden < 2 || n_hidden + 2 > kMaxLayers || n_freqs != 6 ||
      skip_layer < 1 || skip_layer >= n_hidden || width - d0 <= 0 || feature_size < 1 ||
      ceil_div(feature_size + 2, kTileM) * kTileN > kTmemCols) {
    fail(MVSDF_ERR_INVALID, "unsupported SDF net: width=%d hidden=%d skip=%d freqs=%d feat=%d", width, n_hidden,
         skip_layer, n_freqs, feature_size);
    return nullptr;
  }
  mvsdf_net* net = new (std::nothrow) mvsdf_net();
  if (!net) return nullptr;
  NetPlan& p = net->plan;
  memset(&p, 0, sizeof(p));
  p.kind = NET_SDF;
  p.width = width;
  p.n_hidden = n_hidden;
  p.skip_layer = skip_layer;
  p.skip_rows_begin = width - d0;
  p.pe_dim = d0;
  p.feat_size = feature_size;
  p.n_src_layers = n_hidden + 1;
  for (int l = 0; l < n_hidden; ++l) {
    LayerPlan& L = p.L[l];
    L.in_dim = l == 0 ? d0 : width;
    L.out_dim = (l + 1 == skip_layer) ? width - d0 : width;
    L.k_chunks = ceil_div(L.in_dim, kChunkK);
    L.m_tiles = ceil_div(width, kTileM);
    L.act = ACT_SOFTPLUS100;
    L.b_from_pe = l == 0;
    L.row_map = 0;
    L.src_layer = l;
    L.col_scale = l == skip_layer ? (float)(1.0 / std::sqrt(2.0)) : 1.0f;
  }
  for (int hd = 0; hd < 2; ++hd) {
    LayerPlan& L = p.L[n_hidden + hd];
    L.in_dim = width;
    L.out_dim = feature_size + 2;
    L.k_chunks = ceil_div(width, kChunkK);
    L.m_tiles = hd == 0 ? 1 : ceil_div(feature_size + 2, kTileM);
    L.act = ACT_NONE;
    L.b_from_pe = 0;
    L.row_map = hd == 0 ? 1 : 2;
    L.src_layer = n_hidden;
    L.col_scale = 1.0f;
    p.head_index[hd] = n_hidden + hd;
  }
  p.n_layers = n_hidden + 2;
  finalize_plan(p);
  return net;
}

mvsdf_net* mvsdf_render_net_create(int width, int n_hidden, int n_freqs_view, int feature_size) {
  if (width < 128 || width > 512 || width % 128 != 0 || n_hidden < 1 || n_hidden + 1 > kMaxLayers || n_freqs_view != 4 ||
      feature_size < 1) {
    fail(MVSDF_ERR_INVALID, "unsupported rendering net: width=%d hidden=%d freqs=%d feat=%d", width, n_hidden,
         n_freqs_view, feature_size);
    return nullptr;
  }
  mvsdf_net* net = new (std::nothrow) mvsdf_net();
  if (!net) return nullptr;
  NetPlan& p = net->plan;
  memset(&p, 0, sizeof(p));
  p.kind = NET_RENDER;
  p.width = width;
  p.n_hidden = n_hidden;
  p.skip_layer = -1;
  p.skip_rows_begin = 1 << 30;
  p.pe_dim = 0;
  p.feat_size = feature_size;
  p.n_src_layers = n_hidden + 1;
  const int d0 = 3 + (3 + 6 * n_freqs_view) + 3 + feature_size;
  for (int l = 0; l <= n_hidden; ++l) {
    LayerPlan& L = p.L[l];
    L.in_dim = l == 0 ? d0 : width;
    L.out_dim = l == n_hidden ? 3 : width;
    L.k_chunks = ceil_div(L.in_dim, kChunkK);
    L.m_tiles = l == n_hidden ? 1 : ceil_div(width, kTileM);
    L.act = l == n_hidden ? ACT_NONE : ACT_RELU;
    L.b_from_pe = 0;
    L.row_map = 0;
    L.src_layer = l;
    L.col_scale = 1.0f;
  }
  p.head_index[0] = p.head_index[1] = n_hidden;
  p.n_layers = n_hidden + 1;
  finalize_plan(p);
  return net;
}

void mvsdf_net_destroy(mvsdf_net* net) { delete net; }
int mvsdf_net_num_layers(const mvsdf_net* net) { return net ? net->plan.n_src_layers : 0; }
size_t mvsdf_net_packed_bytes(const mvsdf_net* net) { return net ? (size_t)net->plan.total_bytes : 0; }
size_t mvsdf_net_status_offset(const mvsdf_net* net) { return net ? (size_t)net->plan.status_off : 0; }

int mvsdf_pack_weights(const mvsdf_net* net, const float* const* weight_v_host, const float* const* weight_g_host,
                       const float* const* bias_host, void* packed, void* stream) {
  if (!net || !weight_v_host || !bias_host || !packed) return fail(MVSDF_ERR_INVALID, "null argument");
  if (sm_count() <= 0) return fail(MVSDF_ERR_CUDA, "no CUDA device (the product path has no CPU fallback)");
  const NetPlan& p = net->plan;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* blob = static_cast<uint8_t*>(packed);
  float* scale = reinterpret_cast<float*>(blob + p.scale_area_off);
  float* bias_dst = reinterpret_cast<float*>(blob + p.bias_area_off);
  int* status = reinterpret_cast<int*>(blob + p.status_off);
  int rc0 = check_cuda(cudaMemsetAsync(status, 0, kStatusWords * 4, st), "memset status");
  if (rc0) return rc0;
  for (int s = 0; s < p.n_src_layers; ++s) {
    int rows = 0, cols = 0;
    for (int i = 0; i < p.n_layers; ++i)
      if (p.L[i].src_layer == s) {
        rows = p.L[i].out_dim;
        cols = p.L[i].in_dim;
      }
    if (!weight_v_host[s] || !bias_host[s]) return fail(MVSDF_ERR_INVALID, "null weight pointer for layer %d", s);
    const float* g = weight_g_host ? weight_g_host[s] : nullptr;
    note_launch(); row_scale_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(weight_v_host[s], g, rows, cols, scale + p.scale_off[s]);
  }
  for (int i = 0; i < p.n_layers; ++i) {
    const LayerPlan& L = p.L[i];
    const int total = L.m_tiles * kTileM * L.k_chunks * (kChunkK / 8);
    note_launch(); pack_layer_kernel<<<ceil_div(total, 256), 256, 0, st>>>(weight_v_host[L.src_layer], scale + p.scale_off[L.src_layer],
                                                            bias_host[L.src_layer], L, p.feat_size, blob, bias_dst, status);
  }
  return check_cuda(cudaGetLastError(), "pack_weights launch");
}

int mvsdf_sdf_forward(const mvsdf_net* net, const void* packed, const float* x, int64_t n, const int32_t* n_dev,
                      int head, float* out_sdf, float* out_full, void* stream) {
  if (head == MVSDF_HEAD_SDF_SCREEN)
    return mlp_sdf(net, packed, x, n, n_dev, MVSDF_HEAD_SDF_ONLY, out_sdf, nullptr, nullptr, false,
                   static_cast<cudaStream_t>(stream), true);
  return mlp_sdf(net, packed, x, n, n_dev, head, out_sdf, out_full, nullptr, false, static_cast<cudaStream_t>(stream));
}

int mvsdf_sdf_value_grad(const mvsdf_net* net, const void* packed, const float* x, int64_t n, const int32_t* n_dev,
                         int head, float* out_sdf, float* out_full, float* out_grad, void* stream) {
  return mlp_sdf(net, packed, x, n, n_dev, head, out_sdf, out_full, out_grad, true, static_cast<cudaStream_t>(stream));
}

int mvsdf_render_forward(const mvsdf_net* net, const void* packed, const float* points, const float* view_dirs,
                         const float* normals, const float* features, int64_t n, const int32_t* n_dev, float* out_rgb,
                         void* stream) {
  return mlp_render(net, packed, points, view_dirs, normals, features, 0, n, n_dev, out_rgb,
                    static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// C ABI, part 2: the ray tracer (rows a1, a2, a6-a10 of SURVEY.md section 8).
//
// RayTracing.forward (code/model/ray_tracing.py:27-98) re-expressed as a request server around the
// fused SDF-MLP tile kernel: small per-ray state-machine kernels emit "evaluate the SDF at
// (ray, t)" requests into a compacted device-side list, the persistent tcgen05 kernel evaluates
// whatever count the device counter holds, and the next state kernel consumes the results.
// There is no host synchronisation, no boolean-mask indexing and no .sum() test anywhere: the
// reference's global early-exit tests (ray_tracing.py:153,176) are pure optimisations, the per-ray
// semantics below are identical.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "../../include/mvsdf_b200.h"
#include "internal.h"

namespace mvsdf {

constexpr int kBlock = 256;
constexpr int kSteps = 100;          // n_steps must be 100 (confs/mvsdf_dtu.conf:56); checked at the ABI

enum RayFlag : uint32_t {
  F_HIT_SPHERE = 1u << 0,
  F_LIVE_S = 1u << 1,
  F_LIVE_E = 1u << 2,
  F_OVER_S = 1u << 3,
  F_OVER_E = 1u << 4,
  F_SAMPLER = 1u << 5,
  F_NET = 1u << 6,       // network_object_mask
  F_SECANT = 1u << 7,
  F_MINSDF = 1u << 8,
  F_SPEC_S = 1u << 9,    // the start / end march has speculative back-off requests pending (trace_backoff_spec_kernel)
  F_SPEC_E = 1u << 10,
};

struct RayState {
  // all arrays have one entry per ray
  float* dir;        // [R,3]
  float* acc_s;
  float* acc_e;
  float* min_dis;
  float* max_dis;
  float* next_s;
  float* next_e;
  float* cur_s;
  float* cur_e;
  int* slot_s;       // pending request slots (or -1)
  int* slot_e;
  uint32_t* flags;
  // secant state
  float* z_lo;
  float* z_hi;
  float* f_lo;
  float* f_hi;
  float* z;
  int* list;         // compacted ray list (sampler / min-sdf)
  int* list_pos;     // position of a ray inside the list (sampler batches)
};

constexpr int kNumCounters = 256;
// fixed slots of out_counters (include/mvsdf_b200.h): everything below kCtrScreened is an SDF request count
constexpr int kCtrScreened = 251, kCtrSamplerRays = 252, kCtrMinSdfRays = 253, kCtrRefined = 254, kCtrViolations = 255;

struct TraceCtx {
  RayState s;
  const float* cam;  // [B,3]
  float* req_pts;    // [cap,3]
  float* req_val;    // [cap]
  float* ref_pts;    // [cap,3]   prefilter: samples that need the exact evaluation
  float* ref_val;    // [cap]
  int* ref_src;      // [cap]     their index in req_val
  int* counters;     // [kNumCounters]  request counts per phase; the last four slots are fixed (see kCtr*)
  int* ref_counters; // [kNumCounters]  prefilter: refined samples per 100-sample batch
  int* pf_counts;    // [kNumCounters]  prefilter: per (batch, sample chunk) active rays / points of the chunked screening pass
  int* spec_counts;  // [kNumCounters]  sphere tracing: size of the speculative back-off request list of every iteration
  int* mix_counts;   // [kNumCounters]  mixed-precision sphere tracing: size of the screening list of every phase
  int* act_list[2];  // [batch_rays]    rays of the batch that still need their next sample chunk (ping-pong)
  int R, N;
  long long cap;
  float thr, clip, line_step;
  float margin;      // mixed-precision sphere tracing (mvsdf_tracer_params::trace_screen_margin), 0 = off
};

__device__ __forceinline__ float3 ray_point(const float* cam, const float* dir, float t) {
  // cam + t * dir with separate multiply and add, like the reference's elementwise torch ops
  return make_float3(__fadd_rn(cam[0], __fmul_rn(t, dir[0])), __fadd_rn(cam[1], __fmul_rn(t, dir[1])),
                     __fadd_rn(cam[2], __fmul_rn(t, dir[2])));
}

__device__ __forceinline__ int push_request(const TraceCtx& c, int counter, float3 p) {
  const int slot = atomicAdd(c.counters + counter, 1);
  c.req_pts[3 * (size_t)slot + 0] = p.x;
  c.req_pts[3 * (size_t)slot + 1] = p.y;
  c.req_pts[3 * (size_t)slot + 2] = p.z;
  return slot;
}

__device__ __forceinline__ float clampf(float v, float lim) { return fminf(fmaxf(v, -lim), lim); }

// Mixed-precision sphere tracing (trace_screen_margin > 0, off by default): a march position whose last step was long is
// first evaluated at screening precision (list ref_pts, owner (2 ray + end) in ref_src); trace_triage_kernel accepts the
// value when it is clearly outside the margin and queues the position for the exact evaluation otherwise.
__device__ __forceinline__ int push_march(const TraceCtx& c, int counter, int mix_ctr, bool far, int owner, float3 p) {
  if (!far) return push_request(c, counter, p);
  const int i = atomicAdd(c.mix_counts + mix_ctr, 1);
  c.ref_pts[3 * (size_t)i + 0] = p.x;
  c.ref_pts[3 * (size_t)i + 1] = p.y;
  c.ref_pts[3 * (size_t)i + 2] = p.z;
  c.ref_src[i] = owner;
  return -1;
}

// ---- a1 + a2 + tracer initialisation (rend_util.py:48-100, :141-162; ray_tracing.py:104-137)
__global__ void ray_setup_kernel(TraceCtx c, const float* __restrict__ uv, const float* __restrict__ pose,
                                 const float* __restrict__ intr, float* __restrict__ cam_out, float radius, int counter, int mix_ctr) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= c.R) return;
  const int b = r / c.N;
  const float* P = pose + 16 * b;
  const float* K = intr + 16 * b;
  const float fx = K[0], sk = K[1], cx = K[2], fy = K[5], cy = K[6];
  const float x = uv[2 * (size_t)r] + 0.5f, y = uv[2 * (size_t)r + 1] + 0.5f;
  // lift() with z = 1:  (x - cx + cy*sk/fy - sk*y/fy) / fx ,  (y - cy) / fy
  const float xl = __fdiv_rn(__fsub_rn(__fadd_rn(__fsub_rn(x, cx), __fdiv_rn(__fmul_rn(cy, sk), fy)),
                                       __fdiv_rn(__fmul_rn(sk, y), fy)), fx);
  const float yl = __fdiv_rn(__fsub_rn(y, cy), fy);
  float d[3], cam[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    cam[i] = P[4 * i + 3];
    const float w = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[4 * i], xl), __fmul_rn(P[4 * i + 1], yl)), P[4 * i + 2]),
                              P[4 * i + 3]);
    d[i] = __fsub_rn(w, cam[i]);
  }
  const float nrm = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]))), 1e-12f);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    d[i] = __fdiv_rn(d[i], nrm);
    c.s.dir[3 * (size_t)r + i] = d[i];
  }
  if (r % c.N == 0) {
    cam_out[3 * b] = cam[0];
    cam_out[3 * b + 1] = cam[1];
    cam_out[3 * b + 2] = cam[2];
  }
  const float dc = __fadd_rn(__fadd_rn(__fmul_rn(d[0], cam[0]), __fmul_rn(d[1], cam[1])), __fmul_rn(d[2], cam[2]));
  const float cn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(cam[0], cam[0]), __fmul_rn(cam[1], cam[1])), __fmul_rn(cam[2], cam[2])));
  const float under = __fsub_rn(__fmul_rn(dc, dc), __fsub_rn(__fmul_rn(cn, cn), radius * radius));
  uint32_t fl = 0;
  float t0 = 0.f, t1 = 0.f;
  if (under > 0.f) {
    const float rt = sqrtf(under);
    t0 = fmaxf(__fsub_rn(-rt, dc), 0.f);
    t1 = fmaxf(__fsub_rn(rt, dc), 0.f);
    fl = F_HIT_SPHERE | F_LIVE_S | F_LIVE_E;
  }
  c.s.acc_s[r] = t0;
  c.s.acc_e[r] = t1;
  c.s.min_dis[r] = t0;
  c.s.max_dis[r] = t1;
  c.s.next_s[r] = 0.f;
  c.s.next_e[r] = 0.f;
  c.s.cur_s[r] = 0.f;
  c.s.cur_e[r] = 0.f;
  int ss = -1, se = -1;
  if (fl) {
    ss = push_march(c, counter, mix_ctr, c.margin > 0.f, 2 * r, ray_point(cam, d, t0));
    se = push_march(c, counter, mix_ctr, c.margin > 0.f, 2 * r + 1, ray_point(cam, d, t1));
  }
  c.s.slot_s[r] = ss;
  c.s.slot_e[r] = se;
  c.s.flags[r] = fl;
}

__device__ __forceinline__ void collect(const TraceCtx& c, int r) {
  const int ss = c.s.slot_s[r], se = c.s.slot_e[r];
  if (ss >= 0) {
    c.s.next_s[r] = clampf(c.req_val[ss], c.clip);
    c.s.slot_s[r] = -1;
  }
  if (se >= 0) {
    c.s.next_e[r] = clampf(c.req_val[se], c.clip);
    c.s.slot_e[r] = -1;
  }
}

// The overshoot back-off loop (ray_tracing.py:173-191) steps a ray that crossed the surface back by 1/2, 1/4, 1/8 ... of its
// last step until the SDF is non-negative again, at most line_step_iters times.  Every candidate position is known before the
// first of these evaluations (the step length curr_sdf is fixed), so trace_backoff_spec_kernel queues ALL of a ray's
// candidates at once and the walk "take candidate k while the value is still negative" is replayed on the results here:
// identical arithmetic, identical values, one MLP launch per sphere-tracing iteration instead of line_step_iters (each
// launch, however small, streams the whole network through at least one SM pair).  Evaluations the reference would not have
// made are not counted: `used` goes into the iteration's E_trace counter.
__device__ __forceinline__ void resolve_backoff(const TraceCtx& c, int r, uint32_t& fl, int n_k, int count_ctr) {
  int used = 0;
  if (fl & F_SPEC_S) {
    float a = c.s.acc_s[r], v = c.s.next_s[r];
    const float cur = c.s.cur_s[r];
    const int first = c.s.slot_s[r];
    for (int k = 0; k < n_k && v < 0.f; ++k, ++used) {
      a = __fsub_rn(a, __fmul_rn((1.0f - c.line_step) / (float)(1 << k), cur));
      v = clampf(c.req_val[first + k], c.clip);
    }
    c.s.acc_s[r] = a;
    c.s.next_s[r] = v;
    c.s.slot_s[r] = -1;
  }
  if (fl & F_SPEC_E) {
    float a = c.s.acc_e[r], v = c.s.next_e[r];
    const float cur = c.s.cur_e[r];
    const int first = c.s.slot_e[r];
    for (int k = 0; k < n_k && v < 0.f; ++k, ++used) {
      a = __fadd_rn(a, __fmul_rn((1.0f - c.line_step) / (float)(1 << k), cur));
      v = clampf(c.req_val[first + k], c.clip);
    }
    c.s.acc_e[r] = a;
    c.s.next_e[r] = v;
    c.s.slot_e[r] = -1;
  }
  fl &= ~(F_SPEC_S | F_SPEC_E);
  if (used) atomicAdd(c.counters + count_ctr, used);
}

// top of the while-loop body (ray_tracing.py:139-171); `first`: no end-of-body update yet; `last`: iters == max
__global__ void trace_top_kernel(TraceCtx c, int first, int last, int counter, int n_k, int backoff_ctr, int mix_ctr) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= c.R) return;
  uint32_t fl = c.s.flags[r];
  if (!(fl & F_HIT_SPHERE)) return;
  if (fl & (F_SPEC_S | F_SPEC_E)) resolve_backoff(c, r, fl, n_k, backoff_ctr);
  collect(c, r);
  float acc_s = c.s.acc_s[r], acc_e = c.s.acc_e[r];
  if (!first) {   // end of the previous body (:193-194)
    if (!(acc_s < acc_e)) fl &= ~(F_LIVE_S | F_LIVE_E);
  }
  float cur_s = (fl & F_LIVE_S) ? c.s.next_s[r] : 0.f;
  if (cur_s <= c.thr) cur_s = 0.f;
  float cur_e = (fl & F_LIVE_E) ? c.s.next_e[r] : 0.f;
  if (cur_e <= c.thr) cur_e = 0.f;
  if (!(cur_s > c.thr)) fl &= ~F_LIVE_S;
  if (!(cur_e > c.thr)) fl &= ~F_LIVE_E;
  fl &= ~(F_OVER_S | F_OVER_E);
  c.s.cur_s[r] = cur_s;
  c.s.cur_e[r] = cur_e;
  if (!last) {
    acc_s = __fadd_rn(acc_s, cur_s);
    acc_e = __fsub_rn(acc_e, cur_e);
    c.s.acc_s[r] = acc_s;
    c.s.acc_e[r] = acc_e;
    const float* cam = c.cam + 3 * (r / c.N);
    const float* d = c.s.dir + 3 * (size_t)r;
    c.s.next_s[r] = 0.f;
    c.s.next_e[r] = 0.f;
    // mixed precision: the step just taken predicts the next value (it shrinks geometrically towards the surface)
    const float far_step = 2.0f * c.margin;
    if (fl & F_LIVE_S) c.s.slot_s[r] = push_march(c, counter, mix_ctr, c.margin > 0.f && cur_s > far_step, 2 * r, ray_point(cam, d, acc_s));
    if (fl & F_LIVE_E) c.s.slot_e[r] = push_march(c, counter, mix_ctr, c.margin > 0.f && cur_e > far_step, 2 * r + 1, ray_point(cam, d, acc_e));
  }
  c.s.flags[r] = fl;
}

// mixed-precision sphere tracing: screening values clearly outside the margin become the march's next SDF value, the rest
// join the exact request list of the same phase.  `accepted_ctr` is the E_trace slot of the accepted evaluations (the
// reference requests them too; the exact list only counts the others).
__global__ void trace_triage_kernel(TraceCtx c, int mix_ctr, int counter, int accepted_ctr) {
  const int n = c.mix_counts[mix_ctr];
  int accepted = 0, refined = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int owner = c.ref_src[i];
    const int r = owner >> 1;
    const float fs = clampf(c.ref_val[i], c.clip);
    if (fabsf(fs) > c.margin) {
      if (owner & 1) c.s.next_e[r] = fs;
      else c.s.next_s[r] = fs;
      ++accepted;
    } else {
      const int slot = push_request(c, counter, make_float3(c.ref_pts[3 * (size_t)i], c.ref_pts[3 * (size_t)i + 1], c.ref_pts[3 * (size_t)i + 2]));
      if (owner & 1) c.s.slot_e[r] = slot;
      else c.s.slot_s[r] = slot;
      ++refined;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    accepted += __shfl_xor_sync(0xffffffffu, accepted, o);
    refined += __shfl_xor_sync(0xffffffffu, refined, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (accepted) atomicAdd(c.counters + accepted_ctr, accepted);
    if (accepted + refined) atomicAdd(c.counters + kCtrScreened, accepted + refined);
    if (refined) atomicAdd(c.counters + kCtrRefined, refined);
  }
}

// one pass of the overshoot back-off loop (ray_tracing.py:173-191), sequential form: used when the request list cannot hold
// every candidate of every ray at once (2 R line_step_iters entries in the worst case)
__global__ void trace_backoff_kernel(TraceCtx c, int k, int counter) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= c.R) return;
  const uint32_t fl = c.s.flags[r];
  if (!(fl & F_HIT_SPHERE)) return;
  collect(c, r);
  const float back = (1.0f - c.line_step) / (float)(1 << k);
  const float* cam = c.cam + 3 * (r / c.N);
  const float* d = c.s.dir + 3 * (size_t)r;
  if (c.s.next_s[r] < 0.f) {
    const float a = __fsub_rn(c.s.acc_s[r], __fmul_rn(back, c.s.cur_s[r]));
    c.s.acc_s[r] = a;
    c.s.slot_s[r] = push_request(c, counter, ray_point(cam, d, a));
  }
  if (c.s.next_e[r] < 0.f) {
    const float a = __fadd_rn(c.s.acc_e[r], __fmul_rn(back, c.s.cur_e[r]));
    c.s.acc_e[r] = a;
    c.s.slot_e[r] = push_request(c, counter, ray_point(cam, d, a));
  }
}

// all candidates of the overshoot back-off loop of one sphere-tracing iteration (see resolve_backoff)
__global__ void trace_backoff_spec_kernel(TraceCtx c, int n_k, int* __restrict__ spec_count) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= c.R) return;
  uint32_t fl = c.s.flags[r];
  if (!(fl & F_HIT_SPHERE)) return;
  collect(c, r);
  const float* cam = c.cam + 3 * (r / c.N);
  const float* d = c.s.dir + 3 * (size_t)r;
  if (c.s.next_s[r] < 0.f) {
    const int first = atomicAdd(spec_count, n_k);
    float a = c.s.acc_s[r];
    const float cur = c.s.cur_s[r];
    for (int k = 0; k < n_k; ++k) {
      a = __fsub_rn(a, __fmul_rn((1.0f - c.line_step) / (float)(1 << k), cur));
      const float3 p = ray_point(cam, d, a);
      c.req_pts[3 * (size_t)(first + k)] = p.x;
      c.req_pts[3 * (size_t)(first + k) + 1] = p.y;
      c.req_pts[3 * (size_t)(first + k) + 2] = p.z;
    }
    c.s.slot_s[r] = first;
    fl |= F_SPEC_S;
  }
  if (c.s.next_e[r] < 0.f) {
    const int first = atomicAdd(spec_count, n_k);
    float a = c.s.acc_e[r];
    const float cur = c.s.cur_e[r];
    for (int k = 0; k < n_k; ++k) {
      a = __fadd_rn(a, __fmul_rn((1.0f - c.line_step) / (float)(1 << k), cur));
      const float3 p = ray_point(cam, d, a);
      c.req_pts[3 * (size_t)(first + k)] = p.x;
      c.req_pts[3 * (size_t)(first + k) + 1] = p.y;
      c.req_pts[3 * (size_t)(first + k) + 2] = p.z;
    }
    c.s.slot_e[r] = first;
    fl |= F_SPEC_E;
  }
  c.s.flags[r] = fl;
}

// after the loop: network_object_mask = acc_s < acc_e (:41); unconverged start rays go to the sampler (:44)
__global__ void trace_finish_kernel(TraceCtx c, int list_counter) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= c.R) return;
  uint32_t fl = c.s.flags[r];
  fl &= ~(F_NET | F_SAMPLER | F_SECANT | F_MINSDF);
  if (c.s.acc_s[r] < c.s.acc_e[r]) fl |= F_NET;
  c.s.list_pos[r] = -1;
  if (fl & F_LIVE_S) {
    fl |= F_SAMPLER;
    const int pos = atomicAdd(c.counters + list_counter, 1);
    c.s.list[pos] = r;
    c.s.list_pos[r] = pos;
  }
  c.s.flags[r] = fl;
}

// sampler batch: 100 samples on [acc_s, acc_e] for list entries [begin, begin+batch)  (:206-219)
__global__ void sampler_push_kernel(TraceCtx c, const float* __restrict__ lin, int list_counter, int begin, int batch,
                                    int counter) {
  const int total = min(max(c.counters[list_counter] - begin, 0), batch);
  if (blockIdx.x == 0 && threadIdx.x == 0) c.counters[counter] = total * kSteps;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over batch * 100
  if (idx >= total * kSteps) return;
  const int li = idx / kSteps, i = idx - li * kSteps;
  const int r = c.s.list[begin + li];
  const float lo = c.s.acc_s[r], hi = c.s.acc_e[r];
  const float t = __fadd_rn(lo, __fmul_rn(lin[i], __fsub_rn(hi, lo)));
  const float3 p = ray_point(c.cam + 3 * (r / c.N), c.s.dir + 3 * (size_t)r, t);
  c.req_pts[3 * (size_t)idx] = p.x;
  c.req_pts[3 * (size_t)idx + 1] = p.y;
  c.req_pts[3 * (size_t)idx + 2] = p.z;
  c.req_val[idx] = INFINITY;      // chunked screening: a sample that is never evaluated reads as "certainly positive, not the minimum"
}

// first sign change / arg-min selection and secant initialisation (:221-256); one thread per sampler ray
__global__ void sampler_select_kernel(TraceCtx c, const float* __restrict__ lin, const uint8_t* __restrict__ obj_mask,
                                      int training, int list_counter, int begin, int batch) {
  const int total = min(max(c.counters[list_counter] - begin, 0), batch);
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= total) return;
  const int r = c.s.list[begin + li];
  const float* f = c.req_val + (size_t)li * kSteps;
  const float lo = c.s.acc_s[r], hi = c.s.acc_e[r];
  int first_neg = -1, first_zero = -1, amin = 0;
  float fmin_ = f[0];
  for (int i = 0; i < kSteps; ++i) {
    const float v = f[i];
    if (v < 0.f && first_neg < 0) first_neg = i;
    if (v == 0.f && first_zero < 0) first_zero = i;
    if (v < fmin_) {
      fmin_ = v;
      amin = i;
    }
  }
  const int first = first_neg >= 0 ? first_neg : (first_zero >= 0 ? first_zero : kSteps - 1);
  const bool inside_net = f[first] < 0.f;
  const bool inside_true = obj_mask ? obj_mask[r] != 0 : true;
  uint32_t fl = c.s.flags[r];
  fl &= ~F_NET;
  if (inside_net) fl |= F_NET;
  const float span = __fsub_rn(hi, lo);
  int pick = first;
  if (!(inside_true && inside_net)) pick = amin;                     // P_out: sample of minimal SDF (:230-235)
  float t_out = __fadd_rn(lo, __fmul_rn(lin[pick], span));
  const bool sec = training ? (inside_net && inside_true) : inside_net;
  if (sec) {
    const int prev = first == 0 ? kSteps - 1 : first - 1;            // python's [-1] wrap (:248-249)
    const float z_hi = __fadd_rn(lo, __fmul_rn(lin[first], span));
    const float z_lo = __fadd_rn(lo, __fmul_rn(lin[prev], span));
    const float f_hi = f[first], f_lo = f[prev];
    const float z = __fadd_rn(__fdiv_rn(__fmul_rn(-f_lo, __fsub_rn(z_hi, z_lo)), __fsub_rn(f_hi, f_lo)), z_lo);
    c.s.z_lo[r] = z_lo;
    c.s.z_hi[r] = z_hi;
    c.s.f_lo[r] = f_lo;
    c.s.f_hi[r] = f_hi;
    c.s.z[r] = z;
    fl |= F_SECANT;
    t_out = z;
  }
  c.s.acc_s[r] = t_out;
  c.s.flags[r] = fl;
}

// ------------------------------------------------------------------ prefilter of the 100-sample stages
// req_val holds SCREENING values (error < tau / 2).  One thread per ray decides which samples the selection logic below
// can possibly look at with more than their certain sign / certain non-minimality, and queues those for the exact kernel:
//   * sign path (ray_sampler): with k = first certainly negative sample (v <= -tau) and j = first sample that is not
//     certainly positive (v >= tau), the first negative sample lies in [j, k]; it and its predecessor (python's [-1]
//     wrap for index 0) feed the secant -> refine [j-1, k].  Without a certainly negative sample: every uncertain sample
//     and its predecessor;
//   * arg-min path (P_out rays, rays without a negative sample, minimal_sdf_points): every sample within 2 tau of the
//     screening minimum.
// After the merge the selection kernels run unchanged; un-refined samples keep their screening values, of which only
// the (certain) sign and the (certain) fact that they are not the minimum is used.
// Unbiased guard: the refined samples are the near-surface / near-minimum ones, where the screening error is smallest, so
// checking only them would under-report it.  A pseudo-random 1/kAuditPeriod of the screened samples that are NOT refined
// (their screening value is trusted for its sign / non-minimality) is therefore evaluated exactly as well -- only to be
// compared: want[i] == 2 marks such an audit sample, its ref_src is stored negated and the merge kernel leaves req_val alone.
constexpr unsigned kAuditPeriod = 64;
__device__ __forceinline__ bool audit_pick(unsigned sample_index, unsigned salt) {
  unsigned h = (sample_index + salt * 0x9E3779B9u) * 2654435761u;
  h ^= h >> 15;
  h *= 2246822519u;
  h ^= h >> 13;
  return (h % kAuditPeriod) == 0u;
}

__device__ __forceinline__ void push_refine(const TraceCtx& c, int counter, int li, unsigned char* want) {
  const float* f = c.req_val + (size_t)li * kSteps;
  int n = 0;
  for (int i = 0; i < kSteps; ++i) {
    if (!want[i] && f[i] != INFINITY && audit_pick((unsigned)(li * kSteps + i), (unsigned)counter)) want[i] = 2;
    n += want[i] ? 1 : 0;
  }
  if (n == 0) return;
  int slot = atomicAdd(c.ref_counters + counter, n);
  for (int i = 0; i < kSteps; ++i)
    if (want[i]) {
      const size_t src = (size_t)li * kSteps + i;
      c.ref_pts[3 * (size_t)slot] = c.req_pts[3 * src];
      c.ref_pts[3 * (size_t)slot + 1] = c.req_pts[3 * src + 1];
      c.ref_pts[3 * (size_t)slot + 2] = c.req_pts[3 * src + 2];
      c.ref_src[slot] = want[i] == 2 ? -1 - (int)src : (int)src;
      ++slot;
    }
}

// ---- chunked screening pass of ray_sampler's 100 samples --------------------------------------------------------------
// The selection reads nothing beyond the first certainly negative sample k of a ray whose pixel is inside the true mask
// (first sign change <= k; the arg-min is only taken for P_out rays, :230-235).  The screening pass therefore walks the
// samples in chunks (chunk_begin) and drops a ray from the following chunks as soon as a chunk contains a value <= -tau;
// samples never evaluated keep +inf, which every consumer treats as "certainly positive and not the minimum".  Rays
// outside the true mask (training) and rays without a certain negative sample see all chunks.
constexpr int kNumChunks = 7;
// chunk boundaries: sphere tracing leaves acc_start just in front of the surface, so most crossings sit in the first samples
__host__ __device__ constexpr int chunk_begin(int c) {
  return c == 0 ? 0 : c == 1 ? 2 : c == 2 ? 5 : c == 3 ? 10 : c == 4 ? 20 : c == 5 ? 40 : c == 6 ? 70 : kSteps;
}
constexpr int kChunkMax = 30;

// points of chunk `chunk` of the active rays -> compact list (ref_pts, ref_src); act == nullptr: every ray of the batch
__global__ void chunk_gather_kernel(TraceCtx c, const int* __restrict__ act, const int* __restrict__ n_act_ptr, int list_counter,
                                    int begin, int batch, int chunk, int* __restrict__ n_pts_out) {
  const int total = min(max(c.counters[list_counter] - begin, 0), batch);
  const int n_act = act ? *n_act_ptr : total;
  const int c0 = chunk_begin(chunk), len = chunk_begin(chunk + 1) - c0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *n_pts_out = n_act * len;
    c.counters[kCtrScreened] += n_act * len;      // launches of one stream: no race
  }
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_act * len) return;
  const int a = idx / len, i = c0 + (idx - a * len);
  const int li = act ? act[a] : a;
  const size_t src = (size_t)li * kSteps + i;
  c.ref_pts[3 * (size_t)idx] = c.req_pts[3 * src];
  c.ref_pts[3 * (size_t)idx + 1] = c.req_pts[3 * src + 1];
  c.ref_pts[3 * (size_t)idx + 2] = c.req_pts[3 * src + 2];
}

// screening values of the chunk -> req_val; rays that still need the next chunk -> act_next
__global__ void chunk_decide_kernel(TraceCtx c, const uint8_t* __restrict__ obj_mask, float tau, const int* __restrict__ act,
                                    const int* __restrict__ n_act_ptr, int list_counter, int begin, int batch, int chunk,
                                    int* __restrict__ act_next, int* __restrict__ n_act_next) {
  const int total = min(max(c.counters[list_counter] - begin, 0), batch);
  const int n_act = act ? *n_act_ptr : total;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_act) return;
  const int li = act ? act[a] : a;
  const int r = c.s.list[begin + li];
  const bool inside_true = obj_mask ? obj_mask[r] != 0 : true;
  bool found = false;
  const int c0 = chunk_begin(chunk), len = chunk_begin(chunk + 1) - c0;
  for (int i = 0; i < len; ++i) {
    const float v = c.ref_val[(size_t)a * len + i];
    c.req_val[(size_t)li * kSteps + c0 + i] = v;
    if (v <= -tau) found = true;
  }
  if (act_next && !(found && inside_true)) act_next[atomicAdd(n_act_next, 1)] = li;
}

__global__ void prefilter_sampler_kernel(TraceCtx c, const uint8_t* __restrict__ obj_mask, float tau, int list_counter, int begin,
                                         int batch, int counter) {
  const int total = min(max(c.counters[list_counter] - begin, 0), batch);
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= total) return;
  const int r = c.s.list[begin + li];
  const float* f = c.req_val + (size_t)li * kSteps;
  unsigned char want[kSteps];
  int k = -1, j = -1;
  float fm = INFINITY;
  for (int i = 0; i < kSteps; ++i) {
    const float v = f[i];
    want[i] = 0;
    if (k < 0) {
      if (!(v >= tau) && j < 0) j = i;          // (NaN counts as uncertain)
      if (v <= -tau) k = i;
    }
    fm = fminf(fm, v);
  }
  if (k >= 0) {
    for (int i = max(j - 1, 0); i <= k; ++i) want[i] = 1;
    if (j == 0) want[kSteps - 1] = 1;
  } else {
    for (int i = 0; i < kSteps; ++i)
      if (!(fabsf(f[i]) >= tau)) {
        want[i] = 1;
        want[i == 0 ? kSteps - 1 : i - 1] = 1;
      }
  }
  const bool inside_true = obj_mask ? obj_mask[r] != 0 : true;
  if (k < 0 || !inside_true) {
    const float lim = fm + 2.f * tau;
    for (int i = 0; i < kSteps; ++i)
      if (!(f[i] >= lim)) want[i] = 1;
  }
  push_refine(c, counter, li, want);
}

__global__ void prefilter_argmin_kernel(TraceCtx c, float tau, int list_counter, int begin, int batch, int counter) {
  const int total = min(max(c.counters[list_counter] - begin, 0), batch);
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li == 0) c.counters[kCtrScreened] += total * kSteps;
  if (li >= total) return;
  const float* f = c.req_val + (size_t)li * kSteps;
  unsigned char want[kSteps];
  float fm = INFINITY;
  for (int i = 0; i < kSteps; ++i) fm = fminf(fm, f[i]);
  const float lim = fm + 2.f * tau;
  for (int i = 0; i < kSteps; ++i) want[i] = !(f[i] >= lim) ? 1 : 0;
  push_refine(c, counter, li, want);
}

// exact values replace the screening values; kCtrViolations counts samples (refined ones and the audited un-refined ones,
// see audit_pick) whose screening error exceeded tau / 2
__global__ void prefilter_merge_kernel(TraceCtx c, float tau, int counter) {
  const int n = c.ref_counters[counter];
  if (blockIdx.x == 0 && threadIdx.x == 0) c.counters[kCtrRefined] += n;      // launches of one stream: no race
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int raw = c.ref_src[i];
    const int src = raw < 0 ? -1 - raw : raw;
    const float exact = c.ref_val[i], lp = c.req_val[src];
    // soundness needs |error| < tau everywhere; the guard trips at tau/2 on refined samples (near the surface, where the
    // error is smallest) and at 3/4 tau on audited ones (anywhere along the ray)
    const float lim = raw < 0 ? 0.75f * tau : 0.5f * tau;
    if (lp != INFINITY && !(fabsf(exact - lp) <= lim)) atomicAdd(c.counters + kCtrViolations, 1);   // +inf: never screened
    if (raw >= 0) c.req_val[src] = exact;      // audit samples are only compared: the outputs do not depend on the audit
  }
}

// secant iterations (:260-278): `collect_prev` consumes f(z) of the previous request, `push` issues the next
__global__ void secant_kernel(TraceCtx c, int list_counter, int collect_prev, int push, int counter) {
  const int total = c.counters[list_counter];
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= total) return;
  const int r = c.s.list[li];
  if (!(c.s.flags[r] & F_SECANT)) return;
  float z = c.s.z[r];
  if (collect_prev) {
    const float fm = c.req_val[c.s.slot_s[r]];
    float z_lo = c.s.z_lo[r], z_hi = c.s.z_hi[r], f_lo = c.s.f_lo[r], f_hi = c.s.f_hi[r];
    if (fm > 0.f) {
      z_lo = z;
      f_lo = fm;
    }
    if (fm < 0.f) {
      z_hi = z;
      f_hi = fm;
    }
    z = __fadd_rn(__fdiv_rn(__fmul_rn(-f_lo, __fsub_rn(z_hi, z_lo)), __fsub_rn(f_hi, f_lo)), z_lo);
    c.s.z_lo[r] = z_lo;
    c.s.z_hi[r] = z_hi;
    c.s.f_lo[r] = f_lo;
    c.s.f_hi[r] = f_hi;
    c.s.z[r] = z;
    c.s.acc_s[r] = z;
  }
  if (push) c.s.slot_s[r] = push_request(c, counter, ray_point(c.cam + 3 * (r / c.N), c.s.dir + 3 * (size_t)r, z));
}

// training only (:73-94): closest approach for rays that miss the sphere; list the rays for minimal_sdf_points
__global__ void minsdf_prepare_kernel(TraceCtx c, const uint8_t* __restrict__ obj_mask, int list_counter) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= c.R) return;
  uint32_t fl = c.s.flags[r];
  const bool obj = obj_mask ? obj_mask[r] != 0 : true;
  const bool net = fl & F_NET, samp = fl & F_SAMPLER, hs = fl & F_HIT_SPHERE;
  const bool in_m = !net && obj && !samp;
  const bool out_m = !obj && !samp;
  if (!(in_m || out_m)) return;
  if (!hs) {
    const float* cam = c.cam + 3 * (r / c.N);
    const float* d = c.s.dir + 3 * (size_t)r;
    c.s.acc_s[r] = -__fadd_rn(__fadd_rn(__fmul_rn(d[0], cam[0]), __fmul_rn(d[1], cam[1])), __fmul_rn(d[2], cam[2]));
    return;
  }
  if (net && out_m) c.s.min_dis[r] = c.s.acc_s[r];
  fl |= F_MINSDF;
  c.s.flags[r] = fl;
  const int pos = atomicAdd(c.counters + list_counter, 1);
  c.s.list[pos] = r;
}

__global__ void minsdf_push_kernel(TraceCtx c, const float* __restrict__ steps01, int list_counter, int begin, int batch,
                                   int counter) {
  const int total = min(max(c.counters[list_counter] - begin, 0), batch);
  if (blockIdx.x == 0 && threadIdx.x == 0) c.counters[counter] = total * kSteps;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total * kSteps) return;
  const int li = idx / kSteps, i = idx - li * kSteps;
  const int r = c.s.list[begin + li];
  const float lo = c.s.min_dis[r], hi = c.s.max_dis[r];
  const float t = __fadd_rn(__fmul_rn(steps01[i], __fsub_rn(hi, lo)), lo);
  const float3 p = ray_point(c.cam + 3 * (r / c.N), c.s.dir + 3 * (size_t)r, t);
  c.req_pts[3 * (size_t)idx] = p.x;
  c.req_pts[3 * (size_t)idx + 1] = p.y;
  c.req_pts[3 * (size_t)idx + 2] = p.z;
}

__global__ void minsdf_select_kernel(TraceCtx c, const float* __restrict__ steps01, int list_counter, int begin, int batch) {
  const int total = min(max(c.counters[list_counter] - begin, 0), batch);
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= total) return;
  const int r = c.s.list[begin + li];
  const float* f = c.req_val + (size_t)li * kSteps;
  int amin = 0;
  float fm = f[0];
  for (int i = 1; i < kSteps; ++i)
    if (f[i] < fm) {
      fm = f[i];
      amin = i;
    }
  const float lo = c.s.min_dis[r], hi = c.s.max_dis[r];
  c.s.acc_s[r] = __fadd_rn(__fmul_rn(steps01[amin], __fsub_rn(hi, lo)), lo);
}

__global__ void trace_output_kernel(TraceCtx c, float* __restrict__ dists, uint8_t* __restrict__ net_mask,
                                    float* __restrict__ points) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= c.R) return;
  const float t = c.s.acc_s[r];
  dists[r] = t;
  net_mask[r] = (c.s.flags[r] & F_NET) ? 1 : 0;
  if (points) {   // IDRNetwork.forward recomputes points = cam_loc + dists * ray_dirs (:200)
    const float3 p = ray_point(c.cam + 3 * (r / c.N), c.s.dir + 3 * (size_t)r, t);
    points[3 * (size_t)r] = p.x;
    points[3 * (size_t)r + 1] = p.y;
    points[3 * (size_t)r + 2] = p.z;
  }
}


struct WorkspaceLayout {
  size_t off_counters, off_cam, off_f[9], off_i[4], off_flags, off_req_pts, off_req_val, off_ref_pts, off_ref_val, off_ref_src,
      off_act[2], total;
  long long cap;
  int batch_rays;
};

static WorkspaceLayout layout_for(int64_t R, int B, int batch_rays) {
  WorkspaceLayout w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) / 256 * 256;
    return o;
  };
  w.off_counters = take(5 * kNumCounters * 4);
  w.off_cam = take((size_t)B * 3 * 4);
  for (int i = 0; i < 9; ++i) w.off_f[i] = take((size_t)R * 4);       // acc_s acc_e min max next_s next_e cur_s cur_e z
  for (int i = 0; i < 4; ++i) w.off_i[i] = take((size_t)R * 4);       // slot_s slot_e list list_pos
  w.off_flags = take((size_t)R * 4);
  w.batch_rays = (int)std::min<int64_t>(R, batch_rays);
  w.cap = std::max<long long>(2 * R, (long long)w.batch_rays * kSteps);
  w.off_req_pts = take((size_t)w.cap * 12);
  w.off_req_val = take((size_t)w.cap * 4);
  w.off_ref_pts = take((size_t)w.cap * 12);      // prefilter: worst case every sample is refined
  w.off_ref_val = take((size_t)w.cap * 4);
  w.off_ref_src = take((size_t)w.cap * 4);
  for (int i = 0; i < 2; ++i) w.off_act[i] = take((size_t)w.batch_rays * 4);
  w.total = off;
  return w;
}

}  // namespace mvsdf

using namespace mvsdf;

extern "C" {

size_t mvsdf_trace_workspace_bytes(int64_t n_rays, int n_images) {
  // 4 extra float arrays for the secant state live in the tail
  return layout_for(n_rays, n_images, 1 << 18).total + (size_t)n_rays * 4 * 4 + 1024;
}

int mvsdf_trace(const mvsdf_net* net, const void* packed, const float* uv, const float* pose, const float* intrinsics,
                const uint8_t* object_mask, const mvsdf_tracer_params* prm, int n_images, int n_pixels, int training,
                const float* linspace100, const float* steps01, size_t workspace_bytes, void* workspace,
                float* out_ray_dirs, float* out_cam_loc, float* out_dists, uint8_t* out_net_mask, float* out_points,
                int32_t* out_counters, void* stream) {
  if (!net || !packed || !uv || !pose || !intrinsics || !prm || !workspace || !out_ray_dirs || !out_dists ||
      !out_net_mask || !linspace100)
    return fail(MVSDF_ERR_INVALID, "mvsdf_trace: null argument");
  if (prm->n_steps != kSteps) return fail(MVSDF_ERR_INVALID, "mvsdf_trace: n_steps must be %d", kSteps);
  if (prm->line_step_iters < 0 || prm->line_step_iters > 8 || prm->sphere_tracing_iters < 0 ||
      prm->sphere_tracing_iters > 64 || prm->n_secant_steps < 0 || prm->n_secant_steps > 16)
    return fail(MVSDF_ERR_INVALID, "mvsdf_trace: iteration counts out of range");
  if (training && !prm->skip_min_sdf && !steps01)
    return fail(MVSDF_ERR_INVALID, "mvsdf_trace: steps01 is required in training mode (CPU-generator samples, ray_tracing.py:287)");
  const int64_t R = (int64_t)n_images * n_pixels;
  if (R <= 0 || R > (1ll << 30)) return fail(MVSDF_ERR_INVALID, "mvsdf_trace: bad ray count");
  {
    // every request phase owns one device counter below kCtrScreened: refuse configurations that would run into the
    // fixed slots (and from there into the prefilter's counters) BEFORE anything is launched
    const int64_t batches = (R + (1 << 18) - 1) / (1 << 18);
    const int64_t need = 1 + (int64_t)prm->sphere_tracing_iters * (1 + prm->line_step_iters) + batches + prm->n_secant_steps +
                         ((training && !prm->skip_min_sdf) ? batches : 0) +
                         (prm->trace_screen_margin > 0.f ? 1 + prm->sphere_tracing_iters : 0);
    if (need >= kCtrScreened)
      return fail(MVSDF_ERR_INVALID,
                  "mvsdf_trace: %lld request phases (1 + sphere_tracing_iters*(1+line_step_iters) + sampler/min-sdf batches + "
                  "secant steps) exceed the %d request counters",
                  (long long)need, kCtrScreened);
  }
  if (workspace_bytes < mvsdf_trace_workspace_bytes(R, n_images))
    return fail(MVSDF_ERR_WORKSPACE, "mvsdf_trace: workspace too small (%zu < %zu)", workspace_bytes,
                mvsdf_trace_workspace_bytes(R, n_images));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const WorkspaceLayout w = layout_for(R, n_images, 1 << 18);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  TraceCtx c{};
  c.s.dir = out_ray_dirs;
  float** fa[] = {&c.s.acc_s, &c.s.acc_e, &c.s.min_dis, &c.s.max_dis, &c.s.next_s, &c.s.next_e, &c.s.cur_s, &c.s.cur_e, &c.s.z};
  for (int i = 0; i < 9; ++i) *fa[i] = reinterpret_cast<float*>(ws + w.off_f[i]);
  c.s.slot_s = reinterpret_cast<int*>(ws + w.off_i[0]);
  c.s.slot_e = reinterpret_cast<int*>(ws + w.off_i[1]);
  c.s.list = reinterpret_cast<int*>(ws + w.off_i[2]);
  c.s.list_pos = reinterpret_cast<int*>(ws + w.off_i[3]);
  c.s.flags = reinterpret_cast<uint32_t*>(ws + w.off_flags);
  float* tail = reinterpret_cast<float*>(ws + w.total);
  c.s.z_lo = tail;
  c.s.z_hi = tail + R;
  c.s.f_lo = tail + 2 * R;
  c.s.f_hi = tail + 3 * R;
  float* cam = out_cam_loc ? out_cam_loc : reinterpret_cast<float*>(ws + w.off_cam);
  c.cam = cam;
  c.req_pts = reinterpret_cast<float*>(ws + w.off_req_pts);
  c.req_val = reinterpret_cast<float*>(ws + w.off_req_val);
  c.ref_pts = reinterpret_cast<float*>(ws + w.off_ref_pts);
  c.ref_val = reinterpret_cast<float*>(ws + w.off_ref_val);
  c.ref_src = reinterpret_cast<int*>(ws + w.off_ref_src);
  c.counters = reinterpret_cast<int*>(ws + w.off_counters);
  c.ref_counters = c.counters + kNumCounters;
  c.pf_counts = c.counters + 2 * kNumCounters;
  c.spec_counts = c.counters + 3 * kNumCounters;
  c.mix_counts = c.counters + 4 * kNumCounters;
  for (int i = 0; i < 2; ++i) c.act_list[i] = reinterpret_cast<int*>(ws + w.off_act[i]);
  c.R = (int)R;
  c.N = n_pixels;
  c.cap = w.cap;
  c.thr = prm->sdf_threshold;
  c.clip = prm->dist_clip;
  c.line_step = prm->line_search_step;
  c.margin = prm->trace_screen_margin > 0.f ? prm->trace_screen_margin : 0.f;

  int rc = check_cuda(cudaMemsetAsync(c.counters, 0, 5 * kNumCounters * 4, st), "memset counters");
  if (rc) return rc;
  const int grid_r = (int)((R + kBlock - 1) / kBlock);
  int ctr = 0;   // every request phase uses its own counter: no resets, no host round trips
  auto eval = [&](int counter) {
    return mlp_sdf(net, packed, c.req_pts, 0, c.counters + counter, MVSDF_HEAD_SDF_ONLY, c.req_val, nullptr, nullptr,
                   false, st);
  };
  // 100-sample stages with the prefilter: screening pass over all samples, exact pass over the undecidable ones
  // the prefilter keeps two device counters per (sampler batch, sample chunk); beyond that many batches (R > 6.5 M rays
  // per call) the call simply runs without it -- same results, by construction
  const int n_batches_pf = (int)((R + w.batch_rays - 1) / w.batch_rays);
  const float tau = (n_batches_pf * kNumChunks * 2 > kNumCounters) ? 0.f : prm->prefilter_tau;
  int ref_ctr = 0;
  auto eval_screened = [&](int counter, auto&& launch_select_candidates) {
    if (ref_ctr >= kNumCounters) return fail(MVSDF_ERR_INVALID, "mvsdf_trace: too many prefilter batches");
    int e = mlp_sdf(net, packed, c.req_pts, 0, c.counters + counter, MVSDF_HEAD_SDF_ONLY, c.req_val, nullptr, nullptr, false, st,
                    true);
    if (e) return e;
    launch_select_candidates(ref_ctr);
    e = mlp_sdf(net, packed, c.ref_pts, 0, c.ref_counters + ref_ctr, MVSDF_HEAD_SDF_ONLY, c.ref_val, nullptr, nullptr, false, st);
    if (e) return e;
    note_launch(); prefilter_merge_kernel<<<sm_count() * 4, kBlock, 0, st>>>(c, tau, ref_ctr);
    ++ref_ctr;
    return (int)MVSDF_OK;
  };
  // sphere-tracing phase: (mixed precision only: screening pass over the far positions + triage, then) the exact pass
  int mix_ctr = 0;
  auto eval_march = [&](int counter) {
    if (c.margin > 0.f) {
      int e = mlp_sdf(net, packed, c.ref_pts, 0, c.mix_counts + mix_ctr, MVSDF_HEAD_SDF_ONLY, c.ref_val, nullptr, nullptr, false, st, true);
      if (e) return e;
      note_launch(); trace_triage_kernel<<<sm_count() * 4, kBlock, 0, st>>>(c, mix_ctr, counter, ctr++);
      ++mix_ctr;
    }
    return eval(counter);
  };
  note_launch(); ray_setup_kernel<<<grid_r, kBlock, 0, st>>>(c, uv, pose, intrinsics, cam, prm->object_bounding_sphere, ctr, mix_ctr);
  if ((rc = eval_march(ctr++))) return rc;
  const int n_k = prm->line_step_iters;
  // worst case every ray overshoots at both ends: 2 R n_k candidates must fit the request list; MVSDF_SPEC_BACKOFF=0 forces
  // the sequential form (A/B)
  // The speculative form trades evaluations for launches: it evaluates all n_k candidates of an overshooting ray where the
  // walk uses 1-2 of them (measured at cfg2: +3.1 exact evaluations per ray, 576 vs 558 ms per step), and saves n_k - 1
  // dependent launches per iteration (~50 us each: one tile's latency).  Break-even is near 1e5 rays; below 2^17 rays
  // (every training batch, cfg1, cfg3) the launches dominate.  MVSDF_SPEC_BACKOFF=0 / =1 forces one form (A/B).
  const char* spec_e = getenv("MVSDF_SPEC_BACKOFF");
  const bool spec_fits = 2ll * R * n_k <= c.cap;
  const bool speculative = spec_fits && (spec_e ? atoi(spec_e) != 0 : R <= (1ll << 17));
  int backoff_ctr = 0;      // E_trace slot of the previous iteration's back-off evaluations (filled by the next trace_top_kernel)
  for (int it = 0; it < prm->sphere_tracing_iters; ++it) {
    note_launch(); trace_top_kernel<<<grid_r, kBlock, 0, st>>>(c, it == 0, 0, ctr, n_k, backoff_ctr, mix_ctr);
    if ((rc = eval_march(ctr++))) return rc;
    if (n_k > 0 && speculative) {
      // the whole back-off loop of this iteration in one launch (resolve_backoff)
      note_launch(); trace_backoff_spec_kernel<<<grid_r, kBlock, 0, st>>>(c, n_k, c.spec_counts + it);
      if ((rc = mlp_sdf(net, packed, c.req_pts, 0, c.spec_counts + it, MVSDF_HEAD_SDF_ONLY, c.req_val, nullptr, nullptr, false, st)))
        return rc;
      backoff_ctr = ctr++;
    } else {
      for (int k = 0; k < n_k; ++k) {
        note_launch(); trace_backoff_kernel<<<grid_r, kBlock, 0, st>>>(c, k, ctr);
        if ((rc = eval(ctr++))) return rc;
      }
    }
  }
  note_launch(); trace_top_kernel<<<grid_r, kBlock, 0, st>>>(c, prm->sphere_tracing_iters == 0, 1, ctr, n_k, backoff_ctr, mix_ctr);
  const int list_ctr = kCtrSamplerRays;
  note_launch(); trace_finish_kernel<<<grid_r, kBlock, 0, st>>>(c, list_ctr);
  // sampler in batches of batch_rays rays (worst case: every ray unconverged)
  const int n_batches = (int)((R + w.batch_rays - 1) / w.batch_rays);
  for (int b = 0; b < n_batches; ++b) {
    const int begin = b * w.batch_rays;
    const long long items = (long long)w.batch_rays * kSteps;
    note_launch(); sampler_push_kernel<<<(int)((items + kBlock - 1) / kBlock), kBlock, 0, st>>>(c, linspace100, list_ctr, begin,
                                                                                w.batch_rays, ctr);
    if (tau > 0.f) {
      // chunked screening pass (see chunk_gather_kernel), then the exact pass over the undecidable samples
      if ((b + 1) * kNumChunks * 2 > kNumCounters || ref_ctr >= kNumCounters)
        return fail(MVSDF_ERR_INVALID, "mvsdf_trace: too many prefilter batches");
      const int chunk_grid = (int)(((long long)w.batch_rays * kChunkMax + kBlock - 1) / kBlock);
      for (int ch = 0; ch < kNumChunks; ++ch) {
        int* cnt = c.pf_counts + (b * kNumChunks + ch) * 2;             // [0] active rays of this chunk, [1] its points
        const int* act = ch == 0 ? nullptr : c.act_list[ch & 1];
        int* act_next = ch + 1 < kNumChunks ? c.act_list[(ch + 1) & 1] : nullptr;
        note_launch(); chunk_gather_kernel<<<chunk_grid, kBlock, 0, st>>>(c, act, cnt, list_ctr, begin, w.batch_rays, ch, cnt + 1);
        if ((rc = mlp_sdf(net, packed, c.ref_pts, 0, cnt + 1, MVSDF_HEAD_SDF_ONLY, c.ref_val, nullptr, nullptr, false, st, true)))
          return rc;
        note_launch(); chunk_decide_kernel<<<(w.batch_rays + kBlock - 1) / kBlock, kBlock, 0, st>>>(
            c, object_mask, tau, act, cnt, list_ctr, begin, w.batch_rays, ch, act_next, cnt + 2);
      }
      note_launch(); prefilter_sampler_kernel<<<(w.batch_rays + kBlock - 1) / kBlock, kBlock, 0, st>>>(
          c, object_mask, tau, list_ctr, begin, w.batch_rays, ref_ctr);
      if ((rc = mlp_sdf(net, packed, c.ref_pts, 0, c.ref_counters + ref_ctr, MVSDF_HEAD_SDF_ONLY, c.ref_val, nullptr, nullptr, false,
                        st)))
        return rc;
      note_launch(); prefilter_merge_kernel<<<sm_count() * 4, kBlock, 0, st>>>(c, tau, ref_ctr);
      ++ref_ctr;
    } else if ((rc = eval(ctr))) {
      return rc;
    }
    note_launch(); sampler_select_kernel<<<(w.batch_rays + kBlock - 1) / kBlock, kBlock, 0, st>>>(c, linspace100, object_mask, training,
                                                                                   list_ctr, begin, w.batch_rays);
    ctr++;
    if (ctr >= kCtrScreened - 20) return fail(MVSDF_ERR_INVALID, "mvsdf_trace: too many sampler batches");
  }
  for (int i = 0; i <= prm->n_secant_steps; ++i) {
    const int push = i < prm->n_secant_steps;
    note_launch(); secant_kernel<<<grid_r, kBlock, 0, st>>>(c, list_ctr, i > 0, push, ctr);
    if (push) {
      if ((rc = eval(ctr++))) return rc;
    }
  }
  if (training) {
    const int ml_ctr = kCtrMinSdfRays;
    note_launch(); minsdf_prepare_kernel<<<grid_r, kBlock, 0, st>>>(c, object_mask, ml_ctr);
    if (!prm->skip_min_sdf) {
      for (int b = 0; b < n_batches; ++b) {
        const int begin = b * w.batch_rays;
        const long long items = (long long)w.batch_rays * kSteps;
        note_launch(); minsdf_push_kernel<<<(int)((items + kBlock - 1) / kBlock), kBlock, 0, st>>>(c, steps01, ml_ctr, begin,
                                                                                   w.batch_rays, ctr);
        if (tau > 0.f) {
          rc = eval_screened(ctr, [&](int rc_ctr) {
            note_launch(); prefilter_argmin_kernel<<<(w.batch_rays + kBlock - 1) / kBlock, kBlock, 0, st>>>(
                c, tau, ml_ctr, begin, w.batch_rays, rc_ctr);
          });
          if (rc) return rc;
        } else if ((rc = eval(ctr))) {
          return rc;
        }
        note_launch(); minsdf_select_kernel<<<(w.batch_rays + kBlock - 1) / kBlock, kBlock, 0, st>>>(c, steps01, ml_ctr, begin,
                                                                                      w.batch_rays);
        ctr++;
        if (ctr >= kCtrScreened) return fail(MVSDF_ERR_INVALID, "mvsdf_trace: too many min-sdf batches");
      }
    }
  }
  note_launch(); trace_output_kernel<<<grid_r, kBlock, 0, st>>>(c, out_dists, out_net_mask, out_points);
  if (out_counters)
    rc = check_cuda(cudaMemcpyAsync(out_counters, c.counters, kNumCounters * 4, cudaMemcpyDeviceToDevice, st),
                    "copy counters");
  if (rc) return rc;
  return check_cuda(cudaGetLastError(), "mvsdf_trace launches");
}

}  // extern "C"

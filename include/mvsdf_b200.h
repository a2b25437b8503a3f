/*
 * mvsdf_b200 -- C ABI of the B200-native MVSDF differentiable-rendering hot path.
 *
 * The reference (jzhangbs/MVSDF @ a5399816) has no operator / plugin / FFI layer for this path:
 * it is plain nn.Modules and the tracer receives the SDF as a Python closure
 * (code/model/implicit_differentiable_renderer.py:194).  This header is the boundary created one
 * level up, at IDRNetwork.forward / RayTracing.forward / IDRLoss.get_feat_loss_corr: every entry
 * point names the reference interface it replaces.  INTEGRATION.md shows the ctypes binding a
 * maintainer would add on the reference side.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the parameter name ends in _host;
 *   - float = IEEE fp32, masks = uint8 (0/1, the storage of torch.bool), counts = int32;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*), never
 *     synchronises the device, never allocates device memory and never takes ownership;
 *   - return value: 0 on success, negative on error; mvsdf_last_error() returns the message of the
 *     last failing call of the calling thread;
 *   - there is no CPU fallback: without a CUDA device every compute call fails with MVSDF_ERR_CUDA.
 */
#ifndef MVSDF_B200_H_
#define MVSDF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVSDF_OK 0
#define MVSDF_ERR_INVALID (-1)
#define MVSDF_ERR_CUDA (-2)
#define MVSDF_ERR_WORKSPACE (-3)

#define MVSDF_HEAD_SDF_ONLY 0
#define MVSDF_HEAD_FULL 1
#define MVSDF_HEAD_SDF_SCREEN 2   /* mvsdf_sdf_forward only: SDF column at screening precision (single fp16 product, ~1e-3
                                     absolute error; the tracer's sampler prefilter).  Batches too small for the CTA-pair
                                     kernel are evaluated exactly. */

typedef struct mvsdf_net mvsdf_net;   /* host-side layout plan of one packed MLP; owns no device memory */

int mvsdf_abi_version(void);
const char* mvsdf_last_error(void);
/* Instrumentation (no reference counterpart): cumulative number of CUDA kernels this library has launched, and
 * optional CUDA-event timing of the MLP tile launches on their own stream (kinds: 0 SDF-only head, 1 full head,
 * 2 value+gradient, 3 rendering net, 4 SDF-only head at screening precision, 5 backward sweep, 6 dW GEMM of the backward;
 * both arrays hold MVSDF_PROFILE_KINDS entries).  mvsdf_profile_collect synchronises on the recorded events. */
#define MVSDF_PROFILE_KINDS 8
long long mvsdf_launch_count(void);
void mvsdf_launch_count_add(long long n);   /* a caller that replays a captured CUDA graph of library launches reports them */
void mvsdf_profile_enable(int on);
int mvsdf_profile_collect(float* ms_by_kind_host, int* launches_by_kind_host);

/* ---- network plans ------------------------------------------------------------------------------
 * ImplicitNetwork.__init__ (implicit_differentiable_renderer.py:19-75): dims [3+6*n_freqs] + [width]*n_hidden
 * + [2+feature_size], skip connection into layer `skip_layer`, softplus(beta=100), weight_norm on every layer. */
mvsdf_net* mvsdf_sdf_net_create(int width, int n_hidden, int skip_layer, int n_freqs, int feature_size);
/* RenderingNetwork.__init__ (:109-143), mode='idr': dims [9 + feature_size + 6*n_freqs_view] + [width]*n_hidden + [3]. */
mvsdf_net* mvsdf_render_net_create(int width, int n_hidden, int n_freqs_view, int feature_size);
void mvsdf_net_destroy(mvsdf_net* net);
int mvsdf_net_num_layers(const mvsdf_net* net);          /* number of source Linear layers */
size_t mvsdf_net_packed_bytes(const mvsdf_net* net);      /* size of the packed blob the caller must allocate */
/* Numeric-range monitors.  Weights and activations are stored as fp16 hi/lo pairs of 64*x, so |W| or an activation beyond
 * ~1023 cannot be represented.  int32 words at packed + mvsdf_net_status_offset(net), zeroed by every mvsdf_pack_weights:
 *   [MVSDF_STATUS_PACK_RANGE]  packed weight elements outside the representable range (or non-finite),
 *   [MVSDF_STATUS_NONFINITE]   non-finite values written by the heads of the MLP kernels since the last pack (an overflowed
 *                              activation turns into NaN in the hi/lo split and surfaces there).
 * A caller reads them at its next host synchronisation and must treat non-zero as an error (B200IDRNetwork raises). */
#define MVSDF_STATUS_WORDS 4
#define MVSDF_STATUS_PACK_RANGE 0
#define MVSDF_STATUS_NONFINITE 1
size_t mvsdf_net_status_offset(const mvsdf_net* net);

/* Fold weight_norm (W = g * v / ||v||_row, nn.utils.weight_norm dim=0; :70-71, :137-138), scale, split into
 * fp16 hi/lo tiles and write the packed blob.  weight_v_host[l] etc. are HOST arrays of device pointers, one
 * per source layer lin{l}; weight_g_host[l] may be NULL (plain Linear).  Runs every forward (weights change
 * every optimiser step).  Replaces torch's _weight_norm pre-forward hook. */
int mvsdf_pack_weights(const mvsdf_net* net, const float* const* weight_v_host, const float* const* weight_g_host,
                       const float* const* bias_host, void* packed, void* stream);

/* ---- ImplicitNetwork.forward (:77-94) on arbitrary points ------------------------------------------
 * x [n,3].  head = MVSDF_HEAD_SDF_ONLY: out_sdf [n] (the tracer's `sdf=lambda x: net(x)[:,0]`, :194);
 * head = MVSDF_HEAD_FULL: out_full [n, 2+F] (columns: sdf, surface-indicator logit, features), out_sdf optional.
 * n_dev, when non-NULL, is a device int32 that overrides n (count produced on the device, no host sync). */
int mvsdf_sdf_forward(const mvsdf_net* net, const void* packed, const float* x, int64_t n, const int32_t* n_dev,
                      int head, float* out_sdf, float* out_full, void* stream);
/* ImplicitNetwork.forward + .gradient (:96-107) in one pass (forward-mode tangents): out_grad [n,3] = d sdf / d x. */
int mvsdf_sdf_value_grad(const mvsdf_net* net, const void* packed, const float* x, int64_t n, const int32_t* n_dev,
                         int head, float* out_sdf, float* out_full, float* out_grad, void* stream);
/* RenderingNetwork.forward (:145-167): rgb [n,3] = tanh(MLP(cat[points, PE(view), normals, features])). */
int mvsdf_render_forward(const mvsdf_net* net, const void* packed, const float* points, const float* view_dirs,
                         const float* normals, const float* features, int64_t n, const int32_t* n_dev, float* out_rgb,
                         void* stream);

/* ---- RayTracing (code/model/ray_tracing.py) ---------------------------------------------------------
 * Constructor arguments of RayTracing.__init__ (:6-25; values in confs/mvsdf_dtu.conf:49-58) plus the two
 * hard-coded switches of sphere_tracing (:127-131): dist_clip = 0.5 (0.05 and 40 iterations under IDR_RENDER). */
typedef struct mvsdf_tracer_params {
  float object_bounding_sphere;
  float sdf_threshold;
  float line_search_step;
  float dist_clip;
  int line_step_iters;
  int sphere_tracing_iters;
  int n_steps;            /* must be 100 */
  int n_secant_steps;
  int skip_min_sdf;       /* training only: drop minimal_sdf_points (:280-308), whose outputs no MVSDF loss reads */
  float prefilter_tau;    /* 0 = off.  > 0: the 100-sample stages (ray_sampler :206-235, minimal_sdf_points :287-305) first
                             evaluate all samples at screening precision and re-evaluate exactly only the samples whose sign
                             (|sdf| < tau), secant end points or arg-min rank the screening value cannot decide; the
                             screening pass of ray_sampler walks the samples in chunks and stops at the first chunk with a
                             certainly negative sample (nothing behind it is read by :221-249).  Results
                             are bit-identical to tau = 0 as long as the screening error stays below tau; out_counters
                             [MVSDF_CTR_VIOLATIONS] counts refined samples whose screening error exceeded tau / 2 -- a
                             caller that sees it non-zero repeats the call with tau = 0 (B200IDRNetwork does). */
  float trace_screen_margin; /* 0 = off (default: every sphere-tracing value is exact and the march follows the reference's
                             path step for step).  > 0: experiment of profiles/r02/exp_mixed_trace_*: a march position whose
                             last step exceeded 2 * margin is evaluated at screening precision first and the value is used as
                             the step when |value| > margin, otherwise the position is evaluated exactly.  Convergence,
                             overshoot back-off, sampler and secant decisions always see exact values, but the march no
                             longer follows the reference's path: distances agree to the convergence threshold (measured
                             <= 5e-5 at margin 0.02), not to rounding. */
} mvsdf_tracer_params;

#define MVSDF_NUM_TRACE_COUNTERS 256
/* out_counters: slots [0, MVSDF_CTR_SCREENED) are the SDF evaluations the reference algorithm requests per phase (their
 * sum is E_trace of SURVEY 8d, independent of the prefilter); the fixed slots: */
#define MVSDF_CTR_SCREENED 251       /* prefilter: samples evaluated at screening precision (the chunked pass stops early) */
#define MVSDF_CTR_SAMPLER_RAYS 252   /* rays that entered ray_sampler (ray_tracing.py:44-61) */
#define MVSDF_CTR_MINSDF_RAYS 253    /* rays that entered minimal_sdf_points (:86-94), training only */
#define MVSDF_CTR_REFINED 254        /* prefilter: samples evaluated a second time at full precision */
#define MVSDF_CTR_VIOLATIONS 255     /* prefilter: refined samples with |screening - exact| > tau / 2 */

size_t mvsdf_trace_workspace_bytes(int64_t n_rays, int n_images);

/* get_camera_params + lift (code/utils/rend_util.py:48-100), get_sphere_intersection (:141-162) and
 * RayTracing.forward (ray_tracing.py:27-98: sphere_tracing :101-196, ray_sampler :198-258, secant :260-278,
 * minimal_sdf_points :280-308) for n_images x n_pixels rays, with the SDF supplied as packed weights instead of
 * the Python closure of implicit_differentiable_renderer.py:194.
 *   uv [B,N,2] (x=col,y=row), pose [B,4,4] cam->world (the [B,7] quaternion form of rend_util.py:49-54 is expanded to
 *   4x4 by the caller -- B200IDRNetwork does), intrinsics [B,4,4], object_mask [B*N] uint8 or NULL (= all ones);
 *   linspace100 [100] = torch.linspace(0,1,100) (ray_tracing.py:206); steps01 [100] ~ U(0,1) drawn by the caller from
 *   the CPU generator exactly like ray_tracing.py:287 (training only).
 * Outputs: out_ray_dirs [B*N,3], out_cam_loc [B,3] (optional), out_dists [B*N], out_net_mask [B*N] uint8
 *   (network_object_mask), out_points [B*N,3] (optional; cam_loc + dists*ray_dirs, :200 of the renderer),
 *   out_counters [MVSDF_NUM_TRACE_COUNTERS] int32 (optional; see MVSDF_CTR_*). */
int mvsdf_trace(const mvsdf_net* sdf_net, const void* sdf_packed, const float* uv, const float* pose,
                const float* intrinsics, const uint8_t* object_mask, const mvsdf_tracer_params* params, int n_images,
                int n_pixels, int training, const float* linspace100, const float* steps01, size_t workspace_bytes,
                void* workspace, float* out_ray_dirs, float* out_cam_loc, float* out_dists, uint8_t* out_net_mask,
                float* out_points, int32_t* out_counters, void* stream);

/* ---- IDRNetwork.forward after the tracer (implicit_differentiable_renderer.py:200-213, :295-304) and
 * get_rbg_value (:324-338): sdf_output for every ray, order-preserving gather of the surface rays, fused
 * value + analytic normal + feature pass, surface light field, scatter into rgb_values (ones where missed).
 *   surface_mask [R] uint8 (network_object_mask, AND object_mask in training);
 *   out_sdf [R] (optional; a surface ray's entry is the value column of the fused value + normal + feature pass, the other
 *   rays go through one SDF-only launch over a compacted list), out_rgb_values [R,3], out_surf_pts / out_normals [R,3] (first M rows valid),
 *   out_surf_head [R,2] (optional: sdf and surface-indicator logit of the M surface points),
 *   out_hit_index [R] (ray index of each surface point), out_hit_offsets [B+1] (per-image exclusive offsets, [B] = M). */
size_t mvsdf_shade_workspace_bytes(int64_t n_rays, int feature_size);
int mvsdf_shade_rays(const mvsdf_net* sdf_net, const void* sdf_packed, const mvsdf_net* render_net,
                     const void* render_packed, const float* ray_dirs, const float* points, const uint8_t* surface_mask,
                     int n_images, int n_pixels, int feature_size, size_t workspace_bytes, void* workspace,
                     float* out_sdf, float* out_rgb_values, float* out_surf_pts, float* out_normals, float* out_surf_head,
                     int32_t* out_hit_index, int32_t* out_hit_offsets, void* stream);

/* ---- depth-surface samples of training phase 0 (implicit_differentiable_renderer.py:226-239; helpers
 * code/utils/my_utils.py:71-95 get_pixel_grids / idx_img2cam / idx_cam2world): back-projects every pixel of the MVS
 * depth maps into the normalised object frame.
 *   depths [n_maps,h,w]; k_inv [n_maps,3,3] = inverse(depth_cams[.,1,:3,:3]); e_inv [n_maps,4,4] = inverse(depth_cams[.,0])
 *   (the caller inverts, as the reference does with torch.inverse); center [3], size [1];
 *   out_pts [n_maps*h*w,3] = (x_world - center) / size * 2, out_valid [n_maps*h*w] uint8 = depth > 0.
 * The boolean-mask compaction, the +-0.1 jitter and the np.random.choice sub-sampling (:240-247) take host-supplied
 * randomness and stay on the host side (mvsdf_b200/network.py). */
int mvsdf_depth_backproject(const float* depths, const float* k_inv, const float* e_inv, int n_maps, int h, int w,
                            const float* center, const float* size, float* out_pts, uint8_t* out_valid, void* stream);

/* ---- IDRLoss.get_feat_loss_corr (code/model/loss.py:115-165; helpers code/utils/my_utils.py:98-165) ---------
 * Feature maps are constants of the scene (scene_dataset.py:141-149): restack them once into channels-last with
 * mvsdf_feat_nchw_to_nhwc ([n,32,h,w] -> [n,h,w,32]) so that every bilinear tap is one 128-byte load.
 *   surf_pts [M,3] packed by image (diff_surf_pts), hit_offsets [B+1] device int32, cams [B,V,2,4,4] (reference view
 *   first, then the sources; cam[.,0] = 4x4 extrinsic, cam[.,1,:3,:3] = K), maps_nhwc [B,V,h,w,32], size [1], center [3].
 * partials [B,2] float64 = (sum of kept |1-corr|, (V-1)*m_i): this is what a multi-GPU run all-reduces;
 * finalize forms mean_i(sum_i / count_i) exactly like loss.py:155-163. */
int mvsdf_feat_nchw_to_nhwc(const float* src, int n, int channels, int h, int w, float* dst, void* stream);
int mvsdf_feat_loss_partials(const float* surf_pts, const int32_t* hit_offsets, const float* cams, const float* maps_nhwc,
                             int n_images, int n_views, int h, int w, int channels, const float* size,
                             const float* center, double* partials, void* stream);
int mvsdf_feat_loss_finalize(const double* partials, int n_images, float* out_loss, void* stream);
/* Scene feature store variant: maps_nhwc holds ALL feature maps of a scene once, channels-last [n_maps,h,w,32] (the layout
 * the reference's dataset would keep instead of self.feats [n_images,32,h,w] on the host, scene_dataset.py:138-149), and
 * map_index [B,V] (device int32) names the map of every (image, view) -- ground_truth["feat"] = self.feats[idx],
 * ["feat_src"] = self.feats[src_idxs] (scene_dataset.py:205-207) become index lookups, nothing is copied per step. */
int mvsdf_feat_loss_partials_indexed(const float* surf_pts, const int32_t* hit_offsets, const float* cams,
                                     const float* maps_nhwc, const int32_t* map_index, int n_images, int n_views, int h, int w,
                                     int channels, const float* size, const float* center, double* partials, void* stream);
/* Backward of the same term w.r.t. the surface points (what autograd computes through F.grid_sample, the projections and
 * the cosine similarity, loss.py:132-155; the feature maps are constants):  out_grad_pts [M,3] = upstream_grad[0] *
 * d loss / d surf_pts, with partials [B,2] as left by mvsdf_feat_loss_partials (after the all-reduce in a multi-GPU run:
 * the counts are the global denominators).  Every row of out_grad_pts is written (zeros where no term was kept). */
int mvsdf_feat_loss_backward(const float* surf_pts, const int32_t* hit_offsets, const float* cams, const float* maps_nhwc,
                             int n_images, int n_views, int h, int w, int channels, const float* size, const float* center,
                             const double* partials, const float* upstream_grad, float* out_grad_pts, void* stream);
int mvsdf_feat_loss_backward_indexed(const float* surf_pts, const int32_t* hit_offsets, const float* cams,
                                     const float* maps_nhwc, const int32_t* map_index, int n_images, int n_views, int h, int w,
                                     int channels, const float* size, const float* center, const double* partials,
                                     const float* upstream_grad, float* out_grad_pts, void* stream);

/* ---- IDRLoss.get_depth_loss (code/model/loss.py:37-63) with carving_t2 + RunningTopK (code/utils/my_utils.py:168-201,
 * :269-331), use_invalid=False, smooth=None: for every eikonal point, the signed gap to the MVS depth surface along the
 * viewing rays of the n_views depth maps, the inside/outside vote (out_thresh_perc), the clamp to +-1.25 and the far/near
 * attenuation weights.
 *   eik_points [n_points, point_stride] (rows of eikonal_points_hom, normalised object frame; first 3 columns used),
 *   eik_output [n_points] (SDF values), depths [n_views,h,w], depth_cams [n_views,2,4,4], size [1], center [3];
 *   out_target [n_points] = -dist_r, out_weight [n_points] = far_weight * near_weight * in_range (what the backward needs:
 *   d loss / d eik_output = weight * sign(eik_output - target) / n_points);
 *   partials [2] float64 = (sum weight * |eik_output - target|, n_points); finalize with mvsdf_rgb_l1_finalize (sum / count). */
int mvsdf_depth_loss_partials(const float* eik_points, int point_stride, const float* eik_output, int64_t n_points,
                              const float* depths, const float* depth_cams, int n_views, int h, int w, const float* size,
                              const float* center, float out_thresh_perc, float far_thresh, float far_att, float near_thresh,
                              float near_att, float* out_target, float* out_weight, double* partials, void* stream);

/* ---- IDRLoss.get_rgb_loss (loss.py:21-28): sum |rgb - gt| over mask / n_rays.  partials [2] float64 = (sum, n_rays). */
int mvsdf_rgb_l1_partials(const float* rgb_values, const float* rgb_gt, const uint8_t* mask, int64_t n_rays,
                          double* partials, void* stream);
int mvsdf_rgb_l1_finalize(const double* partials, float* out_loss, void* stream);

/* ---- FeatExt (code/utils/my_utils.py:693-708; UNet :595-690, BasicBlock :531-576), eval mode: the CNN whose finest output
 * (32 channels at half resolution) are the feature maps of get_feat_loss_corr; run once per scene over all images
 * (datasets/scene_dataset.py:138-149).  27 convolutions in the order documented in csrc/featext.cu (kFe[]): stem, encoder
 * blocks, decoder, heads.  *_host = HOST arrays (27 entries) of device pointers: conv weights in PyTorch layout
 * ([out,in,k,k]; ConvTranspose2d [in,out,k,k]) and, where a BatchNorm follows, its weight / bias / running_mean /
 * running_var (NULL entries otherwise).  mvsdf_featext_pack folds BatchNorm and writes [tap][in][out] slabs + bias.
 * mvsdf_featext_forward: images_nchw [n,3,H,W] (H, W multiples of 8) -> channels-last maps: out_half_nhwc [n,H/2,W/2,32]
 * (= feat_ext(x)[2], written in the feature store's layout) and optionally the 1/4 and 1/8 maps (= [1], [0]). */
int mvsdf_featext_num_convs(void);
size_t mvsdf_featext_packed_floats(void);
int mvsdf_featext_pack(const float* const* conv_weight_host, const float* const* bn_weight_host, const float* const* bn_bias_host,
                       const float* const* bn_mean_host, const float* const* bn_var_host, float bn_eps, float* packed, void* stream);
size_t mvsdf_featext_workspace_bytes(int n_images, int height, int width);
int mvsdf_featext_forward(const float* packed, const float* images_nchw, int n_images, int height, int width, size_t workspace_bytes,
                          void* workspace, float* out_eighth_nhwc, float* out_quarter_nhwc, float* out_half_nhwc, void* stream);

/* ---- the training step: backward through the two MLPs + optimiser --------------------------
 * What eager autograd does in the reference: loss.backward() (code/training/idr_train.py:287) through
 * ImplicitNetwork.forward / .gradient -- second order, create_graph=True (implicit_differentiable_renderer.py:96-107) --
 * and RenderingNetwork.forward (:145-167), torch.nn.utils.clip_grad_norm_ (:292) and torch.optim.Adam.step (:300).
 *
 * Protocol per step:  mvsdf_pack_weights (+ mvsdf_pack_weights_t: W^T tiles for the reverse sweep)
 *   -> mvsdf_sdf_forward_train / mvsdf_render_forward_train: the fused forward kernels, which additionally write the input
 *      operand of every layer (fp16 hi/lo, mvsdf_train_save_bytes) -- everything the backward needs: with softplus(beta=100),
 *      sp'(z) = 1 - exp(-100 sp(z)) and sp''(z) S = 100 (1 - sp'(z)) T are functions of the saved values
 *   -> mvsdf_sdf_backward / mvsdf_render_backward: reverse sweep on the tcgen05 tile core (D = W^T [dZ | dS], epilogue applies
 *      sp' / sp'' or ReLU', chains the positional encoding into d/dx) and the dW GEMM over the points; results in PLAN
 *      coordinates: out_dw [mvsdf_train_dw_floats] = per layer [padded out rows, padded in], out_db [mvsdf_train_db_floats]
 *   -> mvsdf_weight_grads: weight-norm chain  dg = <dW, v^>,  dv = g/||v|| (dW - <dW, v^> v^)  into tensors shaped like the
 *      parameters (weight_v [out,in], weight_g [out,1], bias [out])
 *   -> mvsdf_adam_step.
 * with_grad = 1: value + d/dx columns (SDF net, 16 points per tile), 0: plain columns (rendering net, 64 points per tile). */
size_t mvsdf_train_packed_t_bytes(const mvsdf_net* net);
size_t mvsdf_train_save_bytes(const mvsdf_net* net, int64_t n, int with_grad);
size_t mvsdf_train_workspace_bytes(const mvsdf_net* net, int64_t n, int with_grad);
size_t mvsdf_train_dw_floats(const mvsdf_net* net);
size_t mvsdf_train_db_floats(const mvsdf_net* net);
int mvsdf_pack_weights_t(const mvsdf_net* net, const float* const* weight_v_host, const float* const* weight_g_host, void* packed_t,
                         void* stream);
/* ImplicitNetwork.forward + .gradient, full head: out_full [n, 2+F], out_grad [n,3]; save [save_bytes]. */
int mvsdf_sdf_forward_train(const mvsdf_net* net, const void* packed, const float* x, int64_t n, size_t save_bytes, void* save,
                            float* out_full, float* out_grad, void* stream);
/* g_full [n, 2+F] = dL/d full, g_grad [n,3] = dL/d grad (either may be NULL = zero); out_dx [n,3] optional = dL/dx.
 * out_dw = out_db = NULL: dx-only sweep (no gradient dumps, no dW GEMM; the workspace needs mvsdf_train_workspace_bytes(net, 0, with_grad)
 *      bytes: the scale header + the CTA-pair sweep's scratch) -- the first of the two
 * sweeps over the surface points, whose only purpose is dL/d x_diff for the implicit-differentiation term. */
int mvsdf_sdf_backward(const mvsdf_net* net, const void* packed_t, const float* x, int64_t n, const void* save, const float* g_full,
                       const float* g_grad, size_t workspace_bytes, void* workspace, float* out_dx, float* out_dw, float* out_db,
                       void* stream);
int mvsdf_render_forward_train(const mvsdf_net* net, const void* packed, const float* points, const float* view_dirs,
                               const float* normals, const float* features, int64_t n, size_t save_bytes, void* save, float* out_rgb,
                               void* stream);
/* rgb [n,3] = the forward output (tanh'), g_rgb [n,3] = dL/d rgb; d_points / d_normals [n,3], d_feats [n,F] optional;
 * d_view [n,3] optional (needs view_dirs [n,3], the forward's input): only trained camera poses make the view direction a
 * function of parameters (utils/rend_util.py:49-57, train_cameras=True). */
int mvsdf_render_backward(const mvsdf_net* net, const void* packed_t, int64_t n, const void* save, const float* rgb, const float* g_rgb,
                          const float* view_dirs, size_t workspace_bytes, void* workspace, float* d_points, float* d_normals,
                          float* d_feats, float* d_view, float* out_dw, float* out_db, void* stream);
/* *_host: HOST arrays of device pointers, one per source layer lin{l} (like mvsdf_pack_weights). */
int mvsdf_weight_grads(const mvsdf_net* net, const float* dw, const float* db, const float* const* weight_v_host,
                       const float* const* weight_g_host, float* const* out_dv_host, float* const* out_dg_host,
                       float* const* out_dbias_host, void* stream);
/* torch.optim.Adam.step over n_tensors fp32 tensors with clip_grad_norm_(max_grad_norm) folded in (<= 0: no clipping);
 * step = 1, 2, ...; scratch_sumsq: device double; out_grad_norm: optional device float = ||g|| before clipping. */
int mvsdf_adam_step(int n_tensors, float* const* params_host, const float* const* grads_host, float* const* exp_avg_host,
                    float* const* exp_avg_sq_host, const int64_t* sizes_host, float lr, float beta1, float beta2, float eps, int step,
                    float max_grad_norm, double* scratch_sumsq, float* out_grad_norm, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVSDF_B200_H_ */

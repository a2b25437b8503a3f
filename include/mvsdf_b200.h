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

typedef struct mvsdf_net mvsdf_net;   /* host-side layout plan of one packed MLP; owns no device memory */

int mvsdf_abi_version(void);
const char* mvsdf_last_error(void);

/* ---- network plans ------------------------------------------------------------------------------
 * ImplicitNetwork.__init__ (implicit_differentiable_renderer.py:19-75): dims [3+6*n_freqs] + [width]*n_hidden
 * + [2+feature_size], skip connection into layer `skip_layer`, softplus(beta=100), weight_norm on every layer. */
mvsdf_net* mvsdf_sdf_net_create(int width, int n_hidden, int skip_layer, int n_freqs, int feature_size);
/* RenderingNetwork.__init__ (:109-143), mode='idr': dims [9 + feature_size + 6*n_freqs_view] + [width]*n_hidden + [3]. */
mvsdf_net* mvsdf_render_net_create(int width, int n_hidden, int n_freqs_view, int feature_size);
void mvsdf_net_destroy(mvsdf_net* net);
int mvsdf_net_num_layers(const mvsdf_net* net);          /* number of source Linear layers */
size_t mvsdf_net_packed_bytes(const mvsdf_net* net);      /* size of the packed blob the caller must allocate */

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

#ifdef __cplusplus
}
#endif
#endif /* MVSDF_B200_H_ */

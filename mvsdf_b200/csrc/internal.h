// Internal (non-ABI) declarations shared by the translation units of libmvsdf_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>

#include "netplan.h"

struct mvsdf_net {
  mvsdf::NetPlan plan;
};

namespace mvsdf {
struct MlpArgs;
int fill_mlp_args(const NetPlan& p, const void* packed, int head, MlpArgs& a);

int fail(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int sm_count();
void note_launch();   // counts kernel launches made by the library (mvsdf_launch_count)
// optional CUDA-event timing of a launch (mvsdf_profile_*); kinds 5 = backward sweep, 6 = dW GEMM
void* prof_begin_ext(int kind, cudaStream_t st);
void prof_end_ext(void* handle, cudaStream_t st);

// MLP tile launches (mlp_abi.cu)
int mlp_sdf(const mvsdf_net* net, const void* packed, const float* x, int64_t n, const int32_t* n_dev, int head,
            float* out_sdf, float* out_full, float* out_grad, bool with_grad, cudaStream_t st, bool screening = false,
            uint8_t* save = nullptr, const long long* save_off = nullptr);
int mlp_render(const mvsdf_net* net, const void* packed, const float* pts, const float* view, const float* normals,
               const float* feats, int feat_stride, int64_t n, const int32_t* n_dev, float* rgb, cudaStream_t st,
               uint8_t* save = nullptr, const long long* save_off = nullptr);

}  // namespace mvsdf

#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "tc_common.cuh"

namespace tcelbo {
cudaError_t launch_density_fwd(bool floored, const float* x, const float* mu, const float* lv, const int64_t* shape,
                               const int64_t* sx, const int64_t* sm, const int64_t* sl, float* out, cudaStream_t st);
cudaError_t launch_density_bwd(bool floored, const float* x, const float* mu, const float* lv, const float* g, const int64_t* shape,
                               const int64_t* sx, const int64_t* sm, const int64_t* sl, float* gx, float* gmu, float* glv, cudaStream_t st);
cudaError_t launch_sampling_fwd(const float* lp, int B, int D, const Weights& w, float lw_n, float lw_s, float post,
                                float* prod, float* joint, float* lse_d, float* srow, cudaStream_t st);
cudaError_t launch_sampling_bwd(const float* lp, int B, int D, const Weights& w, float lw_n, float lw_s, float post,
                                const float* g_prod, const float* g_joint, const float* lse_d, const float* srow, const float* joint,
                                float* glp, cudaStream_t st);
}  // namespace tcelbo

#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace tcelbo {
cudaError_t launch_kl_fwd(const float* lv, int64_t ldlv, const float* mu, int64_t ldmu, int b, int d, float* kl_rows, cudaStream_t st);
cudaError_t launch_kl_bwd(const float* lv, int64_t ldlv, const float* mu, int64_t ldmu, const float* g_rows, int b, int d,
                          float* glv, int64_t ldglv, float* gmu, int64_t ldgmu, cudaStream_t st);
cudaError_t launch_reparam_fwd(const float* mu, int64_t ldmu, const float* lv, int64_t ldlv, const float* eps, int64_t ldeps,
                               int b, int d, float* z, int64_t ldz, cudaStream_t st);
cudaError_t launch_reparam_bwd(const float* lv, int64_t ldlv, const float* eps, int64_t ldeps, const float* gz, int64_t ldgz,
                               int b, int d, float* gmu, int64_t ldgmu, float* glv, int64_t ldglv, bool accumulate, cudaStream_t st);
cudaError_t launch_rowdensity_fwd(const float* x, int64_t ldx, const float* mu, int64_t ldmu, const float* lv, int64_t ldlv,
                                  int b, int d, float* out, cudaStream_t st);
cudaError_t launch_rowdensity_bwd(const float* x, int64_t ldx, const float* mu, int64_t ldmu, const float* lv, int64_t ldlv,
                                  const float* g_rows, int b, int d, float* gx, int64_t ldgx, float* gmu, int64_t ldgmu,
                                  float* glv, int64_t ldglv, cudaStream_t st);
cudaError_t launch_recloss_fwd(const float* x, const float* r, int b, int64_t n, int kind, float* partial, int nchunk, float* out, cudaStream_t st);
cudaError_t launch_recloss_bwd(const float* x, const float* r, const float* g_rows, int b, int64_t n, int kind, float* gr, cudaStream_t st);
cudaError_t launch_expelbo_fwd(const float* rec, const float* kl, int b, float scale, float* out, float* e_rows, cudaStream_t st);
cudaError_t launch_expelbo_bwd(const float* e_rows, const float* g, int b, float scale, float* g_rows, cudaStream_t st);
cudaError_t launch_ex2_peak(float* out, int iters, int ctas, cudaStream_t st);
}  // namespace tcelbo

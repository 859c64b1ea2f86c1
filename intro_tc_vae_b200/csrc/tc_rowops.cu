// Row-wise companions of the TC estimator (all HBM-bound, one warp per row, coalesced over D):
//   KL to the unit Gaussian          ops.py:136-163
//   reparameterisation               ops.py:166-185
//   row-wise Gaussian log-density    ops.py:24-29 (+ .sum(dim=1), solvers/tc.py:107,112)
#include "tc_common.cuh"
#include "tc_rowops.h"
#include "tc_instr.h"

namespace tcelbo {

constexpr int kRowWarps = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void kl_fwd_kernel(const float* __restrict__ lv, int64_t ldlv, const float* __restrict__ mu, int64_t ldmu,
                              int b, int d, float* __restrict__ kl_rows) {
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= b) return;
    float acc = 0.0f;
    for (int dd = lane; dd < d; dd += 32) {
        const float l = lv[(int64_t)row * ldlv + dd], m = mu[(int64_t)row * ldmu + dd];
        acc += 1.0f + l - expf(l) - m * m;
    }
    acc = warp_sum(acc);
    if (lane == 0) kl_rows[row] = -0.5f * acc;
}

__global__ void kl_bwd_kernel(const float* __restrict__ lv, int64_t ldlv, const float* __restrict__ mu, int64_t ldmu,
                              const float* __restrict__ g_rows, int b, int d,
                              float* __restrict__ glv, int64_t ldglv, float* __restrict__ gmu, int64_t ldgmu) {
    const int64_t n = (int64_t)b * d;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / d), dd = (int)(idx % d);
        const float g = g_rows[i];
        glv[(int64_t)i * ldglv + dd] = g * 0.5f * (expf(lv[(int64_t)i * ldlv + dd]) - 1.0f);
        gmu[(int64_t)i * ldgmu + dd] = g * mu[(int64_t)i * ldmu + dd];
    }
}

__global__ void reparam_fwd_kernel(const float* __restrict__ mu, int64_t ldmu, const float* __restrict__ lv, int64_t ldlv,
                                   const float* __restrict__ eps, int64_t ldeps, int b, int d,
                                   float* __restrict__ z, int64_t ldz) {
    const int64_t n = (int64_t)b * d;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / d), dd = (int)(idx % d);
        const float std = expf(0.5f * lv[(int64_t)i * ldlv + dd]);
        z[(int64_t)i * ldz + dd] = mu[(int64_t)i * ldmu + dd] + eps[(int64_t)i * ldeps + dd] * std;
    }
}

__global__ void reparam_bwd_kernel(const float* __restrict__ lv, int64_t ldlv, const float* __restrict__ eps, int64_t ldeps,
                                   const float* __restrict__ gz, int64_t ldgz, int b, int d,
                                   float* __restrict__ gmu, int64_t ldgmu, float* __restrict__ glv, int64_t ldglv) {
    const int64_t n = (int64_t)b * d;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / d), dd = (int)(idx % d);
        const float g = gz[(int64_t)i * ldgz + dd];
        const float std = expf(0.5f * lv[(int64_t)i * ldlv + dd]);
        gmu[(int64_t)i * ldgmu + dd] = g;
        glv[(int64_t)i * ldglv + dd] = g * eps[(int64_t)i * ldeps + dd] * (0.5f * std);
    }
}

// lp = max(-0.5*((x-mu)^2*exp(-lv) + lv + log2pi), -50); the reference's log 2pi is an fp32 constant (ops.py:25)
__global__ void rowdensity_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ mu, int64_t ldmu,
                                      const float* __restrict__ lv, int64_t ldlv, int b, int d, float* __restrict__ out) {
    const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= b) return;
    float acc = 0.0f;
    for (int dd = lane; dd < d; dd += 32) {
        const float xv = x[(int64_t)row * ldx + dd];
        const float m = mu ? mu[(int64_t)row * ldmu + dd] : 0.0f;
        const float l = lv ? lv[(int64_t)row * ldlv + dd] : 0.0f;
        const float t = xv - m;
        const float lp = -0.5f * (t * t * expf(-l) + l + kLog2Pi);
        acc += fmax_nan(lp, kLogpFloor);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[row] = acc;
}

__global__ void rowdensity_bwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ mu, int64_t ldmu,
                                      const float* __restrict__ lv, int64_t ldlv, const float* __restrict__ g_rows, int b, int d,
                                      float* __restrict__ gx, int64_t ldgx, float* __restrict__ gmu, int64_t ldgmu,
                                      float* __restrict__ glv, int64_t ldglv) {
    const int64_t n = (int64_t)b * d;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / d), dd = (int)(idx % d);
        const float xv = x[(int64_t)i * ldx + dd];
        const float m = mu ? mu[(int64_t)i * ldmu + dd] : 0.0f;
        const float l = lv ? lv[(int64_t)i * ldlv + dd] : 0.0f;
        const float t = xv - m, iv = expf(-l);
        const float lp = -0.5f * (t * t * iv + l + kLog2Pi);
        const float g = (lp >= kLogpFloor) ? g_rows[i] : 0.0f;            // clamp passes gradient where lp >= -50
        if (gx) gx[(int64_t)i * ldgx + dd] = -g * t * iv;
        if (gmu) gmu[(int64_t)i * ldgmu + dd] = g * t * iv;
        if (glv) glv[(int64_t)i * ldglv + dd] = g * 0.5f * (t * t * iv - 1.0f);
    }
}

// MUFU.EX2 saturation probe: 8 independent dependency chains per thread, nothing but ex2 in the loop.
__global__ void __launch_bounds__(256) ex2_peak_kernel(float* __restrict__ out, int iters) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = -1.0f - 0.001f * (float)(threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = ex2(-v[k]);          // maps (0,1] <-> [0.5,1): stays finite forever
    }
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += v[k];
    if (acc == 12345.678f) out[0] = acc;                          // keep the chains alive without real traffic
}

cudaError_t launch_ex2_peak(float* out, int iters, int ctas, cudaStream_t st) {
    LaunchScope scope(kKernNone, st);
    ex2_peak_kernel<<<ctas, 256, 0, st>>>(out, iters);
    return cudaGetLastError();
}

static inline int grid1d(int64_t n, int block) {
    int64_t g = (n + block - 1) / block;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    return (int)g;
}

cudaError_t launch_kl_fwd(const float* lv, int64_t ldlv, const float* mu, int64_t ldmu, int b, int d, float* kl_rows, cudaStream_t st) {
    LaunchScope scope(kKernNone, st);
    kl_fwd_kernel<<<(b + kRowWarps - 1) / kRowWarps, kRowWarps * 32, 0, st>>>(lv, ldlv, mu, ldmu, b, d, kl_rows);
    return cudaGetLastError();
}
cudaError_t launch_kl_bwd(const float* lv, int64_t ldlv, const float* mu, int64_t ldmu, const float* g_rows, int b, int d,
                          float* glv, int64_t ldglv, float* gmu, int64_t ldgmu, cudaStream_t st) {
    LaunchScope scope(kKernNone, st);
    kl_bwd_kernel<<<grid1d((int64_t)b * d, 256), 256, 0, st>>>(lv, ldlv, mu, ldmu, g_rows, b, d, glv, ldglv, gmu, ldgmu);
    return cudaGetLastError();
}
cudaError_t launch_reparam_fwd(const float* mu, int64_t ldmu, const float* lv, int64_t ldlv, const float* eps, int64_t ldeps,
                               int b, int d, float* z, int64_t ldz, cudaStream_t st) {
    LaunchScope scope(kKernNone, st);
    reparam_fwd_kernel<<<grid1d((int64_t)b * d, 256), 256, 0, st>>>(mu, ldmu, lv, ldlv, eps, ldeps, b, d, z, ldz);
    return cudaGetLastError();
}
cudaError_t launch_reparam_bwd(const float* lv, int64_t ldlv, const float* eps, int64_t ldeps, const float* gz, int64_t ldgz,
                               int b, int d, float* gmu, int64_t ldgmu, float* glv, int64_t ldglv, cudaStream_t st) {
    LaunchScope scope(kKernNone, st);
    reparam_bwd_kernel<<<grid1d((int64_t)b * d, 256), 256, 0, st>>>(lv, ldlv, eps, ldeps, gz, ldgz, b, d, gmu, ldgmu, glv, ldglv);
    return cudaGetLastError();
}
cudaError_t launch_rowdensity_fwd(const float* x, int64_t ldx, const float* mu, int64_t ldmu, const float* lv, int64_t ldlv,
                                  int b, int d, float* out, cudaStream_t st) {
    LaunchScope scope(kKernNone, st);
    rowdensity_fwd_kernel<<<(b + kRowWarps - 1) / kRowWarps, kRowWarps * 32, 0, st>>>(x, ldx, mu, ldmu, lv, ldlv, b, d, out);
    return cudaGetLastError();
}
cudaError_t launch_rowdensity_bwd(const float* x, int64_t ldx, const float* mu, int64_t ldmu, const float* lv, int64_t ldlv,
                                  const float* g_rows, int b, int d, float* gx, int64_t ldgx, float* gmu, int64_t ldgmu,
                                  float* glv, int64_t ldglv, cudaStream_t st) {
    LaunchScope scope(kKernNone, st);
    rowdensity_bwd_kernel<<<grid1d((int64_t)b * d, 256), 256, 0, st>>>(x, ldx, mu, ldmu, lv, ldlv, g_rows, b, d,
                                                                        gx, ldgx, gmu, ldgmu, glv, ldglv);
    return cudaGetLastError();
}

}  // namespace tcelbo

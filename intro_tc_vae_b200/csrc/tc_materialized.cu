// Materialised-tensor companions of the fused estimator (the reference's own formulation, HBM-bound):
//   pairwise / broadcast Gaussian log-densities   ops.py:15-21 (floored, "torch" variant) and ops.py:24-29
//   minibatch_{stratified,weighted}_sampling on a given [B,B,D] tensor   ops.py:92-115
// Kept for API parity with the reference's ops.py; the fused path never materialises the tensor.
#include "tc_common.cuh"
#include "tc_materialized.h"
#include "tc_instr.h"

namespace tcelbo {

// out[i,j,d] over a broadcast 3-D index space; operand strides may be 0 on broadcast dims.
struct Bcast3 { int64_t n0, n1, n2; int64_t sx[3], sm[3], sl[3]; };

template <bool kFloor>
__device__ __forceinline__ float density_value(float x, float m, float lv, bool& pass, float& iv_out, float& var_out, float& vc_out) {
    float raw;
    if (kFloor) {                                   // ops.py:15-21: var floored at 1e-4 (straight-through), full=True
        const float var = expf(lv);
        const float vc = (var < kVarFloor) ? kVarFloor : var;
        const float t = x - m;
        raw = -(0.5f * (logf(vc) + t * t / vc) + 0.5f * kLog2Pi);
        iv_out = 1.0f / vc; var_out = var; vc_out = vc;
    } else {                                        // ops.py:24-29
        const float iv = expf(-lv);
        const float t = x - m;
        raw = -0.5f * (t * t * iv + lv + kLog2Pi);
        iv_out = iv; var_out = 1.0f; vc_out = 1.0f;
    }
    pass = !(raw < kLogpFloor);
    return fmax_nan(raw, kLogpFloor);
}

template <bool kFloor>
__global__ void density_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mu, const float* __restrict__ lv,
                                   Bcast3 b, float* __restrict__ out) {
    const int64_t n = b.n0 * b.n1 * b.n2;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = idx % b.n2, j = (idx / b.n2) % b.n1, i = idx / (b.n2 * b.n1);
        bool pass; float iv, var, vc;
        out[idx] = density_value<kFloor>(x[i * b.sx[0] + j * b.sx[1] + d * b.sx[2]], mu[i * b.sm[0] + j * b.sm[1] + d * b.sm[2]],
                                         lv[i * b.sl[0] + j * b.sl[1] + d * b.sl[2]], pass, iv, var, vc);
    }
}

// elementwise gradients at the broadcast shape; the caller sums them down to the operand shapes
template <bool kFloor>
__global__ void density_bwd_kernel(const float* __restrict__ x, const float* __restrict__ mu, const float* __restrict__ lv,
                                   const float* __restrict__ g, Bcast3 b, float* __restrict__ gx, float* __restrict__ gmu,
                                   float* __restrict__ glv) {
    const int64_t n = b.n0 * b.n1 * b.n2;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = idx % b.n2, j = (idx / b.n2) % b.n1, i = idx / (b.n2 * b.n1);
        const float xv = x[i * b.sx[0] + j * b.sx[1] + d * b.sx[2]], m = mu[i * b.sm[0] + j * b.sm[1] + d * b.sm[2]];
        const float l = lv[i * b.sl[0] + j * b.sl[1] + d * b.sl[2]];
        bool pass; float iv, var, vc;
        density_value<kFloor>(xv, m, l, pass, iv, var, vc);
        const float gg = pass ? g[idx] : 0.0f;
        const float t = xv - m;
        gx[idx] = -gg * t * iv;
        gmu[idx] = gg * t * iv;
        glv[idx] = kFloor ? gg * var * (-0.5f * iv + 0.5f * t * t * iv * iv) : gg * 0.5f * (t * t * iv - 1.0f);
    }
}

// ---- estimators on a materialised [B,B,D] tensor: one CTA per row i ------------------------------------------
__device__ __forceinline__ float logw_of(const Weights& w, float lw_n, float lw_s, int i, int j) {
    if (!w.mss) return 0.0f;                         // MWS subtracts log(B*N) after the logsumexp (ops.py:96,99)
    if (j == 0) return (i == w.b_glob - 2) ? lw_s : lw_n;
    if (j == 1) return lw_s;
    return w.lw_u;
}

__device__ __forceinline__ float block_reduce(float v, float* sh, bool is_max) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const float t = __shfl_xor_sync(0xffffffffu, v, o); v = is_max ? fmaxf(v, t) : v + t; }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = sh[0];
    for (int k = 1; k < nw; ++k) r = is_max ? fmaxf(r, sh[k]) : r + sh[k];
    return r;
}

// prod[i] = sum_d LSE_j(logw_ij + lp_ijd) ; joint[i] = LSE_j(logw_ij + sum_d lp_ijd); also saves the per-(i,d) LSE and per-(i,j) sums
__global__ void sampling_fwd_kernel(const float* __restrict__ lp, int B, int D, Weights w, float lw_n, float lw_s, float post,
                                    float* __restrict__ prod, float* __restrict__ joint, float* __restrict__ lse_d /*[B,D]*/,
                                    float* __restrict__ srow /*[B,B]*/) {
    __shared__ float sh[32];
    const int i = blockIdx.x;
    const float* base = lp + (size_t)i * B * D;
    float psum = 0.0f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {            // per-dimension logsumexp over j (thread owns a dim)
        float m = -INFINITY;
        for (int j = 0; j < B; ++j) m = fmaxf(m, logw_of(w, lw_n, lw_s, i, j) + base[(size_t)j * D + d]);
        float s = 0.0f;
        for (int j = 0; j < B; ++j) s += expf(logw_of(w, lw_n, lw_s, i, j) + base[(size_t)j * D + d] - m);
        const float l = (m == -INFINITY) ? -INFINITY : m + logf(s);
        lse_d[(size_t)i * D + d] = l;
        psum += l + post;
    }
    psum = block_reduce(psum, sh, false);
    float m = -INFINITY;
    for (int j = threadIdx.x; j < B; j += blockDim.x) {            // joint: thread owns a column
        float s = 0.0f;
        for (int d = 0; d < D; ++d) s += base[(size_t)j * D + d];
        srow[(size_t)i * B + j] = s;
        m = fmaxf(m, logw_of(w, lw_n, lw_s, i, j) + s);
    }
    m = block_reduce(m, sh, true);
    float se = 0.0f;
    for (int j = threadIdx.x; j < B; j += blockDim.x) se += expf(logw_of(w, lw_n, lw_s, i, j) + srow[(size_t)i * B + j] - m);
    se = block_reduce(se, sh, false);
    if (threadIdx.x == 0) { prod[i] = psum; joint[i] = m + logf(se) + post; }
}

__global__ void sampling_bwd_kernel(const float* __restrict__ lp, int B, int D, Weights w, float lw_n, float lw_s, float post,
                                    const float* __restrict__ g_prod, const float* __restrict__ g_joint,
                                    const float* __restrict__ lse_d, const float* __restrict__ srow, const float* __restrict__ joint,
                                    float* __restrict__ glp) {
    const int64_t n = (int64_t)B * B * D;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(idx % D), j = (int)((idx / D) % B), i = (int)(idx / ((int64_t)D * B));
        const float lw = logw_of(w, lw_n, lw_s, i, j);
        const float p = expf(lw + lp[idx] - lse_d[(size_t)i * D + d]);
        const float q = expf(lw + srow[(size_t)i * B + j] - (joint[i] - post));
        glp[idx] = g_prod[i] * p + g_joint[i] * q;
    }
}

static inline int grid1(int64_t n, int block) { int64_t g = (n + block - 1) / block; if (g > 148 * 32) g = 148 * 32; if (g < 1) g = 1; return (int)g; }

cudaError_t launch_density_fwd(bool floored, const float* x, const float* mu, const float* lv, const int64_t* shape,
                               const int64_t* sx, const int64_t* sm, const int64_t* sl, float* out, cudaStream_t st) {
    Bcast3 b; b.n0 = shape[0]; b.n1 = shape[1]; b.n2 = shape[2];
    for (int k = 0; k < 3; ++k) { b.sx[k] = sx[k]; b.sm[k] = sm[k]; b.sl[k] = sl[k]; }
    const int64_t n = b.n0 * b.n1 * b.n2;
    LaunchScope scope(kKernNone, st);
    if (floored) density_fwd_kernel<true><<<grid1(n, 256), 256, 0, st>>>(x, mu, lv, b, out);
    else         density_fwd_kernel<false><<<grid1(n, 256), 256, 0, st>>>(x, mu, lv, b, out);
    return cudaGetLastError();
}

cudaError_t launch_density_bwd(bool floored, const float* x, const float* mu, const float* lv, const float* g, const int64_t* shape,
                               const int64_t* sx, const int64_t* sm, const int64_t* sl, float* gx, float* gmu, float* glv, cudaStream_t st) {
    Bcast3 b; b.n0 = shape[0]; b.n1 = shape[1]; b.n2 = shape[2];
    for (int k = 0; k < 3; ++k) { b.sx[k] = sx[k]; b.sm[k] = sm[k]; b.sl[k] = sl[k]; }
    const int64_t n = b.n0 * b.n1 * b.n2;
    LaunchScope scope(kKernNone, st);
    if (floored) density_bwd_kernel<true><<<grid1(n, 256), 256, 0, st>>>(x, mu, lv, g, b, gx, gmu, glv);
    else         density_bwd_kernel<false><<<grid1(n, 256), 256, 0, st>>>(x, mu, lv, g, b, gx, gmu, glv);
    return cudaGetLastError();
}

cudaError_t launch_sampling_fwd(const float* lp, int B, int D, const Weights& w, float lw_n, float lw_s, float post,
                                float* prod, float* joint, float* lse_d, float* srow, cudaStream_t st) {
    LaunchScope scope(kKernNone, st);
    sampling_fwd_kernel<<<B, 256, 0, st>>>(lp, B, D, w, lw_n, lw_s, post, prod, joint, lse_d, srow);
    return cudaGetLastError();
}

cudaError_t launch_sampling_bwd(const float* lp, int B, int D, const Weights& w, float lw_n, float lw_s, float post,
                                const float* g_prod, const float* g_joint, const float* lse_d, const float* srow, const float* joint,
                                float* glp, cudaStream_t st) {
    LaunchScope scope(kKernNone, st);
    sampling_bwd_kernel<<<grid1((int64_t)B * B * D, 256), 256, 0, st>>>(lp, B, D, w, lw_n, lw_s, post, g_prod, g_joint, lse_d, srow, joint, glp);
    return cudaGetLastError();
}

}  // namespace tcelbo

// Register-file bandwidth probe: FFMA2 / FFMA with three distinct register operands vs constant operands.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// MODE 0: FFMA2 d = d*a + b (a,b loop-invariant regs)    MODE 1: FFMA2 d[k] = d[k]*x[k] + y[k] (3 distinct reg pairs)
// MODE 2: FMUL2 d[k] = d[k]*x[k]                          MODE 3: scalar FFMA d[k] = d[k]*x[k] + y[k]
// MODE 4: MODE 1 + 4 MUFU per 9 FFMA2 (backward-like mix) MODE 5: dependent FFMA2 chain latency (1 chain)
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float seed) {
    u64 d[8], x[8], y[8]; float s[16], sx[16], sy[16], m[4];
#pragma unroll
    for (int k = 0; k < 8; ++k) { d[k] = pack2(1.0f + 0.002f * k, 1.0f + seed * (threadIdx.x & 7)); x[k] = pack2(0.999f + seed * k, 0.998f); y[k] = pack2(0.0007f * k, seed); }
#pragma unroll
    for (int k = 0; k < 16; ++k) { s[k] = 1.0f + seed * k; sx[k] = 0.999f + seed * k; sy[k] = seed * k; }
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = -1.0f - seed * k;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) d[k] = ffma2(d[k], x[0], y[0]);
        } else if (MODE == 1 || MODE == 4) {
#pragma unroll
            for (int k = 0; k < 8; ++k) d[k] = ffma2(d[k], x[k], y[k]);
            if (MODE == 4) {
                d[0] = ffma2(d[0], x[1], y[2]);
#pragma unroll
                for (int k = 0; k < 4; ++k) m[k] = ex2(-m[k]);
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int k = 0; k < 8; ++k) d[k] = fmul2(d[k], x[k]);
        } else if (MODE == 3) {
#pragma unroll
            for (int k = 0; k < 16; ++k) s[k] = ffma(s[k], sx[k], sy[k]);
        } else if (MODE == 5) {
#pragma unroll
            for (int k = 0; k < 8; ++k) d[0] = ffma2(d[0], x[k], y[k]);
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { float lo, hi; unpack2(d[k], lo, hi); acc += lo + hi; unpack2(x[k], lo, hi); acc += lo + hi; unpack2(y[k], lo, hi); acc += lo + hi; }
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += s[k] + sx[k] + sy[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc += m[k];
    if (acc == 1234.5678f) out[0] = acc;
}
template <int MODE> void run(const char* name, double instr_per_it, int warps_per_sm) {
    float* out; cudaMalloc(&out, 64);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int ctas = sms * warps_per_sm / 8, iters = 20000;
    probe<MODE><<<ctas, 256>>>(out, 100, 1e-4f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<ctas, 256>>>(out, iters, 1e-4f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double clk = 1.965e9;
    const double wi = (double)ctas * 8 * iters * instr_per_it / (ms * 1e-3) / sms / clk;     // warp-instr per SM per clk
    printf("%-52s warps/SM %2d  %8.3f ms  warp-instr/clk/SM %5.2f  (cycles per instr per SMSP %5.2f)\n", name, warps_per_sm, ms, wi, 4.0 / wi);
    cudaFree(out);
}
int main() {
    for (int w : {8, 16, 64}) {
        run<0>("FFMA2 d=d*a+b (a,b shared)", 8, w);
        run<1>("FFMA2 d=d*x+y (3 distinct pairs)", 8, w);
        run<2>("FMUL2 d=d*x", 8, w);
        run<3>("FFMA  s=s*x+y (3 distinct regs)", 16, w);
        run<4>("9 FFMA2 (distinct) + 4 MUFU", 13, w);
        run<5>("FFMA2 dependent chain (latency)", 8, w);
    }
    return 0;
}

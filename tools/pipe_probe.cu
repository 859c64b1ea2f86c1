// Pipe-throughput probes for sm_100a: scalar FFMA vs packed FFMA2 vs MUFU.EX2 and mixes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu && ./pipe_probe
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fmin_(float a, float b) { float y; asm volatile("min.NaN.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b)); return y; }

// MODE 0: 16 independent scalar FFMA chains      MODE 1: 8 independent FFMA2 chains (16 lanes-worth)
// MODE 2: 8 MUFU + 16 scalar FFMA per iteration  MODE 3: 8 MUFU + 8 FFMA2 per iteration
// MODE 4: 8 MUFU + 16 FFMA2                      MODE 5: 8 MUFU + 8 FFMA2 + 16 scalar FFMA
// MODE 6: 8 MUFU + 32 scalar FFMA                MODE 7: 8 MUFU + 8 FMNMX + 8 FFMA2 (forward-like)
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters) {
    float s[32]; u64 p[16]; float m[8];
#pragma unroll
    for (int k = 0; k < 32; ++k) s[k] = 1.0f + 0.001f * (threadIdx.x + k);
#pragma unroll
    for (int k = 0; k < 16; ++k) p[k] = pack2(1.0f + 0.002f * k, 1.0f + 0.003f * (threadIdx.x & 7));
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -1.0f - 0.001f * (threadIdx.x + k);
    const float a = 0.999f, b = 0.0007f; const u64 a2 = pack2(a, a), b2 = pack2(b, b);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 16; ++k) s[k] = ffma(s[k], a, b);
        } else if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) p[k] = ffma2(p[k], a2, b2);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) m[k] = ex2(-m[k]);
            if (MODE == 2 || MODE == 5) {
#pragma unroll
                for (int k = 0; k < 16; ++k) s[k] = ffma(s[k], a, b);
            }
            if (MODE == 6) {
#pragma unroll
                for (int k = 0; k < 32; ++k) s[k] = ffma(s[k], a, b);
            }
            if (MODE == 3 || MODE == 5 || MODE == 7) {
#pragma unroll
                for (int k = 0; k < 8; ++k) p[k] = ffma2(p[k], a2, b2);
            }
            if (MODE == 4) {
#pragma unroll
                for (int k = 0; k < 16; ++k) p[k] = ffma2(p[k], a2, b2);
            }
            if (MODE == 7) {
#pragma unroll
                for (int k = 0; k < 8; ++k) s[k] = fmin_(s[k], s[k + 8]);
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) acc += s[k];
#pragma unroll
    for (int k = 0; k < 16; ++k) { float lo, hi; unpack2(p[k], lo, hi); acc += lo + hi; }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += m[k];
    if (acc == 1234.5678f) out[0] = acc;
}

template <int MODE> void run(const char* name, double mufu_per_it, double fma_lanes_per_it, double instr_per_it) {
    float* out; cudaMalloc(&out, 64);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int ctas = sms * 8, iters = 20000;
    probe<MODE><<<ctas, 256>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<ctas, 256>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double thr_it = (double)ctas * 256 * iters / (ms * 1e-3);       // thread-iterations per second
    const double clk = 1.965e9;
    printf("%-44s %8.3f ms | per SM per clk: MUFU lanes %6.2f  FMA lanes %7.2f  warp-instr %5.2f\n", name, ms,
           thr_it * mufu_per_it / sms / clk, thr_it * fma_lanes_per_it / sms / clk, thr_it * instr_per_it / 32 / sms / clk);
    cudaFree(out);
}

int main() {
    run<0>("16 scalar FFMA", 0, 16, 16);
    run<1>("8 FFMA2", 0, 16, 8);
    run<2>("8 MUFU + 16 scalar FFMA", 8, 16, 24);
    run<3>("8 MUFU + 8 FFMA2", 8, 16, 16);
    run<4>("8 MUFU + 16 FFMA2", 8, 32, 24);
    run<5>("8 MUFU + 8 FFMA2 + 16 scalar FFMA", 8, 32, 32);
    run<6>("8 MUFU + 32 scalar FFMA", 8, 32, 40);
    run<7>("8 MUFU + 8 FMNMX + 8 FFMA2", 8, 16, 24);
    return 0;
}

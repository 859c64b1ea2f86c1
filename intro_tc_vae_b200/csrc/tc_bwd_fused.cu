// Fused backward sweep of the TC-ELBO estimator (sm_100a): ONE recomputation of e_ijd = 2^-qc per
// log-density yields all three gradients of ops.py:80-115's autograd graph:
//     r_ijd  = (gJ_i q_ij + gP_i rho_ij e_ijd / S_id) * [q_ijd <= qmax_id]      (mask of the -50 clamp)
//     A_id   = sum_j r dl          -> grad_z        (row-local, registers)
//     CR_id  = sum_j r (2 ln2 qc - 1) -> grad_logvar (row-local, registers)
//     G_jd   = sum_i r dl ns_id    -> grad_mu       (column sum over rows: warp-partial in registers,
//                                                     NW warps reduced through shared memory, one
//                                                     red.global.add.v4.f32 per CTA and 4 dims)
// Thread mapping: the 32 lanes of a warp span the latent dims (VEC consecutive dims per lane and
// 32*VEC-wide chunk), a warp owns RI rows, a CTA NW*RI rows; columns are streamed through a 3-stage
// bulk-TMA pipeline together with the matching slice of the saved joint exponents s2_ij.  The column
// gradient is handed from the warps to the reduction through two staging buffers guarded by mbarriers,
// so warps never wait for each other at a block-wide barrier.
// All FP32-pipe work is packed f32x2 (FFMA2/FMUL2/FADD2): ~7.5 issue slots per log-density.
#include "tc_common.cuh"
#include "tc_kernels.h"
#include "tc_instr.h"

namespace tcelbo {

__device__ __forceinline__ float fset_le(float a, float b) {       // 1.0f if a <= b else 0.0f (FSET.BF)
    float y; asm("set.le.f32.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b)); return y;
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int N> struct VecLd;
template <> struct VecLd<1> { static __device__ __forceinline__ void ld(const float* p, float (&v)[1]) { v[0] = *p; }
                              static __device__ __forceinline__ void st(float* p, const float (&v)[1]) { *p = v[0]; } };
template <> struct VecLd<2> { static __device__ __forceinline__ void ld(const float* p, float (&v)[2]) {
                                  const float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y; }
                              static __device__ __forceinline__ void st(float* p, const float (&v)[2]) {
                                  *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); } };
template <> struct VecLd<4> { static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
                                  const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
                              static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
                                  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); } };

template <int DPT, int JS_, int JTMAX = 16> struct BwdGeom {
    static constexpr int VEC = DPT < 4 ? DPT : 4;
    static constexpr int NCH = DPT / VEC;
    static constexpr int DP = 32 * DPT;
    static constexpr int CH = 32 * VEC;
    static constexpr int JT = (kTileFloats / DP) > JTMAX ? JTMAX : (kTileFloats / DP);  // columns per pipeline stage
    static constexpr int JS = JS_ > JT ? JT : JS_;                                     // columns per G staging buffer
    static constexpr int GV = JS >= 4 ? 4 : JS;                                        // columns per gq vector load
    static constexpr int NP = DPT >= 2 ? DPT / 2 : 1;                                  // f32x2 pairs per row
};

// One (row, column) pair of this lane's DPT dims, packed two dims per instruction.
template <int NP, bool kWeighted, int ABL = 0>
__device__ __forceinline__ void bwd_pairs(const u64 (&mu2)[NP], const u64 (&zs2)[NP], const u64 (&ns2)[NP],
                                          const float (&qmx)[2 * NP], const u64 (&gps2)[NP], float gq, float rho,
                                          u64 (&A2)[NP], u64 (&CR2)[NP], u64 (&G2)[NP]) {
    const u64 gq2 = pack2(gq, gq);
    const u64 rho2 = pack2(rho, rho);
    const u64 nk2 = pack2(-kInvTwoLn2, -kInvTwoLn2);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const u64 dl2 = ffma2(mu2[p], ns2[p], zs2[p]);
        const u64 q2 = ffma2(dl2, dl2, nk2);                                               // q' = q - 1/(2 ln2)
        float q0, q1;
        unpack2(q2, q0, q1);
        const float c0 = fmin_nan(q0, qmx[2 * p]), c1 = fmin_nan(q1, qmx[2 * p + 1]);      // c' = min(q, qmax) - 1/(2 ln2)
        u64 e2 = (ABL == 2) ? pack2(1.0f - c0, 1.0f - c1) : pack2(ex2(-c0), ex2(-c1));     // ABL 2: no MUFU (timing ablation only)
        if (kWeighted) e2 = fmul2(e2, rho2);
        const u64 coef2 = ffma2(e2, gps2[p], gq2);                                         // gps' carries the 2^(-1/(2 ln2)) of e'
        const u64 m2 = pack2(fset_le(q0, qmx[2 * p]), fset_le(q1, qmx[2 * p + 1]));
        const u64 r2 = fmul2(coef2, m2);
        const u64 t2 = fmul2(r2, dl2);
        A2[p] = fadd2(A2[p], t2);
        CR2[p] = ffma2(r2, pack2(c0, c1), CR2[p]);                                         // sum r c' = sum r (2 ln2 c - 1) / (2 ln2)
        G2[p] = ffma2(t2, ns2[p], G2[p]);
    }
}

// Scalar, predicated form of the same arithmetic.  On B200 a packed FFMA2 with three distinct register pairs costs
// ~3.1 dispatch cycles and the mask needs an extra FMUL2, whereas scalar 3-operand FFMAs cost ~1.25 and the clamp mask
// can ride on a predicate (tools/rf_probe.cu, profiles/r1_bwd_variant_sweep.md): ~12 instead of ~14.2 cycles per element.
template <bool kWeighted>
__device__ __forceinline__ void bwd_elem_pred(float mu, float zs, float ns, float qmx, float gps, float gq, float rho,
                                              float& A, float& CR, float& G) {
    const float dl = fmaf(mu, ns, zs);
    const float q = fmaf(dl, dl, -kInvTwoLn2);
    const float c = fmin_nan(q, qmx);
    float e = ex2(-c);
    if (kWeighted) e *= rho;
    const float coef = fmaf(e, gps, gq);
    const float t = coef * dl;
    const float u = c;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.leu.f32 p, %3, %4;\n\t"                 // unmasked (or NaN: propagate)
        "@p add.f32 %0, %0, %5;\n\t"
        "@p fma.rn.f32 %1, %6, %7, %1;\n\t"
        "@p fma.rn.f32 %2, %5, %8, %2;\n\t}"
        : "+f"(A), "+f"(CR), "+f"(G) : "f"(q), "f"(qmx), "f"(t), "f"(coef), "f"(u), "f"(ns));
}

// Same arithmetic for a group of RG rows in two phases: every MUFU of the group is issued before any
// of its results is consumed, so the ex2 latency overlaps the other rows' FP32 work inside one warp.
template <int NP, int RG, bool kWeighted>
__device__ __forceinline__ void bwd_rows_phased(const u64 (&mu2)[NP], const u64 (*zs2)[NP], const u64 (*ns2)[NP],
                                                const float (*qmx)[2 * NP], const u64 (*gps2)[NP],
                                                const float* gq, int gq_stride, const float* rho,
                                                u64 (*A2)[NP], u64 (*CR2)[NP], u64 (&G2)[NP]) {
    const u64 nk2 = pack2(-kInvTwoLn2, -kInvTwoLn2);
    u64 dl2[RG][NP], e2[RG][NP], m2[RG][NP], u2[RG][NP];
#pragma unroll
    for (int r = 0; r < RG; ++r) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            dl2[r][p] = ffma2(mu2[p], ns2[r][p], zs2[r][p]);
            const u64 q2 = ffma2(dl2[r][p], dl2[r][p], nk2);
            float q0, q1;
            unpack2(q2, q0, q1);
            const float c0 = fmin_nan(q0, qmx[r][2 * p]), c1 = fmin_nan(q1, qmx[r][2 * p + 1]);
            e2[r][p] = pack2(ex2(-c0), ex2(-c1));
            m2[r][p] = pack2(fset_le(q0, qmx[r][2 * p]), fset_le(q1, qmx[r][2 * p + 1]));
            u2[r][p] = pack2(c0, c1);
        }
    }
#pragma unroll
    for (int r = 0; r < RG; ++r) {
        const float g = gq[r * gq_stride];
        const u64 gq2 = pack2(g, g);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            u64 e = e2[r][p];
            if (kWeighted) e = fmul2(e, pack2(rho[r], rho[r]));
            const u64 coef2 = ffma2(e, gps2[r][p], gq2);
            const u64 r2 = fmul2(coef2, m2[r][p]);
            const u64 t2 = fmul2(r2, dl2[r][p]);
            A2[r][p] = fadd2(A2[r][p], t2);
            CR2[r][p] = ffma2(r2, u2[r][p], CR2[r][p]);
            G2[p] = ffma2(t2, ns2[r][p], G2[p]);
        }
    }
}

// All JS columns of one staging sub-tile for this warp's RI rows.  Templated on the special-tile case so
// the GV unrolled columns form ONE basic block and independent (row, column) chains can be interleaved.
template <int DPT, int RI, int JS_, bool PHASED, bool kSpecial, int ABL = 0, bool SCALAR = false, int JTMAX = 16>
__device__ __forceinline__ void bwd_subtile(const float* __restrict__ tile, const float* __restrict__ gq, float* __restrict__ gst,
                                            int sub, int lane, int jt0, int i_glob0, const Weights& w,
                                            const u64 (&zs2)[RI][BwdGeom<DPT, JS_, JTMAX>::NP], const u64 (&ns2)[RI][BwdGeom<DPT, JS_, JTMAX>::NP],
                                            const float (&qmx)[RI][2 * BwdGeom<DPT, JS_, JTMAX>::NP], const u64 (&gps2)[RI][BwdGeom<DPT, JS_, JTMAX>::NP],
                                            u64 (&A2)[RI][BwdGeom<DPT, JS_, JTMAX>::NP], u64 (&CR2)[RI][BwdGeom<DPT, JS_, JTMAX>::NP]) {
    using GEO = BwdGeom<DPT, JS_, JTMAX>;
    constexpr int VEC = GEO::VEC, NCH = GEO::NCH, DP = GEO::DP, CH = GEO::CH, JT = GEO::JT, JS = GEO::JS, GV = GEO::GV, NP = GEO::NP;
    constexpr int RG = (RI % 2 == 0) ? 2 : 1;
    u64 mu_next[BwdGeom<DPT, JS_, JTMAX>::NP];
#pragma unroll(ABL == 5 ? 1 : 4)
    for (int g0 = 0; g0 < JS; g0 += GV) {
        float gqv[RI][GV];
#pragma unroll
        for (int r = 0; r < RI; ++r) {
            if (ABL == 3) { for (int u = 0; u < GV; ++u) gqv[r][u] = 1e-3f * (float)(r + 1); }      // ABL 3: no joint-coefficient loads
            else VecLd<GV>::ld(gq + r * JT + sub + g0, gqv[r]);
        }
        auto load_mu = [&](int col, u64 (&dst)[NP]) {
            float vm[DPT];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float v[VEC];
                VecLd<VEC>::ld(tile + col * DP + c * CH + VEC * lane, v);
#pragma unroll
                for (int e = 0; e < VEC; ++e) vm[c * VEC + e] = v[e];
            }
#pragma unroll
            for (int p = 0; p < NP; ++p) dst[p] = pack2(vm[2 * p], DPT >= 2 ? vm[(2 * p + 1) % DPT] : 0.0f);
        };
        // column u+1's operand is loaded while column u computes: the LDS latency (29 cycles, 14 % of the stall samples when
        // loaded just in time) hides behind a whole column of math.  The group loop is fully unrolled so that all JS columns
        // of the sub-tile form one basic block and the prefetch chain runs through it (ABL 5, control: one group per
        // iteration, prefetch restarted per group: +3.5 % time).
        if (ABL == 5 || g0 == 0) load_mu(sub + g0, mu_next);
#pragma unroll
        for (int u = 0; u < GV; ++u) {
            const int jj = sub + g0 + u;
            u64 mu2[NP], G2[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) mu2[p] = mu_next[p];
            if (u + 1 < GV || (ABL != 5 && g0 + GV < JS)) load_mu(jj + 1, mu_next);
#pragma unroll
            for (int p = 0; p < NP; ++p) G2[p] = 0ull;
            float rho[RI];
#pragma unroll
            for (int r = 0; r < RI; ++r) {
                rho[r] = 1.0f;
                if (kSpecial) { float l2; weight_of(w, i_glob0 + r, jt0 + jj, rho[r], l2); }
            }
            if (SCALAR) {
#pragma unroll
                for (int r = 0; r < RI; ++r) {
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        float m0, m1, z0, z1, n0, n1, g0_, g1_, a0, a1, c0, c1, G0, G1;
                        unpack2(mu2[p], m0, m1); unpack2(zs2[r][p], z0, z1); unpack2(ns2[r][p], n0, n1);
                        unpack2(gps2[r][p], g0_, g1_); unpack2(A2[r][p], a0, a1); unpack2(CR2[r][p], c0, c1); unpack2(G2[p], G0, G1);
                        bwd_elem_pred<kSpecial>(m0, z0, n0, qmx[r][2 * p], g0_, gqv[r][u], rho[r], a0, c0, G0);
                        bwd_elem_pred<kSpecial>(m1, z1, n1, qmx[r][2 * p + 1], g1_, gqv[r][u], rho[r], a1, c1, G1);
                        A2[r][p] = pack2(a0, a1); CR2[r][p] = pack2(c0, c1); G2[p] = pack2(G0, G1);
                    }
                }
            } else if (PHASED) {
#pragma unroll
                for (int rg = 0; rg < RI; rg += RG)
                    bwd_rows_phased<NP, RG, kSpecial>(mu2, &zs2[rg], &ns2[rg], &qmx[rg], &gps2[rg], &gqv[rg][u], GV, &rho[rg],
                                                      &A2[rg], &CR2[rg], G2);
            } else {
#pragma unroll
                for (int r = 0; r < RI; ++r)
                    bwd_pairs<NP, kSpecial, ABL>(mu2, zs2[r], ns2[r], qmx[r], gps2[r], gqv[r][u], rho[r], A2[r], CR2[r], G2);
            }
            if (ABL == 1) {                                        // ABL 1: fold G into A, no staging (timing ablation only)
#pragma unroll
                for (int p = 0; p < NP; ++p) A2[0][p] = fadd2(A2[0][p], G2[p]);
                continue;
            }
            // warp-partial column gradient -> staging buffer [warp][column][dim]
            float vg[DPT];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                float lo, hi; unpack2(G2[p], lo, hi);
                vg[2 * p % DPT] = lo; if (DPT >= 2) vg[(2 * p + 1) % DPT] = hi;
            }
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float v[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) v[e] = vg[c * VEC + e];
                VecLd<VEC>::st(gst + (g0 + u) * DP + c * CH + VEC * lane, v);
            }
        }
    }
}

template <int DPT, int RI, int NW, int MINB, int JS_, bool PHASED, int NBUF, int ABL, bool SCALAR, int JTMAX>
__global__ void __launch_bounds__(NW * 32, MINB)
tc_bwd_fused_kernel(const BwdFusedArgs a) {
    using GEO = BwdGeom<DPT, JS_, JTMAX>;
    constexpr int VEC = GEO::VEC, NCH = GEO::NCH, DP = GEO::DP, CH = GEO::CH, JT = GEO::JT, JS = GEO::JS, GV = GEO::GV, NP = GEO::NP;
    constexpr int TILE = JT * DP;
    constexpr int ROWS = NW * RI;
    constexpr int GST = NW * JS * DP;                                        // floats per G staging buffer
    constexpr int RG = (RI % 2 == 0) ? 2 : 1;                                // rows per phase group
    static_assert(DPT == 1 || DPT % 2 == 0, "dims per lane must pair up");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* mu_tiles = reinterpret_cast<float*>(smem_raw);                    // [kStages][TILE]
    float* s2_tiles = mu_tiles + (size_t)kStages * TILE;                     // [kStages][ROWS][JT]
    float* gq_buf = s2_tiles + (size_t)kStages * ROWS * JT;                  // [NW][RI][JT]
    float* gstage = gq_buf + (size_t)NW * RI * JT;                           // [NBUF][NW][JS][DP]
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(gstage + NBUF * (size_t)GST);
    uint64_t* bar_empty = bar_full + kStages;
    uint64_t* g_full = bar_empty + kStages;                                  // [NBUF] staging buffer written by all warps
    uint64_t* g_empty = g_full + NBUF;                                       // [NBUF] staging buffer reduced by all warps

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Work = (latent slice, row block) blocks x T column tiles, linearised block-major and cut into equal contiguous
    // segments of a.seg.base (+1) tiles, one per CTA (tc_layout.h: plan_segments): every CTA carries the same load whatever
    // the batch shape; a segment that crosses a block boundary flushes its row-local sums and reloads the row constants.
    // A block covers DP dims starting at d0 of rows whose pitch is a.pitch floats: for D > 128 the blocks walk 128-dim
    // slices (row-local sums and column sums are independent per dim; only the joint coefficients are shared).
    const int pitch = a.pitch;
    const int T = a.bg_pad / JT;
    const int64_t g_begin = seg_begin(a.seg, blockIdx.x);
    const int ntiles = seg_len(a.seg, blockIdx.x);
    int q = (int)(g_begin / T);                                              // current block = slice * n_rb + row block
    const int t_first = (int)(g_begin - (int64_t)q * T);                     // first column tile inside it
    int rb0 = (q % a.n_rb) * ROWS;                                           // first row of the block
    int row0 = rb0 + warp * RI;
    int d0 = (q / a.n_rb) * DP;

    // ---- row constants and accumulators of this warp's RI rows (this lane's dims) -> registers.
    //      Rows past the padded batch are clamped for loads and carry zero coefficients.
    u64 zs2[RI][NP], ns2[RI][NP], gps2[RI][NP], A2[RI][NP], CR2[RI][NP];
    float qmx[RI][2 * NP];
    auto load_rows = [&]() {
#pragma unroll
        for (int r = 0; r < RI; ++r) {
            const bool valid = (row0 + r) < a.bl_pad;
            const size_t base = (size_t)min(row0 + r, a.bl_pad - 1) * pitch + d0;
            float vz[DPT], vn[DPT], vq[DPT], vg[DPT];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float v[VEC];
                VecLd<VEC>::ld(a.zs + base + c * CH + VEC * lane, v);
#pragma unroll
                for (int e = 0; e < VEC; ++e) vz[c * VEC + e] = v[e];
                VecLd<VEC>::ld(a.ns + base + c * CH + VEC * lane, v);
#pragma unroll
                for (int e = 0; e < VEC; ++e) vn[c * VEC + e] = v[e];
                // The sweep works with q' = q - 1/(2 ln2): then 2 ln2 q - 1 = 2 ln2 q' and the logvar sum needs no extra FFMA;
                // 2^-q' = 2^-q * exp(1/2); the exp(-1/2) is folded into gps.
                VecLd<VEC>::ld(a.qmax + base + c * CH + VEC * lane, v);
#pragma unroll
                for (int e = 0; e < VEC; ++e) vq[c * VEC + e] = v[e] - kInvTwoLn2;
                VecLd<VEC>::ld(a.gps + base + c * CH + VEC * lane, v);
#pragma unroll
                for (int e = 0; e < VEC; ++e) vg[c * VEC + e] = valid ? v[e] * kRsqrtE : 0.0f;
            }
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const int e0 = 2 * p, e1 = (DPT >= 2) ? 2 * p + 1 : 0;
                zs2[r][p] = pack2(vz[e0], DPT >= 2 ? vz[e1] : 0.0f);
                ns2[r][p] = pack2(vn[e0], DPT >= 2 ? vn[e1] : 0.0f);
                gps2[r][p] = pack2(vg[e0], DPT >= 2 ? vg[e1] : 0.0f);
                qmx[r][2 * p] = vq[e0];
                qmx[r][2 * p + 1] = DPT >= 2 ? vq[e1] : 0.0f;
                A2[r][p] = 0ull; CR2[r][p] = 0ull;
            }
        }
    };
    // row-local partial sums of the current block -> slot (this CTA's ordinal among the segments that touch the block)
    auto flush_rows = [&]() {
        const int slot = (int)blockIdx.x - seg_of(a.seg, (int64_t)q * T);
#pragma unroll
        for (int r = 0; r < RI; ++r) {
            if (row0 + r >= a.bl_pad) continue;
            const size_t base = ((size_t)slot * a.bl_pad + row0 + r) * pitch + d0;
            float va[DPT], vc[DPT];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                float lo, hi;
                unpack2(A2[r][p], lo, hi); va[2 * p % DPT] = lo; if (DPT >= 2) va[(2 * p + 1) % DPT] = hi;
                unpack2(CR2[r][p], lo, hi); vc[2 * p % DPT] = lo; if (DPT >= 2) vc[(2 * p + 1) % DPT] = hi;
            }
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float v[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) v[e] = va[c * VEC + e];
                VecLd<VEC>::st(a.Apart + base + c * CH + VEC * lane, v);
#pragma unroll
                for (int e = 0; e < VEC; ++e) v[e] = vc[c * VEC + e];
                VecLd<VEC>::st(a.CRpart + base + c * CH + VEC * lane, v);
            }
        }
    };
    load_rows();

    constexpr uint32_t kTxBytes = (TILE + ROWS * JT) * sizeof(float);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], NW); }
        for (int s = 0; s < NBUF; ++s) { mbar_init(&g_full[s], NW); mbar_init(&g_empty[s], NW); }
        mbar_fence_init();
    }
    __syncthreads();

    // warp 0, all lanes; t = tile ordinal in the segment, tn = its column tile inside block (first row rbn, slice offset dn)
    auto issue = [&](int t, int tn, int rbn, int dn) {
        const int sn = t % kStages;
        const int jn = tn * JT;
        if (lane == 0) {
            if (t >= kStages) mbar_wait(&bar_empty[sn], ((t / kStages) - 1) & 1);
            mbar_arrive_expect_tx(&bar_full[sn], kTxBytes);
            if (pitch == DP)                                                 // whole rows: one contiguous bulk copy
                bulk_g2s(mu_tiles + (size_t)sn * TILE, a.mu_pad + (size_t)jn * DP, TILE * sizeof(float), &bar_full[sn]);
        }
        __syncwarp();
        if (pitch != DP) {                                                   // a 128-dim slice of wider rows: one copy per column
            for (int c = lane; c < JT; c += 32)
                bulk_g2s(mu_tiles + (size_t)sn * TILE + (size_t)c * DP, a.mu_pad + (size_t)(jn + c) * pitch + dn,
                         DP * sizeof(float), &bar_full[sn]);
        }
        for (int r = lane; r < ROWS; r += 32)
            bulk_g2s(s2_tiles + ((size_t)sn * ROWS + r) * JT,
                     a.s2 + (size_t)min(rbn + r, a.bl_pad - 1) * a.ld_s2 + jn, JT * sizeof(float), &bar_full[sn]);
    };
    if (warp == 0 && ntiles > 0) issue(0, t_first, rb0, d0);

    // reduce this thread's share of staging buffer (k & 1) (sub-tile k, columns starting at col0) into Gacc
    // With NBUF >= 3 a warp may run a whole sub-tile ahead of the slowest warp of its CTA before it blocks here.
    auto reduce_share = [&](int k, int col0, int dd0) {
        const int b = k % NBUF;
        mbar_wait(&g_full[b], (k / NBUF) & 1);
        const float* gsb = gstage + (size_t)b * GST;
        for (int f = threadIdx.x; f < JS * DP / 4; f += NW * 32) {
            const int col = f / (DP / 4), chunk = f % (DP / 4);
            float4 acc = *reinterpret_cast<const float4*>(gsb + col * DP + 4 * chunk);
#pragma unroll
            for (int w = 1; w < NW; ++w) {
                const float4 v = *reinterpret_cast<const float4*>(gsb + ((size_t)w * JS + col) * DP + 4 * chunk);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            red_add_v4(a.Gacc + (size_t)(col0 + col) * pitch + dd0 + 4 * chunk, acc.x, acc.y, acc.z, acc.w);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&g_empty[b]);
    };

    float* gq = gq_buf + (size_t)warp * RI * JT;
    int k = 0, prev_col0 = 0, prev_d0 = 0;                                   // running sub-tile index
    int t_in = t_first;                                                      // column tile inside the current block
    for (int t = 0; t < ntiles; ++t, ++t_in) {
        if (t_in == T) {                                                     // segment crosses into the next block
            flush_rows();
            ++q; t_in = 0;
            rb0 = (q % a.n_rb) * ROWS; row0 = rb0 + warp * RI; d0 = (q / a.n_rb) * DP;
            load_rows();
        }
        const int st = t % kStages;
        if (warp == 0 && t + 1 < ntiles) {                                   // prefetch the next tile (possibly of the next block)
            int tn = t_in + 1, rbn = rb0, dn = d0;
            if (tn == T) { tn = 0; rbn = ((q + 1) % a.n_rb) * ROWS; dn = ((q + 1) / a.n_rb) * DP; }
            issue(t + 1, tn, rbn, dn);
        }
        mbar_wait(&bar_full[st], (t / kStages) & 1);
        const float* tile = mu_tiles + (size_t)st * TILE;
        const float* s2t = s2_tiles + ((size_t)st * ROWS + warp * RI) * JT;
        const int jt0 = t_in * JT;
        const bool special = (a.w.mss && jt0 == 0) || (jt0 + JT > a.w.b_glob);

        // joint-term coefficients gJ_i * q_ij of this warp's rows for the tile
        __syncwarp();
        for (int idx = lane; idx < RI * JT; idx += 32) {
            const int r = idx / JT, jj = idx % JT;
            const int row = row0 + r;
            float rho = 1.0f, l2 = 0.0f;
            if (special) weight_of(a.w, a.row_offset + row, jt0 + jj, rho, l2);
            const int rowc = min(row, a.bl_pad - 1);
            const float gjr = __ldg(a.gj + rowc), j2r = __ldg(a.J2 + rowc);
            const float qv = ex2(l2 - s2t[idx] - j2r);
            gq[idx] = (jt0 + jj < a.w.b_glob && row < a.b_loc) ? gjr * qv : 0.0f;
        }
        __syncwarp();

        for (int sub = 0; sub < JT; sub += JS, ++k) {
            const int b = k % NBUF;
            if (ABL != 1 && k >= NBUF) mbar_wait(&g_empty[b], ((k / NBUF) - 1) & 1);   // everyone finished reducing sub-tile k-NBUF
            float* gst = gstage + (size_t)b * GST + (size_t)warp * JS * DP;
            if (special) bwd_subtile<DPT, RI, JS_, PHASED, true, ABL, SCALAR, JTMAX>(tile, gq, gst, sub, lane, jt0, a.row_offset + row0, a.w, zs2, ns2, qmx, gps2, A2, CR2);
            else         bwd_subtile<DPT, RI, JS_, PHASED, false, ABL, SCALAR, JTMAX>(tile, gq, gst, sub, lane, jt0, a.row_offset + row0, a.w, zs2, ns2, qmx, gps2, A2, CR2);
            if (ABL == 1) continue;                                  // timing ablation: no staging hand-off / reduction
            __syncwarp();
            if (lane == 0) mbar_arrive(&g_full[b]);
            if (k >= 1) reduce_share(k - 1, prev_col0, prev_d0);             // the other buffer: its writers are long done
            prev_col0 = jt0 + sub; prev_d0 = d0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[st]);
    }
    if (ABL != 1 && k >= 1) reduce_share(k - 1, prev_col0, prev_d0);
    flush_rows();
}

// Elementwise over [B,D]: sum the column-split partials of the row-local sums (8 independent loads in flight), scale,
// and add the fused KL gradient when asked (ops.py:161-163).
__global__ void bwd_fused_finalize_kernel(const BwdFinArgs a) {
    const int64_t n_row = (int64_t)a.b_loc * a.d;
    const int64_t n_col = a.scratch_parts != nullptr ? 0 : (int64_t)a.b_glob * a.d;
    const size_t split_stride = (size_t)a.bl_pad * a.dp;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_row + n_col; idx += (int64_t)gridDim.x * blockDim.x) {
        if (idx < n_row) {
            const int i = (int)(idx / a.d), dd = (int)(idx % a.d);
            const size_t o = (size_t)i * a.dp + dd;
            if (a.scratch_parts != nullptr) {
                // reduce-scatter of the column gradient as this kernel's load phase: the local rows of every rank's accumulator
                const size_t og = (size_t)(a.row_offset + i) * a.dp + dd;
                float g = 0.0f;
                for (int p0 = 0; p0 < a.n_ranks; p0 += 8) {
                    float v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        v[k] = (p0 + k < a.n_ranks)
                                   ? *reinterpret_cast<const volatile float*>(static_cast<const char*>(a.scratch_parts[p0 + k]) + a.g_off + og * sizeof(float))
                                   : 0.0f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) g += v[k];
                }
                g *= -kTwoLn2;
                if (a.gk != nullptr) g += a.gk[i] * a.mu_all[(int64_t)i * a.ldmu + dd];      // mu_all = this rank's rows here
                a.grad_mu[(int64_t)i * a.ldgmu + dd] = g;
            }
            float sa = 0.0f, sc = 0.0f;
            // partial slots of this (slice, row block): one per segment that touches the block
            const int64_t qb = (int64_t)(dd / a.slice_dp) * a.n_rb + i / a.rows_per_block;
            const int n_slots = seg_slots(a.seg, qb, a.tiles_per_block);
            for (int s0 = 0; s0 < n_slots; s0 += 4) {
                float va[4], vc[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool ok = s0 + k < n_slots;
                    va[k] = ok ? __ldg(a.Apart + (size_t)(s0 + k) * split_stride + o) : 0.0f;
                    vc[k] = ok ? __ldg(a.CRpart + (size_t)(s0 + k) * split_stride + o) : 0.0f;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) { sa += va[k]; sc += vc[k]; }
            }
            float glv = a.vr[o] * (kTwoLn2 * sc);                         // the sweep accumulates sum r (c - 1/(2 ln2))
            if (a.gk != nullptr) glv += a.gk[i] * 0.5f * (expf(a.lv[(int64_t)i * a.ldlv + dd]) - 1.0f);
            a.grad_z[(int64_t)i * a.ldgz + dd] = kTwoLn2 * a.ns[o] * sa;
            a.grad_lv[(int64_t)i * a.ldglv + dd] = glv;
        } else {
            const int64_t k = idx - n_row;
            const int j = (int)(k / a.d), dd = (int)(k % a.d);
            float g = -kTwoLn2 * a.Gpart[(size_t)j * a.dp + dd];
            const int i = j - a.row_offset;
            if (a.gk != nullptr && i >= 0 && i < a.b_loc) g += a.gk[i] * a.mu_all[(int64_t)j * a.ldmu + dd];
            a.grad_mu[(int64_t)j * a.ldgmu + dd] = g;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// launch: the column split is planned here because it depends on the variant's CTA shape
// ------------------------------------------------------------------------------------------------------
static int g_bwd_variant = -1;      // -1: default per shape; set through tcelbo_set_tuning("bwd_variant", v)
void set_bwd_variant(int v) { g_bwd_variant = v; }
static int g_bwd_seg_target = 0;    // 0: default target segment length (column tiles per CTA)
void set_bwd_seg_target(int v) { g_bwd_seg_target = v; }

template <int DPT, int RI, int NW, int MINB, int JS_, bool PHASED, int NBUF = 2, int ABL = 0, bool SCALAR = false, int JTMAX = 16>
static cudaError_t launch_bwd_fused_t(const Plan& p, BwdFusedArgs a, BwdFinArgs* fin, cudaStream_t st) {
    using GEO = BwdGeom<DPT, JS_, JTMAX>;
    constexpr int ROWS = NW * RI;
    const size_t smem = ((size_t)kStages * GEO::JT * GEO::DP + (size_t)kStages * ROWS * GEO::JT + (size_t)NW * RI * GEO::JT
                         + NBUF * (size_t)NW * GEO::JS * GEO::DP) * sizeof(float) + (2 * kStages + 2 * NBUF) * sizeof(uint64_t);
    static PerDevice ctas_on;
    int& ctas_per_sm = ctas_on.cur();
    if (ctas_per_sm == 0) {
        auto kern = tc_bwd_fused_kernel<DPT, RI, NW, MINB, JS_, PHASED, NBUF, ABL, SCALAR, JTMAX>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NW * 32, smem);
        if (e != cudaSuccess) return e;
        ctas_per_sm = occ > 0 ? occ : 1;
    }
    const int n_rb = (p.bl_pad + ROWS - 1) / ROWS;
    const int n_slices = p.dp / GEO::DP;                                      // 1 unless a wide latent is walked in 128-dim slices
    const int T = p.bg_pad / GEO::JT;
    const Segments seg = plan_segments((int64_t)n_rb * n_slices, T, p.sms * ctas_per_sm, g_bwd_seg_target > 0 ? g_bwd_seg_target : 1024 / GEO::JT);
    a.js_len = 0;
    a.pitch = p.dp;
    a.seg = seg; a.n_blocks = n_rb * n_slices; a.n_rb = n_rb;
    fin->seg = seg; fin->tiles_per_block = T; fin->n_rb = n_rb; fin->rows_per_block = ROWS; fin->slice_dp = GEO::DP;
    if (a.plan_only) return cudaSuccess;
    LaunchScope scope(kKernBwdRow, st);
    tc_bwd_fused_kernel<DPT, RI, NW, MINB, JS_, PHASED, NBUF, ABL, SCALAR, JTMAX><<<seg.n_ctas, NW * 32, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_bwd_fused(const Plan& p, const BwdFusedArgs& a, BwdFinArgs* fin, cudaStream_t st) {
    if (g_bwd_variant >= 40 && g_bwd_variant < 50) return launch_bwd_ds(p, a, fin, g_bwd_variant - 40, st);
    switch (p.dpt) {
        case 1:  return launch_bwd_fused_t<1, 4, 12, 1, 8, false>(p, a, fin, st);
        case 2:  return launch_bwd_fused_t<2, 4, 12, 1, 8, false>(p, a, fin, st);
        case 4:
            switch (g_bwd_variant) {                 // tuning matrix for the headline shape (D = 128); see profiles/r1_bwd_variant_sweep.md
                case 0:  return launch_bwd_fused_t<4, 4, 12, 1, 8, false>(p, a, fin, st);            // 12 warps x 168 regs, 1 CTA/SM
                case 1:  return launch_bwd_fused_t<4, 4, 12, 1, 8, true >(p, a, fin, st);            //   + two-phase loop body
                case 2:  return launch_bwd_fused_t<4, 4, 8, 2, 8, false>(p, a, fin, st);             // 4 rows/warp, 2 CTAs/SM (spills)
                case 4:  return launch_bwd_fused_t<4, 2, 8, 3, 4, false>(p, a, fin, st);             // 2 rows/warp, 3 CTAs/SM
                case 11: return launch_bwd_fused_t<4, 3, 8, 2, 4, false, 3>(p, a, fin, st);          // three staging buffers
                case 12: return launch_bwd_fused_t<4, 3, 8, 2, 8, false, 3, 0, false, 8>(p, a, fin, st);   // three 8-column staging buffers, 8-column tiles
                case 13: return launch_bwd_fused_t<4, 3, 8, 2, 8, false, 2, 0, false, 8>(p, a, fin, st);   // 8-column tiles only (control for 12)
                case 14: return launch_bwd_fused_t<4, 3, 8, 2, 8, false, 2, 5>(p, a, fin, st);             // control: 4-column basic blocks, prefetch restarted per group
                case 30: return launch_bwd_fused_t<4, 3, 8, 2, 8, false, 2, 0, true>(p, a, fin, st); // scalar predicated loop
#ifdef TCELBO_ABLATIONS                              // timing ablations return WRONG gradients: never in the shipped library
                case 20: return launch_bwd_fused_t<4, 3, 8, 2, 8, false, 2, 1>(p, a, fin, st);       // ablation: no column-gradient path
                case 21: return launch_bwd_fused_t<4, 3, 8, 2, 8, false, 2, 2>(p, a, fin, st);       // ablation: no MUFU
                case 22: return launch_bwd_fused_t<4, 3, 8, 2, 8, false, 2, 3>(p, a, fin, st);       // ablation: no joint-coefficient loads
#endif
                default: return launch_bwd_fused_t<4, 3, 8, 2, 8, false>(p, a, fin, st);             // best of the sweep
            }
        case 8: case 16:                             // D = 256 / 512: the tuned 128-dim kernel over 2 / 4 slices (grid.z)
            return launch_bwd_fused_t<4, 3, 8, 2, 8, false>(p, a, fin, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_bwd_fused_finalize(const Plan& p, const BwdFinArgs& a, cudaStream_t st) {
    const int64_t n = (int64_t)p.b_loc * p.d + (int64_t)p.b_glob * p.d;
    int64_t g = (n + 255) / 256; if (g > 148 * 16) g = 148 * 16; if (g < 1) g = 1;
    LaunchScope scope(kKernNone, st);
    bwd_fused_finalize_kernel<<<(int)g, 256, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace tcelbo

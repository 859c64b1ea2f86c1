// Host-side plan of one TC-ELBO evaluation: padded sizes, work split and workspace layout.
// Shared by tcelbo_workspace_bytes / tcelbo_forward / tcelbo_backward so the three always agree.
#pragma once
#include <cstddef>
#include <cstdint>

namespace tcelbo {

constexpr int kTileFloats = 4096;     // floats of column data per pipeline stage (16 KiB)
constexpr int kStages     = 3;        // bulk-copy pipeline depth
constexpr int kRowPad     = 128;      // b_loc / b_glob are padded to a multiple of this
constexpr int kFwdWarps   = 4;
constexpr int kBwdWarps   = 8;

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

struct Plan {
    // padded problem
    int d, dp, dpt;            // dp = power of two >= max(d, 32); dpt = dp / 32 (floats per lane / lanes per row)
    int b_loc, b_glob, bl_pad, bg_pad;
    int jt;                    // column-tile height (rows of mu per stage) = kTileFloats / dp
    // forward / backward-row pass: grid (n_rb, n_js)
    int fwd_rows, n_rb_fwd, n_js_fwd, js_len_fwd;
    int bwr_rows, bwr_ri, n_rb_bwr, n_js_bwr, js_len_bwr;
    int bwf_rows, n_js_bwf, js_len_bwf;   // fused backward sweep: 12 warps x bwr_ri rows per CTA, one CTA per SM
    // backward-column pass: grid (n_cb, n_is)
    int bwc_cols, bwc_rj, bwc_it, n_cb, n_is, is_len;
    // workspace (byte offsets)
    size_t off_mu, off_zs, off_ns, off_qmax, off_shift, off_vr;   // [bg_pad|bl_pad][dp]
    size_t off_S, off_J2;                                         // persistent forward results
    size_t off_s2; int64_t ld_s2;                                 // joint exponents [bl_pad][bg_pad]
    size_t off_scratch;                                           // forward split partials (S, J)
    size_t total_bytes;                                           // forward workspace (read-only in backward)
    // backward scratch (its own buffer, so the forward workspace stays immutable and backward can be re-run)
    size_t boff_gps, boff_gj, boff_A, boff_CR, boff_G, bwd_bytes;
    bool save;
};

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// `sms` = multiprocessor count of the current device (148 on B200).
inline bool make_plan(Plan& p, int b_loc, int b_glob, int d, uint32_t flags, int sms) {
    if (b_loc < 1 || b_glob < 1 || d < 1 || d > 512) return false;
    p.d = d;
    int dp = 32; while (dp < d) dp <<= 1;
    p.dp = dp; p.dpt = dp / 32;
    p.b_loc = b_loc; p.b_glob = b_glob;
    p.bl_pad = (int)round_up(b_loc, kRowPad);
    p.bg_pad = (int)round_up(b_glob, kRowPad);
    p.jt = kTileFloats / dp;                                   // 128 .. 8
    if (p.jt > 32) p.jt = 32;
    p.save = (flags & 4u) != 0;

    // ---- forward: a CTA owns fwd_rows rows and a contiguous range of js_len columns
    p.fwd_rows = kFwdWarps * (32 / p.dpt);
    p.n_rb_fwd = p.bl_pad / p.fwd_rows;
    {
        const int slots = sms * 3;                             // resident CTAs at 3 per SM
        const int min_len = p.jt * 4;
        int max_js = p.bg_pad / min_len; if (max_js < 1) max_js = 1;
        int want = (slots * 8 + p.n_rb_fwd - 1) / p.n_rb_fwd;
        if (want < 1) want = 1; if (want > max_js) want = max_js;
        p.js_len_fwd = (int)round_up((p.bg_pad + want - 1) / want, p.jt);
        p.n_js_fwd = (p.bg_pad + p.js_len_fwd - 1) / p.js_len_fwd;
    }
    // ---- backward row pass (row-local gradients)
    p.bwr_ri = p.dpt >= 16 ? 1 : (p.dpt >= 8 ? 2 : 4);
    p.bwr_rows = kBwdWarps * p.bwr_ri;
    p.n_rb_bwr = p.bl_pad / p.bwr_rows;
    {
        const int slots = sms * 2;
        const int min_len = p.jt * 4;
        int max_js = p.bg_pad / min_len; if (max_js < 1) max_js = 1;
        int want = (slots * 8 + p.n_rb_bwr - 1) / p.n_rb_bwr;
        if (want < 1) want = 1; if (want > max_js) want = max_js;
        p.js_len_bwr = (int)round_up((p.bg_pad + want - 1) / want, p.jt);
        p.n_js_bwr = (p.bg_pad + p.js_len_bwr - 1) / p.js_len_bwr;
    }
    // ---- fused backward sweep
    p.bwf_rows = 12 * p.bwr_ri;
    {
        const int n_rb = (p.bl_pad + p.bwf_rows - 1) / p.bwf_rows;
        const int slots = sms;                                 // one resident CTA per SM
        const int min_len = p.jt * 4;
        int max_js = p.bg_pad / min_len; if (max_js < 1) max_js = 1;
        int want = (slots * 8 + n_rb - 1) / n_rb;
        if (want < 1) want = 1; if (want > max_js) want = max_js;
        p.js_len_bwf = (int)round_up((p.bg_pad + want - 1) / want, p.jt);
        p.n_js_bwf = (p.bg_pad + p.js_len_bwf - 1) / p.js_len_bwf;
    }
    // ---- backward column pass (grad_mu): a CTA owns bwc_cols columns and a range of is_len rows
    p.bwc_rj = p.dpt >= 16 ? 2 : (p.dpt >= 8 ? 4 : 8);
    p.bwc_cols = kBwdWarps * p.bwc_rj;
    p.bwc_it = 2048 / dp; if (p.bwc_it < 4) p.bwc_it = 4; if (p.bwc_it > 16) p.bwc_it = 16;
    p.n_cb = p.bg_pad / p.bwc_cols;
    {
        const int slots = sms * 2;
        const int min_len = p.bwc_it * 4;
        int max_is = p.bl_pad / min_len; if (max_is < 1) max_is = 1;
        int want = (slots * 8 + p.n_cb - 1) / p.n_cb;
        if (want < 1) want = 1; if (want > max_is) want = max_is;
        p.is_len = (int)round_up((p.bl_pad + want - 1) / want, p.bwc_it);
        p.n_is = (p.bl_pad + p.is_len - 1) / p.is_len;
    }

    // ---- workspace
    size_t off = 0;
    const size_t row_arr = (size_t)p.bl_pad * dp * sizeof(float);
    const size_t col_arr = (size_t)p.bg_pad * dp * sizeof(float);
    p.off_mu = off;    off = align256(off + col_arr);
    p.off_zs = off;    off = align256(off + row_arr);
    p.off_ns = off;    off = align256(off + row_arr);
    p.off_qmax = off;  off = align256(off + row_arr);
    p.off_shift = off; off = align256(off + row_arr);
    p.off_vr = off;    off = align256(off + row_arr);
    p.off_S = off;     off = align256(off + row_arr);
    p.off_J2 = off;    off = align256(off + (size_t)p.bl_pad * sizeof(float));
    p.ld_s2 = p.bg_pad;
    p.off_s2 = off;    off = align256(off + (p.save ? (size_t)p.bl_pad * p.ld_s2 * sizeof(float) : 0));
    const size_t fwd_scratch = (size_t)p.n_js_fwd * row_arr + (size_t)p.n_js_fwd * p.bl_pad * 2 * sizeof(float);
    p.off_scratch = off; off = align256(off + fwd_scratch);
    p.total_bytes = off;

    size_t b = 0;
    p.boff_gps = b; b = align256(b + row_arr);
    p.boff_gj = b;  b = align256(b + (size_t)p.bl_pad * sizeof(float));
    const int n_js_max = p.n_js_bwr > p.n_js_bwf ? p.n_js_bwr : p.n_js_bwf;
    p.boff_A = b;   b = align256(b + (size_t)n_js_max * row_arr);
    p.boff_CR = b;  b = align256(b + (size_t)n_js_max * row_arr);
    p.boff_G = b;   b = align256(b + (size_t)p.n_is * col_arr);
    p.bwd_bytes = b;
    return true;
}

}  // namespace tcelbo

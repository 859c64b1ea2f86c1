"""NumPy fp32 model of the arithmetic the CUDA kernels perform (TEST INFRASTRUCTURE ONLY).

This is NOT the reference's algorithm restated (that is oracle/tc_oracle.py); it is the GPU
kernels' own formulation -- base-2 exponent domain, fixed per-(row,dim) shift instead of a running
max for the per-dimension logsumexp, importance weights as ratios to the uniform weight, masked
analytic gradients -- written with dense [B,B,D] numpy arrays so the formulation can be checked
against the oracle on the CPU, in this container, before any GPU time is spent.
See DESIGN.md section "Kernel arithmetic" for the derivation; reference semantics from
ops.py:15-29,32-49,92-115 (density variants, weights, estimators).
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32
LOG2E = F(1.4426950408889634)
LN2 = F(0.6931471805599453)
LOG_2PI = F(math.log(2.0 * math.pi))
K50 = F(50.0 * 1.4426950408889634)


def weight_scalars(b_glob: int, dataset_size: int, estimator: str):
    """fp32 log-weights the way the reference forms them (ops.py:42-49: fp32 tensor, then .log())."""
    n, m = dataset_size, b_glob - 1
    if estimator == "mws":
        lw_u = F(-math.log(b_glob * dataset_size))
        return dict(lw_u=lw_u, l2r_n=F(0), l2r_s=F(0), r_n=F(1), r_s=F(1))
    if m == 0:
        raise ZeroDivisionError("float division by zero")
    with np.errstate(all="ignore"):
        lw_u = np.log(F(1.0 / m))
        lw_n = np.log(F(1.0 / n))
        lw_s = np.log(F((n - m) / (n * m)))        # NaN when N < B-1, as in the reference
        l2r_n = (lw_n - lw_u) * LOG2E
        l2r_s = (lw_s - lw_u) * LOG2E
        return dict(lw_u=F(lw_u), l2r_n=F(l2r_n), l2r_s=F(l2r_s), r_n=F(np.exp2(l2r_n)), r_s=F(np.exp2(l2r_s)))


def _ratio_matrices(rows, b_glob, ws):
    """rho_ij = w_ij / w_uniform and log2(rho_ij) for global rows ``rows`` ([R,B])."""
    r = np.ones((len(rows), b_glob), dtype=F)
    l2r = np.zeros((len(rows), b_glob), dtype=F)
    if ws["r_n"] != 1 or ws["r_s"] != 1 or np.isnan(ws["r_s"]):
        r[:, 0], l2r[:, 0] = ws["r_n"], ws["l2r_n"]
        if b_glob > 1:
            r[:, 1], l2r[:, 1] = ws["r_s"], ws["l2r_s"]
        odd = rows == b_glob - 2
        r[odd, 0], l2r[odd, 0] = ws["r_s"], ws["l2r_s"]
    return r, l2r


def row_prologue(z, lv):
    """Per-(i,d) constants of the row-variance ('active') density, ops.py:15-21."""
    z, lv = z.astype(F), lv.astype(F)
    var = np.exp(lv)
    vc = np.maximum(var, F(1e-4))
    iv = F(1) / vc
    c = F(-0.5) * (np.log(vc) + LOG_2PI)
    s = np.sqrt(F(0.5) * LOG2E * iv)
    return dict(zs=z * s, ns=-s, qmax=np.maximum(F(0), (F(50) + c) * LOG2E), shift=np.maximum(c, F(-50)),
                vr=F(0.5) * var * iv)


def forward_rowvar(z, mu_all, lv, dataset_size, estimator="mss", row_offset=0):
    """Returns dict(log_qz_prod, log_qz, S, J2, s2, ...) for the row-variance variant."""
    b_glob = mu_all.shape[0]
    rows = np.arange(row_offset, row_offset + z.shape[0])
    ws = weight_scalars(b_glob, dataset_size, estimator)
    rho, l2rho = _ratio_matrices(rows, b_glob, ws)
    p = row_prologue(z, lv)
    mu_all = mu_all.astype(F)
    dl = p["zs"][:, None, :] + p["ns"][:, None, :] * mu_all[None, :, :]          # FFMA
    q = dl * dl
    qc = np.minimum(q, p["qmax"][:, None, :])
    e = np.exp2(-qc)
    S = (rho[:, :, None] * e).sum(axis=1, dtype=F)
    L = np.log(S) + ws["lw_u"] + p["shift"]
    P = L.sum(axis=1, dtype=F)
    s2 = qc.sum(axis=2, dtype=F)
    x = l2rho - s2
    with np.errstate(all="ignore"):
        xm = x.max(axis=1, keepdims=True)
        J2 = (xm + np.log2(np.exp2(x - xm).sum(axis=1, keepdims=True, dtype=F)))[:, 0]
    C = p["shift"].sum(axis=1, dtype=F)
    J = LN2 * J2 + C + ws["lw_u"]
    return dict(log_qz_prod=P, log_qz=J, S=S, J2=J2, s2=s2, x=x, pro=p, rho=rho, dl=dl, q=q, qc=qc, e=e)


def backward_rowvar(fw, g_log_qz, g_log_qz_prod):
    """(grad_z [R,D], grad_mu_all [B,D], grad_lv [R,D]) from the saved forward quantities."""
    p = fw["pro"]
    gJ = g_log_qz.astype(F)[:, None]
    gP = g_log_qz_prod.astype(F)[:, None, None]
    qw = gJ * np.exp2(fw["x"] - fw["J2"][:, None])                                 # gJ_i * q_ij
    pw = fw["rho"][:, :, None] * fw["e"] / fw["S"][:, None, :]
    coef = qw[:, :, None] + gP * pw
    r = np.where(fw["q"] <= p["qmax"][:, None, :], coef, F(0))
    w1 = r * fw["dl"]
    A = w1.sum(axis=1, dtype=F)
    C2 = (w1 * fw["dl"]).sum(axis=1, dtype=F)
    R = r.sum(axis=1, dtype=F)
    two_ln2 = F(2) * LN2
    grad_z = two_ln2 * p["ns"] * A
    grad_lv = p["vr"] * (two_ln2 * C2 - R)
    grad_mu = -two_ln2 * (w1 * p["ns"][:, None, :]).sum(axis=0, dtype=F)
    return grad_z, grad_mu, grad_lv


def backward_rowvar_sweep(fw, g_log_qz, g_log_qz_prod):
    """The same gradients with the fused sweep's own arithmetic (csrc/tc_bwd_ds.cu: ds_column): the exponent is carried
    shifted, q' = dl^2 - 1/(2 ln2), so that 2 ln2 qc - 1 = 2 ln2 qc' and the logvar sum is sum_j r qc'; e is recomputed as
    2^-qc' = e * exp(1/2) with the exp(-1/2) folded into gP/S."""
    p = fw["pro"]
    k = F(1) / (F(2) * LN2)
    rsqrt_e = F(np.exp(-0.5))
    gJ = g_log_qz.astype(F)[:, None]
    gps = (g_log_qz_prod.astype(F)[:, None] / fw["S"]) * rsqrt_e                    # [R,D]
    qw = gJ * np.exp2(fw["x"] - fw["J2"][:, None])                                 # gJ_i * q_ij
    qs = fw["dl"] * fw["dl"] - k
    qmax_s = p["qmax"][:, None, :] - k
    cs = np.minimum(qs, qmax_s)
    es = np.exp2(-cs) * fw["rho"][:, :, None]
    coef = es * gps[:, None, :] + qw[:, :, None]
    r = coef * (qs <= qmax_s).astype(F)
    t = r * fw["dl"]
    two_ln2 = F(2) * LN2
    grad_z = two_ln2 * p["ns"] * t.sum(axis=1, dtype=F)
    grad_lv = p["vr"] * (two_ln2 * (r * cs).sum(axis=1, dtype=F))
    grad_mu = -two_ln2 * (t * p["ns"][:, None, :]).sum(axis=0, dtype=F)
    return grad_z, grad_mu, grad_lv


def col_prologue(mu_all, lv_all):
    """Per-(j,d) constants of the column-variance ('full') density, ops.py:24-29."""
    mu_all, lv_all = mu_all.astype(F), lv_all.astype(F)
    ivj = np.exp(-lv_all)
    sj = np.sqrt(F(0.5) * LOG2E * ivj)
    return dict(sj=sj, nmus=-(mu_all * sj), c2=F(-0.5) * (lv_all + LOG_2PI) * LOG2E)


def forward_colvar(z, mu_all, lv_all, dataset_size, estimator="mss", row_offset=0):
    b_glob = mu_all.shape[0]
    rows = np.arange(row_offset, row_offset + z.shape[0])
    ws = weight_scalars(b_glob, dataset_size, estimator)
    rho, l2rho = _ratio_matrices(rows, b_glob, ws)
    c = col_prologue(mu_all, lv_all)
    z = z.astype(F)
    dl = z[:, None, :] * c["sj"][None, :, :] + c["nmus"][None, :, :]
    t = c["c2"][None, :, :] - dl * dl
    tcl = np.maximum(t, -K50)
    e = np.exp2(tcl)
    S = (rho[:, :, None] * e).sum(axis=1, dtype=F)
    P = (np.log(S) + ws["lw_u"]).sum(axis=1, dtype=F)
    s2 = tcl.sum(axis=2, dtype=F)
    x = l2rho + s2
    with np.errstate(all="ignore"):
        xm = x.max(axis=1, keepdims=True)
        J2 = (xm + np.log2(np.exp2(x - xm).sum(axis=1, keepdims=True, dtype=F)))[:, 0]
    J = LN2 * J2 + ws["lw_u"]
    return dict(log_qz_prod=P, log_qz=J, S=S, J2=J2, s2=s2, x=x, col=c, rho=rho, dl=dl, t=t, e=e)


def backward_colvar(fw, g_log_qz, g_log_qz_prod):
    """(grad_z [R,D], grad_mu_all [B,D], grad_lv_all [B,D])."""
    c = fw["col"]
    gJ = g_log_qz.astype(F)[:, None]
    gP = g_log_qz_prod.astype(F)[:, None, None]
    qw = gJ * np.exp2(fw["x"] - fw["J2"][:, None])
    pw = fw["rho"][:, :, None] * fw["e"] / fw["S"][:, None, :]
    coef = qw[:, :, None] + gP * pw
    r = np.where(fw["t"] >= -K50, coef, F(0))
    w1 = r * fw["dl"] * c["sj"][None, :, :]                                          # r * Delta' * s_jd
    two_ln2 = F(2) * LN2
    grad_z = -two_ln2 * w1.sum(axis=1, dtype=F)
    grad_mu = two_ln2 * w1.sum(axis=0, dtype=F)
    q = fw["dl"] * fw["dl"]
    grad_lv = (r * (LN2 * q - F(0.5))).sum(axis=0, dtype=F)                          # 0.5*(Delta^2*iv - 1)
    return grad_z, grad_mu, grad_lv

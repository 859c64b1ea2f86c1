"""Generate golden vectors for the TC-ELBO path FROM THE LIVE REFERENCE.

Run in the build container only (needs /root/reference, which does not travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports /root/reference/ops.py and /root/reference/solvers/tc.py unmodified (three absent
third-party modules that the solvers import but the path never uses -- black, matplotlib,
xgboost -- are stubbed in sys.modules), evaluates them on RNG-free inputs in fp32 and fp64 and
writes tests/golden/tc_golden.npz.  Inputs are closed-form functions of (i, d)
so the fixtures stay small: tests regenerate inputs with ``tests/golden/cases.py``.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = "/root/reference"


def _stub_missing_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("black", out=lambda *a, **k: None)
    mpl = mod("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = mod("matplotlib.pyplot")
    mpl.lines = mod("matplotlib.lines", Line2D=object)
    mod("xgboost", XGBClassifier=object)


def main():
    import torch
    from cases import CASES, make_inputs

    _stub_missing_modules()
    sys.path.insert(0, REF)
    import ops as ref_ops                          # /root/reference/ops.py
    from solvers.tc import TCSovler               # /root/reference/solvers/tc.py
    from utils import SingletonWriter              # /root/reference/utils.py
    SingletonWriter().writer = None                # what train.py:100-101 does without TensorBoard
    SingletonWriter().cur_iter = 0

    class _Dataset:                                # only len() is used on the path (solvers/tc.py:81)
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

    class _Self:                                   # duck-typed solver instance: the methods only read these
        def __init__(self, n, beta):
            self.dataset = _Dataset(n)
            self.beta_kl = beta
            self.writer = None

        def write_scalar(self, *a, **k):
            pass

    out = {}
    for name, case in CASES.items():
        B, D, N, beta = case["B"], case["D"], case["N"], case["beta"]
        for dt_name, dt in (("f32", torch.float32), ("f64", torch.float64)):
            mu_np, lv_np, eps_np = make_inputs(case)
            mu = torch.tensor(mu_np, dtype=dt, requires_grad=True)
            lv = torch.tensor(lv_np, dtype=dt, requires_grad=True)
            eps = torch.tensor(eps_np, dtype=dt)
            z = mu + eps * torch.exp(0.5 * lv)      # ops.py:183-185 with a supplied eps
            z.retain_grad()
            pre = f"{name}/{dt_name}/"

            # --- active path: ops.total_correlation / kl_divergence / _compute_kl_loss_simple
            lp = ref_ops.gaussian_log_density_torch(z.unsqueeze(1), mu.unsqueeze(0), lv.unsqueeze(1))
            prod, joint = ref_ops.minibatch_stratified_sampling(lp, B, N)
            out[pre + "log_qz_prod"] = prod.detach().numpy()
            out[pre + "log_qz"] = joint.detach().numpy()
            out[pre + "clamped_frac"] = np.float64((lp.detach() <= -50).double().mean().item())
            tc_none = ref_ops.total_correlation(z, mu, lv, N, reduce="none")
            out[pre + "tc"] = tc_none.detach().numpy()
            out[pre + "tc_mean"] = ref_ops.total_correlation(z, mu, lv, N, reduce="mean").detach().numpy()
            out[pre + "kl"] = ref_ops.kl_divergence(lv, mu, reduce="none").detach().numpy()
            prod_w, joint_w = ref_ops.minibatch_weighted_sampling(lp, B, N)
            out[pre + "mws_log_qz_prod"] = prod_w.detach().numpy()
            out[pre + "mws_log_qz"] = joint_w.detach().numpy()

            slf = _Self(N, beta)
            simple = TCSovler._compute_kl_loss_simple(slf, z, mu, lv, "mean", None, False)
            out[pre + "simple_mean"] = simple.detach().numpy()
            gz, gmu, glv = torch.autograd.grad(simple, [z, mu, lv], retain_graph=True)
            out[pre + "simple_mean_dz_partial"] = gz.numpy()      # z treated as a leaf
            simple.backward(retain_graph=True)                    # total derivative through z = mu + eps*std
            out[pre + "simple_mean_dmu"] = mu.grad.numpy().copy()
            out[pre + "simple_mean_dlv"] = lv.grad.numpy().copy()
            mu.grad = None
            lv.grad = None
            simple_none = TCSovler._compute_kl_loss_simple(slf, z, mu, lv, "none", float(beta), False)
            out[pre + "simple_none"] = simple_none.detach().numpy()
            # soft-intro exp-ELBO on top of the per-sample loss (solvers/intro.py:102-103) with rec_i = 0.3*i
            scale = 1.0 / (3 * 64 * 64)
            rec_i = torch.arange(B, dtype=dt) * 0.3
            expelbo = (-2 * scale * (rec_i + simple_none)).exp().mean()
            out[pre + "expelbo"] = expelbo.detach().numpy()
            expelbo.backward(retain_graph=True)
            out[pre + "expelbo_dmu"] = mu.grad.numpy().copy()
            out[pre + "expelbo_dlv"] = lv.grad.numpy().copy()
            mu.grad = None
            lv.grad = None

            # --- 'full' decomposition (dead code in the reference, solvers/tc.py:91-144)
            full = TCSovler._compute_kl_loss_full(slf, z, mu, lv, "mean", None, False)
            out[pre + "full_mean"] = full.detach().numpy()
            full.backward(retain_graph=True)
            out[pre + "full_mean_dmu"] = mu.grad.numpy().copy()
            out[pre + "full_mean_dlv"] = lv.grad.numpy().copy()
            mu.grad = None
            lv.grad = None
            lpj = ref_ops.gaussian_log_density(z.unsqueeze(1), mu.unsqueeze(0), lv.unsqueeze(0))
            prod_j, joint_j = ref_ops.minibatch_stratified_sampling(lpj, B, N)
            out[pre + "varj_log_qz_prod"] = prod_j.detach().numpy()
            out[pre + "varj_log_qz"] = joint_j.detach().numpy()
            prod_jw, joint_jw = ref_ops.minibatch_weighted_sampling(lpj, B, N)
            out[pre + "varj_mws_log_qz_prod"] = prod_jw.detach().numpy()
            out[pre + "varj_mws_log_qz"] = joint_jw.detach().numpy()

        if B <= 8:
            out[f"{name}/logw"] = ref_ops.log_importance_weight_matrix(B, N).numpy()

    # weight-matrix structure pins (ops.py:32-49) incl. the B == 2 aliasing case
    for b, n in ((2, 10), (3, 3), (5, 100), (8, 3)):
        with np.errstate(all="ignore"):
            out[f"logw/B{b}_N{n}"] = ref_ops.log_importance_weight_matrix(b, n).numpy()

    # keep the fixture small: large fp64 gradient arrays are stored rounded to fp32 (6e-8 relative,
    # far below the 1e-4 gradient tolerance); all loss terms stay fp64
    for k in list(out):
        if out[k].dtype == np.float64 and out[k].size > 8192:
            out[k] = out[k].astype(np.float32)
    path = os.path.join(HERE, "tc_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)/1024:.1f} KiB")

    # sanity print against SURVEY.md 8(c)'s known-answer table
    for name in ("base_B8_D4", "base_B64_D128", "base_B256_D128", "stress_B64_D128"):
        print(name, "tc.mean f64/f32", float(out[name + "/f64/tc_mean"]), float(out[name + "/f32/tc_mean"]),
              "simple", float(out[name + "/f64/simple_mean"]), "full", float(out[name + "/f64/full_mean"]),
              "sum|dmu|", np.abs(out[name + "/f64/simple_mean_dmu"]).sum(),
              "sum|dlv|", np.abs(out[name + "/f64/simple_mean_dlv"]).sum(),
              "clamped", float(out[name + "/f64/clamped_frac"]))


if __name__ == "__main__":
    main()

"""Parity cases added in round 2 (VERDICT r1 "next round" item 1), all through the public ops / the C ABI:

(a) the bench's exact step at the headline size (B=8192, D=128): loss and gradients to mu / logvar against the row-chunked
    CPU oracle, both latent families;
(b) MWS and column-variance + MWS gradients against the oracle at the golden cases;
(c) BASELINE configs[3] shard shape on one GPU (b_loc 4096 of b_glob 32768, D 512, row_offset != 0);
(d) NaN / +Inf / -Inf in mu and logvar against the oracle;
(e) every comparison at the plain north-star tolerances (1e-5 loss terms, 1e-4 gradients, max-norm relative); the errors
    actually achieved are appended to gpurun_out/parity_r2.jsonl (summarised in profiles/r2_parity.md).
"""
import json
import os

import numpy as np
import pytest
import torch

from cases import CASES, make_inputs
from oracle import tc_oracle as O

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FINITE_CASES = [c for c in CASES if not c.startswith("nan_")]


def _ops():
    from intro_tc_vae_b200 import ops
    return ops


def relerr(a, b):
    a = np.asarray(a.detach().cpu() if torch.is_tensor(a) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if torch.is_tensor(b) else b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def record(test, case, **errs):
    """Achieved max relative errors -> gpurun_out/parity_r2.jsonl (travels back from the GPU box)."""
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_r2.jsonl"), "a") as f:
            f.write(json.dumps({"test": test, "case": case, **{k: float(v) for k, v in errs.items()}}) + "\n")
    except OSError:
        pass


def _latents(B, D, family, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    if family == "base":
        mu, lv = torch.randn(B, D, generator=g), -2.0 + torch.randn(B, D, generator=g)
    else:
        mu, lv = 2.0 * torch.randn(B, D, generator=g), -6.0 + 2.0 * torch.randn(B, D, generator=g)
    return mu, lv, torch.randn(B, D, generator=g)


def _oracle_step_chunked(mu_c, lv_c, eps_c, N, beta, g_rows, step=128):
    """sum_i g_i * [(beta-1)*tc_i + kl_i] with z = mu + eps*exp(lv/2), evaluated by the oracle over row chunks (rows are
    independent given all columns); only chunks holding a non-zero g_i are visited.  Returns (loss_rows, dmu, dlv)."""
    B = mu_c.shape[0]
    mu = mu_c.clone().requires_grad_(True)
    lv = lv_c.clone().requires_grad_(True)
    loss_rows = torch.zeros(B)
    for r0 in range(0, B, step):
        sl = slice(r0, min(r0 + step, B))
        if not bool((g_rows[sl] != 0).any()):
            continue
        z = O.reparameterize(mu[sl], lv[sl], eps_c[sl])
        p, j = O.tc_terms_rows(z, lv[sl], mu, r0, B, N)
        rows = (beta - 1.0) * (j - p) + O.kl_no_reduce(lv[sl], mu[sl])
        (rows * g_rows[sl]).sum().backward()
        loss_rows[sl] = rows.detach()
    return loss_rows, mu.grad, lv.grad


# ----------------------------------------------------------------------------------------------------------------
# (a) headline size
# ----------------------------------------------------------------------------------------------------------------
def test_bench_step_at_8192_against_chunked_oracle_all_rows():
    """bench.py's step on bench.py's inputs (synthetic_latents(8192, 128), N=16704, beta=0.5, mean-reduced): the loss and
    the full [8192,128] gradients to mu and logvar against the oracle evaluated over all 64 row chunks."""
    import bench
    ops = _ops()
    B, D, N, beta = 8192, 128, bench.DATASET_SIZE, bench.BETA
    mu_c, lv_c, eps_c = bench.synthetic_latents(B, D)
    loss_rows_o, dmu_o, dlv_o = _oracle_step_chunked(mu_c, lv_c, eps_c, N, beta, torch.full((B,), 1.0 / B))
    loss_o = loss_rows_o.double().mean().item()

    mu = mu_c.cuda().requires_grad_(True)
    lv = lv_c.cuda().requires_grad_(True)
    z = ops.reparameterize(mu, lv, eps_c.cuda())
    rows = ops.kl_tc_loss_terms(z, mu, lv, N, beta, "mss")[0]
    loss = rows.mean()
    loss.backward()
    errs = dict(loss=abs(loss.item() - loss_o) / abs(loss_o), loss_rows=relerr(rows, loss_rows_o),
                dmu=relerr(mu.grad, dmu_o), dlv=relerr(lv.grad, dlv_o))
    record("bench_step_8192_all_rows", "base", **errs)
    assert errs["loss"] < LOSS_RTOL and errs["loss_rows"] < LOSS_RTOL
    assert errs["dmu"] < GRAD_RTOL and errs["dlv"] < GRAD_RTOL

    # the graph-replayed C-ABI step bench.py times returns the same numbers
    from intro_tc_vae_b200.graphs import GraphedKLLoss
    graphed = GraphedKLLoss(B, D, N, beta, "cuda:0")
    l, dmu, dlv = graphed(mu_c.cuda(), lv_c.cuda(), eps_c.cuda())
    errs = dict(loss=abs(l.item() - loss_o) / abs(loss_o), dmu=relerr(dmu, dmu_o), dlv=relerr(dlv, dlv_o))
    record("bench_step_8192_graph_replay", "base", **errs)
    assert errs["loss"] < LOSS_RTOL and errs["dmu"] < GRAD_RTOL and errs["dlv"] < GRAD_RTOL


@pytest.mark.parametrize("family", ["base", "sharp"])
def test_step_at_8192_on_sampled_rows(family):
    """Same step with the upstream gradient restricted to 256 sampled rows (incl. the stratified rows 0, 1, B-2, B-1):
    the loss rows, d/dlogvar of those rows and d/dmu of ALL 8192 columns are then exact functions of the sampled rows,
    which the oracle evaluates in seconds -- for the sharp-posterior family too (40-60 % clamped, variance floor active)."""
    ops = _ops()
    B, D, N, beta = 8192, 128, 16704, 0.5
    mu_c, lv_c, eps_c = _latents(B, D, family, seed=41)
    pick = torch.cat([torch.tensor([0, 1, B - 2, B - 1]), torch.randperm(B - 4, generator=torch.Generator().manual_seed(9))[:252] + 2])
    g_rows = torch.zeros(B)
    g_rows[pick] = torch.linspace(0.5, 1.5, pick.numel()) / pick.numel()
    loss_rows_o, dmu_o, dlv_o = _oracle_step_chunked(mu_c, lv_c, eps_c, N, beta, g_rows, step=64)

    mu = mu_c.cuda().requires_grad_(True)
    lv = lv_c.cuda().requires_grad_(True)
    z = ops.reparameterize(mu, lv, eps_c.cuda())
    rows = ops.kl_tc_loss_terms(z, mu, lv, N, beta, "mss")[0]
    (rows * g_rows.cuda()).sum().backward()
    errs = dict(loss_rows=relerr(rows[pick.cuda()], loss_rows_o[pick]), dmu=relerr(mu.grad, dmu_o), dlv=relerr(lv.grad, dlv_o))
    record("step_8192_sampled_rows", family, **errs)
    assert errs["loss_rows"] < LOSS_RTOL and errs["dmu"] < GRAD_RTOL and errs["dlv"] < GRAD_RTOL


# ----------------------------------------------------------------------------------------------------------------
# (b) MWS / column-variance gradients
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", FINITE_CASES)
@pytest.mark.parametrize("estimator,var_of", [("mws", "row"), ("mws", "col"), ("mss", "col")])
def test_estimator_variants_gradients_against_oracle(name, estimator, var_of):
    """ops.py:92-101 (MWS) and solvers/tc.py:114-116 (column variance): both outputs and the gradients of a weighted sum of
    them to mu / logvar (through z) against the oracle's autograd on the golden cases' inputs."""
    ops = _ops()
    case = CASES[name]
    B, N = case["B"], case["N"]
    mu_np, lv_np, eps_np = make_inputs(case)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)                         # noqa: E731
    w1, w2 = torch.linspace(0.5, 1.5, B), torch.linspace(-0.7, 0.9, B)

    mu_o, lv_o = f32(mu_np).requires_grad_(True), f32(lv_np).requires_grad_(True)
    z_o = O.reparameterize(mu_o, lv_o, f32(eps_np))
    prod_o, joint_o = O.tc_terms(z_o, mu_o, lv_o, N, estimator, var_of)
    ((joint_o * w1).sum() - (prod_o * w2).sum()).backward()

    mu, lv = f32(mu_np).cuda().requires_grad_(True), f32(lv_np).cuda().requires_grad_(True)
    z = ops.reparameterize(mu, lv, f32(eps_np).cuda())
    prod, joint = ops.tc_terms(z, mu, lv, N, estimator, var_of)
    ((joint * w1.cuda()).sum() - (prod * w2.cuda()).sum()).backward()
    errs = dict(prod=relerr(prod, prod_o), joint=relerr(joint, joint_o), dmu=relerr(mu.grad, mu_o.grad), dlv=relerr(lv.grad, lv_o.grad))
    record(f"variant_{estimator}_{var_of}", name, **errs)
    assert errs["prod"] < LOSS_RTOL and errs["joint"] < LOSS_RTOL
    assert errs["dmu"] < GRAD_RTOL and errs["dlv"] < GRAD_RTOL


def test_fused_loss_mws_gradients_against_oracle():
    """The fused (beta-1)*TC + KL op with the MWS estimator (the north-star's named estimator) against the oracle."""
    ops = _ops()
    B, D, N, beta = 384, 20, 16704, 0.5
    mu_c, lv_c, eps_c = _latents(B, D, "base", seed=6)
    mu_o, lv_o = mu_c.clone().requires_grad_(True), lv_c.clone().requires_grad_(True)
    z_o = O.reparameterize(mu_o, lv_o, eps_c)
    p_o, j_o = O.tc_terms(z_o, mu_o, lv_o, N, "mws", "row")
    loss_o = ((beta - 1.0) * (j_o - p_o) + O.kl_no_reduce(lv_o, mu_o)).mean()
    loss_o.backward()
    mu, lv = mu_c.cuda().requires_grad_(True), lv_c.cuda().requires_grad_(True)
    z = ops.reparameterize(mu, lv, eps_c.cuda())
    loss = ops.kl_tc_loss_terms(z, mu, lv, N, beta, "mws")[0].mean()
    loss.backward()
    errs = dict(loss=abs(loss.item() - loss_o.item()) / abs(loss_o.item()), dmu=relerr(mu.grad, mu_o.grad), dlv=relerr(lv.grad, lv_o.grad))
    record("fused_loss_mws", "base_B384_D20", **errs)
    assert errs["loss"] < LOSS_RTOL and errs["dmu"] < GRAD_RTOL and errs["dlv"] < GRAD_RTOL


# ----------------------------------------------------------------------------------------------------------------
# (c) BASELINE configs[3] shard
# ----------------------------------------------------------------------------------------------------------------
def test_cfg4_shard_on_one_gpu():
    """Rank 3 of 8 of the stress config: rows 12288..16383 of a 32768 x 512 batch (N = 737280).  The upstream gradient is
    restricted to 24 of the shard's rows, so d/dz, d/dlogvar of those rows and the shard's contribution to d/dmu of all
    32768 columns can be checked against the row-chunked oracle."""
    from intro_tc_vae_b200 import _lib
    _ops()
    Bg, D, P, r, N = 32768, 512, 8, 3, 737280
    bl = Bg // P
    lo = r * bl
    g = torch.Generator().manual_seed(7)
    mu_c = torch.randn(Bg, D, generator=g)
    lv_c = -2.0 + torch.randn(Bg, D, generator=g)
    z_c = mu_c + torch.randn(Bg, D, generator=g) * torch.exp(0.5 * lv_c)
    pick = torch.cat([torch.tensor([0, 1, bl - 1]), torch.randperm(bl - 3, generator=torch.Generator().manual_seed(2))[:21] + 2])
    w = torch.zeros(bl)
    w[pick] = torch.linspace(0.5, 1.5, pick.numel())

    dev = torch.device("cuda:0")
    mu_all = mu_c.to(dev).requires_grad_(True)
    z = z_c[lo:lo + bl].to(dev).requires_grad_(True)
    lv = lv_c[lo:lo + bl].to(dev).requires_grad_(True)
    flags = _lib.EST_MSS | _lib.VAR_ROW | _lib.SAVE_FOR_BACKWARD
    lq, lqp, _ = torch.ops.tcelbo.tc_forward(z, mu_all, lv, lo, N, flags)
    ((lq - lqp) * w.to(dev)).sum().backward()
    assert torch.isfinite(lq).all() and torch.isfinite(lqp).all()

    mu_o = mu_c.clone().requires_grad_(True)
    worst = dict(prod=0.0, joint=0.0, dz=0.0, dlv=0.0)
    for i in pick.tolist():
        zi = z_c[lo + i:lo + i + 1].clone().requires_grad_(True)
        lvi = lv_c[lo + i:lo + i + 1].clone().requires_grad_(True)
        p_o, j_o = O.tc_terms_rows(zi, lvi, mu_o, lo + i, Bg, N)
        ((j_o - p_o) * w[i]).sum().backward()
        worst["prod"] = max(worst["prod"], abs(lqp[i].item() - p_o.item()) / abs(p_o.item()))
        worst["joint"] = max(worst["joint"], abs(lq[i].item() - j_o.item()) / abs(j_o.item()))
        worst["dz"] = max(worst["dz"], relerr(z.grad[i], zi.grad[0]))
        worst["dlv"] = max(worst["dlv"], relerr(lv.grad[i], lvi.grad[0]))
    worst["dmu"] = relerr(mu_all.grad, mu_o.grad)
    record("cfg4_shard", "b_loc4096_b_glob32768_D512_rank3", **worst)
    assert worst["prod"] < LOSS_RTOL and worst["joint"] < LOSS_RTOL
    assert worst["dz"] < GRAD_RTOL and worst["dlv"] < GRAD_RTOL and worst["dmu"] < GRAD_RTOL


# ----------------------------------------------------------------------------------------------------------------
# (d) non-finite inputs
# ----------------------------------------------------------------------------------------------------------------
def _same_pattern(a, b):
    a, b = a.detach().cpu().numpy(), b.detach().cpu().numpy()
    return (np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.isposinf(a), np.isposinf(b))
            and np.array_equal(np.isneginf(a), np.isneginf(b)))


def _finite_relerr(a, b):
    a, b = a.detach().cpu().double().numpy(), b.detach().cpu().double().numpy()
    fin = np.isfinite(a) & np.isfinite(b)
    if not fin.any():
        return 0.0
    return float(np.abs(a[fin] - b[fin]).max() / max(np.abs(b[fin]).max(), 1e-30))


@pytest.mark.parametrize("B,D", [(48, 32), (200, 128)])
@pytest.mark.parametrize("what", ["mu_nan", "mu_pinf", "mu_ninf", "lv_nan", "lv_pinf", "lv_ninf", "lv_huge", "lv_tiny"])
def test_non_finite_inputs_propagate_like_the_reference(B, D, what):
    """the op must propagate NaN / Inf like the reference, not trap or launder them.
    One poisoned element (row i0, dim d0).  Every loss term and both gradients must carry NaN / +Inf / -Inf at exactly the
    reference's positions and agree elsewhere."""
    ops = _ops()
    N, beta = 16704, 6.0
    mu_c, lv_c, eps_c = _latents(B, D, "base", seed=13)
    i0, d0 = B // 3, D // 2
    tgt, val = what.split("_")
    value = {"nan": float("nan"), "pinf": float("inf"), "ninf": float("-inf"), "huge": 80.0, "tiny": -120.0}[val]
    (mu_c if tgt == "mu" else lv_c)[i0, d0] = value
    w = torch.linspace(0.5, 1.5, B)

    mu_o, lv_o = mu_c.clone().requires_grad_(True), lv_c.clone().requires_grad_(True)
    z_o = O.reparameterize(mu_o, lv_o, eps_c)
    prod_o, joint_o = O.tc_terms(z_o, mu_o, lv_o, N)
    kl_o = O.kl_no_reduce(lv_o, mu_o)
    loss_o = (beta - 1.0) * (joint_o - prod_o) + kl_o
    (loss_o * w).sum().backward()

    mu, lv = mu_c.cuda().requires_grad_(True), lv_c.cuda().requires_grad_(True)
    z = ops.reparameterize(mu, lv, eps_c.cuda())
    loss, kl, joint, prod = ops.kl_tc_loss_terms(z, mu, lv, N, beta)
    (loss * w.cuda()).sum().backward()
    torch.cuda.synchronize()
    for name, got, want in (("z", z, z_o), ("log_qz_prod", prod, prod_o), ("log_qz", joint, joint_o), ("kl", kl, kl_o), ("loss", loss, loss_o)):
        assert _same_pattern(got, want), f"{what}: non-finite pattern of {name} differs from the reference"
        assert _finite_relerr(got, want) < LOSS_RTOL, f"{what}: finite entries of {name}"
    for name, got, want in (("dmu", mu.grad, mu_o.grad), ("dlv", lv.grad, lv_o.grad)):
        # the clamp mask is a select in the sweep (like the where() of torch.clamp's backward), so the gradients carry NaN / Inf at
        # exactly the reference's positions, rows that are already NaN included
        assert _same_pattern(got, want), f"{what}: non-finite pattern of {name} differs from the reference"
        assert _finite_relerr(got, want) < GRAD_RTOL, f"{what}: finite entries of {name}"
    record("non_finite", f"{what}_B{B}_D{D}", loss=_finite_relerr(loss, loss_o), dmu=_finite_relerr(mu.grad, mu_o.grad),
           dlv=_finite_relerr(lv.grad, lv_o.grad), extra_nan_dmu=float((torch.isnan(mu.grad).cpu() & ~torch.isnan(mu_o.grad)).sum()),
           extra_nan_dlv=float((torch.isnan(lv.grad).cpu() & ~torch.isnan(lv_o.grad)).sum()))


# ----------------------------------------------------------------------------------------------------------------
# fused prologue / epilogue of the loss op
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,D,family", [(64, 128, "base"), (200, 20, "sharp"), (130, 256, "base")])
def test_fused_reparameterize_and_batch_mean_against_oracle(B, D, family):
    """tcelbo_klloss_forward_ex / _backward_ex with eps (reparameterize in the prologue, chain rule through z in the backward
    finalize) and loss_mean (reduced in the forward finalize) == the oracle's reparameterize -> compute_kl_loss(mean) -> backward,
    i.e. ops.py:183-185 + solvers/tc.py:69-89.  GraphedKLLoss's direct step is exactly this call sequence (6 launches)."""
    from intro_tc_vae_b200 import _lib
    from intro_tc_vae_b200.graphs import GraphedKLLoss
    N, beta = 16704, 0.5
    mu_c, lv_c, eps_c = _latents(B, D, family, seed=23)
    mu_o, lv_o = mu_c.clone().requires_grad_(True), lv_c.clone().requires_grad_(True)
    z_o = O.reparameterize(mu_o, lv_o, eps_c)
    loss_o = O.kl_loss_simple(z_o, mu_o, lv_o, N, beta, "mean")
    loss_o.backward()
    lib = _lib.load()
    graphed = GraphedKLLoss(B, D, N, beta, "cuda:0", capture=False)
    c0 = lib.tcelbo_launch_count()
    loss, dmu, dlv = graphed(mu_c.cuda(), lv_c.cuda(), eps_c.cuda())
    torch.cuda.synchronize()
    assert lib.tcelbo_launch_count() - c0 == 6
    errs = dict(loss=abs(loss.item() - loss_o.item()) / abs(loss_o.item()), z=relerr(graphed._z, z_o),
                dmu=relerr(dmu, mu_o.grad), dlv=relerr(dlv, lv_o.grad))
    record("fused_reparam_mean", f"{family}_B{B}_D{D}", **errs)
    assert errs["loss"] < LOSS_RTOL and errs["z"] < 1e-6 and errs["dmu"] < GRAD_RTOL and errs["dlv"] < GRAD_RTOL


@pytest.mark.parametrize("B,D,family", [(64, 128, "base"), (96, 32, "sharp")])
def test_mean_reduced_op_equals_reference_mean(B, D, family):
    """ops.kl_tc_loss_mean (what compute_kl_loss(reduce="mean") calls): loss, the KL it logs, and gradients vs the oracle."""
    ops = _ops()
    N, beta = 16704, 6.0
    mu_c, lv_c, eps_c = _latents(B, D, family, seed=29)
    mu_o, lv_o = mu_c.clone().requires_grad_(True), lv_c.clone().requires_grad_(True)
    z_o = O.reparameterize(mu_o, lv_o, eps_c)
    loss_o = O.kl_loss_simple(z_o, mu_o, lv_o, N, beta, "mean")
    kl_o = O.kl_divergence(lv_o, mu_o, "mean")
    (loss_o + 0.25 * kl_o).backward()
    mu, lv = mu_c.cuda().requires_grad_(True), lv_c.cuda().requires_grad_(True)
    z = ops.reparameterize(mu, lv, eps_c.cuda())
    loss, kl = ops.kl_tc_loss_mean(z, mu, lv, N, beta)
    assert loss.dim() == 0 and kl.dim() == 0
    (loss + 0.25 * kl).backward()
    errs = dict(loss=abs(loss.item() - loss_o.item()) / abs(loss_o.item()), kl=abs(kl.item() - kl_o.item()) / abs(kl_o.item()),
                dmu=relerr(mu.grad, mu_o.grad), dlv=relerr(lv.grad, lv_o.grad))
    record("mean_op", f"{family}_B{B}_D{D}", **errs)
    assert errs["loss"] < LOSS_RTOL and errs["kl"] < LOSS_RTOL and errs["dmu"] < GRAD_RTOL and errs["dlv"] < GRAD_RTOL


@pytest.mark.parametrize("B,D,family,beta", [(64, 128, "base", 512.0), (150, 32, "sharp", 256.0)])
def test_fused_exp_elbo_against_reference_formula(B, D, family, beta):
    """solvers/intro.py:84-89,102-103: exp(-2*scale*(rec_i + kl_i)).mean() with kl_i from compute_kl_loss(reduce="none",
    beta=beta_neg), fused into the loss op's finalize / backward prologue, vs the oracle's composition; gradients to mu, logvar
    (through z) and to the per-sample reconstruction loss."""
    from intro_tc_vae_b200.losses import kl_tc_exp_elbo
    ops = _ops()
    N, scale = 16704, 1.0 / (3 * 64 * 64)
    mu_c, lv_c, eps_c = _latents(B, D, family, seed=31)
    rec_c = 40.0 + 25.0 * torch.rand(B, generator=torch.Generator().manual_seed(4))
    mu_o, lv_o, rec_o = mu_c.clone().requires_grad_(True), lv_c.clone().requires_grad_(True), rec_c.clone().requires_grad_(True)
    z_o = O.reparameterize(mu_o, lv_o, eps_c)
    kl_rows_o = O.kl_loss_simple(z_o, mu_o, lv_o, N, beta, "none")
    ee_o = O.exp_elbo(rec_o, kl_rows_o, scale)
    ee_o.backward()
    mu, lv, rec = mu_c.cuda().requires_grad_(True), lv_c.cuda().requires_grad_(True), rec_c.cuda().requires_grad_(True)
    z = ops.reparameterize(mu, lv, eps_c.cuda())
    ee, kl_rows = kl_tc_exp_elbo(z, mu, lv, rec, N, beta, scale)
    ee.backward()
    errs = dict(expelbo=abs(ee.item() - ee_o.item()) / abs(ee_o.item()), kl_rows=relerr(kl_rows, kl_rows_o),
                dmu=relerr(mu.grad, mu_o.grad), dlv=relerr(lv.grad, lv_o.grad), drec=relerr(rec.grad, rec_o.grad))
    record("fused_exp_elbo", f"{family}_B{B}_D{D}", **errs)
    assert errs["expelbo"] < LOSS_RTOL and errs["kl_rows"] < LOSS_RTOL
    assert errs["dmu"] < GRAD_RTOL and errs["dlv"] < GRAD_RTOL and errs["drec"] < GRAD_RTOL

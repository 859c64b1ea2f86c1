"""Pin the CPU oracle (oracle/tc_oracle.py) to outputs of the live reference and to SURVEY.md 8(c)'s
known-answer table.  CPU only."""
import math

import numpy as np
import pytest
import torch

from cases import CASES, make_inputs
from oracle import tc_oracle as O

FINITE_CASES = [c for c in CASES if not c.startswith("nan_")]


def _leafs(case, dt):
    mu_np, lv_np, eps_np = make_inputs(case)
    mu = torch.tensor(mu_np, dtype=dt, requires_grad=True)
    lv = torch.tensor(lv_np, dtype=dt, requires_grad=True)
    eps = torch.tensor(eps_np, dtype=dt)
    z = O.reparameterize(mu, lv, eps)
    return mu, lv, eps, z


def _close(a, b, rtol):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    return np.abs(a - b).max() <= rtol * scale


@pytest.mark.parametrize("name", FINITE_CASES)
@pytest.mark.parametrize("dt_name", ["f32", "f64"])
def test_oracle_matches_reference_outputs(golden, name, dt_name):
    case = CASES[name]
    dt = torch.float32 if dt_name == "f32" else torch.float64
    # identical op sequence in the same dtype -> agreement to rounding noise of the last reduction
    rtol = 2e-6 if dt_name == "f32" else 1e-12
    mu, lv, eps, z = _leafs(case, dt)
    B, N, beta = case["B"], case["N"], case["beta"]
    pre = f"{name}/{dt_name}/"

    prod, joint = O.tc_terms(z, mu, lv, N, "mss", "row")
    assert _close(prod.detach(), golden[pre + "log_qz_prod"], rtol)
    assert _close(joint.detach(), golden[pre + "log_qz"], rtol)
    assert _close(O.total_correlation(z, mu, lv, N, "none").detach(), golden[pre + "tc"], rtol)
    assert _close(O.total_correlation(z, mu, lv, N, "mean").detach(), golden[pre + "tc_mean"], rtol)
    assert _close(O.kl_divergence(lv, mu, "none").detach(), golden[pre + "kl"], rtol)

    prod_w, joint_w = O.tc_terms(z, mu, lv, N, "mws", "row")
    assert _close(prod_w.detach(), golden[pre + "mws_log_qz_prod"], rtol)
    assert _close(joint_w.detach(), golden[pre + "mws_log_qz"], rtol)
    prod_j, joint_j = O.tc_terms(z, mu, lv, N, "mss", "col")
    assert _close(prod_j.detach(), golden[pre + "varj_log_qz_prod"], rtol)
    assert _close(joint_j.detach(), golden[pre + "varj_log_qz"], rtol)
    prod_jw, joint_jw = O.tc_terms(z, mu, lv, N, "mws", "col")
    assert _close(prod_jw.detach(), golden[pre + "varj_mws_log_qz_prod"], rtol)
    assert _close(joint_jw.detach(), golden[pre + "varj_mws_log_qz"], rtol)

    grtol = 2e-5 if dt_name == "f32" else 1e-6      # fp64 grads are stored rounded to fp32 for large cases
    simple = O.kl_loss_simple(z, mu, lv, N, beta, "mean")
    assert _close(simple.detach(), golden[pre + "simple_mean"], rtol)
    gz = torch.autograd.grad(simple, z, retain_graph=True)[0]
    assert _close(gz, golden[pre + "simple_mean_dz_partial"], grtol)
    simple.backward(retain_graph=True)
    assert _close(mu.grad, golden[pre + "simple_mean_dmu"], grtol)
    assert _close(lv.grad, golden[pre + "simple_mean_dlv"], grtol)
    mu.grad = lv.grad = None

    simple_none = O.kl_loss_simple(z, mu, lv, N, beta, "none")
    assert _close(simple_none.detach(), golden[pre + "simple_none"], rtol)
    rec_i = torch.arange(B, dtype=dt) * 0.3
    ee = O.exp_elbo(rec_i, simple_none, 1.0 / (3 * 64 * 64))
    assert _close(ee.detach(), golden[pre + "expelbo"], rtol)
    ee.backward(retain_graph=True)
    assert _close(mu.grad, golden[pre + "expelbo_dmu"], grtol)
    assert _close(lv.grad, golden[pre + "expelbo_dlv"], grtol)
    mu.grad = lv.grad = None

    mu, lv, eps, z = _leafs(case, dt)
    full, mi, tc, dimkl = O.kl_loss_full(z, mu, lv, N, beta, "mean")
    assert _close(full.detach(), golden[pre + "full_mean"], rtol)
    full.backward()
    assert _close(mu.grad, golden[pre + "full_mean_dmu"], grtol)
    assert _close(lv.grad, golden[pre + "full_mean_dlv"], grtol)


def test_known_answer_table_survey_8c(golden):
    """fp64 reference values, reproduced by the oracle in fp64."""
    kat = {
        "base_B8_D4": (0.3138941859, -2.3629473104, -2.6768414963, 2.2881658140, 3.8576367436, 4.4637241331),
        "base_B64_D128": (110.8338816540, -26.0556149486, -136.8894966026, 105.7299408681, 659.8993491381,
                          695.3597833642),
        "base_B256_D128": (111.4621011916, -25.2606032514, -136.7227044430, 106.6098305257, 663.9203364835,
                           697.8451853146),
    }
    for name, (tc_m, j_m, p_m, kl_m, simple, full) in kat.items():
        case = CASES[name]
        mu, lv, eps, z = _leafs(case, torch.float64)
        prod, joint = O.tc_terms(z, mu, lv, case["N"], "mss", "row")
        assert (joint - prod).mean().item() == pytest.approx(tc_m, rel=1e-9)
        assert joint.mean().item() == pytest.approx(j_m, rel=1e-9)
        assert prod.mean().item() == pytest.approx(p_m, rel=1e-9)
        assert O.kl_divergence(lv, mu, "mean").item() == pytest.approx(kl_m, rel=1e-9)
        assert O.kl_loss_simple(z, mu, lv, case["N"], case["beta"]).item() == pytest.approx(simple, rel=1e-9)
        assert O.kl_loss_full(z, mu, lv, case["N"], case["beta"])[0].item() == pytest.approx(full, rel=1e-9)
    # stress row of the table + the extra per-row / per-element pins
    case = CASES["stress_B64_D128"]
    mu, lv, eps, z = _leafs(case, torch.float64)
    tc = O.total_correlation(z, mu, lv, case["N"], "none")
    assert tc.mean().item() == pytest.approx(306.2537420113, rel=1e-9)
    assert [tc[0].item(), tc[1].item(), tc[62].item()] == pytest.approx([335.73011824, 298.19333998, 282.09860799],
                                                                       rel=1e-9)
    simple = O.kl_loss_simple(z, mu, lv, case["N"], case["beta"])
    assert simple.item() == pytest.approx(156939.60951235, rel=1e-9)
    simple.backward()
    assert mu.grad[0, 0].item() == pytest.approx(-3.0452200680, rel=1e-8)
    assert lv.grad[0, 0].item() == pytest.approx(-4.3953192991, rel=1e-8)
    assert mu.grad.abs().sum().item() == pytest.approx(1538712.31723190, rel=1e-9)
    assert lv.grad.abs().sum().item() == pytest.approx(17326.86596151, rel=1e-9)


def test_weight_matrix_structure(golden):
    for b, n in ((2, 10), (3, 3), (5, 100), (8, 3)):
        with np.errstate(all="ignore"):
            got = O.log_importance_weight_matrix(b, n).numpy()
        np.testing.assert_array_equal(got, golden[f"logw/B{b}_N{n}"])
    with pytest.raises(ZeroDivisionError):
        O.log_importance_weight_matrix(1, 10)


def test_n_smaller_than_batch_gives_nan(golden):
    case = CASES["nan_B8_D4"]
    mu, lv, eps, z = _leafs(case, torch.float32)
    tc = O.total_correlation(z, mu, lv, case["N"], "none")
    assert torch.isnan(tc).all()
    assert np.isnan(golden["nan_B8_D4/f32/tc"]).all()


@pytest.mark.parametrize("name", ["base_B64_D128", "stress_B64_D128", "ragged_B37_D20", "pair_B2_D16"])
@pytest.mark.parametrize("estimator,var_of", [("mss", "row"), ("mws", "row"), ("mss", "col")])
def test_row_chunked_oracle_equals_full(name, estimator, var_of):
    """Rows are independent given all columns: the chunked evaluator used for large-B parity agrees."""
    case = CASES[name]
    mu, lv, eps, z = _leafs(case, torch.float64)
    B, N = case["B"], case["N"]
    prod, joint = O.tc_terms(z, mu, lv, N, estimator, var_of)
    step = 5
    for r0 in range(0, B, step):
        r1 = min(B, r0 + step)
        p, j = O.tc_terms_rows(z[r0:r1], lv[r0:r1], mu, r0, B, N, estimator, var_of, logvar_all=lv)
        assert torch.allclose(p, prod[r0:r1], rtol=1e-12, atol=1e-12)
        assert torch.allclose(j, joint[r0:r1], rtol=1e-12, atol=1e-12)

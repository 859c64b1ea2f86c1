"""The reference's op sequence (oracle/tc_oracle.py, plain torch) run EAGERLY ON THE SAME GPU beside the kernels
(SURVEY.md 8d "reference on the same B200"): checks parity at a size the [B,B,D] tensors still fit, and records both
timings in gpurun_out/eager_gpu_baseline.json when that directory exists.  TCELBO_EAGER_B overrides the batch (4096 needs
~60 GiB for the four saved [B,B,D] tensors; 8192 would need 4 x 32 GiB + temporaries and does not fit 180 GB)."""
import json
import os

import pytest
import torch

from oracle import tc_oracle as O

pytestmark = pytest.mark.gpu


def _time(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def test_eager_torch_on_gpu_beside_the_kernels():
    from intro_tc_vae_b200 import ops
    B, D, N, beta = int(os.environ.get("TCELBO_EAGER_B", "2048")), 128, 16704, 0.5
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1234)
    mu_c, lv_c, eps_c = torch.randn(B, D, generator=g), -2.0 + torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
    res = {}

    def eager():
        mu, lv = mu_c.to(dev).requires_grad_(True), lv_c.to(dev).requires_grad_(True)
        z = O.reparameterize(mu, lv, eps_c.to(dev))
        loss = ((beta - 1.0) * O.total_correlation(z, mu, lv, N, reduce="none") + O.kl_divergence(lv, mu, reduce="none")).mean()
        loss.backward()
        res["eager"] = (loss.detach(), mu.grad, lv.grad)

    def ours():
        mu, lv = mu_c.to(dev).requires_grad_(True), lv_c.to(dev).requires_grad_(True)
        z = ops.reparameterize(mu, lv, eps_c.to(dev))
        loss = ops.kl_tc_loss_terms(z, mu, lv, N, beta)[0].mean()
        loss.backward()
        res["ours"] = (loss.detach(), mu.grad, lv.grad)

    t_eager, t_ours = _time(eager), _time(ours)
    peak_gib = torch.cuda.max_memory_allocated(dev) / 2**30
    (l0, gm0, gl0), (l1, gm1, gl1) = res["eager"], res["ours"]
    assert abs(l1.item() - l0.item()) <= 1e-5 * abs(l0.item())
    assert ((gm1 - gm0).abs().max() / gm0.abs().max()).item() < 1e-4
    assert ((gl1 - gl0).abs().max() / gl0.abs().max()).item() < 1e-4
    line = {"B": B, "D": D, "eager_torch_gpu_ms": t_eager, "kernels_eager_launch_ms": t_ours, "speedup": t_eager / t_ours,
            "log_densities_per_s_eager": B * B * D / (t_eager * 1e-3), "log_densities_per_s_kernels": B * B * D / (t_ours * 1e-3),
            "peak_memory_gib_incl_eager": peak_gib}
    print(json.dumps(line))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"eager_gpu_baseline_B{B}.json"), "w") as f:
            json.dump(line, f)

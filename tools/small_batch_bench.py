"""Small-batch timing of the TC-ELBO step (BASELINE configs[0] / [1] latent shapes: B = 3 / 64, z_dim 128).

    python tools/small_batch_bench.py [--batch 64] [--zdim 128]

Prints one JSON line: microseconds per compute_kl_loss-equivalent step (reparameterize + KL + TC + backward) replayed from a
CUDA graph (6 library launches) and issued eagerly through the drop-in signatures, plus the two sweeps' own durations.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from intro_tc_vae_b200 import _lib, ops
from intro_tc_vae_b200.graphs import GraphedKLLoss
from intro_tc_vae_b200.solvers import TCLossMixin


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--zdim", type=int, default=128)
    ap.add_argument("--reps", type=int, default=200)
    args = ap.parse_args()
    B, D, N, beta = args.batch, args.zdim, 16704, 0.5
    dev = torch.device("cuda:0")
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    mu_c, lv_c, eps_c = torch.randn(B, D, generator=g), -2.0 + torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
    mu, lv, eps = mu_c.to(dev).requires_grad_(True), lv_c.to(dev).requires_grad_(True), eps_c.to(dev)

    class _Dataset:
        def __len__(self):
            return N

    class _Solver(TCLossMixin):
        beta_kl = beta
        dataset = _Dataset()

        def write_scalar(self, *a, **k):
            pass

    solver = _Solver()

    def eager():
        mu.grad = lv.grad = None
        z = ops.reparameterize(mu, lv, eps)
        loss = solver.compute_kl_loss(z, mu, lv)
        loss.backward()
        return loss

    graphed = GraphedKLLoss(B, D, N, beta, dev)
    graphed(mu.detach(), lv.detach(), eps)
    ref = eager()
    torch.cuda.synchronize()
    assert abs(graphed.loss.item() - ref.item()) <= 1e-5 * abs(ref.item())

    def timed(fn, reps):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / reps

    out = {"batch": B, "z_dim": D, "graph_step_us": timed(graphed.replay, args.reps), "eager_dropin_step_us": timed(eager, args.reps)}
    cur = torch.cuda.current_stream(dev)
    for kid, name in ((1, "fwd_sweep_us"), (2, "bwd_sweep_us")):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(cur); ev1.record(cur)
        ts = []
        for _ in range(20):
            _lib.check(lib.tcelbo_profile_events(kid, ev0.cuda_event, ev1.cuda_event), "profile_events")
            eager()
            torch.cuda.synchronize()
            lib.tcelbo_profile_events(0, None, None)
            ts.append(1e3 * ev0.elapsed_time(ev1))
        ts.sort()
        out[name] = ts[len(ts) // 2]
    print(json.dumps(out))


if __name__ == "__main__":
    main()

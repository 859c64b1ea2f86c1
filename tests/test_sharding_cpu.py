"""world_size-2 gloo test of the row-sharding plumbing (all-gather forward, reduce-scatter backward,
global row offsets).  The CPU oracle stands in for the kernels so the collective logic is exercised
without a GPU: each rank evaluates ITS rows against the gathered columns, and the results/gradients
must equal the single-process evaluation of the whole batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cases import CASES, make_inputs
        from oracle import tc_oracle as O
        from intro_tc_vae_b200.sharding import gather_rows, shard_rows

        case = CASES["base_B64_D128"]
        mu_np, lv_np, eps_np = make_inputs(case)
        B, N = case["B"], case["N"]
        b_loc = B // world
        lo = rank * b_loc
        mu = torch.tensor(mu_np[lo:lo + b_loc], dtype=torch.float64, requires_grad=True)
        lv = torch.tensor(lv_np[lo:lo + b_loc], dtype=torch.float64, requires_grad=True)
        eps = torch.tensor(eps_np[lo:lo + b_loc], dtype=torch.float64)
        z = O.reparameterize(mu, lv, eps)
        row_offset, b_glob = shard_rows(dist.group.WORLD, b_loc)
        assert (row_offset, b_glob) == (lo, B)
        mu_all = gather_rows(mu, dist.group.WORLD)
        prod, joint = O.tc_terms_rows(z, lv, mu_all, row_offset, b_glob, N)
        # per-rank mean over local rows; DDP-style averaging of the gradients across ranks gives the global mean
        loss = (joint - prod).mean()
        loss.backward()
        ret[rank] = dict(tc=(joint - prod).detach().numpy(), dmu=mu.grad.numpy() / world, dlv=lv.grad.numpy() / world)
    finally:
        dist.destroy_process_group()


def test_two_rank_row_sharding_equals_single_process():
    from cases import CASES, make_inputs
    from oracle import tc_oracle as O

    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)

    case = CASES["base_B64_D128"]
    mu_np, lv_np, eps_np = make_inputs(case)
    mu = torch.tensor(mu_np, dtype=torch.float64, requires_grad=True)
    lv = torch.tensor(lv_np, dtype=torch.float64, requires_grad=True)
    z = O.reparameterize(mu, lv, torch.tensor(eps_np, dtype=torch.float64))
    tc = O.total_correlation(z, mu, lv, case["N"], "none")
    tc.mean().backward()
    got_tc = np.concatenate([ret[r]["tc"] for r in range(world)])
    got_dmu = np.concatenate([ret[r]["dmu"] for r in range(world)])
    got_dlv = np.concatenate([ret[r]["dlv"] for r in range(world)])
    np.testing.assert_allclose(got_tc, tc.detach().numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(got_dmu, mu.grad.numpy(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(got_dlv, lv.grad.numpy(), rtol=1e-10, atol=1e-14)

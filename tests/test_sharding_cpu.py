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
        row_offset, b_glob = shard_rows(dist.group.WORLD, b_loc, torch.device("cpu"))   # incl. the equal-shards assertion
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


def _worker_unequal(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from intro_tc_vae_b200.sharding import shard_rows
        try:
            shard_rows(dist.group.WORLD, 32 if rank == 0 else 24, torch.device("cpu"))
            ret[rank] = "no error"
        except Exception as exc:                       # torch._assert_async raises on the CPU
            ret[rank] = f"{type(exc).__name__}: {exc}"
    finally:
        dist.destroy_process_group()


def test_unequal_shards_fail_loudly():
    """a ragged last batch that differs across ranks must not silently compute with wrong row offsets."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29900 + (os.getpid() % 90)
    mp.spawn(_worker_unequal, args=(world, port, ret), nprocs=world, join=True)
    assert all("different numbers of rows" in ret[r] for r in range(world)), dict(ret)


def _worker_gradsync(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from intro_tc_vae_b200.ddp import GradSync
        torch.manual_seed(0)                                   # same initial weights on every rank
        enc, dec = torch.nn.Linear(6, 4), torch.nn.Linear(4, 6)
        sync = GradSync(list(enc.parameters()) + list(dec.parameters()), dist.group.WORLD)
        x = torch.randn(8, 6, generator=torch.Generator().manual_seed(100 + rank))
        # half step 1: only the encoder trains (solvers/intro.py:66-69); half step 2: only the decoder (:119-122)
        for p in dec.parameters():
            p.requires_grad = False
        (dec(enc(x)) - x).pow(2).mean().backward()
        g_enc = [p.grad.clone() for p in enc.parameters()]
        assert all(p.grad is None for p in dec.parameters())
        for p in enc.parameters():
            p.requires_grad = False
        for p in dec.parameters():
            p.requires_grad = True
        (dec(enc(x)) - x).pow(2).mean().backward()
        ret[rank] = dict(enc=[g.numpy() for g in g_enc], dec=[p.grad.numpy() for p in dec.parameters()], n=sync.n_allreduces, x=x.numpy())
    finally:
        dist.destroy_process_group()


def test_gradsync_averages_gradients_at_the_end_of_every_backward():
    """intro_tc_vae_b200.ddp.GradSync under the Soft-Intro requires_grad toggling: one flattened all-reduce per backward, and the
    averaged gradients equal the gradients of the mean loss over the concatenated batch."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29700 + (os.getpid() % 190)
    mp.spawn(_worker_gradsync, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0]["n"] == 2 and ret[1]["n"] == 2
    torch.manual_seed(0)
    enc, dec = torch.nn.Linear(6, 4), torch.nn.Linear(4, 6)
    x = torch.tensor(np.concatenate([ret[0]["x"], ret[1]["x"]]))
    (dec(enc(x)) - x).pow(2).mean().backward()
    for r in range(world):
        for got, p in zip(ret[r]["enc"], enc.parameters()):
            np.testing.assert_allclose(got, p.grad.numpy(), rtol=1e-5, atol=1e-7)
        for got, p in zip(ret[r]["dec"], dec.parameters()):
            np.testing.assert_allclose(got, p.grad.numpy(), rtol=1e-5, atol=1e-7)

"""Row-sharded TC-ELBO over NCCL vs the same global batch on one GPU (run under torchrun, >= 2 GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from intro_tc_vae_b200 import ops


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for B, D, N in ((256, 128, 16704), (1024, 128, 16704), (512, 64, 737280)):
        g = torch.Generator().manual_seed(B + D)
        mu_c = torch.randn(B, D, generator=g)
        lv_c = -2.0 + torch.randn(B, D, generator=g)
        eps_c = torch.randn(B, D, generator=g)
        w_c = torch.linspace(0.5, 1.5, B)
        b_loc = B // world
        lo = rank * b_loc
        # sharded: each rank owns rows [lo, lo + b_loc)
        mu = mu_c[lo:lo + b_loc].to(dev).requires_grad_(True)
        lv = lv_c[lo:lo + b_loc].to(dev).requires_grad_(True)
        z = ops.reparameterize(mu, lv, eps_c[lo:lo + b_loc].to(dev))
        tc = ops.total_correlation(z, mu, lv, N, reduce="none", group=dist.group.WORLD)
        kl = ops.kl_divergence(lv, mu, reduce="none")
        ((5.0 * tc + kl) * w_c[lo:lo + b_loc].to(dev)).sum().backward()
        # single-GPU evaluation of the whole batch (every rank does it redundantly)
        mu_f = mu_c.to(dev).requires_grad_(True)
        lv_f = lv_c.to(dev).requires_grad_(True)
        z_f = ops.reparameterize(mu_f, lv_f, eps_c.to(dev))
        tc_f = ops.total_correlation(z_f, mu_f, lv_f, N, reduce="none")
        kl_f = ops.kl_divergence(lv_f, mu_f, reduce="none")
        ((5.0 * tc_f + kl_f) * w_c.to(dev)).sum().backward()

        def rel(a, b):
            return ((a - b).abs().max() / b.abs().max()).item()
        e_tc = rel(tc.detach(), tc_f.detach()[lo:lo + b_loc])
        e_mu = rel(mu.grad, mu_f.grad[lo:lo + b_loc])
        e_lv = rel(lv.grad, lv_f.grad[lo:lo + b_loc])
        good = e_tc < 1e-6 and e_mu < 1e-5 and e_lv < 1e-5
        ok = ok and good
        print(f"rank {rank} B={B} D={D}: tc rel {e_tc:.1e}  dmu rel {e_mu:.1e}  dlv rel {e_lv:.1e}  {'OK' if good else 'MISMATCH'}", flush=True)
    # ---- the same shard through the peer-memory exchange (fused loss; three rounds exercise the buffer rotation)
    from intro_tc_vae_b200 import peer
    from intro_tc_vae_b200.graphs import GraphedKLLoss
    for B, D, N, beta in ((256, 128, 16704, 6.0), (1024, 128, 16704, 0.5), (512, 64, 737280, 2.0), (296, 20, 5000, 4.0)):
        b_loc = B // world
        lo = rank * b_loc
        exch = peer.PeerExchange(b_loc, D, dist.group.WORLD, dev)
        for rnd in range(3):
            g = torch.Generator().manual_seed(B + D + 17 * rnd)
            mu_c = torch.randn(B, D, generator=g)
            lv_c = -2.0 + torch.randn(B, D, generator=g)
            eps_c = torch.randn(B, D, generator=g)
            w_c = torch.linspace(0.5, 1.5, B)
            mu = mu_c[lo:lo + b_loc].to(dev).requires_grad_(True)
            lv = lv_c[lo:lo + b_loc].to(dev).requires_grad_(True)
            z = ops.reparameterize(mu, lv, eps_c[lo:lo + b_loc].to(dev))
            loss, kl, lqz, lqzp = ops.kl_tc_loss_terms(z, mu, lv, N, beta, exchange=exch)
            ((loss + 0.25 * kl + 0.5 * lqz - 0.125 * lqzp) * w_c[lo:lo + b_loc].to(dev)).sum().backward()
            mu_f = mu_c.to(dev).requires_grad_(True)
            lv_f = lv_c.to(dev).requires_grad_(True)
            z_f = ops.reparameterize(mu_f, lv_f, eps_c.to(dev))
            loss_f, kl_f, lqz_f, lqzp_f = ops.kl_tc_loss_terms(z_f, mu_f, lv_f, N, beta)
            ((loss_f + 0.25 * kl_f + 0.5 * lqz_f - 0.125 * lqzp_f) * w_c.to(dev)).sum().backward()

            def rel(a, b):
                return ((a - b).abs().max() / b.abs().max()).item()
            e_l = rel(loss.detach(), loss_f.detach()[lo:lo + b_loc])
            e_mu = rel(mu.grad, mu_f.grad[lo:lo + b_loc])
            e_lv = rel(lv.grad, lv_f.grad[lo:lo + b_loc])
            good = e_l < 1e-6 and e_mu < 1e-5 and e_lv < 1e-5
            ok = ok and good
            print(f"rank {rank} peer B={B} D={D} round {rnd}: loss rel {e_l:.1e}  dmu rel {e_mu:.1e}  dlv rel {e_lv:.1e}  {'OK' if good else 'MISMATCH'}", flush=True)
        del exch
    # ---- graph replay: peer exchange vs NCCL exchange, fresh inputs on every replay
    B, D, N, beta = 2048, 128, 16704, 0.5
    b_loc = B // world
    lo = rank * b_loc
    g_nccl = GraphedKLLoss(b_loc, D, N, beta, dev, group=dist.group.WORLD, exchange="nccl")
    g_peer = GraphedKLLoss(b_loc, D, N, beta, dev, group=dist.group.WORLD, exchange="peer")
    for rnd in range(4):
        g = torch.Generator().manual_seed(99 + rnd)
        mu_c, lv_c, eps_c = torch.randn(B, D, generator=g), -2.0 + torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
        a = [t[lo:lo + b_loc].to(dev) for t in (mu_c, lv_c, eps_c)]
        l0, dm0, dl0 = [t.clone() for t in g_nccl(*a)]
        l1, dm1, dl1 = [t.clone() for t in g_peer(*a)]

        def rel(a, b):
            return ((a - b).abs().max() / b.abs().max()).item()
        e = (abs(l0.item() - l1.item()) / abs(l0.item()), rel(dm1, dm0), rel(dl1, dl0))
        good = max(e) < 1e-5
        ok = ok and good
        print(f"rank {rank} graph peer-vs-nccl round {rnd}: loss rel {e[0]:.1e}  dmu rel {e[1]:.1e}  dlv rel {e[2]:.1e}  {'OK' if good else 'MISMATCH'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res = int(flag.item())
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if res == 1 else 1)       # graphs that captured collectives make destroy_process_group hang


if __name__ == "__main__":
    main()

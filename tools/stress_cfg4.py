"""BASELINE configs[3] on one GPU: global batch 32768, z_dim 512, rank 3 of 8 (rows 12288..16383).
Forward + backward through the raw op with row_offset, a few rows checked against the row-chunked CPU oracle."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from intro_tc_vae_b200 import _lib, ops  # noqa: F401
from oracle import tc_oracle as O


def main():
    dev = torch.device("cuda:0")
    Bg, D, P, r, N = 32768, 512, 8, 3, 737280
    bl = Bg // P
    lo = r * bl
    g = torch.Generator().manual_seed(7)
    mu_c = torch.randn(Bg, D, generator=g)
    lv_c = -2.0 + torch.randn(Bg, D, generator=g)
    z_c = mu_c + torch.randn(Bg, D, generator=g) * torch.exp(0.5 * lv_c)
    mu_all = mu_c.to(dev).requires_grad_(True)
    z = z_c[lo:lo + bl].to(dev).requires_grad_(True)
    lv = lv_c[lo:lo + bl].to(dev).requires_grad_(True)
    flags = _lib.EST_MSS | _lib.VAR_ROW | _lib.SAVE_FOR_BACKWARD
    torch.cuda.synchronize()
    iters = int(os.environ.get("STRESS_ITERS", "5"))          # the first iterations include lazy initialisation and cudaMalloc
    for it in range(iters):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        lq, lqp, _ = torch.ops.tcelbo.tc_forward(z, mu_all, lv, lo, N, flags)
        e1.record()
        w = torch.linspace(0.5, 1.5, bl, device=dev)
        ((lq - lqp) * w).sum().backward()
        e2.record()
        torch.cuda.synchronize()
        n = bl * Bg * D
        print(f"iter {it}: fwd {e0.elapsed_time(e1):.2f} ms, bwd {e1.elapsed_time(e2):.2f} ms -> "
              f"{n / (e0.elapsed_time(e2) * 1e-3) / 1e12:.3f} T log-densities/s; peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
        if it + 1 < iters:
            z.grad = lv.grad = mu_all.grad = None
    assert torch.isfinite(lq).all() and torch.isfinite(lqp).all()
    assert torch.isfinite(z.grad).all() and torch.isfinite(lv.grad).all() and torch.isfinite(mu_all.grad).all()
    ok = True
    t0 = time.time()
    for i in (0, 1, bl // 2, bl - 1):
        zi = z_c[lo + i:lo + i + 1].clone().requires_grad_(True)
        lvi = lv_c[lo + i:lo + i + 1].clone().requires_grad_(True)
        p_o, j_o = O.tc_terms_rows(zi, lvi, mu_c, lo + i, Bg, N)
        ((j_o - p_o) * w[i].item()).sum().backward()
        ep = abs(lqp[i].item() - p_o.item()) / abs(p_o.item())
        ej = abs(lq[i].item() - j_o.item()) / abs(j_o.item())
        egz = ((z.grad[i].cpu() - zi.grad[0]).abs().max() / zi.grad.abs().max()).item()
        egl = ((lv.grad[i].cpu() - lvi.grad[0]).abs().max() / lvi.grad.abs().max()).item()
        good = ep < 1e-5 and ej < 1e-5 and egz < 1e-4 and egl < 1e-4
        ok = ok and good
        print(f"row {lo + i}: log_qz_prod rel {ep:.1e}, log_qz rel {ej:.1e}, grad_z rel {egz:.1e}, grad_logvar rel {egl:.1e} {'OK' if good else 'MISMATCH'}")
    print(f"oracle rows took {time.time() - t0:.1f} s")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

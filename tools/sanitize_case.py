"""Small driver for compute-sanitizer (one tool per run):  compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_case.py
Runs the whole step (fused prologue -> forward sweep -> finalize -> backward prologue -> fused backward sweep -> finalize) and the
row-wise companions on the smoke shape, a ragged shape and a wide-latent shape; shapes are small because the tools slow kernels ~100x."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from intro_tc_vae_b200 import ops
from intro_tc_vae_b200.graphs import GraphedKLLoss
from intro_tc_vae_b200.losses import kl_tc_exp_elbo, reconstruction_loss


def main():
    dev = torch.device("cuda:0")
    for B, D in ((64, 128), (37, 20), (24, 512), (300, 64)):
        g = torch.Generator().manual_seed(B)
        mu = torch.randn(B, D, generator=g).to(dev).requires_grad_(True)
        lv = (-2.0 + torch.randn(B, D, generator=g)).to(dev).requires_grad_(True)
        eps = torch.randn(B, D, generator=g).to(dev)
        z = ops.reparameterize(mu, lv, eps)
        loss, kl = ops.kl_tc_loss_mean(z, mu, lv, 16704, 0.5)
        rows = ops.kl_tc_loss_terms(z, mu, lv, 16704, 512.0, "mws")[0]
        rec = torch.rand(B, device=dev, requires_grad=True)
        ee, _ = kl_tc_exp_elbo(z, mu, lv, rec, 16704, 512.0, 1.0 / 12288)
        full = ops.tc_terms(z, mu, lv, 16704, "mss", "col")
        x = torch.rand(B, 3, 8, 8, device=dev)
        r = torch.rand(B, 3, 8, 8, device=dev, requires_grad=True)
        (loss + kl + rows.mean() + ee + (full[1] - full[0]).mean() + reconstruction_loss(x, r, "mse", "mean")).backward()
        graphed = GraphedKLLoss(B, D, 16704, 0.5, dev, capture=False)
        graphed(mu.detach(), lv.detach(), eps)
        torch.cuda.synchronize()
        print(f"B={B} D={D}: loss {loss.item():.5f} direct {graphed.loss.item():.5f} grad norm {mu.grad.norm().item():.4f}", flush=True)


if __name__ == "__main__":
    main()

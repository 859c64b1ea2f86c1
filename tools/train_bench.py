"""Soft-Intro-TC training-step throughput on synthetic images (BASELINE configs[1] / [4] shapes).

    python tools/train_bench.py [--image 64] [--zdim 128] [--batch 64] [--steps 30] [--warmup 10]
    python -m torch.distributed.run --nproc-per-node N ... tools/train_bench.py   # data parallel, TC estimator row-sharded

The conv encoder/decoder are stock torch modules in the reference's "conv" recipe (models.py:8-55,190-300: 5x5 stem,
double-3x3 conv blocks with BatchNorm + LeakyReLU, AvgPool down / nearest-neighbour up, fc -> chunked mu/logvar); they
are NOT part of the accelerated path.  Every loss term (reparameterize, KL + TC, reconstruction, exp-ELBO) runs through
libtcelbo.so via intro_tc_vae_b200.solvers.IntroTCSovler.  Prints one JSON line (images/s over all ranks).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.distributed as dist
from intro_tc_vae_b200 import ops
from intro_tc_vae_b200.solvers import IntroTCSovler


def block(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout, eps=1e-4), nn.LeakyReLU(0.2, inplace=True),
                         nn.Conv2d(cout, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout, eps=1e-4), nn.LeakyReLU(0.2, inplace=True))


class Encoder(nn.Module):
    def __init__(self, cdim, zdim, channels, image_size):
        super().__init__()
        self.image_size = image_size
        layers = [nn.Conv2d(cdim, channels[0], 5, 1, 2, bias=False), nn.BatchNorm2d(channels[0], eps=1e-4), nn.LeakyReLU(0.2, inplace=True), nn.AvgPool2d(2)]
        cc, sz = channels[0], image_size // 2
        for ch in channels[1:]:
            layers += [block(cc, ch), nn.AvgPool2d(2)]
            cc, sz = ch, sz // 2
        layers.append(block(cc, cc))
        self.main = nn.Sequential(*layers)
        self.out_shape = (cc, sz, sz)
        self.fc = nn.Linear(cc * sz * sz, 2 * zdim)

    def forward(self, x):
        return self.fc(self.main(x).flatten(1)).chunk(2, dim=1)


class Decoder(nn.Module):
    def __init__(self, cdim, zdim, channels, in_shape):
        super().__init__()
        self.in_shape = in_shape
        cc = channels[-1]
        self.fc = nn.Sequential(nn.Linear(zdim, in_shape[0] * in_shape[1] * in_shape[2]), nn.LeakyReLU(0.2, inplace=True))
        layers = []
        for ch in channels[::-1]:
            layers += [block(cc, ch), nn.Upsample(scale_factor=2, mode="nearest")]
            cc = ch
        layers += [block(cc, cc), nn.Conv2d(cc, cdim, 5, 1, 2), nn.Sigmoid()]
        self.main = nn.Sequential(*layers)

    def forward(self, z):
        return self.main(self.fc(z).view(z.size(0), *self.in_shape))


class ConvVAE(nn.Module):
    def __init__(self, cdim, zdim, channels, image_size):
        super().__init__()
        self.cdim, self.zdim = cdim, zdim
        self.encoder = Encoder(cdim, zdim, channels, image_size)
        self.decoder = Decoder(cdim, zdim, channels, self.encoder.out_shape)

    def encode(self, x):
        return self.encoder(x)

    def decode(self, z):
        return self.decoder(z)

    def sample(self, z):
        return self.decoder(z)

    def forward(self, x, deterministic=False):
        mu, logvar = self.encode(x)
        z = mu if deterministic else ops.reparameterize(mu, logvar)
        return mu, logvar, z, self.decode(z)


class _Dataset:
    def __len__(self):
        return 16704


class DPSolver(IntroTCSovler):
    """Data-parallel harness: average .grad over the ranks after every backward (one flattened all-reduce)."""

    def sync_gradients(self, params):
        if self.process_group is None:
            return
        grads = [p.grad for p in params if p.grad is not None]
        if not grads:
            return
        flat = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.process_group)
        for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
            g.copy_(f)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--image", type=int, default=64)
    ap.add_argument("--zdim", type=int, default=128)
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--peer", action="store_true", help="N>1: run the estimator's exchange steps over NVLink peer memory instead of NCCL")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    torch.manual_seed(0)
    channels = [64, 128, 256, 512] if args.image == 64 else [64, 128, 256, 512, 512]       # train.py:61-70
    model = ConvVAE(3, args.zdim, channels, args.image).to(dev)
    opt_e = torch.optim.Adam(model.encoder.parameters(), lr=2e-4)
    opt_d = torch.optim.Adam(model.decoder.parameters(), lr=2e-4)
    solver = DPSolver(dataset=_Dataset(), model=model, batch_size=args.batch, optimizer_e=opt_e, optimizer_d=opt_d, recon_loss_type="mse",
                      beta_kl=0.5, beta_rec=0.75, beta_neg=512.0, gamma_r=1e-8, device=dev, use_amp=False, grad_scaler=None,
                      writer=None, test_iter=1000, clip=100.0)
    solver.process_group = group
    if group is not None and args.peer:
        from intro_tc_vae_b200.peer import PeerExchange
        solver.peer_exchange = PeerExchange(args.batch, args.zdim, group, dev)
    batch = torch.rand(args.batch, 3, args.image, args.image, device=dev)
    out = None
    for it in range(args.warmup):
        out = solver.train_step(batch, it)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(args.steps):
        out = solver.train_step(batch, args.warmup + it)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "soft_intro_tc_train_images_per_s", "value": args.batch * world / (t.item() * 1e-3), "unit": "images/s",
                          "n_gpus": world, "ms_per_step": t.item(), "per_gpu_batch": args.batch, "image": args.image, "z_dim": args.zdim,
                          "last_losses": {k: (round(v, 5) if v is not None else None) for k, v in out.items()},
                          "params_M": round(sum(p.numel() for p in model.parameters()) / 1e6, 2),
                          "exchange": ("peer memory" if (group is not None and args.peer) else ("nccl" if group is not None else "none")),
                          "note": "fp32 (TF32 convs off), stock torch conv modules, all loss terms through libtcelbo.so, eager launches"}), flush=True)
    if world > 1:
        dist.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()

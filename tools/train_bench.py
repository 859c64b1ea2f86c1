"""Soft-Intro-TC training-step throughput on synthetic images (BASELINE configs[1] / [4] shapes; BASELINE.json metric 2).

    python tools/train_bench.py [--image 64] [--zdim 128] [--batch 64] [--steps 30] [--warmup 10] [--eager]
    python -m torch.distributed.run --nproc-per-node N ... tools/train_bench.py   # data parallel, TC estimator row-sharded

Harness only.  The conv encoder/decoder are stock torch modules in the reference's "conv" recipe (models.py:8-55,190-300:
5x5 stem, double-3x3 conv blocks with BatchNorm + LeakyReLU, AvgPool down / nearest-neighbour up, fc -> chunked mu/logvar);
they are NOT part of the accelerated path and stand in for the reference's models.SoftIntroVAE, which is not on the GPU box.
The update itself is intro_tc_vae_b200.train_step.SoftIntroTCStep: the Soft-Intro-TC step of solvers/intro.py:56-196 with every
loss term (reparameterize, KL + TC, reconstruction, exp-ELBO) on libtcelbo.so, captured in two CUDA graphs, no host
synchronisation.  Prints one JSON line (images/s over all ranks); bench.py calls measure() for its train_images_per_s keys.
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
from intro_tc_vae_b200.train_step import SoftIntroTCStep


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


def measure(image, zdim, batch, steps, warmup, dev, group=None, peer=False, capture=True):
    """images/s of the Soft-Intro-TC update on a fixed synthetic batch [batch, 3, image, image] per rank (max over ranks)."""
    world = dist.get_world_size(group) if group is not None else 1
    torch.manual_seed(0)
    channels = [64, 128, 256, 512] if image == 64 else [64, 128, 256, 512, 512]       # train.py:61-70
    model = ConvVAE(3, zdim, channels, image).to(dev)
    opt_e = torch.optim.Adam(model.encoder.parameters(), lr=2e-4, capturable=capture)
    opt_d = torch.optim.Adam(model.decoder.parameters(), lr=2e-4, capturable=capture)
    exchange = None
    if group is not None and peer:
        from intro_tc_vae_b200.peer import PeerExchange
        exchange = PeerExchange(batch, zdim, group, dev)
    real = torch.rand(batch, 3, image, image, device=dev)
    noise = torch.randn(batch, zdim, device=dev)
    step = SoftIntroTCStep(model, opt_e, opt_d, 16704, real.shape, recon_loss_type="mse", beta_kl=0.5, beta_rec=0.75, beta_neg=512.0,
                           gamma_r=1e-8, clip=100.0, group=group, exchange=exchange, capture=capture)
    for _ in range(warmup):
        out = step(real, noise)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step(real, noise)
    e1.record()
    torch.cuda.synchronize()
    step.check_finite()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    return {"metric": "soft_intro_tc_train_images_per_s", "value": batch * world / (ms * 1e-3), "unit": "images/s", "n_gpus": world,
            "ms_per_step": ms, "per_gpu_batch": batch, "image": image, "z_dim": zdim,
            "last_losses": {k: round(v.item(), 5) for k, v in out.items()},
            "params_M": round(sum(p.numel() for p in model.parameters()) / 1e6, 2),
            "exchange": ("peer memory" if exchange is not None else ("nccl" if group is not None else "none")),
            "launch": "two CUDA graphs per step (encoder phase, decoder phase), no host synchronisation" if capture else "eager",
            "note": "fp32 (cuDNN TF32 convs as in the reference's defaults), stock torch conv modules, all loss terms through libtcelbo.so"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--image", type=int, default=64)
    ap.add_argument("--zdim", type=int, default=128)
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--peer", action="store_true", help="N>1: run the estimator's exchange steps over NVLink peer memory instead of NCCL")
    ap.add_argument("--eager", action="store_true", help="issue the step eagerly instead of replaying the two captured graphs")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    line = measure(args.image, args.zdim, args.batch, args.steps, args.warmup, dev, group, args.peer, capture=not args.eager)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()

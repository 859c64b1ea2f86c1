"""Sweep the fused-backward kernel variants (tcelbo_set_tuning) on one GPU and print per-variant times.

    python tools/tune_bwd.py [--batch 8192] [--zdim 128] [--variants 0,1,2,...]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from intro_tc_vae_b200 import _lib, ops  # noqa: F401  (registers torch.ops.tcelbo)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--zdim", type=int, default=128)
    ap.add_argument("--variants", default="0,1,2,3,4", help="tuning points of csrc/tc_bwd_ds.cu (0 = shipped)")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--rows", type=int, default=0, help="rows of this shard (default: the whole batch); emulates one rank of a row-sharded job")
    ap.add_argument("--fwd-seg", default="", help="comma list of forward segment-length targets (column tiles per CTA)")
    ap.add_argument("--fwd-wave", default="", help="comma list: CTAs per SM the forward grid is sized for (0 = as resident)")
    ap.add_argument("--fwd-map", default="", help="comma list: 0 = shipped forward mapping (16 dims per lane), 1 = round 1's 32 dims per lane")
    ap.add_argument("--seg", default="0", help="comma list of segment-length targets (column tiles per CTA; 0 = default)")
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    B, D, N = args.batch, args.zdim, 16704
    g = torch.Generator().manual_seed(1234)
    mu = torch.randn(B, D, generator=g).to(dev)
    lv = (-2.0 + torch.randn(B, D, generator=g)).to(dev)
    z = mu + torch.randn(B, D, generator=g).to(dev) * torch.exp(0.5 * lv)
    R = args.rows if args.rows > 0 else B
    z, lv = z[:R].contiguous(), lv[:R].contiguous()
    flags = _lib.EST_MSS | _lib.VAR_ROW | _lib.SAVE_FOR_BACKWARD
    lq, lqp, ws = torch.ops.tcelbo.tc_forward(z, mu, lv, 0, N, flags)
    gj = torch.full((R,), 0.5 / B, device=dev)
    gp = -gj
    ref = None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for v, seg in [(int(x), int(y)) for x in args.variants.split(",") for y in args.seg.split(",")]:
        _lib.check(lib.tcelbo_set_tuning(b"bwd_variant", v), "set_tuning")
        _lib.check(lib.tcelbo_set_tuning(b"bwd_seg_tiles", seg), "set_tuning")
        try:
            for _ in range(2):
                out = torch.ops.tcelbo.tc_backward(z, mu, lv, 0, N, flags, gj, gp, ws)
            torch.cuda.synchronize()
        except Exception as e:                                  # a variant that cannot launch (smem / regs)
            print(f"variant {v:2d}: FAILED {e}")
            continue
        ts = []
        for _ in range(args.reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = torch.ops.tcelbo.tc_backward(z, mu, lv, 0, N, flags, gj, gp, ws)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        if ref is None:
            ref = [t.clone() for t in out]
            err = 0.0
        else:
            err = max(((a - b).abs().max() / b.abs().max()).item() for a, b in zip(out, ref))
        ts.sort()
        print(f"variant {v:2d} seg {seg:3d}: backward total {ts[len(ts)//2]:.3f} ms (min {ts[0]:.3f})  max rel diff vs first {err:.1e}")
    lib.tcelbo_set_tuning(b"bwd_variant", 0)
    lib.tcelbo_set_tuning(b"bwd_seg_tiles", 0)
    # forward sweep: segment-length targets, lane mapping, grid wave size
    def sweep_fwd(key, label, values):
        ref = None
        for v in values:
            _lib.check(lib.tcelbo_set_tuning(key, v), "set_tuning")
            for _ in range(2):
                out = torch.ops.tcelbo.tc_forward(z, mu, lv, 0, N, flags)
            ts = []
            for _ in range(args.reps):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = torch.ops.tcelbo.tc_forward(z, mu, lv, 0, N, flags)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            if ref is None:
                ref = [t.clone() for t in out[:2]]
            err = max(((a - b).abs().max() / b.abs().max()).item() for a, b in zip(out[:2], ref))
            ts.sort()
            print(f"forward {label} {v:3d}: forward total {ts[len(ts)//2]:.3f} ms (min {ts[0]:.3f})  max rel diff vs first {err:.1e}")
        lib.tcelbo_set_tuning(key, 0)

    sweep_fwd(b"fwd_seg_tiles", "seg", [int(y) for y in args.fwd_seg.split(",") if y])
    sweep_fwd(b"fwd_map", "map", [int(y) for y in args.fwd_map.split(",") if y])
    sweep_fwd(b"fwd_wave", "wave", [int(y) for y in args.fwd_wave.split(",") if y])

if __name__ == "__main__":
    main()

"""Time the forward-sweep variants (tcelbo_set_tuning("fwd_variant", v)) and check they agree."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from intro_tc_vae_b200 import _lib, ops  # noqa: F401


def main():
    lib = _lib.load()
    dev = torch.device("cuda:0")
    B, D, N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 128, 16704
    g = torch.Generator().manual_seed(1234)
    mu = torch.randn(B, D, generator=g).to(dev)
    lv = (-2.0 + torch.randn(B, D, generator=g)).to(dev)
    z = mu + torch.randn(B, D, generator=g).to(dev) * torch.exp(0.5 * lv)
    flags = _lib.EST_MSS | _lib.VAR_ROW | _lib.SAVE_FOR_BACKWARD
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ref = None
    for v in (0, 1):
        lib.tcelbo_set_tuning(b"fwd_variant", v)
        for _ in range(2):
            out = torch.ops.tcelbo.tc_forward(z, mu, lv, 0, N, flags)
        ts = []
        for _ in range(5):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = torch.ops.tcelbo.tc_forward(z, mu, lv, 0, N, flags)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        if ref is None:
            ref = [o.clone() for o in out[:2]]
            err = 0.0
        else:
            err = max(((a - b).abs().max() / b.abs().max()).item() for a, b in zip(out[:2], ref))
        print(f"fwd variant {v}: forward total {ts[2]:.3f} ms (min {ts[0]:.3f})  max rel diff vs variant 0 {err:.1e}")
    lib.tcelbo_set_tuning(b"fwd_variant", 0)


if __name__ == "__main__":
    main()

"""Probe: torch symmetric memory on this box (rendezvous, peer reads from a custom kernel via ctypes, barrier in a CUDA graph)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    os.makedirs("gpurun_out", exist_ok=True)
    log = open(f"gpurun_out/symm_probe_rank{rank}.log", "w")
    global print
    _print = print
    def print(*a, **k):
        k.pop("flush", None)
        _print(*a, file=log, flush=True)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    t = symm_mem.empty(1024, 128, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "rendezvous ok; world", hdl.world_size, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast_ptr", hex(hdl.multicast_ptr) if hdl.multicast_ptr else None, flush=True)
    t.fill_(float(rank + 1))
    hdl.barrier(channel=0)
    peer = hdl.get_buffer((rank + 1) % world, (1024, 128), torch.float32)
    print(rank, "peer value", peer[5, 7].item(), flush=True)
    # timing of barrier and of a peer read
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        hdl.barrier(channel=0)
    e1.record()
    torch.cuda.synchronize()
    print(rank, "barrier us", e0.elapsed_time(e1) * 10, flush=True)
    dst = torch.empty(1024, 128, device=dev)
    e0.record()
    for _ in range(100):
        dst.copy_(peer)
    e1.record()
    torch.cuda.synchronize()
    print(rank, "peer copy 512 KiB us", e0.elapsed_time(e1) * 10, flush=True)
    # graph capture of barrier + peer copy
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            hdl.barrier(channel=0); dst.copy_(peer)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            hdl.barrier(channel=0)
            dst.copy_(peer)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        print(rank, "graph capture of barrier+peer copy ok", dst[0, 0].item(), flush=True)
    except Exception as exc:
        print(rank, "graph capture failed:", type(exc).__name__, exc, flush=True)
    # NCCL reference timings for the same message sizes
    x = torch.ones(1024, 128, device=dev); out = torch.empty(world * 1024, 128, device=dev)
    for _ in range(5):
        dist.all_gather_into_tensor(out, x)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(100):
        dist.all_gather_into_tensor(out, x)
    e1.record()
    torch.cuda.synchronize()
    print(rank, "nccl all_gather 512 KiB/rank us", e0.elapsed_time(e1) * 10, flush=True)
    y = torch.empty(1024, 128, device=dev)
    for _ in range(5):
        dist.reduce_scatter_tensor(y, out)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(100):
        dist.reduce_scatter_tensor(y, out)
    e1.record()
    torch.cuda.synchronize()
    print(rank, "nccl reduce_scatter us", e0.elapsed_time(e1) * 10, flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()

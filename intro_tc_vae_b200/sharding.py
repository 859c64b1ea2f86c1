"""Row sharding of one global batch across the ranks of a process group.

Rank r of P owns global rows [r*B_loc, (r+1)*B_loc).  Rows are independent given all columns, so the
only exchange steps are an all-gather of the column operand before the forward sweep and the matching
reduce-scatter of its gradient after the backward sweep; both run over NCCL (NVLink/NVSwitch) in
production and over gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor


class GatherRows(torch.autograd.Function):
    """``[B_loc, D] -> [P*B_loc, D]`` all-gather whose backward is reduce-scatter(sum)."""

    @staticmethod
    def forward(ctx, x: Tensor, group) -> Tensor:
        import torch.distributed as dist
        ctx.group = group
        ctx.rows = x.shape[0]
        world = dist.get_world_size(group)
        out = torch.empty(world * x.shape[0], x.shape[1], dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        import torch.distributed as dist
        out = torch.empty(ctx.rows, g.shape[1], dtype=g.dtype, device=g.device)
        dist.reduce_scatter_tensor(out, g.contiguous(), op=dist.ReduceOp.SUM, group=ctx.group)
        return out, None


def gather_rows(x: Tensor, group) -> Tensor:
    return GatherRows.apply(x, group)


def shard_rows(group, b_loc: int, device=None) -> Tuple[int, int]:
    """(row_offset, b_glob) of this rank for equal shards of ``b_loc`` rows.

    The row offsets and the importance weights assume every rank holds the same number of rows.  With ``device`` given the
    shard sizes are all-gathered (one 8-byte collective, no host synchronisation) and a device-side assertion fails the
    run loudly when they differ (e.g. a ragged last batch that is not ragged the same way on every rank) instead of
    computing with wrong offsets or hanging in the all-gather of the rows."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if device is not None and world > 1:
        mine = torch.tensor([b_loc], dtype=torch.int64, device=device)
        sizes = torch.empty(world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(sizes, mine, group=group)
        torch._assert_async((sizes == b_loc).all(), "tcelbo: the ranks of the process group hold different numbers of rows")
    return dist.get_rank(group) * b_loc, world * b_loc

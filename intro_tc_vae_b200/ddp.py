"""Gradient averaging for data-parallel runs of the reference's UNMODIFIED ``train_step``.

The reference has no data parallelism, and ``torch.nn.parallel.DistributedDataParallel`` does not
fit its Soft-Intro step: ``requires_grad`` is toggled on the encoder / decoder every half step (solvers/intro.py:66-69,
119-122) and sub-modules are called directly (``model.encoder``, ``model.decoder``, ``model.sample``), so DDP's forward
hooks never arm its reducer.  :class:`GradSync` hooks the autograd engine instead: every parameter gets a
post-accumulate-grad hook; the first one to fire in a backward pass queues ONE end-of-backward callback, which flattens
the gradients that pass produced, all-reduces them once (NCCL over NVLink in production, gloo in the CPU tests), and
writes the averages back -- before ``clip_grad_norm_`` and ``optimizer.step()`` run, exactly where the reference
expects finished gradients (solvers/intro.py:110-116, 153-160).
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors


class GradSync:
    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self._ready: List[torch.nn.Parameter] = []
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        self.n_allreduces = 0

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if self.world == 1:
            return
        if not self._ready:
            torch.autograd.Variable._execution_engine.queue_callback(self._finish)
        self._ready.append(p)

    def _finish(self) -> None:
        params, self._ready = self._ready, []
        grads = [p.grad for p in params if p.grad is not None]
        if not grads:
            return
        flat = _flatten_dense_tensors(grads)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(self.world)
        for g, f in zip(grads, _unflatten_dense_tensors(flat, grads)):
            g.copy_(f)
        self.n_allreduces += 1

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []

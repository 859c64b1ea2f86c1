"""CUDA-graph capture of one TC-ELBO loss evaluation (forward + backward) for fixed shapes.

A row-sharded step at 8 GPUs is ~1 ms of kernels behind ~10 launches and two NCCL collectives, which
eager PyTorch cannot issue fast enough from one Python thread; replaying a captured graph removes the
launch gaps.  Everything the library launches is capturable by construction (no host synchronisation,
no allocation inside the C ABI, explicit stream argument)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import ops


class GraphedKLLoss:
    """``loss, dmu, dlogvar = graphed(mu, logvar, eps)`` with

        z    = mu + eps * exp(0.5 * logvar)                       (ops.py:183-185)
        loss = mean_i[(beta - 1) * tc_i + kl_i]                   (solvers/tc.py:69-89, reduce="mean")

    evaluated and differentiated inside one CUDA graph.  ``mu``/``logvar``/``eps`` are ``[b_loc, d]`` fp32 CUDA
    tensors (or pinned host tensors: they are copied into the graph's static inputs on the current stream);
    the returned tensors are the graph's static outputs and are overwritten by the next call.
    ``group``: row-shard over the ranks of a process group (NCCL all-gather / reduce-scatter are captured too).
    ``exchange="peer"``: do the two exchange steps over NVLink peer memory inside the library's kernels
    (:mod:`intro_tc_vae_b200.peer`) instead of NCCL; ``"nccl"`` keeps the collectives; ``"auto"`` picks peer memory
    when the group has more than one rank and symmetric memory can be set up, NCCL otherwise.
    """

    def __init__(self, b_loc: int, d: int, dataset_size: int, beta: float, device, group=None,
                 estimator: str = "mss", warmup: int = 3, exchange: str = "nccl"):
        self.device = torch.device(device)
        self.exchange = None
        self.exchange_kind = "none"
        if group is not None:
            import torch.distributed as dist
            if dist.get_world_size(group) > 1:
                self.exchange_kind = "nccl"
                if exchange not in ("nccl", "peer", "auto"):
                    raise ValueError(f"exchange must be 'nccl', 'peer' or 'auto', got {exchange!r}")
                if exchange in ("peer", "auto"):
                    from . import peer
                    try:
                        self.exchange = peer.PeerExchange(b_loc, d, group, self.device)
                        self.exchange_kind = "peer"
                    except Exception:
                        if exchange == "peer":
                            raise
        self.mu = torch.zeros(b_loc, d, device=self.device, requires_grad=True)
        self.logvar = torch.zeros(b_loc, d, device=self.device, requires_grad=True)
        self.eps = torch.zeros(b_loc, d, device=self.device)
        self._args = (int(dataset_size), float(beta), estimator, group)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.loss: Optional[Tensor] = None
        with torch.no_grad():                              # benign values for the warm-up / capture passes
            self.logvar.fill_(-2.0)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._eager_step()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.mu.grad = None
        self.logvar.grad = None
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.loss = self._eager_step()
        self.graph = graph

    def _eager_step(self) -> Tensor:
        n, beta, estimator, group = self._args
        self.mu.grad = None
        self.logvar.grad = None
        z = ops.reparameterize(self.mu, self.logvar, self.eps)
        loss = ops.kl_tc_loss_terms(z, self.mu, self.logvar, n, beta, estimator, group, self.exchange)[0].mean()
        loss.backward()
        return loss.detach()

    def __call__(self, mu: Tensor, logvar: Tensor, eps: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        with torch.no_grad():
            self.mu.copy_(mu, non_blocking=True)
            self.logvar.copy_(logvar, non_blocking=True)
            self.eps.copy_(eps, non_blocking=True)
        self.graph.replay()
        return self.loss, self.mu.grad, self.logvar.grad

    def replay(self) -> Tuple[Tensor, Tensor, Tensor]:
        """Re-run on the inputs already resident in the static buffers."""
        self.graph.replay()
        return self.loss, self.mu.grad, self.logvar.grad

"""Encoder head layout: ``mu`` and ``logvar`` as two contiguous ``[B, D]`` buffers straight from the
head GEMMs, instead of pitch-``2*D`` chunk views of one ``[B, 2*D]`` output (reference ``models.py:240-244``:
``y = self.fc(y); mu, logvar = y.chunk(2, dim=1)``).

The head stays a plain library GEMM (cuBLAS through ``torch.addmm``) -- the conv encoder is outside the accelerated path -- but
it is issued as two GEMMs over the two halves of the SAME ``nn.Linear`` parameters, each writing its own dense output:

* the TC kernels then read unit-pitch rows (the chunk views cost them a 2x row pitch, not a copy);
* ``mu_out`` lets the GEMM write ``mu`` into a caller-provided buffer (e.g. a staging buffer that is published to the other
  ranks).  Writing straight into the peer exchange's symmetric buffer is deliberately NOT wired up: that buffer may be
  rewritten only two exchanges after it was read (``peer.py``), and the Soft-Intro step encodes two batches back to back before
  either is exchanged (solvers/intro.py:81-89), so a GEMM-side write would need its own cross-rank barrier -- more than the
  4 us copy it would save.

``attach(model.encoder)`` swaps the head in without touching the reference's module definition or its parameters
(checkpoints stay interchangeable: the weights are still ``encoder.fc.weight`` / ``.bias``).
"""
from __future__ import annotations

import types
from typing import Optional, Tuple

import torch
from torch import Tensor


class _SplitLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y: Tensor, weight: Tensor, bias: Optional[Tensor], mu_out: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
        d = weight.shape[0] // 2
        b = y.shape[0]
        mu = mu_out if mu_out is not None else torch.empty(b, d, dtype=y.dtype, device=y.device)
        logvar = torch.empty(b, d, dtype=y.dtype, device=y.device)
        if bias is not None:
            torch.addmm(bias[:d], y, weight[:d].t(), out=mu)
            torch.addmm(bias[d:], y, weight[d:].t(), out=logvar)
        else:
            torch.mm(y, weight[:d].t(), out=mu)
            torch.mm(y, weight[d:].t(), out=logvar)
        ctx.save_for_backward(y, weight)
        ctx.has_bias = bias is not None
        if mu_out is not None:
            ctx.mark_dirty(mu_out)
        return mu, logvar

    @staticmethod
    def backward(ctx, g_mu: Tensor, g_lv: Tensor):
        y, weight = ctx.saved_tensors
        d = weight.shape[0] // 2
        g_y = g_w = g_b = None
        if ctx.needs_input_grad[0]:
            g_y = torch.addmm(g_mu @ weight[:d], g_lv, weight[d:])
        if ctx.needs_input_grad[1]:
            g_w = torch.cat([g_mu.t() @ y, g_lv.t() @ y], dim=0)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            g_b = torch.cat([g_mu.sum(0), g_lv.sum(0)])
        return g_y, g_w, g_b, None


def split_head(y: Tensor, fc: torch.nn.Linear, mu_out: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """``fc(y).chunk(2, dim=1)`` (models.py:242-244) with both halves dense; ``mu_out`` optionally receives ``mu`` in place."""
    if fc.out_features % 2:
        raise ValueError("the head must emit 2 * z_dim features")
    return _SplitLinear.apply(y, fc.weight, fc.bias, mu_out)


def attach(encoder: torch.nn.Module) -> torch.nn.Module:
    """Give ``encoder`` (reference ``models.Encoder``: ``main`` conv stack + ``fc``) a forward that ends in :func:`split_head`
    instead of ``fc`` + ``chunk`` (models.py:240-244).  Parameters and state dict are unchanged."""
    def forward(self, x: Tensor):
        y = self.main(x).view(x.size(0), -1)
        return split_head(y, self.fc)
    encoder.forward = types.MethodType(forward, encoder)
    return encoder

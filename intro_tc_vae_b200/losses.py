"""Reconstruction loss with the reference's signature (reference ``ops.py:188-236``).

Adjacent to the TC path: a streaming per-sample reduction over C*H*W pixels that
feeds the exp-ELBO terms.  Runs as torch ops on the caller's device for now.
"""
from __future__ import annotations

import torch.nn.functional as F
from torch import Tensor


def reconstruction_loss(x: Tensor, recon_x: Tensor, loss_type: str = "mse", reduction: str = "sum") -> Tensor:
    if x.size(0) == 0:
        raise AssertionError("empty batch")
    if reduction not in ("sum", "mean", "none"):
        raise NotImplementedError(reduction)
    recon_x = recon_x.reshape(recon_x.size(0), -1)
    x = x.reshape(x.size(0), -1).detach()
    if loss_type == "mse":
        err = F.mse_loss(recon_x, x, reduction="none")
    elif loss_type == "l1":
        err = F.l1_loss(recon_x, x, reduction="none")
    elif loss_type == "bce":
        err = F.binary_cross_entropy(recon_x, x, reduction="none")
    else:
        raise NotImplementedError(loss_type)
    err = err.sum(1)
    if reduction == "sum":
        return err.sum()
    if reduction == "mean":
        return err.mean()
    return err

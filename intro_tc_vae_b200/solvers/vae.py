"""Base solver: loss composition + the plain VAE update (reference ``solvers/vae.py:26-136``).

Kept to what the TC-ELBO path needs: constructor signature, ``compute_kl_loss`` / ``compute_rec_loss``
(the two virtuals the TC solvers override or call), the VAE ``train_step`` and scalar logging.  Image
dumps and disentanglement metrics of the reference (solvers/vae.py:138-254) are out of scope.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
from torch import Tensor

from .. import ops
from ..losses import reconstruction_loss
from ..utils import SingletonWriter


class VAESolver:
    def __init__(self, dataset, model, batch_size: int, optimizer_e, optimizer_d, recon_loss_type: str,
                 beta_kl: float, beta_rec: float, device, use_amp: bool, grad_scaler,
                 writer=None, test_iter: int = 1000, clip: Optional[float] = None):
        self.dataset = dataset
        self.model = model
        self.batch_size = batch_size
        self.optimizer_e = optimizer_e
        self.optimizer_d = optimizer_d
        self.recon_loss_type = recon_loss_type
        self.beta_kl = beta_kl
        self.beta_rec = beta_rec
        self.device = device
        self.use_amp = use_amp            # plumbed through and unused, exactly like the reference
        self.grad_scaler = grad_scaler
        self.writer = writer
        self.test_iter = test_iter
        self.clip = clip
        # losses are normalised by the image size C*H*W (solvers/vae.py:61)
        self.scale = 1 / (self.model.cdim * self.model.encoder.image_size ** 2)

    # ---- the two loss virtuals ---------------------------------------------------------------------
    def compute_kl_loss(self, z: Optional[Tensor], mu: Tensor, logvar: Tensor, reduce: str = "mean",
                        beta: float = None, write: bool = False) -> Tensor:
        """solvers/vae.py:63-77: ``beta * KL``."""
        if beta is None:
            beta = self.beta_kl
        kl = ops.kl_divergence(logvar, mu, reduce=reduce)
        if write:
            self.write_scalar(SingletonWriter().cur_iter, "kl_loss_unscaled", kl)
        return beta * kl

    def compute_rec_loss(self, x, recon_x, reduction="sum", beta: float = None, write: bool = False) -> Tensor:
        """solvers/vae.py:79-87: ``beta * reconstruction_loss``."""
        if beta is None:
            beta = self.beta_rec
        rec = reconstruction_loss(x, recon_x, self.recon_loss_type, reduction)
        if write:
            self.write_scalar(SingletonWriter().cur_iter, "r_loss_unscaled", rec)
        return beta * rec

    # ---- hooks for data-parallel training (no-ops on one GPU) -----------------------------------------
    def sync_gradients(self, params: Iterable[torch.nn.Parameter]) -> None:
        """Called after every ``backward()``; a data-parallel harness all-reduces ``.grad`` here."""

    # ---- VAE update (solvers/vae.py:89-136) -----------------------------------------------------------
    def train_step(self, batch: Tensor, cur_iter: int) -> dict:
        if batch.dim() == 3:
            batch = batch.unsqueeze(0)
        real = batch.to(self.device)

        mu, logvar, z, rec = self.model(real)
        loss_rec = self.compute_rec_loss(real, rec, reduction="mean", write=True)
        loss_kl = self.compute_kl_loss(z, mu, logvar, write=True)
        loss = self.scale * (loss_rec + loss_kl)

        self.optimizer_d.zero_grad()
        self.optimizer_e.zero_grad()
        loss.backward()
        self.sync_gradients(self.model.parameters())
        total_norm = None
        if self.clip:
            total_norm = torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip).item()
        self.optimizer_e.step()
        self.optimizer_d.step()

        if torch.isnan(loss):
            raise RuntimeError("NaN loss")
        if self.writer:
            self.write_scalars(cur_iter, losses=dict(r_loss=loss_rec.item(), kl_loss=loss_kl.item()))
            if self.clip:
                self.writer.add_scalar("total_norm", total_norm, global_step=cur_iter)
            self.writer.flush()
        return {"loss_enc": loss.item(), "loss_dec": loss.item(), "loss_kl": loss_kl.item(),
                "loss_rec": loss_rec.item(), "L2": total_norm}

    # ---- scalar logging (solvers/vae.py:173-187) ---------------------------------------------------------
    def write_scalar(self, cur_iter: int, tag: str, value: Tensor):
        if self.writer and value.dim() == 0:
            self.writer.add_scalar(tag, value.data.item(), global_step=cur_iter)

    def write_scalars(self, cur_iter: int, losses: dict, **kwargs):
        if self.writer is not None:
            self.write_losses(cur_iter, losses)
            for name, value in kwargs.items():
                self.writer.add_scalar(name, value, global_step=cur_iter)

    def write_losses(self, cur_iter: int, losses: dict):
        if self.writer is not None:
            self.writer.add_scalars("losses", losses, global_step=cur_iter)

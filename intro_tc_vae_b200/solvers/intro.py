"""Soft-Intro two-phase update (reference ``solvers/intro.py:17-196``).

The introspective objective adds ``exp(-2*scale*(rec_i + kl_i))`` terms (per sample, then batch mean)
for reconstructed and generated batches to the encoder loss (solvers/intro.py:84-108); ``kl_i`` comes
from ``compute_kl_loss(reduce="none", beta=beta_neg)``, i.e. the TC estimator in ``IntroTCSovler``.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from .. import ops
from ..losses import exp_elbo
from .vae import VAESolver


class IntroSolver(VAESolver):
    def __init__(self, dataset, model, batch_size: int, optimizer_e, optimizer_d, recon_loss_type: str,
                 beta_kl: float, beta_rec: float, beta_neg: float, gamma_r: float, device, use_amp: bool,
                 grad_scaler, writer=None, test_iter: int = 1000, clip: Optional[float] = None):
        super().__init__(dataset, model, batch_size, optimizer_e, optimizer_d, recon_loss_type, beta_kl, beta_rec,
                         device, use_amp, grad_scaler, writer, test_iter, clip)
        self.beta_neg = beta_neg
        self.gamma_r = gamma_r

    def _set_trainable(self, encoder: bool, decoder: bool) -> None:
        for p in self.model.encoder.parameters():
            p.requires_grad = encoder
        for p in self.model.decoder.parameters():
            p.requires_grad = decoder

    def _per_sample(self, rec_loss: Tensor) -> Tensor:
        while rec_loss.dim() > 1:                         # solvers/intro.py:94-100
            rec_loss = rec_loss.sum(-1)
        return rec_loss

    def exp_elbo(self, rec_per_sample: Tensor, kl_per_sample: Tensor) -> Tensor:
        """solvers/intro.py:102-103."""
        return exp_elbo(rec_per_sample, kl_per_sample, self.scale)

    def train_step(self, batch: Tensor, cur_iter: int) -> dict:
        if batch.dim() == 3:
            batch = batch.unsqueeze(0)
        b_size = batch.size(0)
        noise = torch.randn(size=(b_size, self.model.zdim)).to(self.device)     # CPU generator, as in the reference
        real = batch.to(self.device)
        model = self.model

        # ------------------------------ encoder update (solvers/intro.py:65-116)
        self._set_trainable(encoder=True, decoder=False)
        fake = model.sample(noise)
        real_mu, real_logvar = model.encode(real)
        z = ops.reparameterize(real_mu, real_logvar)
        rec = model.decoder(z)

        loss_rec = self.compute_rec_loss(real, rec, reduction="mean")
        lossE_real_kl = self.compute_kl_loss(z, real_mu, real_logvar, write=True)

        rec_mu, rec_logvar, z_rec, rec_rec = model(rec.detach())
        fake_mu, fake_logvar, z_fake, rec_fake = model(fake.detach())
        kl_rec = self.compute_kl_loss(z_rec, rec_mu, rec_logvar, reduce="none", beta=self.beta_neg)
        kl_fake = self.compute_kl_loss(z_fake, fake_mu, fake_logvar, reduce="none", beta=self.beta_neg)
        rec_rec_e = self._per_sample(self.compute_rec_loss(rec, rec_rec, reduction="none"))
        rec_fake_e = self._per_sample(self.compute_rec_loss(fake, rec_fake, reduction="none"))

        expelbo_rec = self.exp_elbo(rec_rec_e, kl_rec)
        expelbo_fake = self.exp_elbo(rec_fake_e, kl_fake)
        lossE = self.scale * (loss_rec + lossE_real_kl) + 0.25 * (expelbo_rec + expelbo_fake)

        self.optimizer_e.zero_grad()
        lossE.backward()
        self.sync_gradients(model.encoder.parameters())
        total_norm_E = total_norm_D = None
        if self.clip:
            total_norm_E = torch.nn.utils.clip_grad_norm_(model.parameters(), self.clip).item()
        self.optimizer_e.step()

        # ------------------------------ decoder update (solvers/intro.py:118-160)
        self._set_trainable(encoder=False, decoder=True)
        fake = model.sample(noise)
        rec = model.decoder(z.detach())
        loss_rec = self.compute_rec_loss(real, rec, reduction="mean", write=True)

        rec_mu, rec_logvar = model.encode(rec)
        z_rec = ops.reparameterize(rec_mu, rec_logvar)
        fake_mu, fake_logvar = model.encode(fake)
        z_fake = ops.reparameterize(fake_mu, fake_logvar)
        rec_rec = model.decode(z_rec.detach())
        rec_fake = model.decode(z_fake.detach())

        gamma = self.gamma_r * self.beta_rec
        loss_rec_rec = self.compute_rec_loss(rec.detach(), rec_rec, reduction="mean", beta=gamma)
        loss_fake_rec = self.compute_rec_loss(fake.detach(), rec_fake, reduction="mean", beta=gamma)
        lossD_rec_kl = self.compute_kl_loss(z_rec, rec_mu, rec_logvar)
        lossD_fake_kl = self.compute_kl_loss(z_fake, fake_mu, fake_logvar)
        lossD = self.scale * (loss_rec + 0.5 * (lossD_rec_kl + lossD_fake_kl) + 0.5 * (loss_rec_rec + loss_fake_rec))

        self.optimizer_d.zero_grad()
        lossD.backward()
        self.sync_gradients(model.decoder.parameters())
        if self.clip:
            total_norm_D = torch.nn.utils.clip_grad_norm_(model.parameters(), self.clip).item()
        self.optimizer_d.step()

        if torch.isnan(lossD) or torch.isnan(lossE):
            raise RuntimeError("NaN loss")

        if self.writer:
            self.write_scalars(cur_iter,
                               losses=dict(r_loss=loss_rec.item(), kl_loss=lossE_real_kl.item(),
                                           expelbo_f=expelbo_fake.item()),
                               diff_kl=(lossD_fake_kl - lossE_real_kl).item())
            if self.clip:
                self.writer.add_scalars("total_norm", {"E": total_norm_E, "D": total_norm_D}, global_step=cur_iter)
            self.writer.add_scalar("lossE", lossE, global_step=cur_iter)
            self.writer.add_scalar("lossD", lossD, global_step=cur_iter)
            self.writer.flush()

        norms = [n for n in (total_norm_E, total_norm_D) if n is not None]
        return {"loss_enc": lossE.item(), "loss_dec": lossD.item(), "loss_kl": lossE_real_kl.item(),
                "loss_rec": loss_rec.item(), "L2": max(norms) if norms else None}

"""CPU oracle for the TC-ELBO hot path of meffmadd/intro-tc-vae.

TEST INFRASTRUCTURE ONLY.  Nothing under ``intro_tc_vae_b200/`` may import this
module; it is used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker and
the CPU timing baseline, never as the product path.

What it is: a restatement, in plain torch tensor ops that run on the host CPU,
of the reference's loss helpers.  Every function names the reference lines it
follows (paths relative to /root/reference).  The restatement materialises the
full [B, B, D] log-density tensor exactly as the reference does (same aten op
sequence, so its CPU cost is representative of the reference's), and gradients
come from torch autograd exactly as in the reference.

Parity pinning: the reference's own tests never exercise this path
(tests/test_ops.py:10-66 covers reconstruction loss, reparameterize shape and KL
shape only), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF:
``tests/golden/make_golden.py`` imports /root/reference/ops.py and
/root/reference/solvers/tc.py in the build container, evaluates them on
RNG-free inputs and commits the results under ``tests/golden/``;
``tests/test_oracle.py`` checks this module against those fixtures (fp32 and
fp64) and against the known-answer table of SURVEY.md section 8(c).

Third-party arithmetic on the path: ``torch.nn.functional.gaussian_nll_loss``
(reference pins torch==1.9.1 in requirements.txt:1; this image has 2.11.0, same
formula).  Its published algorithm is restated in ``gaussian_log_density_torch``
below: clone var, clamp_(min=eps) under no_grad (a straight-through floor),
0.5*(log var + (input-target)^2/var) + 0.5*log(2*pi).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
from torch import Tensor

LOG_2PI = math.log(2.0 * math.pi)
VAR_FLOOR = 1e-4          # eps passed at ops.py:18
LOGP_FLOOR = -50.0        # clamp at ops.py:21 and ops.py:29


class _StraightThroughFloor(torch.autograd.Function):
    """value = max(var, eps); d value / d var = 1.

    Mirrors gaussian_nll_loss's ``var = var.clone(); with no_grad: var.clamp_(min=eps)``.
    """

    @staticmethod
    def forward(ctx, var: Tensor, eps: float) -> Tensor:
        return var.clamp(min=eps)

    @staticmethod
    def backward(ctx, g: Tensor):
        return g, None


def gaussian_log_density_torch(x: Tensor, mu: Tensor, logvar: Tensor) -> Tensor:
    """Active density variant, ops.py:15-21 (via F.gaussian_nll_loss, eps=1e-4, full=True)."""
    var = torch.exp(logvar)
    if torch.any(var < 0):                      # gaussian_nll_loss raises here (host sync in the reference too)
        raise ValueError("var has negative entry/entries")
    vc = _StraightThroughFloor.apply(var, VAR_FLOOR)
    nll = 0.5 * (torch.log(vc) + (x - mu) ** 2 / vc)
    nll = nll + 0.5 * LOG_2PI
    return torch.clamp(-nll, min=LOGP_FLOOR)


def gaussian_log_density(x: Tensor, mu: Tensor, logvar: Tensor) -> Tensor:
    """'full'-path density variant, ops.py:24-29 (no variance floor; log 2*pi is an fp32 constant there)."""
    norm = torch.log(torch.tensor([2.0 * math.pi], dtype=torch.float32)).to(x.device)
    d = x - mu
    lp = -0.5 * (d * d * torch.exp(-logvar) + logvar + norm)
    return torch.clamp(lp, min=LOGP_FLOOR)


def log_importance_weight_matrix(batch_size: int, dataset_size: int) -> Tensor:
    """ops.py:32-49.  fp32, host.  Flat stride-(M+1)=B writes hit columns 0 and 1, then W[M-1, 0]."""
    n = dataset_size
    m = batch_size - 1
    strat = (n - m) / (n * m)                   # ZeroDivisionError at B == 1, as in the reference
    w = torch.full((batch_size, batch_size), 1.0 / m, dtype=torch.float32)
    flat = w.view(-1)
    flat[0::m + 1] = 1.0 / n
    flat[1::m + 1] = strat
    w[m - 1, 0] = strat
    return w.log()


def minibatch_stratified_sampling(log_qz_prob: Tensor, batch_size: int, dataset_size: int) -> Tuple[Tensor, Tensor]:
    """ops.py:104-115.  Returns (sum_d LSE_j(logW + lp), LSE_j(logW + sum_d lp))."""
    logw = log_importance_weight_matrix(batch_size, dataset_size).to(log_qz_prob.device)
    prod_marginals = torch.logsumexp(logw.view(batch_size, batch_size, 1) + log_qz_prob, dim=1).sum(dim=1)
    joint = torch.logsumexp(logw + log_qz_prob.sum(dim=2), dim=1)
    return prod_marginals, joint


def minibatch_weighted_sampling(log_qz_prob: Tensor, batch_size: int, dataset_size: int) -> Tuple[Tensor, Tensor]:
    """ops.py:92-101."""
    c = math.log(batch_size * dataset_size)
    prod_marginals = (torch.logsumexp(log_qz_prob, dim=1) - c).sum(dim=1)
    joint = torch.logsumexp(log_qz_prob.sum(dim=2), dim=1) - c
    return prod_marginals, joint


def pairwise_log_density(z: Tensor, mu: Tensor, logvar: Tensor, var_of: str = "row") -> Tensor:
    """[B,B,D] tensor indexed [i (z), j (mu), d].

    var_of == "row": ops.py:80-82 (row's own logvar, floor-clamped density);
    var_of == "col": solvers/tc.py:114-116 (column's logvar, un-floored density).
    """
    if var_of == "row":
        return gaussian_log_density_torch(z.unsqueeze(1), mu.unsqueeze(0), logvar.unsqueeze(1))
    if var_of == "col":
        return gaussian_log_density(z.unsqueeze(1), mu.unsqueeze(0), logvar.unsqueeze(0))
    raise ValueError(var_of)


def tc_terms(z: Tensor, mu: Tensor, logvar: Tensor, dataset_size: int,
             estimator: str = "mss", var_of: str = "row") -> Tuple[Tensor, Tensor]:
    """(log_qz_prod [B], log_qz [B]) for either estimator / density variant."""
    b = z.size(0)
    lp = pairwise_log_density(z, mu, logvar, var_of)
    if estimator == "mss":
        return minibatch_stratified_sampling(lp, b, dataset_size)
    if estimator == "mws":
        return minibatch_weighted_sampling(lp, b, dataset_size)
    raise ValueError(estimator)


def total_correlation(z: Tensor, mu: Tensor, logvar: Tensor, dataset_size: int, reduce: str = "mean") -> Tensor:
    """ops.py:52-89."""
    prod_marginals, joint = tc_terms(z, mu, logvar, dataset_size, "mss", "row")
    tc = joint - prod_marginals
    return tc.mean() if reduce == "mean" else tc


def kl_no_reduce(logvar: Tensor, mu: Tensor) -> Tensor:
    """ops.py:161-163."""
    return -0.5 * (1 + logvar - logvar.exp() - mu.pow(2)).sum(1)


def kl_divergence(logvar: Tensor, mu: Tensor, reduce: str = "sum") -> Tensor:
    """ops.py:136-158.  Argument order is (logvar, mu)."""
    kl = kl_no_reduce(logvar, mu)
    if reduce == "sum":
        return kl.sum()
    if reduce == "mean":
        return kl.mean()
    return kl


def reparameterize(mu: Tensor, logvar: Tensor, eps: Optional[Tensor] = None) -> Tensor:
    """ops.py:166-185; ``eps`` may be supplied so that results are RNG-independent."""
    std = torch.exp(0.5 * logvar)
    if eps is None:
        eps = torch.randn_like(std)
    return mu + eps * std


def kl_loss_simple(z: Tensor, mu: Tensor, logvar: Tensor, dataset_size: int, beta: float,
                   reduce: str = "mean") -> Tensor:
    """TCSovler._compute_kl_loss_simple, solvers/tc.py:69-89: (beta-1)*tc + kl."""
    kl = kl_divergence(logvar, mu, reduce=reduce)
    tc = total_correlation(z, mu, logvar, dataset_size, reduce=reduce)
    return (beta - 1.0) * tc + kl


def kl_loss_full(z: Tensor, mu: Tensor, logvar: Tensor, dataset_size: int, beta: float,
                 reduce: str = "mean"):
    """TCSovler._compute_kl_loss_full, solvers/tc.py:91-144: returns (loss, mi, tc, dimkl)."""
    condx = gaussian_log_density(z, mu, logvar).sum(dim=1)
    zeros = torch.zeros_like(z)
    pz = gaussian_log_density(z, zeros, zeros).sum(dim=1)
    prod_marginals, joint = tc_terms(z, mu, logvar, dataset_size, "mss", "col")
    mi = condx - joint
    tc = joint - prod_marginals
    dimkl = prod_marginals - pz
    if reduce == "mean":
        mi, tc, dimkl = mi.mean(), tc.mean(), dimkl.mean()
    return mi + beta * tc + dimkl, mi, tc, dimkl


def exp_elbo(rec_per_sample: Tensor, kl_per_sample: Tensor, scale: float) -> Tensor:
    """Soft-intro term, solvers/intro.py:102-103: mean_i exp(-2*scale*(rec_i + kl_i))."""
    return (-2 * scale * (rec_per_sample + kl_per_sample)).exp().mean()


# ---------------------------------------------------------------------------------------------
# Row-chunked evaluation: rows i are independent given all columns j and the global weight
# matrix, so large batches can be checked without a [B,B,D] tensor.
# ---------------------------------------------------------------------------------------------

def _logw_rows(rows: Tensor, batch_size: int, dataset_size: int, estimator: str, dtype) -> Tensor:
    """Rows ``rows`` of the log weight matrix ([len(rows), B]) without building all of it."""
    n, m = dataset_size, batch_size - 1
    strat = (n - m) / (n * m)
    w = torch.full((rows.numel(), batch_size), 1.0 / m, dtype=torch.float32)
    w[:, 0] = 1.0 / n
    w[:, 1] = strat
    w[rows == m - 1, 0] = strat                 # also covers B == 2, where M-1 == 0 (W = [[s, s], [1/N, s]])
    return w.log().to(dtype)


def tc_terms_rows(z_rows: Tensor, logvar_rows: Tensor, mu_all: Tensor, row_offset: int, batch_size: int,
                  dataset_size: int, estimator: str = "mss", var_of: str = "row",
                  logvar_all: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """(log_qz_prod, log_qz) for global rows [row_offset, row_offset+len(z_rows)) of a batch of
    ``batch_size`` columns.  Same arithmetic as ``tc_terms`` restricted to those rows."""
    r = z_rows.size(0)
    rows = torch.arange(row_offset, row_offset + r)
    if var_of == "row":
        lp = gaussian_log_density_torch(z_rows.unsqueeze(1), mu_all.unsqueeze(0), logvar_rows.unsqueeze(1))
    else:
        lp = gaussian_log_density(z_rows.unsqueeze(1), mu_all.unsqueeze(0), logvar_all.unsqueeze(0))
    if estimator == "mws":
        c = math.log(batch_size * dataset_size)
        return (torch.logsumexp(lp, dim=1) - c).sum(dim=1), torch.logsumexp(lp.sum(dim=2), dim=1) - c
    logw = _logw_rows(rows, batch_size, dataset_size, estimator, lp.dtype)
    prod_marginals = torch.logsumexp(logw.unsqueeze(2) + lp, dim=1).sum(dim=1)
    joint = torch.logsumexp(logw + lp.sum(dim=2), dim=1)
    return prod_marginals, joint

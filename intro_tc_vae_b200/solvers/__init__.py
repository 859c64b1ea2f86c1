"""Solver-side boundary of the TC-ELBO path.

This package ships only :class:`TCLossMixin` (the loss methods of reference ``solvers/tc.py:58-144`` on top of the fused
kernels).  The solver classes themselves -- ``VAESolver``, ``IntroSolver``, ``TCSovler``, ``IntroTCSovler`` -- are the
REFERENCE's own, unmodified: asking this module for one of them runs :func:`intro_tc_vae_b200.install` on the importable
reference checkout (``ops``, ``solvers`` on ``sys.path``) and returns the reference's class, whose ``train_step``
(solvers/vae.py:89-136, solvers/intro.py:56-196) then drives the kernels through ``compute_kl_loss`` /
``compute_rec_loss`` / ``reparameterize``.
"""
from .tc import TCLossMixin

__all__ = ["TCLossMixin", "VAESolver", "IntroSolver", "TCSovler", "IntroTCSovler"]

_REFERENCE_CLASSES = {"VAESolver": "solvers.vae", "IntroSolver": "solvers.intro", "TCSovler": "solvers.tc",
                      "IntroTCSovler": "solvers.intro_tc"}


def __getattr__(name):
    if name in _REFERENCE_CLASSES:
        import importlib
        from .. import install
        try:
            install()
            return getattr(importlib.import_module(_REFERENCE_CLASSES[name]), name)
        except ImportError as exc:
            raise ImportError(f"intro_tc_vae_b200.solvers.{name} is the reference's own class: put the intro-tc-vae checkout "
                              f"on sys.path (its ops.py / solvers/ must be importable) -- {exc}") from exc
    raise AttributeError(name)

"""B200-native total-correlation ELBO path for Intro-TC-VAE (drop-in for the reference's
``ops.total_correlation`` / ``TCSovler.compute_kl_loss`` / ``IntroTCSovler.compute_kl_loss``).

    from intro_tc_vae_b200 import ops              # same function names/signatures as the reference's ops.py
    from intro_tc_vae_b200.solvers import TCSovler, IntroTCSovler
    intro_tc_vae_b200.install()                    # or: patch an importable reference checkout in place

The compute lives in ``libtcelbo.so`` (CUDA, sm_100a, C ABI in include/tcelbo.h); importing this
package never falls back to another implementation when the library is missing.
"""
from __future__ import annotations

from . import _lib
from ._lib import LIB_PATH, TcelboError

__version__ = "0.1.0"


def library_available() -> bool:
    import os
    return os.path.exists(LIB_PATH)


def install(reference_modules: bool = True) -> None:
    """Route an importable reference checkout (``ops``, ``solvers.tc`` on sys.path) through this package.

    ``solvers/tc.py:5-11`` binds ``total_correlation`` and ``kl_divergence`` by name at import time, so
    the names are replaced in that module's namespace as well as in ``ops``.
    """
    from . import ops as fast_ops
    import importlib

    ref_ops = importlib.import_module("ops")
    ref_tc = importlib.import_module("solvers.tc")
    for name in ("total_correlation", "kl_divergence"):
        setattr(ref_ops, name, getattr(fast_ops, name))
        setattr(ref_tc, name, getattr(fast_ops, name))
    from .solvers.tc import TCLossMixin
    ref_tc.TCSovler._compute_kl_loss_full = TCLossMixin._compute_kl_loss_full

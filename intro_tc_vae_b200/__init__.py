"""B200-native total-correlation ELBO path for Intro-TC-VAE (drop-in for the reference's
``ops.total_correlation`` / ``TCSovler.compute_kl_loss`` / ``IntroTCSovler.compute_kl_loss``).

    import intro_tc_vae_b200
    intro_tc_vae_b200.install()                    # patch an importable reference checkout in place; train.py / main.py run as is
    from intro_tc_vae_b200 import ops              # or call the same function names/signatures as the reference's ops.py directly

The compute lives in ``libtcelbo.so`` (CUDA, sm_100a, C ABI in include/tcelbo.h); importing this
package never falls back to another implementation when the library is missing.
"""
from __future__ import annotations

from . import _lib
from ._lib import LIB_PATH, TcelboError

__version__ = "0.2.0"

# names of the reference's ops.py that this package implements on the GPU; install() rebinds every one of them in every
# reference module that imported it by name
_OPS_NAMES = ("total_correlation", "kl_divergence", "kl_no_reduce", "reparameterize", "reconstruction_loss",
              "gaussian_log_density", "gaussian_log_density_torch", "minibatch_stratified_sampling",
              "minibatch_weighted_sampling")
_REFERENCE_MODULES = ("ops", "models", "solvers.vae", "solvers.intro", "solvers.tc", "solvers.intro_tc")


def library_available() -> bool:
    import os
    return os.path.exists(LIB_PATH)


def install() -> dict:
    """Route an importable reference checkout (``ops``, ``models``, ``solvers.*`` on sys.path) through this package.

    * The reference binds loss helpers by name at import time (``from ops import reparameterize`` in models.py:5 and
      solvers/intro.py:14, ``kl_divergence`` / ``reconstruction_loss`` in solvers/vae.py:22, the TC helpers in
      solvers/tc.py:5-11), so each name is replaced in ``ops`` AND in every module that imported it.
    * ``TCSovler._compute_kl_loss_simple`` / ``_compute_kl_loss_full`` (solvers/tc.py:69-144) become the fused-kernel
      versions of :class:`intro_tc_vae_b200.solvers.TCLossMixin`; ``TCSovler.compute_kl_loss`` and the
      ``IntroTCSovler`` forwarder (solvers/intro_tc.py:8-17) are left alone -- they resolve these at call time.
    * This package's ``utils.SingletonWriter`` becomes an alias of the reference's, so the iteration counter and
      TensorBoard writer that train.py:100-103,212 set are the ones the patched loss methods log to.

    Returns ``{module name: [patched names]}``.  After install() the reference's loss path accepts fp32 CUDA tensors only
    (no CPU fallback: ``--device -1`` runs must not install).
    """
    import importlib

    from . import losses as fast_losses
    from . import ops as fast_ops
    from . import utils as fast_utils
    from .solvers.tc import TCLossMixin

    fast = {name: getattr(fast_ops, name) for name in _OPS_NAMES if hasattr(fast_ops, name)}
    fast["reconstruction_loss"] = fast_losses.reconstruction_loss
    patched = {}
    for mod_name in _REFERENCE_MODULES:
        mod = importlib.import_module(mod_name)
        done = []
        for name, fn in fast.items():
            if hasattr(mod, name):
                setattr(mod, name, fn)
                done.append(name)
        patched[mod_name] = done
    ref_tc = importlib.import_module("solvers.tc")
    ref_tc.TCSovler._compute_kl_loss_simple = TCLossMixin._compute_kl_loss_simple
    ref_tc.TCSovler._compute_kl_loss_full = TCLossMixin._compute_kl_loss_full
    patched["solvers.tc"] += ["TCSovler._compute_kl_loss_simple", "TCSovler._compute_kl_loss_full"]
    ref_utils = importlib.import_module("utils")
    fast_utils.SingletonWriter = ref_utils.SingletonWriter
    patched["intro_tc_vae_b200.utils"] = ["SingletonWriter"]
    return patched

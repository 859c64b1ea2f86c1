"""install(): an importable reference checkout is routed through this package (CPU-side check of the patching;
the reference is only present in the build container, so the test skips elsewhere)."""
import os
import sys
import types

import pytest

REF = "/root/reference"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "ops.py")), reason="reference checkout not present")
def test_install_patches_reference_modules():
    sys.dont_write_bytecode = True
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    try:
        for name, attrs in (("black", {"out": None}), ("matplotlib", {"use": lambda *a, **k: None}),
                            ("matplotlib.pyplot", {}), ("matplotlib.lines", {"Line2D": object}), ("xgboost", {"XGBClassifier": object})):
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.modules["matplotlib"].lines = sys.modules["matplotlib.lines"]
        sys.path.insert(0, REF)
        import intro_tc_vae_b200
        from intro_tc_vae_b200 import ops as fast_ops
        intro_tc_vae_b200.install()
        import ops as ref_ops
        import solvers.tc as ref_tc
        from solvers.intro_tc import IntroTCSovler
        assert ref_ops.total_correlation is fast_ops.total_correlation
        assert ref_tc.total_correlation is fast_ops.total_correlation      # solvers/tc.py:5-11 binds by name
        assert ref_tc.kl_divergence is fast_ops.kl_divergence
        assert ref_tc.TCSovler._compute_kl_loss_full.__module__ == "intro_tc_vae_b200.solvers.tc"
        assert IntroTCSovler.compute_kl_loss.__module__ == "solvers.intro_tc"   # forwarder untouched, resolves TCSovler at call time
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k not in saved_mods and (k in ("ops", "utils", "models", "dataset", "config", "train") or k.startswith(("solvers", "evaluation", "black", "matplotlib", "xgboost"))):
                del sys.modules[k]

"""The UNMODIFIED reference as a checker: stage it under ``oracle/_ref/`` and import it from there.

TEST INFRASTRUCTURE ONLY (same rule as ``oracle/tc_oracle.py``): nothing under ``intro_tc_vae_b200/`` imports
this module.  Users: ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py``.

The reference is a flat directory of Python modules (no package, no build), so "building" it is staging: when
``/root/reference`` is present (the build container), :func:`stage` copies the modules the TC-ELBO path touches --
``ops.py``, ``utils.py``, ``models.py``, ``dataset.py``, ``solvers/*.py``, ``evaluation/*.py`` -- into the git-ignored
``oracle/_ref/`` (never committed; it travels to the GPU box with the repo snapshot like the built ``.so``) and writes
three stub packages for third-party modules the reference imports but the path never calls and this image lacks
(``black``: models.py:2, ``matplotlib``: solvers/vae.py:4,19, ``xgboost``: evaluation/utils.py:7).  On the GPU box
``/root/reference`` does not exist and the staged copy is what runs.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DIR = os.path.join(HERE, "_ref")
STUB_DIR = os.path.join(REF_DIR, "_stubs")

_FILES = ["ops.py", "utils.py", "models.py", "dataset.py", "config.py",
          "solvers/__init__.py", "solvers/vae.py", "solvers/intro.py", "solvers/tc.py", "solvers/intro_tc.py",
          "evaluation/__init__.py", "evaluation/generator.py", "evaluation/metrics.py", "evaluation/utils.py"]
_REF_MODULES = ("ops", "utils", "models", "dataset", "config", "solvers", "evaluation")
_STUBS = {
    "black/__init__.py": "out = None\n",
    "matplotlib/__init__.py": "def use(*a, **k):\n    pass\n",
    "matplotlib/pyplot.py": "",
    "matplotlib/lines.py": "class Line2D:\n    pass\n",
    "xgboost/__init__.py": "class XGBClassifier:\n    pass\n",
}


def stage(force: bool = False) -> bool:
    """Copy the reference modules into oracle/_ref/ (build container only).  Returns True when a staged copy exists."""
    if os.path.isdir(REF_SRC):
        for rel in _FILES:
            src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DIR, rel)
            if not os.path.exists(src):
                continue
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if force or not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst):
                shutil.copyfile(src, dst)
                os.chmod(dst, 0o644)
        for rel, text in _STUBS.items():
            dst = os.path.join(STUB_DIR, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            with open(dst, "w") as f:
                f.write("# stub written by oracle/ref_loader.py: absent third-party module the TC-ELBO path never calls\n" + text)
    return available()


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "ops.py")) and os.path.exists(os.path.join(REF_DIR, "solvers", "tc.py"))


def _missing(name: str) -> bool:
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


@contextlib.contextmanager
def on_path():
    """Put the staged reference (and the stubs for modules this image lacks) first on sys.path; on exit restore sys.path and
    drop the reference's top-level modules from sys.modules so that two tests never share a patched copy."""
    if not available():
        raise ImportError("oracle/_ref is not staged: run `python -c 'import __graft_entry__ as g; g.build()'` where /root/reference exists")
    sys.dont_write_bytecode = True
    saved_path = list(sys.path)
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _REF_MODULES}
    for k in saved:
        del sys.modules[k]
    stubbed = [m for m in ("black", "matplotlib", "xgboost") if _missing(m)]
    try:
        sys.path.insert(0, REF_DIR)
        if stubbed:
            sys.path.append(STUB_DIR)              # last: a real installation of any of the three wins
        yield REF_DIR
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            top = k.split(".")[0]
            if top in _REF_MODULES or (top in stubbed and STUB_DIR in (getattr(sys.modules[k], "__file__", "") or "")):
                del sys.modules[k]
        sys.modules.update(saved)


def load_ops():
    """The reference's ``ops`` module (ops.py), imported from the staged copy; the module object stays usable after return."""
    with on_path():
        return importlib.import_module("ops")

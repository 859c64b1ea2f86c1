"""Soft-Intro solver with the TC loss (reference ``solvers/intro_tc.py:7-17``)."""
from __future__ import annotations

from .intro import IntroSolver
from .tc import TCLossMixin


class IntroTCSovler(TCLossMixin, IntroSolver):
    """``IntroSolver.train_step`` with ``compute_kl_loss`` taken from the TC solver."""

"""RNG-free input families shared by the golden generator and the tests.

Inputs are closed-form functions of (i, d) evaluated in fp64, so a fixture only needs to store
the reference's OUTPUTS.  ``base``: moderate posteriors (about 5 % of the B*B*D log-densities hit
the -50 clamp); ``stress``: sharp posteriors (about 20 % of variances below the 1e-4 floor, about
60 % of log-densities clamped).
"""
from __future__ import annotations

import numpy as np

CASES = {
    # name: B, D, N (dataset size), beta, family
    "base_B8_D4":       dict(B=8,   D=4,   N=100,    beta=6.0,   family="base"),
    "base_B64_D128":    dict(B=64,  D=128, N=16704,  beta=6.0,   family="base"),
    "base_B256_D128":   dict(B=256, D=128, N=16704,  beta=6.0,   family="base"),
    "stress_B64_D128":  dict(B=64,  D=128, N=16704,  beta=512.0, family="stress"),
    "tiny_B3_D128":     dict(B=3,   D=128, N=3,      beta=0.5,   family="base"),     # BASELINE cfg 1 (B=3, N=3)
    "pair_B2_D16":      dict(B=2,   D=16,  N=50,     beta=2.0,   family="base"),     # weight-matrix aliasing case
    "ragged_B37_D20":   dict(B=37,  D=20,  N=1000,   beta=4.0,   family="stress"),   # nothing a multiple of anything
    "wide_B24_D512":    dict(B=24,  D=512, N=737280, beta=512.0, family="base"),     # cfg 4's z_dim, dSprites-sized N
    "mid_B130_D256":    dict(B=130, D=256, N=16704,  beta=0.5,   family="stress"),   # cfg 5's z_dim, B-2 in a later tile
    "nan_B8_D4":        dict(B=8,   D=4,   N=3,      beta=6.0,   family="base"),     # N < B-1 -> NaN everywhere
}


def make_inputs(case):
    """(mu, logvar, eps) as fp64 numpy arrays of shape [B, D]."""
    B, D, family = case["B"], case["D"], case["family"]
    i = np.arange(B, dtype=np.float64)[:, None]
    d = np.arange(D, dtype=np.float64)[None, :]
    eps = np.sin(1.3 * i + 0.7 * d + 0.5)
    if family == "base":
        mu = np.sin(0.37 * i + 0.11 * d)
        lv = -2.0 + np.cos(0.23 * i - 0.07 * d)
    elif family == "stress":
        mu = 2.0 * np.sin(0.37 * i + 0.11 * d)
        lv = -6.0 + 4.0 * np.cos(0.23 * i - 0.07 * d)
    else:
        raise ValueError(family)
    return mu, lv, eps

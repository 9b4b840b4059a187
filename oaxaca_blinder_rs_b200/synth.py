"""Deterministic synthetic "wage" data of the shapes BASELINE.json names.

Per row: g ~ Bernoulli(.5) -> "M" (group A) / "F" (group B = reference_group); latent f ~ N(0,1);
continuous x_j = 0.3 f + sqrt(.91) e_j + 0.2 [g = M]; categorical levels with P_F / P_M;
y = b0_g + sum_j b_j x_j + gamma_level + eps, b0_M = 2.9, b0_F = 2.7, b_j = 0.05 (1 + j mod 5)/5,
eps ~ N(0, .5^2); weights ~ U(.5, 3).  Well conditioned, no nulls, every level in both groups.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

SEED = 0x0B200


def make_wage(n: int, n_cont: int, cat_levels: Sequence[int] = (), weights: bool = False,
              seed: int = SEED, chunk: int = 1 << 20) -> dict:
    """Returns host columns ready for ob_design_pack: cont [n_cont][n] f64, cat codes int32, y, w, group u8."""
    rng = np.random.Generator(np.random.Philox(seed))
    group = (rng.random(n) < 0.5).astype(np.uint8)          # 1 = "F" = group B (reference); 0 = "M" = group A
    is_m = group == 0
    f = rng.standard_normal(n)
    y = np.where(is_m, 2.9, 2.7) + 0.5 * rng.standard_normal(n)
    cont = []
    for j in range(n_cont):
        x = 0.3 * f + np.sqrt(0.91) * rng.standard_normal(n) + 0.2 * is_m
        y += 0.05 * (1 + j % 5) / 5.0 * x
        cont.append(x)
    cats = []
    for q, m in enumerate(cat_levels):
        pf = np.linspace(m, 1, m); pf /= pf.sum()            # F favours low levels, e.g. (.4,.3,.2,.1)
        pm = pf.copy(); pm[0] -= 0.1 * pf[0] * 2.5; pm[-1] += 0.1 * pf[0] * 2.5   # e.g. (.3,.3,.2,.2)
        u = rng.random(n)
        code = np.where(is_m, np.searchsorted(np.cumsum(pm), u), np.searchsorted(np.cumsum(pf), u)).astype(np.int32)
        code = np.minimum(code, m - 1)
        gamma = np.concatenate([[0.0], 0.05 * 2 ** np.arange(m - 1)])      # (0, .05, .10, .20)
        y += gamma[code]
        cats.append(code)
    w = rng.uniform(0.5, 3.0, n) if weights else None
    return dict(n=n, cont=cont, cat_codes=cats, cat_levels=list(cat_levels), outcome=y, weights=w, group=group)


def dense_design(d: dict):
    """Host mirror of the pack (prepare_data, builder.rs:294-378) for tests: (Xa, ya, wa, Xb, yb, wb)."""
    n = d["n"]
    cols = [np.ones(n)] + list(d["cont"])
    for code, m in zip(d["cat_codes"], d["cat_levels"]):
        for lv in range(1, m):
            cols.append((code == lv).astype(np.float64))
    X = np.stack(cols, 1)
    A, B = d["group"] == 0, d["group"] == 1
    w = d["weights"]
    return (X[A], d["outcome"][A], None if w is None else w[A], X[B], d["outcome"][B], None if w is None else w[B])


def norm_spec(d: dict, normalize: Optional[Sequence[int]] = None):
    """(m, dummy column indices) for the categorical variables listed in `normalize` (indices into cat_levels)."""
    out = []
    col = 1 + len(d["cont"])
    for q, m in enumerate(d["cat_levels"]):
        if normalize is None or q in normalize:
            out.append((m, list(range(col, col + m - 1))))
        col += m - 1
    return out

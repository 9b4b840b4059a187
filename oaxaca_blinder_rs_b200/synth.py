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


def make_wage_rows(n: int, n_cont: int, cat_levels: Sequence[int] = (), weights: bool = False, seed: int = SEED,
                   rank: int = 0, world: int = 1, plan=None, chunk: int = 1 << 20) -> dict:
    """Same population as make_wage, generated chunk by chunk (every chunk from its own Philox key), keeping only
    the rows row sharding (SURVEY.md 8e mode N) assigns to `rank`: plan(n_group, world, rank) -> [begin, end) positions
    within the group.  A rank never materialises more than one chunk of foreign rows, so n = 1e8 frames can be
    produced per GPU process.  world = 1 returns the whole frame (tests compare shards against it)."""
    if plan is None:
        from .core import row_shard_plan as plan
    nchunks = (n + chunk - 1) // chunk

    def chunk_rng(c):
        return np.random.Generator(np.random.Philox(key=[seed, c]))

    # pass 1: group membership only (first draw of every chunk) -> group sizes and this rank's ranges
    counts = np.zeros((nchunks, 2), dtype=np.int64)
    for c in range(nchunks):
        m = min(chunk, n - c * chunk)
        g = chunk_rng(c).random(m) < 0.5
        counts[c] = (m - int(g.sum()), int(g.sum()))
    na, nb = int(counts[:, 0].sum()), int(counts[:, 1].sum())
    a0, a1 = plan(na, world, rank)
    b0, b1 = plan(nb, world, rank)
    before = np.vstack([np.zeros((1, 2), dtype=np.int64), np.cumsum(counts, axis=0)])
    parts = []
    for c in range(nchunks):
        ca0, cb0 = before[c]
        ca1, cb1 = before[c + 1]
        if (ca1 <= a0 or ca0 >= a1) and (cb1 <= b0 or cb0 >= b1):
            continue                                         # no row of this chunk lives on this rank
        m = min(chunk, n - c * chunk)
        d = _wage_chunk(chunk_rng(c), m, n_cont, cat_levels, weights)
        pos_a = ca0 + np.cumsum(d["group"] == 0) - 1
        pos_b = cb0 + np.cumsum(d["group"] == 1) - 1
        keep = ((d["group"] == 0) & (pos_a >= a0) & (pos_a < a1)) | ((d["group"] == 1) & (pos_b >= b0) & (pos_b < b1))
        parts.append({k: ([x[keep] for x in v] if isinstance(v, list) else (None if v is None else v[keep]))
                      for k, v in d.items()})
    def cat(key, dtype):
        xs = [p[key] for p in parts]
        return np.concatenate(xs) if xs else np.empty(0, dtype=dtype)
    out = dict(cont=[np.concatenate([p["cont"][j] for p in parts]) if parts else np.empty(0) for j in range(n_cont)],
               cat_codes=[np.concatenate([p["cat_codes"][q] for p in parts]) if parts else np.empty(0, dtype=np.int32)
                          for q in range(len(cat_levels))],
               cat_levels=list(cat_levels), outcome=cat("outcome", np.float64),
               weights=cat("weights", np.float64) if weights else None, group=cat("group", np.uint8),
               n_a_global=na, n_b_global=nb, n_global=n)
    out["n"] = int(out["group"].shape[0])
    return out


def _wage_chunk(rng, n, n_cont, cat_levels, weights):
    group = (rng.random(n) < 0.5).astype(np.uint8)
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
        pf = np.linspace(m, 1, m); pf /= pf.sum()
        pm = pf.copy(); pm[0] -= 0.1 * pf[0] * 2.5; pm[-1] += 0.1 * pf[0] * 2.5
        u = rng.random(n)
        code = np.where(is_m, np.searchsorted(np.cumsum(pm), u), np.searchsorted(np.cumsum(pf), u)).astype(np.int32)
        code = np.minimum(code, m - 1)
        gamma = np.concatenate([[0.0], 0.05 * 2 ** np.arange(m - 1)])
        y += gamma[code]
        cats.append(code)
    w = rng.uniform(0.5, 3.0, n) if weights else None
    return dict(cont=cont, cat_codes=cats, outcome=y, weights=w, group=group)


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

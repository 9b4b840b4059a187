#!/usr/bin/env python
"""Golden vectors for the Machado-Mata path, independent of oracle/.

Every quantile regression is solved as the reference's LP (math/quantile_regression.rs:34-112:
min tau 1'u + (1 - tau) 1'v  s.t.  X beta + u - v = y, u, v >= 0) by scipy.optimize.linprog's HiGHS dual simplex, which
returns the LP's optimal vertex; run_single_pass (quantile_decomposition.rs:173-279) and the bootstrap loop (:337-354)
are restated in numpy on top of those coefficients.  The reference draws its random quantiles, simulated rows and
resamples from unseeded thread_rng streams, so the streams are part of the fixture.

    python tests/golden/make_mm_golden.py      # writes tests/golden/mm_fixture.json
"""
import json
import os

import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog


def solve_qr(X, y, tau):
    n, K = X.shape
    cost = np.r_[np.zeros(K), tau * np.ones(n), (1.0 - tau) * np.ones(n)]
    A = sp.hstack([sp.csr_matrix(X), sp.eye(n), -sp.eye(n)]).tocsr()
    r = linprog(cost, A_eq=A, b_eq=y, bounds=[(None, None)] * K + [(0, None)] * (2 * n), method="highs-ds")
    assert r.status == 0
    return r.x[:K]


def empirical_quantile(v, q):                      # quantile_decomposition.rs:164-171
    v = np.sort(v)
    return 0.0 if len(v) == 0 else v[min(int(len(v) * q), len(v) - 1)]


def single_pass(Xa, ya, Xb, yb, taus, da, db, quantiles):
    ba = np.array([solve_qr(Xa, ya, t) for t in taus])
    bb = np.array([solve_qr(Xb, yb, t) for t in taus])
    yaa = np.einsum("ij,ij->i", Xa[da], ba)
    ybb = np.einsum("ij,ij->i", Xb[db], bb)
    yab = np.einsum("ij,ij->i", Xa[da], bb)
    out = []
    for q in quantiles:
        qaa, qbb, qab = (empirical_quantile(v, q) for v in (yaa, ybb, yab))
        out.append([qaa - qbb, qab - qbb, qaa - qab])
    return np.array(out), ba, bb


def main():
    rng = np.random.default_rng(20261018)
    na, nb, sims, reps = 70, 55, 12, 3
    quantiles = [0.1, 0.25, 0.5, 0.75, 0.9]

    def group(n, shift):
        sector = rng.integers(0, 3, size=n)
        X = np.c_[np.ones(n), rng.normal(12, 2, size=n), rng.uniform(0, 30, size=n), sector == 1, sector == 2].astype(float)
        y = 1.2 + shift + (0.07 + 0.02 * shift) * X[:, 1] + 0.015 * X[:, 2] + 0.1 * X[:, 3] - 0.15 * X[:, 4] \
            + rng.standard_t(4, size=n) * 0.3 * (1 + 0.03 * X[:, 1])
        return X, y

    Xa, ya = group(na, 0.3)
    Xb, yb = group(nb, 0.0)
    taus = rng.uniform(0.01, 0.99, size=(reps + 1, sims))
    draw_a = rng.integers(0, na, size=(reps + 1, sims))
    draw_b = rng.integers(0, nb, size=(reps + 1, sims))
    idx_a = rng.integers(0, na, size=(reps, na))
    idx_b = rng.integers(0, nb, size=(reps, nb))
    point, ba, bb = single_pass(Xa, ya, Xb, yb, taus[0], draw_a[0], draw_b[0], quantiles)
    rep = []
    for r in range(reps):
        st, _, _ = single_pass(Xa[idx_a[r]], ya[idx_a[r]], Xb[idx_b[r]], yb[idx_b[r]], taus[r + 1], draw_a[r + 1], draw_b[r + 1], quantiles)
        rep.append(st)
    # the reference's own known-answer data (quantile_regression.rs:137-170): perfectly linear, every quantile = [0, 1]
    kat = dict(X=[[1, 1], [1, 2], [1, 3], [1, 4], [1, 5]], y=[1, 2, 3, 4, 5], taus=[0.5, 0.25], beta=[0.0, 1.0], tol=1e-4)
    fx = dict(quantiles=quantiles, sims=sims, reps=reps,
              Xa=Xa.tolist(), ya=ya.tolist(), Xb=Xb.tolist(), yb=yb.tolist(),
              taus=taus.tolist(), draw_a=draw_a.tolist(), draw_b=draw_b.tolist(), idx_a=idx_a.tolist(), idx_b=idx_b.tolist(),
              point_stats=point.tolist(), point_betas_a=ba.tolist(), point_betas_b=bb.tolist(),
              rep_stats=[s.tolist() for s in rep], reference_kat=kat,
              solver="scipy.optimize.linprog(method='highs-ds')")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mm_fixture.json")
    with open(path, "w") as f:
        json.dump(fx, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Golden vectors for the Heckman two-step pass, computed with numpy / scipy ONLY (independent of oracle/): restates
math/probit.rs:25-175, heckman.rs:38-108, estimation.rs:114-269 and builder.rs:464-534, :538-699 of the reference on a
small synthetic frame (the reference's own tests/heckman_test.rs draws its data from an unseeded-in-spirit StdRng and
asserts only that a row named "IMR" exists, so it pins no number).  Writes tests/golden/heckman_fixture.json.

    python tests/golden/make_heckman_golden.py
"""
import json
import os

import numpy as np
from scipy.stats import norm

HERE = os.path.dirname(os.path.abspath(__file__))


def probit(y, X, max_iter=100, tol=1e-6):
    b = np.zeros(X.shape[1])
    for _ in range(max_iter):
        z = X @ b
        phi = norm.pdf(z)
        Phi = np.clip(norm.cdf(z), 1e-10, 1 - 1e-10)
        lam = np.where(y > 0.5, phi / Phi, -phi / (1 - Phi))
        w = phi * phi / (Phi * (1 - Phi))
        H = -(X.T * w) @ X - 1e-9 * np.eye(X.shape[1])
        step = np.linalg.solve(-H, X.T @ lam)
        b = b + step
        if np.linalg.norm(step) < tol:
            break
    return b


def group(X, y, Z, s):
    gamma = probit(s, Z)
    sel = s == 1.0
    zg = Z[sel] @ gamma
    Phi = norm.cdf(zg)
    imr = np.where(Phi < 1e-10, 0.0, norm.pdf(zg) / Phi)
    Xa = np.c_[X[sel], imr]
    beta = np.linalg.solve(Xa.T @ Xa, Xa.T @ y[sel])
    return dict(beta=beta, xmean=Xa.mean(axis=0), gamma=gamma, zmean=Z.mean(axis=0), delta=np.mean(-imr * (imr + zg)))


def heckman_pass(ref, Xa, ya, Za, sa, Xb, yb, Zb, sb):
    A, B = group(Xa, ya, Za, sa), group(Xb, yb, Zb, sb)
    R = A if ref == 0 else B
    sel = R["beta"][-1] * R["delta"] * R["gamma"] * (A["zmean"] - B["zmean"])
    if ref == 0:
        bs = A["beta"]
    elif ref == 1:
        bs = B["beta"]
    else:
        wA = len(ya) / (len(ya) + len(yb))
        bs = A["beta"] * wA + B["beta"] * (1 - wA)
    xa, xb, ba, bb = A["xmean"], B["xmean"], A["beta"], B["beta"]
    dx, db = xa - xb, ba - bb
    three = [dx @ bb, xb @ db, dx @ db]
    expl = dx @ bs
    two = [expl, (xa @ ba - xb @ bb) - expl]
    det_e = dx * bs
    det_u = xa * (ba - bs) + xb * (bs - bb)
    stats = np.concatenate([two, three, det_e, det_u, sel])
    return dict(stats=stats.tolist(), beta_a=ba.tolist(), beta_b=bb.tolist(), gamma_a=A["gamma"].tolist(), gamma_b=B["gamma"].tolist(),
                total_gap=float(ya.mean() - yb.mean()))


def main():
    rng = np.random.default_rng(20261018)
    n = 1200
    grp = rng.integers(0, 2, n)
    z = rng.normal(size=n)
    x = z + 0.5 * rng.normal(size=n) + 0.3 * (grp == 0)
    u = rng.normal(size=n)
    e = 0.8 * u + 0.6 * rng.normal(size=n)
    s = (0.4 + 0.5 * z + 0.2 * (grp == 0) + u > 0).astype(float)
    d = (rng.random(n) < 0.4).astype(float)                  # a dummy in the outcome design
    y = 1.0 + 2.0 * x + 0.5 * d + 0.4 * (grp == 0) + e        # observed for everyone (the reference drops null outcomes anyway)
    X = np.c_[np.ones(n), x, d]
    Z = np.c_[np.ones(n), z]
    A, B = grp == 0, grp == 1
    out = {"columns": {"group": grp.tolist(), "x": x.tolist(), "d": d.tolist(), "z": z.tolist(), "s": s.tolist(), "y": y.tolist()},
           "design": "X = [1, x, d], Z = [1, z], selection = s; group A = (group == 0), reference group B = (group == 1)",
           "expected": {name: heckman_pass(ref, X[A], y[A], Z[A], s[A], X[B], y[B], Z[B], s[B])
                        for name, ref in (("A", 0), ("B", 1), ("weighted", 3))}}
    with open(os.path.join(HERE, "heckman_fixture.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote heckman_fixture.json; explained/unexplained (B):", out["expected"]["B"]["stats"][:2])


if __name__ == "__main__":
    main()

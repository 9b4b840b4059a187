"""Generates tests/golden/reference_fixtures.json.

The reference cannot be built or imported here (Rust, no toolchain), so the golden vectors are the
reference's OWN test fixtures and asserted values (file:line cited per entry), extended with the
intermediate quantities obtained by applying the cited formulas with numpy.linalg (independent of
oracle/): least squares via lstsq, Yun shift per normalization.rs:26-47, decomposition per
decomposition.rs:56-122, base rows per builder.rs:634-674.

Run:  python tests/golden/make_golden.py   (deterministic; rewrites the JSON)
"""
import json
import os

import numpy as np


def lstsq(X, y, w=None):
    if w is not None:
        sw = np.sqrt(w)
        X, y = X * sw[:, None], y * sw
    return np.linalg.lstsq(X, y, rcond=None)[0]


def mean(X, w=None):
    return X.mean(0) if w is None else (X * w[:, None]).sum(0) / w.sum()


def decompose(Xa, ya, wa, Xb, yb, wb, ref, norm=(), n_cont=None):
    """norm: list of (m, idx list).  ref in A|B|pooled|weighted."""
    K = Xa.shape[1]
    ba, bb = lstsq(Xa, ya, wa), lstsq(Xb, yb, wb)
    raw_a, raw_b = ba.copy(), bb.copy()
    xa, xb = mean(Xa, wa), mean(Xb, wb)

    def yun(beta, shift=None):
        base = []
        for m, idx in norm:
            idx = [i + 1 if (shift is not None and i >= shift) else i for i in idx]
            mu = beta[idx].sum() / m
            beta[0] += mu
            beta[idx] -= mu
            base.append(-mu)
        return base

    base_a, base_b = yun(ba), yun(bb)
    if ref == "A":
        bs, base_s = ba.copy(), list(base_a)
    elif ref == "B":
        bs, base_s = bb.copy(), list(base_b)
    elif ref == "pooled":
        ind = 1 + n_cont
        Xp = np.vstack([Xa, Xb])
        g = np.r_[np.ones(len(Xa)), np.zeros(len(Xb))]
        Xp = np.c_[Xp[:, :ind], g, Xp[:, ind:]]
        wp = None if wa is None else np.r_[wa, wb]
        bp = lstsq(Xp, np.r_[ya, yb], wp)
        base_s = yun(bp, shift=ind)
        bs = np.delete(bp, ind)
    else:
        na = len(Xa) if wa is None else wa.sum()
        nb = len(Xb) if wb is None else wb.sum()
        wA = na / (na + nb)
        bs = ba * wA + bb * (1 - wA)
        base_s = [a * wA + b * (1 - wA) for a, b in zip(base_a, base_b)]
    dx, db = xa - xb, ba - bb
    three = [dx @ bb, xb @ db, dx @ db]
    expl = dx @ bs
    two = [expl, (xa @ ba - xb @ bb) - expl]
    de = list(dx * bs)
    du = list(xa * (ba - bs) + xb * (bs - bb))
    for v, (m, idx) in enumerate(norm):
        xab, xbb = 1 - xa[idx].sum(), 1 - xb[idx].sum()
        un = xab * (base_a[v] - base_s[v]) + xbb * (base_s[v] - base_b[v])
        ex = (xab - xbb) * base_s[v]
        du.append(un); de.append(ex)
        two[0] += ex; two[1] += un
    gap = (ya.mean() if wa is None else ya @ wa / wa.sum()) - (yb.mean() if wb is None else yb @ wb / wb.sum())
    return dict(raw_beta_a=list(raw_a), raw_beta_b=list(raw_b), beta_a=list(ba), beta_b=list(bb),
                base_a=base_a, base_b=base_b, base_star=base_s, xa_mean=list(xa), xb_mean=list(xb),
                beta_star=list(bs), two_fold=two, three_fold=three, det_expl=de, det_unexpl=du,
                total_gap=gap, resid_b=list(yb - Xb @ raw_b))


def main():
    out = {}
    # ---- F1: tests/integration_test.rs:4-10 (fixture), :105-144 (asserts gap==10, additivity, n=10/10)
    wage = np.array([10, 12, 11, 13, 15, 20, 22, 21, 23, 25] * 2, float)
    edu = np.array([12, 16, 14, 16, 18] * 4, float)
    gender = ["F"] * 5 + ["M"] * 5 + ["F"] * 5 + ["M"] * 5
    g = np.array(gender)
    A, B = g == "M", g == "F"   # reference_group = "F" is group B; A = first other sorted value
    Xa, Xb = np.c_[np.ones(A.sum()), edu[A]], np.c_[np.ones(B.sum()), edu[B]]
    f1 = dict(source="tests/integration_test.rs:4-10,105-144",
              columns=dict(wage=list(wage), education=list(edu), gender=gender),
              outcome="wage", group="gender", reference_group="F", predictors=["education"],
              asserted=dict(total_gap=10.0, n_a=10, n_b=10, tol=1e-9), expected={})
    for ref in ("A", "B", "pooled", "weighted"):
        f1["expected"][ref] = decompose(Xa, wage[A], None, Xb, wage[B], None, ref, n_cont=1)
    out["F1"] = f1

    # ---- F2: tests/integration_test.rs:146-163 (C(union), normalize, default beta* = GroupA builder.rs:123)
    union = ["none", "union", "union_plus", "none", "union", "union_plus", "none", "union", "union_plus", "none"] * 2
    u = np.array(union)
    d1, d2 = (u == "union").astype(float), (u == "union_plus").astype(float)
    Xa = np.c_[np.ones(A.sum()), edu[A], d1[A], d2[A]]
    Xb = np.c_[np.ones(B.sum()), edu[B], d1[B], d2[B]]
    f2 = dict(source="tests/integration_test.rs:146-163",
              columns=dict(wage=list(wage), education=list(edu), gender=gender, union=union),
              outcome="wage", group="gender", reference_group="F", predictors=["education"],
              categorical=["union"], normalize=["union"],
              names=["__ob_intercept__", "education", "union_union", "union_union_plus"],
              base_names=["union_none"],
              asserted=dict(total_gap=10.0, n_a=10, n_b=10, tol=1e-9), expected={})
    for ref in ("A", "B", "pooled", "weighted"):
        f2["expected"][ref] = decompose(Xa, wage[A], None, Xb, wage[B], None, ref, norm=[(3, [2, 3])], n_cont=1)
    out["F2"] = f2

    # ---- F3: tests/weights_test.rs:19-46 (unweighted gap 0.666 +- 0.01, weighted gap -3.333 +- 0.01)
    y = np.array([10.0, 10.0, 2.0, 5.0, 7.0, 8.0])
    grp = ["A", "A", "A", "B", "B", "B"]
    w = np.array([1.0, 1.0, 10.0, 1.0, 1.0, 1.0])
    x = np.array([1.0, 1.0, 0.0, 0.0, 1.0, 1.0])
    Xa, Xb = np.c_[np.ones(3), x[:3]], np.c_[np.ones(3), x[3:]]
    out["F3"] = dict(source="tests/weights_test.rs:19-46",
                     columns=dict(outcome=list(y), group=grp, weight=list(w), x=list(x)),
                     outcome="outcome", group="group", reference_group="B", predictors=["x"],
                     asserted=dict(unweighted_gap=0.666, weighted_gap=-3.333, tol=0.01),
                     expected=dict(unweighted=decompose(Xa, y[:3], None, Xb, y[3:], None, "A", n_cont=1),
                                   weighted=decompose(Xa, y[:3], w[:3], Xb, y[3:], w[3:], "A", n_cont=1)))

    # ---- F4: tests/optimize_budget_test.rs:4-34 (gap 16; point residuals of group B are -5/0/+5)
    wage4 = np.array([30.0, 32.0, 34.0, 10.0, 15.0, 20.0, 12.0, 17.0, 22.0])
    edu4 = np.array([10.0, 12.0, 14.0, 10.0, 10.0, 10.0, 12.0, 12.0, 12.0])
    grp4 = ["A"] * 3 + ["B"] * 6
    Xa, Xb = np.c_[np.ones(3), edu4[:3]], np.c_[np.ones(6), edu4[3:]]
    out["F4"] = dict(source="tests/optimize_budget_test.rs:4-34",
                     columns=dict(wage=list(wage4), education=list(edu4), group=grp4),
                     outcome="wage", group="group", reference_group="B", predictors=["education"],
                     asserted=dict(total_gap=16.0, residuals_b=[-5.0, 0.0, 5.0, -5.0, 0.0, 5.0], tol=1e-9),
                     expected=dict(A=decompose(Xa, wage4[:3], None, Xb, wage4[3:], None, "A", n_cont=1)))

    # ---- unit known answers
    out["KAT"] = dict(
        ols=dict(source="math/ols.rs:151-162", X=[[1, 0], [1, 1], [1, 2], [1, 3], [1, 4]], y=[1, 3, 5, 7, 9],
                 beta=[1.0, 2.0], tol=1e-9),
        ols_collinear=dict(source="math/ols.rs:164-181", X=[[1, 2, 4], [1, 3, 6], [1, 4, 8], [1, 5, 10]],
                           y=[1, 2, 3, 4], error="NalgebraError"),
        ols_insufficient=dict(source="math/ols.rs:183-209", n=2, K=5, error="InsufficientData"),
        yun=dict(source="math/normalization.rs:58-111", beta=[10.0, 2.0, 4.0], m=3, idx=[1, 2],
                 beta_out=[12.0, 0.0, 2.0], base=-2.0),
        three_fold=dict(source="decomposition.rs:129-139", xa=[1, 5], xb=[1, 3], ba=[2, 4], bb=[1, 3],
                        out=[6.0, 4.0, 2.0]),
        bootstrap_p=dict(source="inference.rs:41-57",
                         cases=[dict(est=[1, 2, 3, 4, 5], p=0.0), dict(est=[-2, -1, 0, 1, 2], p=1.0),
                                dict(est=[-1, 1, 2, 3, 4], p=0.4)], tol=1e-9),
        null_handling=dict(source="tests/null_handling_test.rs:4-64",
                           outcome=[10.0, 12.0, 11.0, None, 15.0, 16.0, 17.0, 18.0],
                           group=["A", "A", "A", "A", "B", "B", "B", "B"],
                           education=[10.0, 12.0, 11.0, 12.0, 14.0, 16.0, 15.0, None], n_a=3, n_b=3),
    )
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fixtures.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()

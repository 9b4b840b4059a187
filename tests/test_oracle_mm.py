"""Pins the Machado-Mata oracle (oracle/ob_oracle_mm.c) to vectors it did not produce: the LP vertices of an
independent solver (HiGHS dual simplex, tests/golden/make_mm_golden.py) and the reference's own known-answer tests
(math/quantile_regression.rs:137-170).  CPU only."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def mm_fixture():
    with open(os.path.join(HERE, "golden", "mm_fixture.json")) as f:
        return json.load(f)


def test_reference_known_answers(orc, mm_fixture):
    """test_solve_qr_median / test_solve_qr_quartile: perfectly linear data, beta = [0, 1] within 1e-4."""
    kat = mm_fixture["reference_kat"]
    X, y = np.array(kat["X"], dtype=float), np.array(kat["y"], dtype=float)
    for tau in kat["taus"]:
        beta, info = orc.qr(X, y, tau)
        assert info["status"] == orc.QR_VERTEX
        assert len(beta) == 2
        np.testing.assert_allclose(beta, kat["beta"], rtol=0, atol=kat["tol"])
        np.testing.assert_allclose(beta, kat["beta"], rtol=0, atol=1e-13)     # the vertex itself


def test_qr_vertices_match_highs(orc, mm_fixture):
    fx = mm_fixture
    Xa, ya, Xb, yb = (np.array(fx[k]) for k in ("Xa", "ya", "Xb", "yb"))
    taus = np.array(fx["taus"])[0]
    for X, y, key in ((Xa, ya, "point_betas_a"), (Xb, yb, "point_betas_b")):
        exp = np.array(fx[key])
        for s, tau in enumerate(taus):
            beta, info = orc.qr(X, y, tau)
            assert info["status"] == orc.QR_VERTEX, (key, s, info)
            assert info["ncand"] >= X.shape[1]
            np.testing.assert_allclose(beta, exp[s], rtol=1e-10, atol=1e-11, err_msg=f"{key}[{s}] tau={tau}")


def test_multiplicities_equal_gathered_rows(orc, mm_fixture):
    """The weighted regression the GPU path runs on the multiplicity matrix is the LP of the gathered frame."""
    fx = mm_fixture
    X, y = np.array(fx["Xa"]), np.array(fx["ya"])
    idx = np.array(fx["idx_a"])[0]
    c = np.bincount(idx, minlength=len(y)).astype(float)
    for tau in (0.07, 0.31, 0.5, 0.88):
        bg, ig = orc.qr(X[idx], y[idx], tau)
        bw, iw = orc.qr(X, y, tau, c=c)
        assert ig["status"] == iw["status"] == orc.QR_VERTEX
        np.testing.assert_allclose(bw, bg, rtol=1e-11, atol=1e-12)


def test_single_pass_and_bootstrap_match_numpy_restatement(orc, mm_fixture):
    fx = mm_fixture
    Xa, ya, Xb, yb = (np.array(fx[k]) for k in ("Xa", "ya", "Xb", "yb"))
    out = orc.mm_run(Xa, ya, Xb, yb, fx["sims"], fx["quantiles"], fx["reps"], fx["idx_a"], fx["idx_b"], fx["taus"],
                     fx["draw_a"], fx["draw_b"], nthreads=2)
    nq = len(fx["quantiles"])
    np.testing.assert_allclose(out["point_stats"].reshape(nq, 3), fx["point_stats"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(out["betas_a"], fx["point_betas_a"], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(out["rep_stats"].reshape(fx["reps"], nq, 3), fx["rep_stats"], rtol=0, atol=1e-10)
    assert out["n_ok"] == fx["reps"] and (out["rep_status"] == 0).all()
    # decomposition identities of every pass: gap = characteristics + coefficients (quantile_decomposition.rs:271-275)
    st = out["rep_stats"].reshape(fx["reps"], nq, 3)
    np.testing.assert_allclose(st[..., 0], st[..., 1] + st[..., 2], rtol=0, atol=1e-12)
    # the reduction: bootstrap_stats over the passes (inference.rs:4-34), t = point / se when |se| > 1e-9
    for j in range(3 * nq):
        col = out["rep_stats"][:, j]
        se = np.sqrt(np.sum((col - col.mean()) ** 2) / (len(col) - 1))
        assert abs(out["se"][j] - se) < 1e-12
        assert abs(out["t"][j] - (out["point_stats"][j] / se if abs(se) > 1e-9 else 0.0)) < 1e-9


def test_failed_regressions_shift_the_pairing(orc):
    """filter_map(.ok()) keeps the successful fits of each group in order and min(len) pairs them up
    (quantile_decomposition.rs:227-244); a group with fewer than simulations / 2 fits fails the pass (:238-242)."""
    rng = np.random.default_rng(3)
    n, K = 40, 2
    Xa = np.c_[np.ones(n), rng.normal(size=n)]
    ya = Xa @ [1.0, 0.5] + rng.normal(size=n)
    Xb = np.c_[np.ones(n), np.ones(n)]          # collinear: every regression of group B fails
    yb = rng.normal(size=n)
    out = orc.mm_pass(Xa, ya, Xb, yb, [0.2, 0.5, 0.8, 0.6], [0, 1, 2, 3], [0, 1, 2, 3], [0.5])
    assert (out["status_b"] == orc.QR_FAILED).all() and (out["status_a"] == orc.QR_VERTEX).all()
    assert out["rc"] == 4 and out["nsucc"] == 0

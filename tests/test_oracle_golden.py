"""Pins the oracle (oracle/ob_oracle.c) to the reference's own fixtures and known answers
(tests/golden/reference_fixtures.json).  CPU only."""
import numpy as np
import pytest

from helpers import fixture_design as _design

REFS = {"A": 0, "B": 1, "pooled": 2, "weighted": 3}


def _check_pass(out, exp, tol=1e-9):
    for k in ("beta_a", "beta_b", "xa_mean", "xb_mean", "beta_star", "two_fold", "three_fold",
              "det_expl", "det_unexpl"):
        np.testing.assert_allclose(out[k], exp[k], rtol=0, atol=tol, err_msg=k)
    assert abs(out["total_gap"] - exp["total_gap"]) < tol
    np.testing.assert_allclose(out["resid_b"], exp["resid_b"], rtol=0, atol=tol)


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("ref", ["A", "B", "pooled", "weighted"])
@pytest.mark.parametrize("fx", ["F1", "F2"])
def test_integration_fixtures(orc, golden, fx, ref, precise):
    fix = golden[fx]
    Xa, ya, wa, Xb, yb, wb, norm, n_cont = _design(fix)
    spec = orc.Spec(K=Xa.shape[1], n_cont=n_cont, ref_kind=REFS[ref],
                    norm=[orc.NormVar(m, idx, True) for m, idx in norm])
    out = orc.single_pass(spec, Xa, ya, wa, Xb, yb, wb, precise=precise)
    _check_pass(out, fix["expected"][ref])
    # the reference's own assertions: integration_test.rs:13-49
    a = fix["asserted"]
    assert abs(out["total_gap"] - a["total_gap"]) < a["tol"]
    assert abs(out["two_fold"].sum() - out["total_gap"]) < a["tol"]
    assert (len(ya), len(yb)) == (a["n_a"], a["n_b"])


def test_f2_three_fold_is_not_yun_corrected(orc, golden):
    """with Yun the three-fold sum is 10.28, not the gap 10 (builder.rs:623)."""
    fix = golden["F2"]
    Xa, ya, wa, Xb, yb, wb, norm, n_cont = _design(fix)
    spec = orc.Spec(K=4, n_cont=1, ref_kind=0, norm=[orc.NormVar(m, i) for m, i in norm])
    out = orc.single_pass(spec, Xa, ya, wa, Xb, yb, wb)
    assert abs(out["three_fold"].sum() - 10.28) < 1e-9
    np.testing.assert_allclose(out["beta_a"], [8.5, 0.9, -0.1, 0.4], atol=1e-9)
    np.testing.assert_allclose(out["beta_b"], [-1.5, 0.9, -0.3, -0.1], atol=1e-9)


def test_weights_fixture(orc, golden):
    fix = golden["F3"]
    a = fix["asserted"]
    for key, weighted in (("unweighted", False), ("weighted", True)):
        Xa, ya, wa, Xb, yb, wb, norm, n_cont = _design(fix, weighted)
        out = orc.single_pass(orc.Spec(K=2, n_cont=1, ref_kind=0), Xa, ya, wa, Xb, yb, wb)
        _check_pass(out, fix["expected"][key])
        assert abs(out["total_gap"] - a[key + "_gap"]) < a["tol"]   # weights_test.rs:36,46
    np.testing.assert_allclose(out["beta_a"], [2.0, 8.0], atol=1e-9)
    np.testing.assert_allclose(out["beta_b"], [5.0, 2.5], atol=1e-9)


def test_optimize_budget_fixture(orc, golden):
    fix = golden["F4"]
    Xa, ya, wa, Xb, yb, wb, norm, n_cont = _design(fix)
    out = orc.single_pass(orc.Spec(K=2, n_cont=1, ref_kind=0), Xa, ya, wa, Xb, yb, wb)
    assert abs(out["total_gap"] - fix["asserted"]["total_gap"]) < 1e-9     # optimize_budget_test.rs:34
    np.testing.assert_allclose(out["resid_b"], fix["asserted"]["residuals_b"], atol=1e-9)


def test_ols_known_answers(orc, golden):
    k = golden["KAT"]
    beta, resid = orc.ols(k["ols"]["y"], np.array(k["ols"]["X"], float))
    np.testing.assert_allclose(beta, k["ols"]["beta"], atol=k["ols"]["tol"])          # ols.rs:151-162
    with pytest.raises(orc.OracleError) as e:                                        # ols.rs:164-181
        orc.ols(k["ols_collinear"]["y"], np.array(k["ols_collinear"]["X"], float))
    assert e.value.code == 4
    with pytest.raises(orc.OracleError) as e:                                        # ols.rs:183-209
        orc.ols(np.ones(2), np.arange(10, dtype=float).reshape(2, 5) + 1.0)
    assert e.value.code == 6
    with pytest.raises(orc.OracleError) as e:                                        # ols.rs:60-66
        orc.ols(np.ones(3), np.ones((3, 1)), w=np.array([1.0, -1.0, 1.0]))
    assert e.value.code == 3


def test_yun_and_three_fold_known_answers(orc, golden):
    k = golden["KAT"]["yun"]
    spec = orc.Spec(K=3, n_cont=0, norm=[orc.NormVar(k["m"], k["idx"])])
    beta, base = orc.yun(spec, k["beta"])
    np.testing.assert_allclose(beta, k["beta_out"], atol=1e-12)                      # normalization.rs:58-111
    assert abs(base[0] - k["base"]) < 1e-12
    t = golden["KAT"]["three_fold"]
    import ctypes as C
    out = (C.c_double * 3)()
    arr = lambda v: np.array(v, float).ctypes.data_as(C.POINTER(C.c_double))
    orc.lib().orc_three_fold(arr(t["xa"]), arr(t["xb"]), arr(t["ba"]), arr(t["bb"]), C.c_int32(2), out)
    np.testing.assert_allclose(list(out), t["out"], atol=1e-12)                      # decomposition.rs:129-139


def test_bootstrap_stats_known_answers(orc, golden):
    for case in golden["KAT"]["bootstrap_p"]["cases"]:                               # inference.rs:41-57
        assert abs(orc.bootstrap_stats(case["est"])["p"] - case["p"]) < 1e-9
    r = orc.bootstrap_stats(np.arange(1, 101, dtype=float))
    assert r["ci_lo"] == 3.0 and r["ci_hi"] == 98.0          # floor(.025*100)=2 -> 3.0 ; floor(97.5)=97 -> 98.0
    assert abs(r["se"] - np.std(np.arange(1, 101), ddof=1)) < 1e-12
    e = orc.bootstrap_stats(np.array([]))
    assert all(np.isnan(v) for v in e.values())              # inference.rs:5-7
    one = orc.bootstrap_stats(np.array([2.0]))
    assert np.isnan(one["se"]) and one["ci_lo"] == 2.0 and one["ci_hi"] == 2.0


def test_rif_matches_formula(orc):
    """math/rif.rs:14-88 restated independently with numpy (type-7 quantile == np.quantile default)."""
    rng = np.random.default_rng(7)
    y = rng.normal(3.0, 0.7, 501)
    for tau in (0.1, 0.5, 0.9):
        n = len(y)
        q = np.quantile(y, tau)
        s = np.sort(y)
        sd = y.std(ddof=1)
        iqr = s[int(np.ceil(.75 * n)) - 1] - s[int(np.ceil(.25 * n)) - 1]
        h = 0.9 * min(sd, iqr / 1.34) * n ** -0.2
        f = np.exp(-0.5 * ((q - y) / h) ** 2).sum() / np.sqrt(2 * np.pi) / (n * h)
        exp = q + (tau - (y <= q)) / f
        np.testing.assert_allclose(orc.rif(y, tau), exp, rtol=1e-12)
    np.testing.assert_array_equal(orc.rif(np.array([5.0]), 0.5), [5.0])   # rif.rs:18-20
    const = orc.rif(np.full(10, 2.0), 0.5)                                 # spread fallback rif.rs:57
    assert np.all(np.isfinite(const))


def test_run_with_index_stream_matches_manual_gather(orc, golden):
    """orc_run == per-replicate single_pass on gathered rows, failures dropped in order."""
    fix = golden["F2"]
    Xa, ya, wa, Xb, yb, wb, norm, n_cont = _design(fix)
    spec = orc.Spec(K=4, n_cont=1, ref_kind=2, norm=[orc.NormVar(m, i) for m, i in norm])
    reps = 40
    ia, ib = orc.index_stream(11, reps, 0, len(ya)), orc.index_stream(11, reps, 1, len(yb))
    res = orc.run(spec, Xa, ya, wa, Xb, yb, wb, reps, ia, ib, nthreads=2)
    ok_rows = []
    for b in range(reps):
        try:
            p = orc.single_pass(spec, Xa[ia[b]], ya[ia[b]], None, Xb[ib[b]], yb[ib[b]], None)
            assert res["rep_status"][b] == 0
            np.testing.assert_allclose(res["rep_stats"][b], p["stats"], rtol=0, atol=1e-12)
            ok_rows.append(p["stats"])
        except orc.OracleError as e:
            assert res["rep_status"][b] == e.code
    assert res["n_ok"] == len(ok_rows) and 0 < res["n_ok"]
    ok_rows = np.array(ok_rows)
    for j in range(spec.n_stats):
        st = orc.bootstrap_stats(ok_rows[:, j])
        assert np.isclose(res["se"][j], st["se"], rtol=1e-12, atol=0, equal_nan=True)
        assert res["ci_lo"][j] == st["ci_lo"] and res["ci_hi"][j] == st["ci_hi"]
    # same stream from the oracle's internal generator
    res2 = orc.run(spec, Xa, ya, wa, Xb, yb, wb, reps, None, None, seed=11, nthreads=1)
    np.testing.assert_array_equal(res2["rep_status"], res["rep_status"])
    np.testing.assert_allclose(res2["se"], res["se"], rtol=1e-12, equal_nan=True)


def test_zero_reps(orc, golden):
    """bootstrap_reps = 0 -> SE fields NaN, t = 0 (builder.rs:851-855)."""
    fix = golden["F1"]
    Xa, ya, wa, Xb, yb, wb, norm, n_cont = _design(fix)
    res = orc.run(orc.Spec(K=2, n_cont=1), Xa, ya, wa, Xb, yb, wb, 0)
    assert res["n_ok"] == 0 and np.all(np.isnan(res["se"])) and np.all(res["t"] == 0.0)


def test_heckman_oracle_matches_the_independent_numpy_restatement():
    """oracle/ob_oracle_heckman.c (probit.rs:25-175, heckman.rs:38-108, estimation.rs:114-269, builder.rs:464-534) against
    tests/golden/heckman_fixture.json, which tests/golden/make_heckman_golden.py computed with numpy / scipy alone."""
    import json
    import os
    from oracle import pyoracle as orc
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "heckman_fixture.json")))
    c = {k: np.array(v, float) for k, v in fx["columns"].items()}
    X = np.c_[np.ones(len(c["y"])), c["x"], c["d"]]
    Z = np.c_[np.ones(len(c["y"])), c["z"]]
    A, B = c["group"] == 0, c["group"] == 1
    for name, ref in (("A", 0), ("B", 1), ("weighted", 3)):
        out = orc.heckman_run(ref, X[A], c["y"][A], Z[A], c["s"][A], X[B], c["y"][B], Z[B], c["s"][B], 0, None, None)
        exp = fx["expected"][name]
        np.testing.assert_allclose(out["point_stats"], exp["stats"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(out["beta_a"], exp["beta_a"], rtol=1e-10)
        np.testing.assert_allclose(out["gamma_b"], exp["gamma_b"], rtol=1e-10)
        assert abs(out["total_gap"] - exp["total_gap"]) < 1e-12
    # the reference's own probit test (probit.rs:180-211): converges, positive slope
    beta, conv, it = orc.probit([0, 1, 0, 1, 0, 1], [[1, -1.5], [1, -0.5], [1, 0], [1, 0.5], [1, 1], [1, 1.5]])
    assert conv and it > 0 and beta[1] > 0
    beta, conv, it = orc.probit([0, 0, 1, 1], [[1, -1], [1, -0.5], [1, 0.5], [1, 1]], max_iter=1, tol=1e-15)   # probit.rs:213-227
    assert not conv and it == 1
    with pytest.raises(RuntimeError):                   # Pooled: K vs K+1 coefficient vectors in the reference
        orc.heckman_run(2, X[A], c["y"][A], Z[A], c["s"][A], X[B], c["y"][B], Z[B], c["s"][B], 0, None, None)

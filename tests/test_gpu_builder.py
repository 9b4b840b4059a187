"""The reference's own integration tests, replayed through the full stack on the GPU:
OaxacaBuilder (C++ host layer) -> C ABI -> CUDA kernels.  Each test cites the reference test it mirrors.
Run on the B200 box: pytest -m gpu."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "oaxaca_blinder_rs_b200", "_lib", "oaxaca-cli")
WAGE_CSV = os.path.join(ROOT, "tests", "golden", "wage.csv")


@pytest.fixture(scope="module")
def ob():
    import oaxaca_blinder_rs_b200 as ob
    return ob


def sample_frame(golden):
    return dict(golden["F1"]["columns"])           # integration_test.rs:4-10


def run_and_check(results, expected_gap):           # integration_test.rs:13-53
    assert abs(results.total_gap - expected_gap) < 1e-9
    e, u = results.explained().estimate, results.unexplained().estimate
    assert abs(e + u - results.total_gap) < 1e-9, "Decomposition does not sum to total gap"
    assert results.n_a == 10 and results.n_b == 10
    s = results.summary()
    assert "Oaxaca-Blinder Decomposition Results" in s


@pytest.mark.parametrize("ref", ["GroupB", "GroupA", "Pooled", "Weighted", "Cotton", "Neumark"])
def test_full_run_reference_kinds(ob, golden, ref):                # integration_test.rs:105-144, features_test.rs
    b = ob.OaxacaBuilder(sample_frame(golden), "wage", "gender", "F")
    b.predictors(["education"]).bootstrap_reps(5).reference_coefficients(ob.ReferenceCoefficients[ref])
    r = b.run()
    run_and_check(r, 10.0)
    key = {"GroupA": "A", "GroupB": "B", "Pooled": "pooled", "Neumark": "pooled", "Weighted": "weighted", "Cotton": "weighted"}[ref]
    exp = golden["F1"]["expected"][key]
    np.testing.assert_allclose(r.beta_star, exp["beta_star"], atol=1e-9)
    np.testing.assert_allclose([c.estimate for c in r.three_fold.aggregate], exp["three_fold"], atol=1e-9)
    assert [c.name for c in r.two_fold.aggregate] == ["explained", "unexplained"]
    assert [c.name for c in r.three_fold.aggregate] == ["endowments", "coefficients", "interaction"]
    assert r.three_fold.detailed == []                             # builder.rs:942
    assert [c.name for c in r.two_fold.detailed_explained] == ["__ob_intercept__", "education"]


def test_default_reference_is_group_a(ob, golden):                 # builder.rs:123 (the code, not the doc comment)
    r = ob.OaxacaBuilder(sample_frame(golden), "wage", "gender", "F").predictors(["education"]).bootstrap_reps(0).run()
    np.testing.assert_allclose(r.beta_star, golden["F1"]["expected"]["A"]["beta_star"], atol=1e-9)


def test_with_categorical_variable(ob, golden):                    # integration_test.rs:146-163
    fix = golden["F2"]
    b = ob.OaxacaBuilder(dict(fix["columns"]), "wage", "gender", "F")
    b.predictors(["education"]).categorical_predictors(["union"]).normalize(["union"]).bootstrap_reps(5)
    r = b.run()
    run_and_check(r, 10.0)
    exp = fix["expected"]["A"]
    names = [c.name for c in r.two_fold.detailed_unexplained]
    assert names == fix["names"] + fix["base_names"]                # base row "union_none" appended (builder.rs:661-669)
    np.testing.assert_allclose([c.estimate for c in r.two_fold.detailed_explained], exp["det_expl"], atol=1e-9)
    np.testing.assert_allclose([c.estimate for c in r.two_fold.detailed_unexplained], exp["det_unexpl"], atol=1e-9)
    assert abs(sum(c.estimate for c in r.three_fold.aggregate) - 10.28) < 1e-9   # three-fold not Yun-corrected


def test_weighted_decomposition(ob, golden):                       # weights_test.rs:4-49
    fix = golden["F3"]
    f = dict(fix["columns"])
    r0 = ob.OaxacaBuilder(f, "outcome", "group", "B").predictors(["x"]).bootstrap_reps(0).run()
    assert abs(r0.total_gap - 0.666) < 0.01
    r1 = ob.OaxacaBuilder(f, "outcome", "group", "B").predictors(["x"]).weights("weight").bootstrap_reps(0).run()
    assert abs(r1.total_gap - (-3.333)) < 0.01
    # bootstrap_reps = 0: SE fields NaN, t = 0
    assert all(np.isnan(c.std_err) and c.t_stat == 0.0 for c in r1.two_fold.aggregate)


def test_null_handling(ob, golden):                                # null_handling_test.rs:4-67
    k = golden["KAT"]["null_handling"]
    f = {"outcome": k["outcome"], "group": k["group"], "education": k["education"]}
    r = ob.OaxacaBuilder(f, "outcome", "group", "B").predictors(["education"]).run()
    assert (r.n_a, r.n_b) == (3, 3)


def test_optimize_budget_inputs(ob, golden):                       # optimize_budget_test.rs:4-34
    fix = golden["F4"]
    r = ob.OaxacaBuilder(dict(fix["columns"]), "wage", "group", "B").predictors(["education"]).run()
    assert abs(r.total_gap - 16.0) < 1e-9
    np.testing.assert_allclose(r.residuals, fix["asserted"]["residuals_b"], atol=1e-9)


def test_rif_decomposition(ob, orc):                               # rif_test.rs:4-53
    wage, group, edu = [], [], []
    for i in range(100):
        wage.append(20.0 + (i % 5)); group.append("F"); edu.append(12.0 + (i % 4))
    for i in range(100):
        wage.append(15.0 + (i % 15)); group.append("M"); edu.append(12.0 + (i % 4))
    f = {"wage": wage, "group": group, "education": edu}
    r = ob.OaxacaBuilder(f, "wage", "group", "F").predictors(["education"]).bootstrap_reps(10).decompose_quantile(0.9)
    assert r.total_gap > 0.0
    # stronger than the reference's assertion: the gap equals mean(RIF_A) - mean(RIF_B) of the oracle's RIF
    w = np.array(wage)
    exp = orc.rif(w[100:], 0.9).mean() - orc.rif(w[:100], 0.9).mean()
    assert abs(r.total_gap - exp) < 1e-9 * max(1, abs(exp))


def test_formula_path(ob, golden):                                 # formula_test.rs:4-27
    fix = golden["F2"]
    r = ob.OaxacaBuilder.from_formula(dict(fix["columns"]), "wage ~ education + C(union)", "gender", "F").bootstrap_reps(3).run()
    assert abs(r.total_gap - 10.0) < 1e-9
    assert [c.name for c in r.two_fold.detailed_explained] == fix["names"]


def test_builder_vs_oracle_with_index_stream(ob, orc):
    """Full stack parity: builder -> pack -> bootstrap with an explicit index stream vs builder oracle + C oracle."""
    from oracle import builder_oracle as bo
    from test_builder_host import frame_with_everything
    f = frame_with_everything()
    preds, cats, norm = ["education", "experience"], ["sector", "region"], ["sector", "region"]
    e = bo.prepare(f, "wage", "gender", "F", preds, cats, norm, "w")
    reps = 80
    ia, ib = orc.index_stream(3, reps, 0, len(e["ya"])), orc.index_stream(3, reps, 1, len(e["yb"]))
    b = ob.OaxacaBuilder(f, "wage", "gender", "F")
    b.predictors(preds).categorical_predictors(cats).normalize(norm).weights("w").bootstrap_reps(reps)
    b.reference_coefficients(ob.ReferenceCoefficients.Pooled).index_stream(ia, ib)
    r = b.run()
    spec = orc.Spec(K=len(e["names"]), n_cont=2, ref_kind=orc.REF_POOLED,
                    norm=[orc.NormVar(v["m"], v["idx"], v["has_base"]) for v in e["norm"]])
    o = orc.run(spec, e["Xa"], e["ya"], e["wa"], e["Xb"], e["yb"], e["wb"], reps, ia, ib, nthreads=4)
    well = o["rep_min_pivot"] >= 1e-9
    assert well.all() and r.successful_bootstraps == o["n_ok"]
    D = len(e["names"]) + len(e["base_names"])
    se = np.array([c.std_err for c in r.two_fold.aggregate + r.three_fold.aggregate + r.two_fold.detailed_explained
                   + r.two_fold.detailed_unexplained])
    est = np.array([c.estimate for c in r.two_fold.aggregate + r.three_fold.aggregate + r.two_fold.detailed_explained
                    + r.two_fold.detailed_unexplained])
    assert len(se) == 5 + 2 * D
    np.testing.assert_allclose(est, o["point"]["stats"], rtol=0, atol=1e-10 * max(1, np.abs(o["point"]["stats"]).max()))
    np.testing.assert_allclose(se, o["se"], rtol=0, atol=1e-10 * max(1, np.abs(o["se"]).max()))
    xa, ya, xb, yb = b.get_data_matrices()                          # builder.rs:252-291
    np.testing.assert_array_equal(xa, e["Xa"]); np.testing.assert_array_equal(yb, e["yb"])


def test_heckman_selection_through_the_builder(ob):
    """tests/heckman_test.rs: OaxacaBuilder + .heckman_selection(..) + bootstrap_reps(0) -> an "IMR" row in the detailed
    decomposition.  Here also the numbers: against the oracle on the cleaned frame (nulls in the selection columns are part
    of clean_dataframe's filter, builder.rs:760-784), plus the detailed_selection rows (builder.rs:507-534)."""
    from oracle import pyoracle as orc
    rng = np.random.default_rng(42)
    n = 2000
    z = rng.normal(size=n)
    x = z + 0.5 * rng.normal(size=n)
    u = rng.normal(size=n)
    e = 0.8 * u + 0.6 * rng.normal(size=n)
    s = (0.5 * z + u > 0).astype(float)
    y = 1.0 + 2.0 * x + e
    grp = np.where(rng.random(n) < 0.5, "A", "B")
    zl = [None if i in (5, 77) else float(v) for i, v in enumerate(z)]        # nulls: those rows are dropped
    sl = [None if i == 123 else float(v) for i, v in enumerate(s)]
    frame = {"outcome": y.tolist(), "x": x.tolist(), "z": zl, "selection": sl, "group": grp.tolist()}
    b = ob.OaxacaBuilder(frame, "outcome", "group", "B").predictors(["x"]).heckman_selection("selection", ["z"]).bootstrap_reps(40).seed(3)
    r = b.run()
    names = [c.name for c in r.two_fold.detailed_explained]
    assert names == ["__ob_intercept__", "x", "IMR"]                          # heckman_test.rs: has_imr
    assert [c.name for c in r.two_fold.detailed_selection] == ["__ob_intercept__", "z"]
    keep = np.ones(n, bool); keep[[5, 77, 123]] = False
    A, B = keep & (grp == "A"), keep & (grp == "B")
    X = np.c_[np.ones(n), x]; Z = np.c_[np.ones(n), z]
    o = orc.heckman_run(0, X[A], y[A], Z[A], s[A], X[B], y[B], Z[B], s[B], 0, None, None)     # builder default: GroupA
    est = np.array([c.estimate for c in r.two_fold.aggregate + r.three_fold.aggregate + r.two_fold.detailed_explained
                    + r.two_fold.detailed_unexplained + r.two_fold.detailed_selection])
    np.testing.assert_allclose(est, o["point_stats"], rtol=1e-9, atol=1e-12)
    assert abs(r.total_gap - o["total_gap"]) < 1e-12 and (r.n_a, r.n_b) == (int(A.sum()), int(B.sum()))
    assert all(np.isfinite(c.std_err) and c.std_err > 0 for c in r.two_fold.aggregate)
    assert all(v == 0.0 for v in r.residuals)                                 # estimation.rs:156-157
    # Pooled with Heckman: K vs K+1 coefficient vectors in the reference -> refused, not faked
    with pytest.raises(ob.OaxacaError) as e:
        ob.OaxacaBuilder(frame, "outcome", "group", "B").predictors(["x"]).heckman_selection("selection", ["z"]) \
            .reference_coefficients(ob.ReferenceCoefficients.Pooled).bootstrap_reps(0).run()
    assert e.value.kind == "Unsupported"


def test_pyo3_surface(ob, golden):                                 # python.rs:193-256
    m = ob.OaxacaBlinder(sample_frame(golden), "wage", "gender", "F", ["education"], bootstrap_reps=4)
    r = m.fit()
    assert abs(r.total_gap - 10.0) < 1e-9 and r.two_fold.aggregate[0].name == "explained" and (r.n_a, r.n_b) == (10, 10)
    q = m.fit_quantile(0.5)
    assert np.isfinite(q.total_gap)
    with pytest.raises(RuntimeError):
        ob.OaxacaBlinder(sample_frame(golden), "wage", "gender", "F", ["nope"]).fit()


# ---- process-level CLI tests (tests/cli_test.rs:6-101) ----
def cli(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True, timeout=120)


def test_cli_mean_decomposition(tmp_path):                         # cli_test.rs:7-34
    p = cli("--data", WAGE_CSV, "--outcome", "wage", "--group", "gender", "--reference", "F", "--predictors", "education",
            "--bootstrap-reps", "2", "--output-json", str(tmp_path / "o.json"), "--output-markdown", str(tmp_path / "o.md"))
    assert p.returncode == 0, p.stderr
    for h in ("Oaxaca-Blinder Decomposition Results", "Two-Fold Decomposition", "Detailed Decomposition (Explained)",
              "Detailed Decomposition (Unexplained)"):
        assert h in p.stdout
    d = json.load(open(tmp_path / "o.json"))                        # export_test.rs: keys of to_json
    assert {"total_gap", "two_fold", "three_fold", "n_a", "n_b", "residuals"} <= set(d) and abs(d["total_gap"] - 10.0) < 1e-9
    assert "Two-Fold Decomposition" in open(tmp_path / "o.md").read()


def test_cli_with_categorical():                                   # cli_test.rs:36-58
    p = cli("--data", WAGE_CSV, "--outcome", "wage", "--group", "gender", "--reference", "F", "--predictors", "education",
            "--categorical", "sector", "--bootstrap-reps", "2")
    assert p.returncode == 0 and "Oaxaca-Blinder Decomposition Results" in p.stdout


def test_cli_formula_and_rif():
    p = cli("--data", WAGE_CSV, "--formula", "wage ~ education + C(sector)", "--group", "gender", "--reference", "F",
            "--bootstrap-reps", "3", "--ref-coeffs", "pooled", "--rif-quantile", "0.5")
    assert p.returncode == 0, p.stderr


def test_cli_invalid_argument():                                   # cli_test.rs:86-101
    p = cli("--data", os.path.join(ROOT, "tests", "golden", "non_existent_file.csv"), "--outcome", "wage", "--group", "gender",
            "--reference", "F", "--predictors", "education")
    assert p.returncode != 0 and "Error:" in p.stderr


def test_design_outliving_its_context_is_harmless():
    """Finalisers run in arbitrary order (and on arbitrary threads): closing a context first must orphan its designs,
    not leave them pointing at freed streams / pools."""
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(3_000, 2, seed=2)
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    ctx.close()
    ctx2 = ob.Context(0)
    with pytest.raises(ob.OaxacaError) as e:
        des.ctx = ctx2
        ob.bootstrap(des, 4, seed=1)
    assert e.value.kind == "InvalidArgument"
    des.close()          # only frees the host struct
    ctx2.close()


def test_cli_heckman_flags(tmp_path):
    """--selection-outcome / --selection-predictors (main.rs RunArgs) reach heckman_selection: an IMR row is printed."""
    rng = np.random.default_rng(9)
    n = 1500
    z = rng.normal(size=n); x = z + 0.5 * rng.normal(size=n); u = rng.normal(size=n)
    s = (0.5 * z + u > 0).astype(int)
    y = 1.0 + 2.0 * x + 0.8 * u + 0.6 * rng.normal(size=n)
    g = np.where(rng.random(n) < 0.5, "A", "B")
    path = tmp_path / "sel.csv"
    with open(path, "w") as fh:
        fh.write("outcome,x,z,selection,group\n")
        for i in range(n):
            fh.write(f"{y[i]:.10g},{x[i]:.10g},{z[i]:.10g},{s[i]},{g[i]}\n")
    r = subprocess.run([CLI, "--data", str(path), "--outcome", "outcome", "--group", "group", "--reference", "B", "--predictors", "x",
                        "--selection-outcome", "selection", "--selection-predictors", "z", "--bootstrap-reps", "20", "--seed", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "IMR" in r.stdout and "explained" in r.stdout

"""Machado-Mata quantile decomposition (quantile_decomposition.rs:173-421, math/quantile_regression.rs:22-135)
on the GPU through the C ABI (ob_mm_run) vs the oracle (oracle/ob_oracle_mm.c, itself pinned against HiGHS dual-simplex
vertices and the reference's known answers), under explicit streams: resample indices, random quantiles, simulated rows.
Every quantile regression must land on the same LP vertex (coefficients within 1e-10), hence every effect, SE and CI."""
import json
import os

import numpy as np
import pytest

from helpers import relerr

pytestmark = pytest.mark.gpu
RTOL = 1e-10
HERE = os.path.dirname(os.path.abspath(__file__))


def make_frame(n, n_x, seed, round_y=None):
    rng = np.random.default_rng(seed)
    grp = rng.integers(0, 2, n).astype(np.uint8)
    xs = [rng.normal(12, 2, n) + 0.5 * (grp == 0)] + [rng.uniform(0, 30, n) for _ in range(n_x - 1)]
    cat = rng.integers(0, 3, n).astype(np.int32)
    y = 1.0 + 0.3 * (grp == 0) + (0.07 + 0.02 * (grp == 0)) * xs[0] + sum(0.01 * x for x in xs[1:]) + 0.1 * (cat == 1) \
        - 0.15 * (cat == 2) + rng.standard_t(4, n) * 0.3 * (1 + 0.03 * xs[0])
    if round_y is not None:
        y = y.round(round_y)
    return dict(group=grp, cont=xs, cat=cat, y=y)


def dense(fr):
    n = len(fr["y"])
    X = np.c_[np.ones(n), np.stack(fr["cont"], 1), (fr["cat"] == 1).astype(float), (fr["cat"] == 2).astype(float)]
    A, B = fr["group"] == 0, fr["group"] == 1
    return (X[A], fr["y"][A]), (X[B], fr["y"][B])


def streams(seed, reps, sims, na, nb):
    rng = np.random.default_rng(seed)
    return dict(taus=rng.uniform(0.01, 0.99, size=(reps + 1, sims)),
                draw_a=rng.integers(0, na, size=(reps + 1, sims)).astype(np.uint32),
                draw_b=rng.integers(0, nb, size=(reps + 1, sims)).astype(np.uint32),
                idx_a=rng.integers(0, na, size=(reps, na)).astype(np.uint32),
                idx_b=rng.integers(0, nb, size=(reps, nb)).astype(np.uint32))


@pytest.mark.parametrize("n,n_x,sims,reps,round_y", [(600, 2, 40, 6, None), (4000, 3, 60, 8, None), (2500, 2, 50, 5, 2),
                                                     (1500, 12, 30, 4, None), (3000, 20, 20, 2, None), (3000, 28, 16, 2, None),
                                                     (4000, 36, 12, 1, None), (5000, 44, 12, 1, None)])
def test_mm_matches_oracle(orc, n, n_x, sims, reps, round_y):
    import oaxaca_blinder_rs_b200 as ob
    fr = make_frame(n, n_x, seed=10 + n_x, round_y=round_y)
    (Xa, ya), (Xb, yb) = dense(fr)
    st = streams(5, reps, sims, len(ya), len(yb))
    q = [0.1, 0.25, 0.5, 0.75, 0.9]
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, fr["cont"], [fr["cat"]], [3], fr["y"], None, fr["group"])
    gpu = ob.machado_mata(des, q, simulations=sims, reps=reps, want_rep=True, want_betas=True, **st)
    des.close(); ctx.close()
    o = orc.mm_run(Xa, ya, Xb, yb, sims, q, reps, st["idx_a"], st["idx_b"], st["taus"], st["draw_a"], st["draw_b"], nthreads=8)
    assert gpu["qr"]["failed"] == 0 and gpu["qr"]["total"] == 2 * (reps + 1) * sims, (gpu["qr"], gpu["point_qr_info_a"], gpu["point_qr_info_b"])
    # the point pass's regressions, coefficient by coefficient: the same LP vertex
    assert ((gpu["point_qr_info_a"] & 0xff) == ob.core.QR_VERTEX).all() and ((gpu["point_qr_info_b"] & 0xff) == ob.core.QR_VERTEX).all()
    assert relerr(gpu["point_betas_a"], o["betas_a"]) <= RTOL, relerr(gpu["point_betas_a"], o["betas_a"])
    assert relerr(gpu["point_betas_b"], o["betas_b"]) <= RTOL
    nq = len(q)
    assert relerr(gpu["point_stats"], o["point_stats"].reshape(nq, 3)) <= RTOL
    assert np.array_equal(gpu["rep_status"], o["rep_status"]) and gpu["n_ok"] == o["n_ok"] == reps
    assert relerr(gpu["rep_stats"], o["rep_stats"].reshape(reps, nq, 3)) <= RTOL
    for k, ko in (("std_err", "se"), ("p_value", "p"), ("ci_lower", "ci_lo"), ("ci_upper", "ci_hi"), ("t_stat", "t")):
        assert relerr(gpu[k], o[ko].reshape(nq, 3)) <= RTOL, k
    # gap = characteristics + coefficients in every pass (quantile_decomposition.rs:271-275)
    np.testing.assert_allclose(gpu["rep_stats"][..., 0], gpu["rep_stats"][..., 1] + gpu["rep_stats"][..., 2], rtol=0, atol=1e-12)


def test_mm_golden_fixture_on_gpu():
    """tests/golden/mm_fixture.json (HiGHS dual simplex + numpy, independent of the oracle) straight against the GPU."""
    import oaxaca_blinder_rs_b200 as ob
    fx = json.load(open(os.path.join(HERE, "golden", "mm_fixture.json")))
    Xa, ya, Xb, yb = (np.array(fx[k]) for k in ("Xa", "ya", "Xb", "yb"))
    ctx = ob.Context(0)
    des = ob.Design.from_dense(ctx, Xa, ya, None, Xb, yb, None, n_cont=2)
    out = ob.machado_mata(des, fx["quantiles"], simulations=fx["sims"], reps=fx["reps"], idx_a=fx["idx_a"], idx_b=fx["idx_b"],
                          taus=fx["taus"], draw_a=fx["draw_a"], draw_b=fx["draw_b"], want_rep=True, want_betas=True)
    des.close(); ctx.close()
    np.testing.assert_allclose(out["point_betas_a"], fx["point_betas_a"], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(out["point_betas_b"], fx["point_betas_b"], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(out["point_stats"], fx["point_stats"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(out["rep_stats"], fx["rep_stats"], rtol=0, atol=1e-10)


def test_reference_known_answer_on_gpu():
    """math/quantile_regression.rs:137-170: perfectly linear data, every regression quantile is [0, 1] (tolerance 1e-4
    there); both groups get the same data, so every effect is 0."""
    import oaxaca_blinder_rs_b200 as ob
    X = np.array([[1, 1], [1, 2], [1, 3], [1, 4], [1, 5]], dtype=float)
    y = np.array([1, 2, 3, 4, 5], dtype=float)
    ctx = ob.Context(0)
    des = ob.Design.from_dense(ctx, X, y, None, X, y, None, n_cont=1)
    taus = np.array([[0.5, 0.25, 0.1, 0.9]])
    out = ob.machado_mata(des, [0.5], simulations=4, reps=0, taus=taus, draw_a=[[0, 1, 2, 3]], draw_b=[[0, 1, 2, 3]], want_betas=True)
    des.close(); ctx.close()
    np.testing.assert_allclose(out["point_betas_a"], np.tile([0.0, 1.0], (4, 1)), rtol=0, atol=1e-12)
    np.testing.assert_allclose(out["point_stats"], 0.0, atol=1e-12)
    assert np.isnan(out["std_err"]).all()


def test_native_streams_are_deterministic_and_shard_invariant():
    """Native Philox streams (resamples, random quantiles, simulated rows) are keyed by global pass ids: the same seed gives
    the same bits, a replicate shard computes exactly its rows of the full run, another seed gives other draws; and the
    decomposition identity holds in every pass."""
    import oaxaca_blinder_rs_b200 as ob
    fr = make_frame(3000, 2, seed=77)
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, fr["cont"], [fr["cat"]], [3], fr["y"], None, fr["group"])
    kw = dict(quantiles=[0.25, 0.5, 0.75], simulations=64, reps=12, want_rep=True)
    a = ob.machado_mata(des, seed=9, **kw)
    b = ob.machado_mata(des, seed=9, **kw)
    c = ob.machado_mata(des, seed=10, **kw)
    part = ob.machado_mata(des, seed=9, rep_begin=5, rep_end=9, skip_reduce=True, **kw)
    des.close(); ctx.close()
    assert np.array_equal(a["rep_stats"], b["rep_stats"]) and np.array_equal(a["std_err"], b["std_err"])
    assert not np.array_equal(a["rep_stats"], c["rep_stats"])
    assert np.array_equal(part["rep_stats"], a["rep_stats"][5:9]) and np.array_equal(part["point_stats"], a["point_stats"])
    assert a["n_ok"] == 12 and a["qr"]["failed"] == 0
    np.testing.assert_allclose(a["rep_stats"][..., 0], a["rep_stats"][..., 1] + a["rep_stats"][..., 2], rtol=0, atol=1e-12)
    assert np.isfinite(a["std_err"]).all() and (a["std_err"] > 0).all()
    # group A's outcome is shifted up: a positive gap at every quantile, in the point pass and in every bootstrap pass
    assert (a["point_stats"][:, 0] > 0).all() and (a["rep_stats"][..., 0] > 0).all()
    assert (a["ci_lower"][:, 0] > 0).all() and (a["ci_lower"] <= a["ci_upper"]).all()


def test_failed_pass_and_refusals():
    import oaxaca_blinder_rs_b200 as ob
    rng = np.random.default_rng(2)
    n = 60
    Xa = np.c_[np.ones(n), rng.normal(size=n)]
    ya = Xa @ [1.0, 0.5] + rng.normal(size=n)
    Xb = np.c_[np.ones(n), np.ones(n)]          # collinear: every regression of group B fails -> the point pass fails
    ctx = ob.Context(0)
    des = ob.Design.from_dense(ctx, Xa, ya, None, Xb, rng.normal(size=n), None, n_cont=1)
    with pytest.raises(ob.OaxacaError) as e:
        ob.machado_mata(des, [0.5], simulations=8, reps=0)
    assert e.value.kind == "NalgebraError"       # quantile_decomposition.rs:238-242
    des.close()
    des = ob.Design.from_dense(ctx, Xa, ya, np.ones(n), Xa, ya, np.ones(n), n_cont=1)
    with pytest.raises(ob.OaxacaError) as e:
        ob.machado_mata(des, [0.5], simulations=8, reps=0)
    assert e.value.kind == "Unsupported"
    des.close(); ctx.close()


# ---- builder mirrors: QuantileDecompositionBuilder (quantile_decomposition.rs:21-100) ----
def reference_frame():
    """tests/integration_test.rs:166-172"""
    return {"wage": [10.0, 12.0, 11.0, 13.0, 15.0, 20.0, 22.0, 21.0, 23.0, 25.0, 9.0, 18.0],
            "education": [12.0, 16.0, 14.0, 16.0, 18.0, 12.0, 16.0, 14.0, 16.0, 18.0, 10.0, 20.0],
            "gender": ["F", "F", "F", "F", "F", "F", "M", "M", "M", "M", "M", "M"]}


def test_quantile_decomposition_builder_reference_test():
    """tests/integration_test.rs:166-197, assertion for assertion."""
    import oaxaca_blinder_rs_b200 as ob
    b = ob.QuantileDecompositionBuilder(reference_frame(), "wage", "gender", "F")
    results = b.predictors(["education"]).quantiles([0.25, 0.5, 0.75]).simulations(10).bootstrap_reps(2).run()
    for key in ("q25", "q50", "q75"):
        assert key in results.results_by_quantile
        d = results.results_by_quantile[key]
        assert abs(d.characteristics_effect.estimate + d.coefficients_effect.estimate - d.total_gap.estimate) < 1e-9
        assert (d.total_gap.name, d.characteristics_effect.name, d.coefficients_effect.name) == ("Total Gap", "Characteristics", "Coefficients")
    assert (results.n_a, results.n_b) == (6, 6)
    s = results.summary()
    assert "Machado-Mata Quantile Decomposition Results" in s and "--- Decomposition for Quantile: q25 ---" in s
    assert "Group A (Advantaged): 6 observations" in s


def test_builder_with_streams_matches_oracle(orc):
    """The whole builder path (string group / categorical columns -> device ingest -> ob_mm_run) under explicit streams vs the
    oracle on the matrices get_data_matrices() returns; and the reference's key format "q{(tau * 100) as u32}"
    (quantile_decomposition.rs:277: 0.29 * 100 = 28.999999999999996 -> "q28")."""
    import oaxaca_blinder_rs_b200 as ob
    fr = make_frame(1200, 2, seed=31)
    frame = {"y": fr["y"], "x0": fr["cont"][0], "x1": fr["cont"][1], "sector": np.array(["s%d" % c for c in fr["cat"]]),
             "g": np.where(fr["group"] == 0, "A", "B")}
    xa, ya, xb, yb = ob.OaxacaBuilder(frame, "y", "g", "B").predictors(["x0", "x1"]).categorical_predictors(["sector"]).get_data_matrices()
    sims, reps, q = 24, 3, [0.29, 0.5, 0.9]
    st = streams(8, reps, sims, len(ya), len(yb))
    b = ob.QuantileDecompositionBuilder(frame, "y", "g", "B").predictors(["x0", "x1"]).categorical_predictors(["sector"])
    res = b.quantiles(q).simulations(sims).bootstrap_reps(reps).streams(**st).run()
    o = orc.mm_run(xa, ya, xb, yb, sims, q, reps, st["idx_a"], st["idx_b"], st["taus"], st["draw_a"], st["draw_b"], nthreads=4)
    assert sorted(res.results_by_quantile) == ["q28", "q50", "q90"]
    assert res.successful_bootstraps == reps and res.qr["failed"] == 0
    for k, key in enumerate(("q28", "q50", "q90")):
        d = res.results_by_quantile[key]
        for j, comp in enumerate((d.total_gap, d.characteristics_effect, d.coefficients_effect)):
            i = 3 * k + j
            assert abs(comp.estimate - o["point_stats"][i]) <= RTOL * max(abs(o["point_stats"][i]), 1e-3)
            assert abs(comp.std_err - o["se"][i]) <= RTOL * max(abs(o["se"][i]), 1e-3)
            assert abs(comp.ci_lower - o["ci_lo"][i]) <= RTOL * max(abs(o["ci_lo"][i]), 1e-3)
            assert abs(comp.t_stat - o["t"][i]) <= 1e-8 * max(abs(o["t"][i]), 1.0)


def test_builder_null_outcome_is_an_error():
    """prepare_data, quantile_decomposition.rs:111-118: a null outcome in one of the two groups is an error (the
    Machado-Mata builder does not clean the frame)."""
    import oaxaca_blinder_rs_b200 as ob
    f = reference_frame()
    f["wage"] = list(f["wage"]); f["wage"][3] = None
    with pytest.raises(ob.OaxacaError) as e:
        ob.QuantileDecompositionBuilder(f, "wage", "gender", "F").predictors(["education"]).simulations(10).bootstrap_reps(1).run()
    assert e.value.kind == "InvalidGroupVariable" and "Null outcome encountered" in str(e.value)


def test_cli_quantile_decomposition(tmp_path):                     # tests/cli_test.rs:60-83
    import subprocess
    root = os.path.dirname(HERE)
    cli = os.path.join(root, "oaxaca_blinder_rs_b200", "_lib", "oaxaca-cli")
    p = subprocess.run([cli, "--data", os.path.join(HERE, "golden", "wage.csv"), "--outcome", "wage", "--group", "gender", "--reference", "F",
                        "--predictors", "education", "--analysis-type", "quantile", "--bootstrap-reps", "2", "--simulations", "10",
                        "--output-json", str(tmp_path / "q.json")], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    assert "Machado-Mata Quantile Decomposition Results" in p.stdout
    d = json.load(open(tmp_path / "q.json"))
    assert sorted(d["results_by_quantile"]) == ["q10", "q25", "q50", "q75", "q90"]       # main.rs:236-239 defaults


@pytest.mark.parametrize("world,reps", [(2, 9), (3, 4)])
def test_mm_library_replicate_sharding_is_bit_identical(world, reps):
    """Mode R for the Machado-Mata passes (ob_mm_opts.shard_replicates): every rank holds the design, takes its share of
    the bootstrap passes, the pass rows are all-gathered over the context's communicator and reduced on every rank:
    bit-identical to one GPU.  `world` contexts of one process joined by the in-process communicator (the transport is
    the only difference to NCCL)."""
    import threading
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import core
    fr = make_frame(2000, 2, seed=5)
    kw = dict(quantiles=[0.1, 0.5, 0.9], simulations=32, reps=reps, seed=4, want_rep=True)

    def pack(ctx):
        return ob.Design.pack(ctx, fr["cont"], [fr["cat"]], [3], fr["y"], None, fr["group"])
    ctx = ob.Context(0)
    des = pack(ctx)
    one = ob.machado_mata(des, **kw)
    des.close(); ctx.close()
    grp = core.LocalGroup(world)
    outs, errs = [None] * world, [None] * world

    def work(r):
        try:
            c = ob.Context(0)
            c.init_local(grp, r)
            d = pack(c)
            outs[r] = ob.machado_mata(d, shard_replicates=True, **kw)
            d.close(); c.close()
        except Exception as e:  # noqa: BLE001
            errs[r] = e
    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert all(e is None for e in errs), errs
    for o in outs:
        assert o["n_ok"] == one["n_ok"] == reps and o["rep_stats"].shape == one["rep_stats"].shape
        for k in ("point_stats", "rep_stats", "rep_status", "std_err", "p_value", "ci_lower", "ci_upper", "t_stat"):
            assert np.array_equal(np.nan_to_num(o[k], nan=-7.0), np.nan_to_num(one[k], nan=-7.0)), k


def test_mm_fullsize_point_pass_matches_oracle(orc):
    """At the size the bench line quotes (n = 2e5 rows, 1e5 per group, K = 10): every regression of the point pass and of
    one bootstrap pass against the oracle -- long rows are where the summation order of the Gram sweeps and the
    interior-point path differ most, and the polish has to pick K rows out of 1e5."""
    import oaxaca_blinder_rs_b200 as ob
    fr = make_frame(200_000, 7, seed=1)
    (Xa, ya), (Xb, yb) = dense(fr)
    sims, reps, q = 16, 1, [0.1, 0.5, 0.9]
    st = streams(12, reps, sims, len(ya), len(yb))
    st["taus"][0, :4] = [0.0102, 0.0209, 0.9791, 0.9898]          # the extreme quantiles: the hardest for the interior point
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, fr["cont"], [fr["cat"]], [3], fr["y"], None, fr["group"])
    gpu = ob.machado_mata(des, q, simulations=sims, reps=reps, want_rep=True, want_betas=True, **st)
    des.close(); ctx.close()
    o = orc.mm_run(Xa, ya, Xb, yb, sims, q, reps, st["idx_a"], st["idx_b"], st["taus"], st["draw_a"], st["draw_b"], nthreads=16)
    assert gpu["qr"] == dict(total=4 * sims, vertex=4 * sims, approx=0, failed=0, iterations=gpu["qr"]["iterations"])
    assert relerr(gpu["point_betas_a"], o["betas_a"]) <= RTOL, relerr(gpu["point_betas_a"], o["betas_a"])
    assert relerr(gpu["point_betas_b"], o["betas_b"]) <= RTOL, relerr(gpu["point_betas_b"], o["betas_b"])
    assert relerr(gpu["point_stats"], o["point_stats"].reshape(3, 3)) <= RTOL
    assert relerr(gpu["rep_stats"], o["rep_stats"].reshape(reps, 3, 3)) <= RTOL
    iters = (gpu["point_qr_info_a"] >> 8) & 0xff
    assert iters.max() <= 40, iters                              # incl. tau = 0.0102 and 0.9898 (one step length for both iterates)


def _edge_cases():
    rng = np.random.default_rng(0)
    n = 200
    X = np.c_[np.ones(n), rng.normal(size=n)]
    y = X @ [1.0, 2.0] + rng.normal(size=n)
    d = rng.integers(0, 3, size=n)
    Xd = np.c_[np.ones(n), d == 1, d == 2].astype(float)
    return {
        "constant outcome": (X, np.full(n, 3.0)),
        "zero outcome": (X, np.zeros(n)),
        "two rows": (np.array([[1, 1.0], [1, 2.0]]), np.array([1.0, 3.0])),
        "three rows": (np.array([[1, 1.0], [1, 2.0], [1, 4.0]]), np.array([1.0, 3.0, 2.0])),
        "outcome scale 1e9": (X, 1e9 * y),
        "outcome scale 1e-9": (X, 1e-9 * y),
        "dummies only, integer outcome (ties)": (Xd, rng.integers(0, 5, size=n).astype(float)),
        "intercept only": (np.ones((n, 1)), y),
    }


@pytest.mark.parametrize("name", list(_edge_cases()))
def test_mm_edge_shapes_match_oracle(orc, name):
    """Degenerate and badly scaled inputs: perfect fits (every row has a zero residual), groups of two and three rows,
    outcomes of size 1e9 and 1e-9 (every tolerance of the solver is relative), tie-heavy data (dozens of zero-residual rows
    per regression), an intercept-only design.  Group B is a well-behaved frame throughout."""
    import oaxaca_blinder_rs_b200 as ob
    Xa, ya = _edge_cases()[name]
    rng = np.random.default_rng(1)
    nb = 150
    Xb = np.c_[np.ones(nb), rng.normal(size=(nb, Xa.shape[1] - 1))]
    yb = Xb @ np.arange(1, Xa.shape[1] + 1) + rng.normal(size=nb)
    sims, reps = 12, 0
    st = streams(3, reps, sims, len(ya), len(yb))
    # (not 0.05 / 0.3 / ..: with tau n an integer the tau-th quantile of an intercept-only design is a whole interval, the LP
    # has no unique vertex, and both solvers report "interior-point solution only" -- correctly)
    st["taus"][0, :4] = [0.0503, 0.3011, 0.8017, 0.9707]
    q = [0.25, 0.5, 0.75]
    ctx = ob.Context(0)
    des = ob.Design.from_dense(ctx, Xa, ya, None, Xb, yb, None, n_cont=Xa.shape[1] - 1)
    gpu = ob.machado_mata(des, q, simulations=sims, reps=reps, want_betas=True, taus=st["taus"], draw_a=st["draw_a"], draw_b=st["draw_b"])
    des.close(); ctx.close()
    o = orc.mm_pass(Xa, ya, Xb, yb, st["taus"][0], st["draw_a"][0], st["draw_b"][0], q)
    assert o["rc"] == 0 and np.array_equal(gpu["point_qr_info_a"] & 0xff, o["status_a"]) and (o["status_a"] == orc.QR_VERTEX).all()
    scale = np.abs(ya).max() or 1.0          # (an all-zero outcome: coefficients are 0 up to denormal noise)
    assert np.abs(gpu["point_betas_a"] - o["betas_a"]).max() <= 1e-10 * max(scale, np.abs(o["betas_a"]).max())
    assert relerr(gpu["point_betas_b"], o["betas_b"]) <= RTOL
    assert np.abs(gpu["point_stats"] - o["stats"]).max() <= 1e-10 * max(np.abs(o["stats"]).max(), scale, np.abs(yb).max())


def test_mm_rank_deficient_group_fails_like_a_dropped_fit(orc):
    """More columns than rows in group A: every regression of A fails the rank test (status 2), on the device as in the
    oracle, and the point pass is an error (quantile_decomposition.rs:238-242).  (clarabel would return one of the LP's
    non-unique solutions here)"""
    import oaxaca_blinder_rs_b200 as ob
    Xa, ya = np.array([[1, 1.0, 2.0], [1, 2.0, 1.0]]), np.array([1.0, 3.0])
    rng = np.random.default_rng(1)
    Xb = np.c_[np.ones(50), rng.normal(size=(50, 2))]
    yb = rng.normal(size=50)
    o = orc.mm_pass(Xa, ya, Xb, yb, [0.3, 0.6], [0, 1], [0, 1], [0.5])
    assert o["rc"] == 4 and (o["status_a"] == orc.QR_FAILED).all() and (o["status_b"] == orc.QR_VERTEX).all()
    ctx = ob.Context(0)
    des = ob.Design.from_dense(ctx, Xa, ya, None, Xb, yb, None, n_cont=2)
    with pytest.raises(ob.OaxacaError) as e:
        ob.machado_mata(des, [0.5], simulations=2, reps=0, taus=[[0.3, 0.6]], draw_a=[[0, 1]], draw_b=[[0, 1]])
    assert e.value.kind == "NalgebraError"
    des.close(); ctx.close()


def test_mm_batching_and_count_width_do_not_change_a_bit():
    """More passes than one workspace batch holds (panels of 128 multiplicity columns, processed under a workspace budget)
    and 16-bit multiplicities: the same bits as the single-batch 8-bit run, native streams (keyed by global pass ids)."""
    import oaxaca_blinder_rs_b200 as ob
    fr = make_frame(1500, 2, seed=3)
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, fr["cont"], [fr["cat"]], [3], fr["y"], None, fr["group"])
    kw = dict(quantiles=[0.25, 0.5, 0.75], simulations=8, reps=300, seed=21, want_rep=True)
    one = ob.machado_mata(des, **kw)
    small = ob.machado_mata(des, max_workspace_bytes=1_000_000, **kw)         # one panel per batch (three batches), 13 blocks
    wide = ob.machado_mata(des, count_bits=16, **kw)
    des.close(); ctx.close()
    assert one["n_ok"] == 300 and one["qr"]["total"] == 2 * 301 * 8
    for other in (small, wide):
        for k in ("point_stats", "rep_stats", "rep_status", "std_err", "p_value", "ci_lower", "ci_upper", "t_stat"):
            assert np.array_equal(np.nan_to_num(other[k], nan=-7.0), np.nan_to_num(one[k], nan=-7.0)), k
        assert other["qr"] == one["qr"]
    assert small["gpu_launches"] > one["gpu_launches"]


def test_mm_many_simulations(orc):
    """2000 simulations per pass: the effects kernel sorts them in more than 48 KB of (opt-in) shared memory; the sorted
    quantiles must still equal the oracle's.  4096 is the limit of the C ABI, 4097 is refused."""
    import oaxaca_blinder_rs_b200 as ob
    fr = make_frame(400, 2, seed=8)
    (Xa, ya), (Xb, yb) = dense(fr)
    sims, q = 2000, [0.05, 0.5, 0.95]
    st = streams(2, 0, sims, len(ya), len(yb))
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, fr["cont"], [fr["cat"]], [3], fr["y"], None, fr["group"])
    gpu = ob.machado_mata(des, q, simulations=sims, reps=0, taus=st["taus"], draw_a=st["draw_a"], draw_b=st["draw_b"])
    o = orc.mm_pass(Xa, ya, Xb, yb, st["taus"][0], st["draw_a"][0], st["draw_b"][0], q)
    assert o["rc"] == 0 and gpu["qr"]["failed"] == 0 and gpu["qr"]["total"] == 2 * sims
    assert relerr(gpu["point_stats"], o["stats"]) <= RTOL
    big = ob.machado_mata(des, q, simulations=4096, reps=0, seed=1)
    assert big["qr"]["total"] == 2 * 4096 and np.isfinite(big["point_stats"]).all()
    with pytest.raises(ob.OaxacaError) as e:
        ob.machado_mata(des, q, simulations=4097, reps=0)
    assert e.value.kind == "InvalidArgument"
    des.close(); ctx.close()

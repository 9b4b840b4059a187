"""GPU-vs-oracle parity through the C ABI (libobboot).  Run on the B200 box: pytest -m gpu.

Same seeded inputs, same explicit resample index stream on both sides:
  * multiplicities: bit-exact (checked via the replicate status/statistics and the count tests)
  * per-replicate coefficients and statistics, SEs, CIs: <= 1e-10 relative (north_star)
  * replicate failures (singular resamples) dropped identically.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-10   # north_star: "within a relative 1e-10 (fp64 with a different summation order)"
REFS = {"A": 0, "B": 1, "pooled": 2, "weighted": 3}


from helpers import relerr, relerr_to_scale, worst  # noqa: E402  (elementwise-relative metric, tests/helpers.py)


@pytest.fixture(scope="module")
def ob():
    import oaxaca_blinder_rs_b200 as ob
    return ob


@pytest.fixture(scope="module")
def ctx(ob):
    c = ob.Context(0)
    yield c
    c.close()


def run_both(ob, orc, ctx, Xa, ya, wa, Xb, yb, wb, n_cont, ref, norm, reps, seed=5, **kw):
    K = Xa.shape[1]
    spec = orc.Spec(K=K, n_cont=n_cont, ref_kind=ref, norm=[orc.NormVar(m, i, True) for m, i in norm])
    ia, ib = orc.index_stream(seed, reps, 0, len(ya)), orc.index_stream(seed, reps, 1, len(yb))
    ref_out = orc.run(spec, Xa, ya, wa, Xb, yb, wb, reps, ia, ib, nthreads=8, precise=True)
    des = ob.Design.from_dense(ctx, Xa, ya, wa, Xb, yb, wb, n_cont)
    gpu = ob.bootstrap(des, reps, ref_kind=ref, norm=[ob.NormVar(m, i, True) for m, i in norm],
                       idx_a=ia, idx_b=ib, want_rep=True, **kw)
    des.close()
    return gpu, ref_out


def compare(gpu, ref, tol=RTOL, orc=None, cancel_floor=0.0):
    """Point estimates, every numerically well-posed replicate, and the reduction.

    A resample whose design is rank-deficient up to rounding (oracle's smallest relative Cholesky pivot
    < 1e-9, i.e. a mathematically zero pivot) succeeds or fails by the sign of rounding noise in ANY
    fp64 implementation, the reference's included; those replicates are excluded from the per-replicate
    comparison and the reduction is then checked on the GPU's own replicate set."""
    p = ref["point"]
    assert abs(gpu["total_gap"] - p["total_gap"]) <= tol * max(1, abs(p["total_gap"]))
    for k_gpu, k_ref in (("point_stats", "stats"), ("xa_mean", "xa_mean"), ("xb_mean", "xb_mean"),
                         ("beta_star", "beta_star"), ("beta_a", "beta_a"), ("beta_b", "beta_b")):
        assert relerr(gpu[k_gpu], p[k_ref]) <= tol, k_gpu
    # residuals y - x.beta are differences of O(|y|) numbers: their error is relative to that scale
    yscale = max(1.0, float(np.max(np.abs(p["resid_b"]))), abs(p["total_gap"]))
    assert relerr_to_scale(gpu["residuals_b"], p["resid_b"], yscale) <= tol, "residuals_b"
    well = ref["rep_min_pivot"] >= 1e-9
    np.testing.assert_array_equal(gpu["rep_status"][well], ref["rep_status"][well])
    assert relerr(gpu["rep_stats"][well], ref["rep_stats"][well], cancel_floor=cancel_floor) <= tol, \
        worst(gpu["rep_stats"][well], ref["rep_stats"][well], cancel_floor=cancel_floor)
    assert relerr(gpu["rep_beta_a"][well], ref["rep_beta_a"][well]) <= tol, worst(gpu["rep_beta_a"][well], ref["rep_beta_a"][well])
    assert relerr(gpu["rep_beta_b"][well], ref["rep_beta_b"][well]) <= tol, worst(gpu["rep_beta_b"][well], ref["rep_beta_b"][well])
    if np.array_equal(gpu["rep_status"], ref["rep_status"]) and well.all():
        red = ref
    else:   # same replicate set as the GPU: oracle's bootstrap_stats over the GPU's successful replicates
        import oracle.pyoracle as po
        red = po.reduce(np.where(np.isnan(gpu["rep_stats"]), 0.0, gpu["rep_stats"]), gpu["rep_status"], gpu["point_stats"])
        red = dict(se=red["se"], ci_lo=red["ci_lo"], ci_hi=red["ci_hi"], p=red["p"], t=red["t"], n_ok=red["n_ok"])
    assert gpu["n_ok"] == red["n_ok"]
    assert relerr(gpu["std_err"], red["se"]) <= tol
    assert relerr(gpu["ci_lower"], red["ci_lo"]) <= tol
    assert relerr(gpu["ci_upper"], red["ci_hi"]) <= tol
    np.testing.assert_allclose(gpu["p_value"], red["p"], rtol=0, atol=1e-15)
    ok = np.abs(red["se"]) > 1e-8        # away from the |se| > 1e-9 switch of builder.rs:851
    assert relerr(gpu["t_stat"][ok], red["t"][ok]) <= 1e-9
    return int((~well).sum())


def _fixture_design(fix, weighted=False):
    from helpers import fixture_design
    return fixture_design(fix, weighted)


@pytest.mark.parametrize("ref", ["A", "B", "pooled", "weighted"])
@pytest.mark.parametrize("fx", ["F1", "F2"])
def test_golden_fixtures_on_gpu(ob, orc, ctx, golden, fx, ref):
    """The reference's integration fixtures (integration_test.rs:105-163) through the CUDA path."""
    fix = golden[fx]
    Xa, ya, wa, Xb, yb, wb, norm, n_cont = _fixture_design(fix)
    gpu, ref_out = run_both(ob, orc, ctx, Xa, ya, wa, Xb, yb, wb, n_cont, REFS[ref], norm, reps=60)
    exp = fix["expected"][ref]
    for k in ("beta_a", "beta_b", "xa_mean", "xb_mean", "beta_star", "two_fold", "three_fold", "det_expl", "det_unexpl"):
        np.testing.assert_allclose(gpu[k], exp[k], rtol=0, atol=1e-9, err_msg=k)
    assert abs(gpu["total_gap"] - 10.0) < 1e-9 and abs(gpu["two_fold"].sum() - gpu["total_gap"]) < 1e-9
    np.testing.assert_allclose(gpu["residuals_b"], exp["resid_b"], atol=1e-9)
    # 5 rows per group: resampled coefficients of the two groups can agree to ~6 digits, so a detailed term
    # xbar_j (beta_a_j - beta*_j) cancels that far below its operands -- see helpers.relerr(cancel_floor)
    compare(gpu, ref_out, cancel_floor=1e-3)
    assert 0 < gpu["n_ok"] <= 60


def test_weights_fixture_on_gpu(ob, orc, ctx, golden):
    fix = golden["F3"]                                      # weights_test.rs:19-46
    for key, weighted in (("unweighted", False), ("weighted", True)):
        Xa, ya, wa, Xb, yb, wb, norm, n_cont = _fixture_design(fix, weighted)
        des = ob.Design.from_dense(ctx, Xa, ya, wa, Xb, yb, wb, n_cont)
        out = ob.bootstrap(des, 0)                          # bootstrap_reps(0): weights_test.rs:32
        exp = fix["expected"][key]
        assert abs(out["total_gap"] - exp["total_gap"]) < 1e-9
        np.testing.assert_allclose(out["beta_a"], exp["beta_a"], atol=1e-9)
        np.testing.assert_allclose(out["beta_b"], exp["beta_b"], atol=1e-9)
        assert np.all(np.isnan(out["std_err"])) and np.all(out["t_stat"] == 0.0) and out["n_ok"] == 0
        des.close()


@pytest.mark.parametrize("ref", ["A", "B", "pooled", "weighted"])
@pytest.mark.parametrize("weighted", [False, True])
def test_synthetic_wage_yun(ob, orc, ctx, ref, weighted):
    """config-1 shaped: continuous + C(sector 4 levels) + C(region 3 levels), Yun on both, all beta* kinds."""
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(6000, 3, cat_levels=(4, 3), weights=weighted, seed=11)
    Xa, ya, wa, Xb, yb, wb = synth.dense_design(d)
    gpu, ref_out = run_both(ob, orc, ctx, Xa, ya, wa, Xb, yb, wb, 3, REFS[ref], synth.norm_spec(d), reps=150)
    compare(gpu, ref_out)
    assert gpu["n_ok"] == 150


def test_multi_panel_multi_tile(ob, orc, ctx):
    """K=21 (two column tiles), 300 replicates (three panels), many row segments."""
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(60000, 20, seed=3)
    Xa, ya, wa, Xb, yb, wb = synth.dense_design(d)
    gpu, ref_out = run_both(ob, orc, ctx, Xa, ya, wa, Xb, yb, wb, 20, 0, [], reps=300)
    compare(gpu, ref_out)


def test_wide_design_wls_yun(ob, orc, ctx):
    """headline-shaped columns: 44 continuous + 2 x C(4 levels), WLS + Yun (K=51, 11 column tiles), small n."""
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(20000, 44, cat_levels=(4, 4), weights=True, seed=4)
    Xa, ya, wa, Xb, yb, wb = synth.dense_design(d)
    gpu, ref_out = run_both(ob, orc, ctx, Xa, ya, wa, Xb, yb, wb, 44, 2, synth.norm_spec(d), reps=40)
    compare(gpu, ref_out)


def test_failed_replicates_are_dropped_identically(ob, orc, ctx):
    """Rare category: resamples that miss it are singular -> dropped on both sides."""
    rng = np.random.default_rng(2)
    n = 60
    edu = rng.normal(13, 2, n)
    rare = np.zeros(n); rare[[3, 40]] = 1.0      # one rare-level row per group
    grp = np.arange(n) >= 30
    y = 1 + 0.5 * edu + rare + rng.normal(0, 1, n) + grp
    X = np.c_[np.ones(n), edu, rare]
    Xa, ya, Xb, yb = X[grp], y[grp], X[~grp], y[~grp]
    gpu, ref_out = run_both(ob, orc, ctx, Xa, ya, None, Xb, yb, None, 1, 0, [], reps=200)
    assert compare(gpu, ref_out) == 0          # a missing category gives an exactly zero pivot: no ambiguity
    assert 0 < gpu["n_ok"] < 200
    assert set(np.unique(gpu["rep_status"])) == {0, 4}


def test_count_width_and_saturation(ob, orc, ctx):
    """uint8 multiplicities saturate (one row drawn 300 times) -> automatic uint16 rerun, still exact."""
    rng = np.random.default_rng(8)
    n = 400
    X = np.c_[np.ones(n), rng.normal(size=(n, 2))]
    y = X @ [1.0, 2.0, -1.0] + rng.normal(size=n)
    Xa, ya, Xb, yb = X[:200], y[:200], X[200:], y[200:]
    reps = 5
    ia = orc.index_stream(1, reps, 0, 200); ib = orc.index_stream(1, reps, 1, 200)
    ia[2, :] = np.r_[np.full(180, 7), np.arange(20)]      # multiplicity 181 for row 7 (+1 from arange) fits uint8
    ib[3, :] = np.r_[np.full(190, 5), np.arange(10) * 3]  # still < 256
    ia[4, :150] = 9
    ia[4, 150:] = np.arange(50) + 20
    ia[1, :] = 11                                          # 200 draws of one row -> singular replicate
    big = np.tile(np.arange(200), 2)[:200]
    spec = orc.Spec(K=3, n_cont=2)
    for bits in (0, 8, 16):
        des = ob.Design.from_dense(ctx, Xa, ya, None, Xb, yb, None, 2)
        gpu = ob.bootstrap(des, reps, idx_a=ia, idx_b=ib, want_rep=True, count_bits=bits)
        ref_out = orc.run(spec, Xa, ya, None, Xb, yb, None, reps, ia, ib)
        compare(gpu, ref_out)
        des.close()
    # > 255 draws of one row: auto widens, forced 8-bit refuses
    n2 = 600
    X2 = np.c_[np.ones(n2), rng.normal(size=(n2, 1))]
    y2 = X2 @ [1.0, 0.5] + rng.normal(size=n2)
    ia2 = orc.index_stream(2, 3, 0, 300); ib2 = orc.index_stream(2, 3, 1, 300)
    ia2[1, :280] = 4
    des = ob.Design.from_dense(ctx, X2[:300], y2[:300], None, X2[300:], y2[300:], None, 1)
    gpu = ob.bootstrap(des, 3, idx_a=ia2, idx_b=ib2, want_rep=True)
    ref_out = orc.run(orc.Spec(K=2, n_cont=1), X2[:300], y2[:300], None, X2[300:], y2[300:], None, 3, ia2, ib2)
    compare(gpu, ref_out)
    with pytest.raises(ob.OaxacaError) as e:
        ob.bootstrap(des, 3, idx_a=ia2, idx_b=ib2, count_bits=8)
    assert e.value.kind == "Unsupported"
    des.close()


def test_index_stream_multiplicities_are_bit_exact(ob, orc, ctx):
    """north_star's first correctness clause, asserted directly: fed an explicit resample index stream
    (builder.rs:822-827), the multiplicity matrix the bootstrap contracts with equals np.bincount of that stream --
    uint8 and uint16 widths, several panels (> 127 replicates), unsharded and on both halves of a 2-way row split."""
    rng = np.random.default_rng(21)
    na, nb = 5003, 4099                       # not multiples of the 32-row stage: padding rows must stay zero
    X = np.c_[np.ones(na + nb), rng.normal(size=(na + nb, 2))]
    y = rng.normal(size=na + nb)
    des = ob.Design.from_dense(ctx, X[:na], y[:na], None, X[na:], y[na:], None, 2)
    reps = 300                                # three panels of 128 slots (slot 0 = point estimate)
    streams = {0: orc.index_stream(77, reps, 0, na), 1: orc.index_stream(77, reps, 1, nb)}
    streams[0][5, :200] = 17                  # a heavy row: multiplicity ~ 200 still fits uint8
    for g, n in ((0, na), (1, nb)):
        want = np.stack([np.bincount(streams[g][r], minlength=n) for r in range(reps)])
        for bits in (8, 16):
            got, flags = des.debug_counts_from_indices(streams[g], g, count_bits=bits)
            assert flags == 0
            assert got.dtype == np.uint16 and np.array_equal(got, want), (g, bits)
    # saturation is reported, not wrapped: 300 draws of one row at 8 bits
    sat = streams[0][:2].copy(); sat[1, :300] = 3
    got8, f8 = des.debug_counts_from_indices(sat, 0, count_bits=8)
    got16, f16 = des.debug_counts_from_indices(sat, 0, count_bits=16)
    assert f8 & 1 and f16 == 0 and np.array_equal(got16[1], np.bincount(sat[1], minlength=na))
    # out-of-range index is flagged
    bad = streams[1][:1].copy(); bad[0, 0] = nb
    assert des.debug_counts_from_indices(bad, 1)[1] & 2
    des.close()
    # row-sharded: each shard histograms the GLOBAL stream and keeps its own rows
    from oaxaca_blinder_rs_b200 import core
    for world in (2, 4):
        for rank in range(world):
            a0, a1 = core.row_shard_plan(na, world, rank)
            b0, b1 = core.row_shard_plan(nb, world, rank)
            sh = ob.Design.from_dense(ctx, X[:na][a0:a1], y[:na][a0:a1], None, X[na:][b0:b1], y[na:][b0:b1], None, 2)
            sh.set_row_shard(na, nb, world, rank)
            for g, (lo, hi, n) in ((0, (a0, a1, na)), (1, (b0, b1, nb))):
                got, flags = sh.debug_counts_from_indices(streams[g][:130], g)
                want = np.stack([np.bincount(streams[g][r], minlength=n)[lo:hi] for r in range(130)])
                assert flags == 0 and np.array_equal(got, want), (world, rank, g)
            sh.close()


def test_errors_match_reference_variants(ob, ctx):
    rng = np.random.default_rng(0)
    X = np.c_[np.ones(8), rng.normal(size=(8, 1))]
    y = rng.normal(size=8)
    # collinear point estimate -> NalgebraError (ols.rs:164-181)
    Xc = np.c_[np.ones(8), np.arange(8.0), 2 * np.arange(8.0)]
    des = ob.Design.from_dense(ctx, Xc, y, None, Xc, y, None, 2)
    with pytest.raises(ob.OaxacaError) as e:
        ob.bootstrap(des, 5)
    assert e.value.kind == "NalgebraError"
    des.close()
    # n <= k -> InsufficientData (ols.rs:183-209)
    Xw = np.c_[np.ones(3), rng.normal(size=(3, 4))]
    des = ob.Design.from_dense(ctx, Xw, y[:3], None, Xw, y[:3], None, 4)
    with pytest.raises(ob.OaxacaError) as e:
        ob.bootstrap(des, 5)
    assert e.value.kind == "InsufficientData"
    des.close()
    # negative weights -> InvalidGroupVariable (ols.rs:60-66)
    with pytest.raises(ob.OaxacaError) as e:
        ob.Design.from_dense(ctx, X, y, -np.ones(8), X, y, np.ones(8), 1)
    assert e.value.kind == "InvalidGroupVariable"
    # empty group -> InvalidGroupVariable (builder.rs:431-435)
    des = ob.Design.from_dense(ctx, X, y, None, X[:0], y[:0], None, 1)
    with pytest.raises(ob.OaxacaError) as e:
        ob.bootstrap(des, 5)
    assert e.value.kind == "InvalidGroupVariable"
    des.close()


def test_batching_and_sharding_are_bit_identical(ob, orc, ctx):
    """Fixed-order split-n: results do not depend on workspace batching nor on the replicate shard."""
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(30000, 6, cat_levels=(3,), weights=True, seed=21)
    Xa, ya, wa, Xb, yb, wb = synth.dense_design(d)
    norm = [ob.NormVar(m, i) for m, i in synth.norm_spec(d)]
    reps = 300
    des = ob.Design.from_dense(ctx, Xa, ya, wa, Xb, yb, wb, 6)
    full = ob.bootstrap(des, reps, ref_kind=3, norm=norm, seed=99, want_rep=True)
    small = ob.bootstrap(des, reps, ref_kind=3, norm=norm, seed=99, want_rep=True, max_workspace_bytes=24 << 20)
    np.testing.assert_array_equal(full["rep_stats"], small["rep_stats"])
    np.testing.assert_array_equal(full["std_err"], small["std_err"])
    parts = [ob.bootstrap(des, reps, ref_kind=3, norm=norm, seed=99, rep_begin=a, rep_end=b, skip_reduce=True)
             for a, b in ((0, 77), (77, 200), (200, 300))]
    stats = np.concatenate([p["rep_stats"] for p in parts])
    status = np.concatenate([p["rep_status"] for p in parts])
    np.testing.assert_array_equal(stats, full["rep_stats"])
    red = ob.reduce_stats(ctx, stats, status, full["point_stats"])
    for k in ("std_err", "p_value", "ci_lower", "ci_upper", "t_stat"):
        np.testing.assert_array_equal(red[k], full[k])
    # the reduction itself against the oracle's bootstrap_stats
    oref = orc.reduce(stats, status, full["point_stats"])
    assert relerr(red["std_err"], oref["se"]) <= RTOL and red["n_ok"] == oref["n_ok"]
    np.testing.assert_array_equal(red["ci_lower"], oref["ci_lo"])
    np.testing.assert_array_equal(red["ci_upper"], oref["ci_hi"])
    des.close()


def test_native_philox_stream(ob, orc, ctx):
    """The GPU's own resampling stream: exact multinomial shape (sum = n), determinism, and SE agreement
    with the oracle's independent stream within Monte-Carlo error (SE of an SE ~ SE / sqrt(2B))."""
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(20000, 4, cat_levels=(4,), seed=31)
    Xa, ya, wa, Xb, yb, wb = synth.dense_design(d)
    des = ob.Design.from_dense(ctx, Xa, ya, wa, Xb, yb, wb, 4)
    c0 = des.debug_counts(7, 0, 0).astype(np.int64)
    c1 = des.debug_counts(7, 1, 0).astype(np.int64)
    cb = des.debug_counts(7, 0, 1).astype(np.int64)
    assert c0.sum() == len(ya) and c1.sum() == len(ya) and cb.sum() == len(yb)
    assert not np.array_equal(c0, c1)
    np.testing.assert_array_equal(c0, des.debug_counts(7, 0, 0))
    # multinomial(n, 1/n) marginals: mean 1, variance 1 - 1/n; P(0) ~ e^-1
    allc = np.concatenate([des.debug_counts(7, r, 0).astype(np.int64) for r in range(20)])
    assert abs(allc.mean() - 1.0) < 1e-12
    assert abs(allc.var() - 1.0) < 0.02 and abs((allc == 0).mean() - np.exp(-1)) < 0.005
    # full marginal law: each count ~ Binomial(n, 1/n); observed frequencies of 0..7 within 5 sigma
    from scipy import stats as sps
    n_a, N = len(ya), allc.size
    for k in range(8):
        pk = sps.binom.pmf(k, n_a, 1.0 / n_a)
        assert abs((allc == k).mean() - pk) < 5.0 * np.sqrt(pk * (1 - pk) / N) + 1e-9, (k, (allc == k).mean(), pk)
    # replicates are independent of each other and of the row: correlation of two count vectors ~ N(0, 1/n)
    assert abs(np.corrcoef(c0, c1)[0, 1]) < 5.0 / np.sqrt(n_a)
    # streams do not depend on which replicate octet / batch position a replicate is generated in
    np.testing.assert_array_equal(des.debug_counts(7, 13, 0), des.debug_counts(7, 13, 0))
    B = 800
    norm = [ob.NormVar(m, i) for m, i in synth.norm_spec(d)]
    g1 = ob.bootstrap(des, B, ref_kind=1, norm=norm, seed=1234)
    g2 = ob.bootstrap(des, B, ref_kind=1, norm=norm, seed=1234)
    np.testing.assert_array_equal(g1["std_err"], g2["std_err"])          # deterministic
    g3 = ob.bootstrap(des, B, ref_kind=1, norm=norm, seed=77)
    assert not np.array_equal(g1["std_err"], g3["std_err"])
    spec = orc.Spec(K=des.K, n_cont=4, ref_kind=1, norm=[orc.NormVar(m, i) for m, i in synth.norm_spec(d)])
    o = orc.run(spec, Xa, ya, wa, Xb, yb, wb, B, None, None, seed=5, nthreads=8, precise=False)
    assert g1["n_ok"] == B
    z = (g1["std_err"] - o["se"]) / (o["se"] / np.sqrt(2 * B) * np.sqrt(2) + 1e-300)
    big = np.abs(o["se"]) > 1e-12
    assert np.all(np.abs(z[big]) < 5.0), z[big]
    des.close()


def test_pack_matches_host_design(ob, ctx):
    """ob_design_pack (CUDA) == prepare_data restated on the host, bit for bit, rows in frame order."""
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(10007, 5, cat_levels=(4, 2, 3), weights=True, seed=77)
    d["group"][::17] = 2                       # rows of a third group are ignored (builder.rs:73-94)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    Xa, ya, wa, Xb, yb, wb = des.download()
    eXa, eya, ewa, eXb, eyb, ewb = synth.dense_design(d)
    for got, exp in ((Xa, eXa), (ya, eya), (wa, ewa), (Xb, eXb), (yb, eyb), (wb, ewb)):
        np.testing.assert_array_equal(got, exp)
    assert des.K == 1 + 5 + 3 + 1 + 2 and des.n_cont == 5
    des.close()


@pytest.mark.parametrize("tau", [0.1, 0.5, 0.9])
def test_rif_prestep(ob, orc, ctx, tau):
    """decompose_quantile pre-step on the device vs math/rif.rs restated in the oracle."""
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(50001, 2, seed=5)
    Xa, ya, wa, Xb, yb, wb = synth.dense_design(d)
    des = ob.Design.from_dense(ctx, Xa, ya, None, Xb, yb, None, 2)
    des.apply_rif(tau)
    _, ga, _, _, gb, _ = des.download()
    assert relerr(ga, orc.rif(ya, tau)) <= RTOL and relerr(gb, orc.rif(yb, tau)) <= RTOL
    des.close()


def test_outcome_refresh_equals_repack(ob, ctx):
    """ob_design_update_outcome (callers re-running on the same X with another y: jmp.rs:44-106,
    engine/src/analysis.rs:871-914): bit-identical to packing the frame again with the new outcome."""
    from oaxaca_blinder_rs_b200 import synth
    for weighted in (False, True):
        d = synth.make_wage(40_003, 5, cat_levels=(3,), weights=weighted, seed=13)
        norm = [ob.NormVar(m, i) for m, i in synth.norm_spec(d)]
        y2 = d["outcome"] * 1.5 - np.sin(d["cont"][0])
        des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
        first = ob.bootstrap(des, 60, ref_kind=2, norm=norm, seed=4, want_rep=True)
        des.update_outcome(y2)
        refreshed = ob.bootstrap(des, 60, ref_kind=2, norm=norm, seed=4, want_rep=True)
        fresh_des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], y2, d["weights"], d["group"])
        fresh = ob.bootstrap(fresh_des, 60, ref_kind=2, norm=norm, seed=4, want_rep=True)
        for k in ("point_stats", "rep_stats", "std_err", "ci_lower", "ci_upper", "residuals_b"):
            np.testing.assert_array_equal(refreshed[k], fresh[k])
        assert not np.array_equal(first["point_stats"], refreshed["point_stats"])
        # back to the original outcome: identical to the first run
        des.update_outcome(d["outcome"])
        again = ob.bootstrap(des, 60, ref_kind=2, norm=norm, seed=4, want_rep=True)
        np.testing.assert_array_equal(again["rep_stats"], first["rep_stats"])
        with pytest.raises(ob.OaxacaError):
            des.update_outcome(y2[:-1])
        des.close(); fresh_des.close()
    # dense designs: the frame is [y_a ; y_b]
    d = synth.make_wage(5_000, 2, seed=2)
    Xa, ya, wa, Xb, yb, wb = synth.dense_design(d)
    des = ob.Design.from_dense(ctx, Xa, ya, wa, Xb, yb, wb, 2)
    des.update_outcome(np.concatenate([2 * ya, 2 * yb]))
    got = des.download()
    np.testing.assert_array_equal(got[1], 2 * ya)
    np.testing.assert_array_equal(got[4], 2 * yb)
    des.close()


def test_ingest_matches_host_cleaning(ob, ctx):
    """ob_ingest_begin/finish (device-side clean_dataframe + create_dummies_manual + split_groups coding,
    builder.rs:760-806, :61-102) against the same steps restated with numpy/python on the host: nulls in every kind
    of column, an unsorted dictionary, a level that only occurs in dropped rows, a third group, unused entries."""
    from oaxaca_blinder_rs_b200 import core
    rng = np.random.default_rng(17)
    n = 50_021
    x1 = rng.normal(size=n); x2 = rng.normal(size=n)
    y = 1 + x1 - 0.5 * x2 + rng.normal(size=n)
    w = rng.uniform(0.5, 2.0, n)
    sector_dict = ["tech", "agri", "zzz_unused", "serv", "manu", "only_in_null_rows"]       # arbitrary order
    sector = rng.choice([0, 1, 3, 4], size=n).astype(np.int32)
    group_dict = ["M", "other", "F"]
    group = rng.choice([0, 1, 2], size=n, p=[0.45, 0.1, 0.45]).astype(np.int32)
    # nulls
    x1[rng.choice(n, 300, replace=False)] = np.nan
    y[rng.choice(n, 200, replace=False)] = np.nan
    w[rng.choice(n, 100, replace=False)] = np.nan
    sector[rng.choice(n, 250, replace=False)] = -1
    group[rng.choice(n, 150, replace=False)] = -1
    rare = rng.choice(n, 40, replace=False)
    sector[rare] = 5; x2[rare] = np.nan                                       # that level exists only in dropped rows
    des, meta = core.ingest(ctx, [x1, x2], [(sector, sector_dict)], y, w, (group, group_dict), reference_group="F")
    # host restatement
    ok = ~(np.isnan(x1) | np.isnan(x2) | np.isnan(y) | np.isnan(w)) & (sector >= 0) & (group >= 0)
    assert meta["rows_kept"] == int(ok.sum())
    lv = sorted({sector_dict[c] for c in sector[ok]})
    assert meta["levels"] == [lv] and lv == ["agri", "manu", "serv", "tech"] and meta["group_a"] == "M"
    sec_names = np.array(sector_dict)[np.where(sector >= 0, sector, 0)]
    cols = [np.ones(n), x1, x2] + [(sec_names == s).astype(float) for s in lv[1:]]
    X = np.stack(cols, 1)
    gname = np.array(group_dict)[np.where(group >= 0, group, 0)]
    A, B = ok & (gname == "M"), ok & (gname == "F")
    Xa, ya, wa, Xb, yb, wb = des.download()
    np.testing.assert_array_equal(Xa, X[A]); np.testing.assert_array_equal(Xb, X[B])
    np.testing.assert_array_equal(ya, y[A]); np.testing.assert_array_equal(yb, y[B])
    np.testing.assert_array_equal(wa, w[A]); np.testing.assert_array_equal(wb, w[B])
    # and the bootstrap on it equals the one on a host-cleaned, host-coded frame packed the ordinary way
    code = {s: i for i, s in enumerate(lv)}
    keep = np.flatnonzero(ok)
    des2 = ob.Design.pack(ctx, [x1[keep], x2[keep]], [np.array([code[s] for s in sec_names[keep]], dtype=np.int32)], [len(lv)],
                          y[keep], w[keep], np.where(gname[keep] == "M", 0, np.where(gname[keep] == "F", 1, 2)).astype(np.uint8))
    r1 = ob.bootstrap(des, 50, seed=3, want_rep=True)
    r2 = ob.bootstrap(des2, 50, seed=3, want_rep=True)
    np.testing.assert_array_equal(r1["rep_stats"], r2["rep_stats"])
    des.close(); des2.close()
    # fewer than two groups after cleaning -> InvalidGroupVariable (builder.rs:67-71)
    with pytest.raises(ob.OaxacaError) as e:
        core.ingest(ctx, [x1, x2], [], y, None, (np.where(group == 2, 2, -1).astype(np.int32), group_dict), reference_group="F")
    assert e.value.kind == "InvalidGroupVariable"


def test_reduction_beyond_16384_replicates(ob, orc, ctx):
    """bootstrap_stats (inference.rs:4-34) over more replicates than the shared-memory sort holds: the per-statistic
    sort then runs in a global scratch row; same numbers as the oracle's reduction."""
    from oaxaca_blinder_rs_b200 import core
    rng = np.random.default_rng(3)
    reps, S = 40_000, 9
    stats = rng.normal(size=(reps, S)) * np.arange(1, S + 1) + np.linspace(-1, 1, S)
    status = (rng.random(reps) < 0.01).astype(np.int32) * 4            # ~1 % failed replicates (NalgebraError)
    stats[status != 0] = np.nan
    point = stats[0].copy(); point[np.isnan(point)] = 0.0
    got = core.reduce_stats(ctx, np.nan_to_num(stats), status, point)
    ref = orc.reduce(np.nan_to_num(stats), status, point)
    assert got["n_ok"] == ref["n_ok"] == int((status == 0).sum())
    for k, rk in (("std_err", "se"), ("p_value", "p"), ("ci_lower", "ci_lo"), ("ci_upper", "ci_hi"), ("t_stat", "t")):
        assert relerr(got[k], ref[rk]) <= RTOL, k


def test_rif_quantile_sweep_on_one_design(ob, orc, ctx):
    """BASELINE configs[3]: tau = 0.1 / 0.5 / 0.9 on the same frame.  apply_rif always transforms the RAW outcome, so one
    packed design serves the whole sweep; update_outcome installs a new raw outcome."""
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(30_001, 3, weights=True, seed=8)
    Xa, ya, wa, Xb, yb, wb = synth.dense_design(d)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    for tau in (0.1, 0.5, 0.9, 0.5):
        des.apply_rif(tau)
        _, ga, _, _, gb, _ = des.download()
        assert relerr(ga, orc.rif(ya, tau)) <= RTOL and relerr(gb, orc.rif(yb, tau)) <= RTOL, tau
    y2 = 2.0 * d["outcome"] + 1.0
    des.update_outcome(y2)
    des.apply_rif(0.9)
    _, ga, _, _, gb, _ = des.download()
    assert relerr(ga, orc.rif(2.0 * ya + 1.0, 0.9)) <= RTOL
    des.close()


@pytest.mark.parametrize("n_cont,ref", [(100, "pooled"), (120, "A"), (170, "pooled"), (230, "weighted")])
def test_wide_designs_beyond_90_columns(ob, orc, ctx, n_cont, ref):
    """No K <= 90 limit any more: the Gram kernel takes any row stride (two ring stages once three no longer fit shared
    memory), the solve keeps one system at a time in shared memory (global scratch beyond ~160 columns)."""
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(6_000, n_cont, cat_levels=(3,), weights=True, seed=n_cont)
    Xa, ya, wa, Xb, yb, wb = synth.dense_design(d)
    norm = synth.norm_spec(d)
    assert Xa.shape[1] == n_cont + 3
    gpu, ref_out = run_both(ob, orc, ctx, Xa, ya, wa, Xb, yb, wb, n_cont, REFS[ref], norm, reps=12)
    compare(gpu, ref_out)
    assert gpu["n_ok"] == 12


def test_design_wider_than_the_kernels_is_refused(ob, ctx):
    X = np.ones((50, 300))
    with pytest.raises(ob.OaxacaError) as e:
        ob.Design.from_dense(ctx, X, np.ones(50), None, X, np.ones(50), None, 10)
    assert e.value.kind == "Unsupported"

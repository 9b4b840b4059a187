"""Heckman two-step replicate (estimation.rs:114-269, heckman.rs:38-108, math/probit.rs:25-175,
builder.rs:464-534) on the GPU vs the oracle (oracle/ob_oracle_heckman.c, itself pinned against the numpy/scipy golden
fixture), through the C ABI, under an explicit resample index stream: probit coefficients, augmented OLS coefficients,
decomposition incl. the IMR rows and detailed_selection, SEs and CIs within 1e-10."""
import json
import os

import numpy as np
import pytest

from helpers import relerr

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def make_selection_frame(n, n_x, seed, sel_shift=0.4):
    rng = np.random.default_rng(seed)
    grp = rng.integers(0, 2, n).astype(np.uint8)
    z1, z2 = rng.normal(size=n), rng.normal(size=n)
    xs = [z1 + 0.5 * rng.normal(size=n) + 0.3 * (grp == 0)] + [rng.normal(size=n) for _ in range(n_x - 1)]
    u = rng.normal(size=n)
    e = 0.8 * u + 0.6 * rng.normal(size=n)
    s = (sel_shift + 0.5 * z1 - 0.3 * z2 + 0.2 * (grp == 0) + u > 0).astype(np.float64)
    cat = rng.integers(0, 3, n).astype(np.int32)
    y = 1.0 + sum((0.5 + 0.1 * j) * x for j, x in enumerate(xs)) + 0.3 * cat + 0.4 * (grp == 0) + e
    return dict(group=grp, cont=xs, cat=cat, y=y, s=s, z=[z1, z2])


def dense(fr):
    n = len(fr["y"])
    X = np.c_[np.ones(n), np.stack(fr["cont"], 1), (fr["cat"] == 1).astype(float), (fr["cat"] == 2).astype(float)]
    Z = np.c_[np.ones(n), np.stack(fr["z"], 1)]
    A, B = fr["group"] == 0, fr["group"] == 1
    return (X[A], fr["y"][A], Z[A], fr["s"][A]), (X[B], fr["y"][B], Z[B], fr["s"][B])


@pytest.mark.parametrize("ref,n,n_x,reps", [(1, 6_000, 2, 150), (0, 20_000, 3, 40), (3, 6_000, 2, 150), (1, 3_000, 20, 30)])
def test_heckman_matches_oracle(orc, ref, n, n_x, reps):
    import oaxaca_blinder_rs_b200 as ob
    fr = make_selection_frame(n, n_x, seed=100 + n_x)
    (Xa, ya, Za, sa), (Xb, yb, Zb, sb) = dense(fr)
    ia, ib = orc.index_stream(3, reps, 0, len(ya)), orc.index_stream(3, reps, 1, len(yb))
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, fr["cont"], [fr["cat"]], [3], fr["y"], None, fr["group"])
    des.attach_selection(fr["s"], fr["z"])
    assert des.selection_cols == 3
    gpu = ob.bootstrap(des, reps, ref_kind=ref, idx_a=ia, idx_b=ib, want_rep=True, max_workspace_bytes=200_000_000 if n == 20_000 else 0)
    des.close(); ctx.close()
    o = orc.heckman_run(ref, Xa, ya, Za, sa, Xb, yb, Zb, sb, reps, ia, ib, nthreads=8)
    K = Xa.shape[1]
    assert gpu["S"] == o["S"] == 5 + 2 * (K + 1) + 3
    assert relerr(gpu["point_stats"], o["point_stats"]) <= RTOL
    assert relerr(gpu["beta_a"], o["beta_a"]) <= RTOL and relerr(gpu["beta_b"], o["beta_b"]) <= RTOL
    assert relerr(gpu["sel_gamma_a"], o["gamma_a"]) <= RTOL and relerr(gpu["sel_gamma_b"], o["gamma_b"]) <= RTOL
    assert abs(gpu["total_gap"] - o["total_gap"]) <= RTOL * abs(o["total_gap"])
    assert np.array_equal(gpu["rep_status"], o["rep_status"]) and gpu["n_ok"] == o["n_ok"] == reps
    assert relerr(gpu["rep_stats"], o["rep_stats"]) <= RTOL
    assert relerr(gpu["std_err"], o["se"]) <= RTOL and relerr(gpu["ci_lower"], o["ci_lo"]) <= RTOL and relerr(gpu["ci_upper"], o["ci_hi"]) <= RTOL
    assert gpu["det_selection"].shape == (3,) and gpu["det_expl"].shape == (K + 1,)
    assert np.all(gpu["residuals_b"] == 0.0)                     # estimation.rs:156-157
    # the IMR row is there and matters (tests/heckman_test.rs asserts its presence)
    assert abs(gpu["beta_b"][-1]) > 0.05


def test_heckman_golden_fixture_on_gpu():
    """tests/golden/heckman_fixture.json (numpy / scipy, independent of the oracle) straight against the GPU."""
    import oaxaca_blinder_rs_b200 as ob
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "heckman_fixture.json")))
    c = {k: np.array(v, float) for k, v in fx["columns"].items()}
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, [c["x"], c["d"]], [], [], c["y"], None, c["group"].astype(np.uint8))
    des.attach_selection(c["s"], [c["z"]])
    for name, ref in (("A", 0), ("B", 1), ("weighted", 3)):
        out = ob.bootstrap(des, 0, ref_kind=ref)
        exp = fx["expected"][name]
        np.testing.assert_allclose(out["point_stats"], exp["stats"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(out["beta_a"], exp["beta_a"], rtol=1e-10)
        np.testing.assert_allclose(out["sel_gamma_b"], exp["gamma_b"], rtol=1e-10)
        assert abs(out["total_gap"] - exp["total_gap"]) < 1e-12
    with pytest.raises(ob.OaxacaError) as e:                      # Pooled: K vs K+1 coefficient vectors in the reference
        ob.bootstrap(des, 0, ref_kind=ob.REF_POOLED)
    assert e.value.kind == "Unsupported"
    des.close(); ctx.close()


def test_heckman_native_stream_and_failures():
    """Native Philox stream: deterministic, batching-independent; a replicate whose resample holds too few selected
    rows fails (InsufficientData / NalgebraError) and is dropped like any failed replicate."""
    import oaxaca_blinder_rs_b200 as ob
    fr = make_selection_frame(4_000, 2, seed=5)
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, fr["cont"], [fr["cat"]], [3], fr["y"], None, fr["group"])
    des.attach_selection(fr["s"], fr["z"])
    a = ob.bootstrap(des, 300, ref_kind=1, seed=11, want_rep=True)
    b = ob.bootstrap(des, 300, ref_kind=1, seed=11, want_rep=True, max_workspace_bytes=30_000_000)
    assert a["n_ok"] == 300 and np.array_equal(a["rep_stats"], b["rep_stats"]) and np.array_equal(a["std_err"], b["std_err"])
    assert np.all(np.isfinite(a["std_err"])) and np.all(a["std_err"][:2] > 0)
    des.close()
    tiny = make_selection_frame(60, 2, seed=6, sel_shift=-1.6)     # few selected rows: many resamples cannot fit K + 1 = 6 columns
    des = ob.Design.pack(ctx, tiny["cont"], [tiny["cat"]], [3], tiny["y"], None, tiny["group"])
    des.attach_selection(tiny["s"], tiny["z"])
    try:
        out = ob.bootstrap(des, 200, ref_kind=1, seed=2, want_rep=True)
        assert out["n_ok"] < 200 and set(np.unique(out["rep_status"])) <= {0, 3, 4, 6}
        assert np.all(np.isnan(out["rep_stats"][out["rep_status"] != 0]))
    except ob.OaxacaError as e:                                    # the point estimate itself may be infeasible on 60 rows
        assert e.kind in ("InsufficientData", "NalgebraError", "InvalidGroupVariable")
    des.close(); ctx.close()


def test_heckman_argument_checks():
    import oaxaca_blinder_rs_b200 as ob
    fr = make_selection_frame(2_000, 2, seed=7)
    ctx = ob.Context(0)
    w = np.ones(2_000)
    des = ob.Design.pack(ctx, fr["cont"], [fr["cat"]], [3], fr["y"], w, fr["group"])
    with pytest.raises(ob.OaxacaError) as e:
        des.attach_selection(fr["s"], fr["z"])
    assert e.value.kind == "Unsupported"                          # weights
    des.close()
    des = ob.Design.pack(ctx, fr["cont"], [fr["cat"]], [3], fr["y"], None, fr["group"])
    with pytest.raises(ob.OaxacaError):
        des.attach_selection(fr["s"][:-1], [z[:-1] for z in fr["z"]])      # wrong length
    s_nan = fr["s"].copy(); s_nan[3] = np.nan
    with pytest.raises(ob.OaxacaError) as e:
        des.attach_selection(s_nan, fr["z"])
    assert e.value.kind == "InvalidGroupVariable" and des.selection_cols == 0
    with pytest.raises(ob.OaxacaError) as e:
        des.attach_selection(fr["s"], [fr["z"][0]] * 8)            # too many selection predictors
    assert e.value.kind == "Unsupported"
    # without a selection equation the design still runs the ordinary path
    out = ob.bootstrap(des, 8, seed=1)
    assert out["S"] == 5 + 2 * des.K
    des.close(); ctx.close()


@pytest.mark.parametrize("world", [2, 3])
def test_heckman_replicate_sharded_is_bit_identical(world):
    """Mode R inside the library carries the Heckman replicate rows like any others: every rank returns the one-GPU result."""
    import threading
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import core
    fr = make_selection_frame(5_000, 2, seed=21)
    reps = 131
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, fr["cont"], [fr["cat"]], [3], fr["y"], None, fr["group"])
    des.attach_selection(fr["s"], fr["z"])
    one = ob.bootstrap(des, reps, ref_kind=3, seed=4, want_rep=True)
    des.close(); ctx.close()
    grp = core.LocalGroup(world)
    outs, errs = [None] * world, [None] * world

    def work(r):
        try:
            c = ob.Context(0)
            c.init_local(grp, r)
            dd = ob.Design.pack(c, fr["cont"], [fr["cat"]], [3], fr["y"], None, fr["group"])
            dd.attach_selection(fr["s"], fr["z"])
            outs[r] = ob.bootstrap(dd, reps, ref_kind=3, seed=4, want_rep=True, shard_replicates=True)
            dd.close(); c.close()
        except Exception as e:  # noqa: BLE001
            errs[r] = e
    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert all(e is None for e in errs), errs
    same = lambda a, b: np.array_equal(np.nan_to_num(a, nan=-7.0), np.nan_to_num(b, nan=-7.0))   # noqa: E731
    for o in outs:
        for k in ("point_stats", "rep_stats", "rep_status", "rep_beta_a", "std_err", "ci_lower", "ci_upper", "p_value", "sel_gamma_a"):
            assert same(o[k], one[k]), k

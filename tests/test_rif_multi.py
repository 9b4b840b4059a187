"""Quantile sweep in one pass (ob_design_apply_rif_multi; decompose_quantile builder.rs:711-757 called per tau): the
three RIF outcomes ride as three x.y column sets on ONE Gram pass, the Gram is factored once per replicate and solved
for three right-hand sides.  Every quantile's results must equal the single-quantile path BIT FOR BIT (same resamples
under the same seed; each Gram column is an independent fixed-order sum), and the oracle within 1e-10."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TAUS = (0.1, 0.5, 0.9)
PER_OUTCOME = ("point_stats", "std_err", "p_value", "ci_lower", "ci_upper", "t_stat", "beta_star", "beta_a", "beta_b", "residuals_b")


def _same(a, b):
    return np.array_equal(np.nan_to_num(np.asarray(a), nan=-7.0), np.nan_to_num(np.asarray(b), nan=-7.0))


@pytest.mark.parametrize("n_cont,cats,weighted,ref_kind", [(4, (3,), False, 0), (50, (), True, 2), (30, (), False, 3), (6, (4, 3), True, 1)])
def test_multi_quantile_pass_equals_single_quantile_runs(n_cont, cats, weighted, ref_kind):
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(30_000, n_cont, cat_levels=cats, weights=weighted, seed=12)
    norm = [ob.NormVar(m, i) for m, i in synth.norm_spec(d)]
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    kw = dict(ref_kind=ref_kind, norm=norm, seed=23, want_rep=True)
    singles = []
    for tau in TAUS:
        des.apply_rif(tau)
        singles.append(ob.bootstrap(des, 140, **kw))
    des.apply_rif_multi(TAUS)
    assert des.n_outcomes == 3
    multi = ob.bootstrap(des, 140, **kw)
    assert multi["n_outcomes"] == 3 and multi["point_stats"].shape == (3, multi["S"]) and multi["rep_stats"].shape == (140, 3, multi["S"])
    for t, one in enumerate(singles):
        for k in PER_OUTCOME:
            assert _same(multi[k][t], one[k]), (t, k)
        assert _same(multi["rep_stats"][:, t], one["rep_stats"]) and _same(multi["rep_beta_a"][:, t], one["rep_beta_a"])
        assert multi["total_gap_multi"][t] == one["total_gap"]
        assert _same(multi["rep_status"], one["rep_status"]) and multi["n_ok"] == one["n_ok"]
    assert _same(multi["xa_mean"], singles[0]["xa_mean"]) and multi["total_gap"] == singles[0]["total_gap"]
    # back to one quantile on the widened design, and to the raw outcome
    des.apply_rif(0.5)
    assert des.n_outcomes == 1
    again = ob.bootstrap(des, 140, **kw)
    for k in PER_OUTCOME + ("rep_stats",):
        assert _same(again[k], singles[1][k]), k
    des.update_outcome(d["outcome"])
    raw = ob.bootstrap(des, 20, **kw)
    fresh = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    ref = ob.bootstrap(fresh, 20, **kw)
    assert _same(raw["rep_stats"], ref["rep_stats"]) and _same(raw["std_err"], ref["std_err"])
    fresh.close(); des.close(); ctx.close()


def test_multi_quantile_pass_vs_oracle(orc):
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import synth
    from helpers import relerr
    d = synth.make_wage(20_000, 5, cat_levels=(3,), weights=True, seed=4)
    norm = synth.norm_spec(d)
    Xa, ya, wa, Xb, yb, wb = synth.dense_design(d)
    reps = 30
    ia, ib = orc.index_stream(6, reps, 0, len(ya)), orc.index_stream(6, reps, 1, len(yb))
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    des.apply_rif_multi(TAUS)
    gpu = ob.bootstrap(des, reps, ref_kind=ob.REF_POOLED, norm=[ob.NormVar(m, i) for m, i in norm], idx_a=ia, idx_b=ib, want_rep=True)
    des.close(); ctx.close()
    spec = orc.Spec(K=Xa.shape[1], n_cont=5, ref_kind=orc.REF_POOLED, norm=[orc.NormVar(m, i) for m, i in norm])
    for t, tau in enumerate(TAUS):
        ref = orc.run(spec, Xa, orc.rif(ya, tau), wa, Xb, orc.rif(yb, tau), wb, reps, ia, ib, nthreads=8, precise=True)
        assert np.array_equal(gpu["rep_status"], ref["rep_status"])
        for k_gpu, k_ref in (("point_stats", "stats"), ("beta_star", "beta_star"), ("beta_a", "beta_a")):
            assert relerr(gpu[k_gpu][t], ref["point"][k_ref]) <= 1e-10, (tau, k_gpu)
        assert relerr(gpu["rep_stats"][:, t], ref["rep_stats"]) <= 1e-10, tau
        assert relerr(gpu["std_err"][t], ref["se"]) <= 1e-10 and relerr(gpu["ci_lower"][t], ref["ci_lo"]) <= 1e-10
        assert abs(gpu["total_gap_multi"][t] - ref["point"]["total_gap"]) <= 1e-10 * abs(ref["point"]["total_gap"])


def test_multi_quantile_pass_replicate_sharded():
    """Mode R inside the library carries the [T x S] statistics rows unchanged."""
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import core, synth
    d = synth.make_wage(25_000, 3, cat_levels=(), weights=False, seed=7)
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    des.apply_rif_multi(TAUS)
    one = ob.bootstrap(des, 90, seed=3, want_rep=True)
    des.close(); ctx.close()
    world = 3
    grp = core.LocalGroup(world)
    outs, errs = [None] * world, [None] * world

    def work(r):
        try:
            c = ob.Context(0)
            c.init_local(grp, r)
            dd = ob.Design.pack(c, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
            dd.apply_rif_multi(TAUS)
            outs[r] = ob.bootstrap(dd, 90, seed=3, want_rep=True, shard_replicates=True)
            dd.close(); c.close()
        except Exception as e:  # noqa: BLE001
            errs[r] = e
    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert all(e is None for e in errs), errs
    for o in outs:
        for k in PER_OUTCOME + ("rep_stats", "rep_beta_b"):
            assert _same(o[k], one[k]), k

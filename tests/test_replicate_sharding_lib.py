"""Mode R inside the library (ob_boot_opts.shard_replicates): every rank holds the whole design, the
library shards the global replicate ids, all-gathers the statistics device to device over the context's communicator
and reduces on every rank.  Results must be BIT-IDENTICAL to one GPU on every rank.

GPU part runs on one device: `world` contexts of one process (threads) joined by the in-process communicator -- the
same code path as NCCL except for the transport (tests/test_row_sharding_nccl.py covers NCCL on >= 2 GPUs)."""
import threading

import numpy as np
import pytest


def test_replicate_shard_tiles_the_replicates():
    from oaxaca_blinder_rs_b200 import core, distributed as obd, _native
    _native.build()
    for reps in (0, 1, 2, 7, 250, 2000, 10_000):
        for world in (1, 2, 3, 4, 7, 8, 64):
            cuts = [core.replicate_shard(reps, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == reps
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
            assert cuts == [obd.shard_range(r, world, reps) for r in range(world)]     # the host-gather path cuts the same
    for bad in ((10, 0, 0), (10, 2, 2), (-1, 2, 0)):
        with pytest.raises(core.OaxacaError):
            core.replicate_shard(*bad)


def _same(a, b):
    return np.array_equal(np.nan_to_num(np.asarray(a), nan=-7.0), np.nan_to_num(np.asarray(b), nan=-7.0))


def _run_world(d, world, reps, **kw):
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import core
    grp = core.LocalGroup(world)
    outs, errs = [None] * world, [None] * world

    def work(r):
        try:
            ctx = ob.Context(0)
            ctx.init_local(grp, r)
            des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
            outs[r] = ob.bootstrap(des, reps, shard_replicates=True, want_rep=True, **kw)
            des.close()
            ctx.close()
        except Exception as e:  # noqa: BLE001
            errs[r] = e
    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert all(e is None for e in errs), errs
    return outs


KEYS = ("point_stats", "rep_stats", "rep_status", "rep_beta_a", "rep_beta_b", "std_err", "ci_lower", "ci_upper", "p_value",
        "t_stat", "xa_mean", "xb_mean", "beta_star", "residuals_b")


@pytest.mark.gpu
@pytest.mark.parametrize("world,reps", [(2, 300), (3, 301), (4, 2), (8, 130)])
def test_library_replicate_sharding_is_bit_identical(world, reps):
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(40_000, 4, cat_levels=(4,), weights=True, seed=21)
    norm = [ob.NormVar(m, i) for m, i in synth.norm_spec(d)]
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    one = ob.bootstrap(des, reps, ref_kind=ob.REF_WEIGHTED, norm=norm, seed=99, want_rep=True)
    des.close(); ctx.close()
    outs = _run_world(d, world, reps, ref_kind=ob.REF_WEIGHTED, norm=norm, seed=99)
    for o in outs:
        assert o["n_ok"] == one["n_ok"] == reps and o["total_gap"] == one["total_gap"]
        assert o["rep_stats"].shape == (reps, one["S"])
        for k in KEYS:
            assert _same(o[k], one[k]), k
        assert o["timings_ms"]["comm"] > 0.0


@pytest.mark.gpu
def test_library_replicate_sharding_index_stream_and_failures(orc):
    """Explicit index stream (the full stream on every rank) with replicates that fail (a dummy level absent from a
    resample): the drop set, the statistics and the reduction equal the one-GPU run and the oracle."""
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import synth
    from helpers import relerr
    d = synth.make_wage(600, 2, cat_levels=(4,), weights=False, seed=3)
    for g in (0, 1):                                   # level 3 survives in two rows per group: many resamples lose it
        rare = np.flatnonzero((d["cat_codes"][0] == 3) & (d["group"] == g))
        d["cat_codes"][0][rare[2:]] = 0
    Xa, ya, wa, Xb, yb, wb = synth.dense_design(d)
    reps = 150
    ia, ib = orc.index_stream(8, reps, 0, len(ya)), orc.index_stream(8, reps, 1, len(yb))
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    one = ob.bootstrap(des, reps, idx_a=ia, idx_b=ib, want_rep=True)
    des.close(); ctx.close()
    assert 0 < one["n_ok"] < reps
    outs = _run_world(d, 3, reps, idx_a=ia, idx_b=ib)
    for o in outs:
        assert o["n_ok"] == one["n_ok"]
        for k in KEYS:
            assert _same(o[k], one[k]), k
    spec = orc.Spec(K=Xa.shape[1], n_cont=2)
    ref = orc.run(spec, Xa, ya, None, Xb, yb, None, reps, ia, ib, nthreads=4, precise=True)
    well = ref["rep_min_pivot"] >= 1e-9
    assert np.array_equal(outs[0]["rep_status"][well], ref["rep_status"][well])
    assert relerr(outs[0]["rep_stats"][well], ref["rep_stats"][well]) <= 1e-10


@pytest.mark.gpu
def test_shard_replicates_argument_checks():
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import synth, distributed as obd
    d = synth.make_wage(5_000, 2, seed=1)
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    # no communicator: the flag is a no-op (world of one)
    a = ob.bootstrap(des, 16, seed=4, shard_replicates=True)
    b = ob.bootstrap(des, 16, seed=4)
    assert _same(a["std_err"], b["std_err"])
    des.close()
    # a row shard's communicator shards rows, not replicates
    sh = obd.pack_row_shard(ctx, d, 0, 2)
    with pytest.raises(ob.OaxacaError) as e:
        ob.bootstrap(sh, 8, seed=1, shard_replicates=True)
    assert e.value.kind == "InvalidArgument"
    sh.close(); ctx.close()

"""Row sharding (mode N).

CPU part: ob_row_shard_plan (pure host function of the C ABI) tiles every group contiguously, follows the
fixed summation tree (cuts at multiples of the leaf size, identical for every world size), and
distributed.shard_frame selects exactly those rows -- also checked across 2 gloo ranks.

GPU part (one device is enough): `world` contexts of one process (threads, in-process communicator) each hold
one row shard; statistics must be BIT-IDENTICAL to the unsharded run on the same device, for the native
Philox stream and for an explicit index stream, including when the workspace budget forces several batches.
"""
import os
import sys
import threading

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_version_and_plan_tiles_rows():
    from oaxaca_blinder_rs_b200 import core, _native
    _native.build()
    for n in (0, 1, 31, 32, 33, 1000, 4096, 8191, 8192, 8193, 100_003, 5_000_000, 49_999_871):
        for world in (1, 2, 4, 8, 16, 64):
            cuts = [core.row_shard_plan(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:])), (n, world, cuts)
            assert all(0 <= b <= e <= n for b, e in cuts)
            # cuts of a coarser world are a subset of the cuts of a finer one (aligned subtrees)
            if world > 1:
                coarse = {c[0] for c in (core.row_shard_plan(n, world // 2, r) for r in range(world // 2))}
                assert coarse <= {c[0] for c in cuts}
            # interior cuts are multiples of the pipeline stage (32 rows)
            assert all(b % 32 == 0 or b == n for b, _ in cuts)
    # large groups are spread evenly (within one leaf)
    cuts = [core.row_shard_plan(50_000_000, 8, r) for r in range(8)]
    sizes = [e - b for b, e in cuts]
    assert max(sizes) - min(sizes) <= 50_000_000 // 64 + 64      # within one leaf (leaves are 1/256 of the group)


def test_plan_rejects_bad_worlds():
    from oaxaca_blinder_rs_b200 import core
    for world, rank in ((3, 0), (0, 0), (128, 0), (2, 2), (2, -1)):
        with pytest.raises(core.OaxacaError):
            core.row_shard_plan(1000, world, rank)


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oaxaca_blinder_rs_b200 import distributed as obd, synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    d = synth.make_wage(20_011, 3, cat_levels=(3,), weights=True, seed=5)
    loc = obd.shard_frame(d, rank, world)
    # every rank reports its local group sizes and a checksum of its rows; the union must be the frame
    t = torch.tensor([float((loc["group"] == 0).sum()), float((loc["group"] == 1).sum()), float(loc["outcome"].sum()),
                      float(loc["weights"].sum())], dtype=torch.float64)
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    q.put((rank, [p.tolist() for p in parts], loc["n_a_global"], loc["n_b_global"]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_frame_partitions_the_frame_gloo():
    from oaxaca_blinder_rs_b200 import synth
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = synth.make_wage(20_011, 3, cat_levels=(3,), weights=True, seed=5)
    na, nb = int((d["group"] == 0).sum()), int((d["group"] == 1).sum())
    for rank, parts, nag, nbg in got:
        assert (nag, nbg) == (na, nb)
        assert sum(p[0] for p in parts) == na and sum(p[1] for p in parts) == nb
        assert abs(sum(p[2] for p in parts) - d["outcome"].sum()) < 1e-6
        assert abs(sum(p[3] for p in parts) - d["weights"].sum()) < 1e-6


# ------------------------------------------------------------------------------------------- GPU
def _run_sharded(d, world, reps, norm, ref_kind, seed=None, idx=None, max_ws=0):
    """`world` threads, one context + one row shard each, in-process communicator; returns rank outputs."""
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import core, distributed as obd
    grp = core.LocalGroup(world)
    outs, errs = [None] * world, [None] * world

    def work(r):
        try:
            ctx = ob.Context(0)
            ctx.init_local(grp, r)
            des = obd.pack_row_shard(ctx, d, r, world)
            kw = dict(seed=seed) if idx is None else dict(idx_a=idx[0], idx_b=idx[1])
            outs[r] = ob.bootstrap(des, reps, ref_kind=ref_kind, norm=norm, want_rep=True, max_workspace_bytes=max_ws, **kw)
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


def _same(a, b):
    return np.array_equal(np.nan_to_num(np.asarray(a), nan=-7.0), np.nan_to_num(np.asarray(b), nan=-7.0))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_row_sharded_native_stream_is_bit_identical(world):
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(60_000, 4, cat_levels=(4,), weights=True, seed=11)
    norm = [ob.NormVar(m, i) for m, i in synth.norm_spec(d)]
    reps = 300
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    one = ob.bootstrap(des, reps, ref_kind=ob.REF_POOLED, norm=norm, seed=77, want_rep=True)
    des.close(); ctx.close()
    # small budget -> 1-2 panels per batch on the shards -> 2-3 batches, exercising the per-batch collectives
    outs = _run_sharded(d, world, reps, norm, ob.REF_POOLED, seed=77, max_ws=16_000_000)
    for o in outs:
        assert o["n_ok"] == one["n_ok"] == reps
        for k in ("point_stats", "rep_stats", "rep_beta_a", "rep_beta_b", "std_err", "ci_lower", "ci_upper", "p_value"):
            assert _same(o[k], one[k]), k
        assert o["total_gap"] == one["total_gap"]
    # residuals_b of the shards concatenate to the unsharded vector
    assert _same(np.concatenate([o["residuals_b"] for o in outs]), one["residuals_b"])


@pytest.mark.gpu
def test_row_sharded_batched_and_uneven(orc):
    """Tiny uneven groups (the last ranks hold few or no rows), several panel batches, explicit index stream:
    bit-identical to one GPU and within 1e-10 of the oracle."""
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(3_001, 2, cat_levels=(), weights=False, seed=3)
    Xa, ya, wa, Xb, yb, wb = synth.dense_design(d)
    reps = 260
    ia, ib = orc.index_stream(5, reps, 0, len(ya)), orc.index_stream(5, reps, 1, len(yb))
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    one = ob.bootstrap(des, reps, idx_a=ia, idx_b=ib, want_rep=True)
    des.close(); ctx.close()
    outs = _run_sharded(d, 4, reps, [], ob.REF_GROUP_A, idx=(ia, ib), max_ws=6_000_000)
    for o in outs:
        for k in ("point_stats", "rep_stats", "std_err", "ci_lower", "ci_upper"):
            assert _same(o[k], one[k]), k
    spec = orc.Spec(K=des.K, n_cont=2)
    ref = orc.run(spec, Xa, ya, None, Xb, yb, None, reps, ia, ib, nthreads=4)
    err = np.max(np.abs(outs[0]["rep_stats"] - ref["rep_stats"])) / max(1.0, np.max(np.abs(ref["rep_stats"])))
    assert err <= 1e-10, err      # north_star tolerance: 1e-10 relative


@pytest.mark.gpu
def test_sharded_design_without_communicator_fails_loudly():
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import synth, distributed as obd
    d = synth.make_wage(5_000, 2, seed=1)
    ctx = ob.Context(0)
    des = obd.pack_row_shard(ctx, d, 0, 2)
    with pytest.raises(ob.OaxacaError) as e:
        ob.bootstrap(des, 8, seed=1)
    assert e.value.kind == "NcclError"
    with pytest.raises(ob.OaxacaError):
        des.set_row_shard(des.n_a_global + 5, des.n_b_global, 2, 0)     # local rows no longer match the plan
    des.close(); ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("weighted", [False, True])
def test_allgather_rows_equals_whole_frame_pack(weighted):
    """Mode R upload: each context packs a contiguous frame slice; the gathered design equals the whole-frame pack."""
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import core, synth, distributed as obd
    d = synth.make_wage(30_011, 3, cat_levels=(3,), weights=weighted, seed=9)
    ctx = ob.Context(0)
    whole = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    ref = whole.download()
    ref_run = ob.bootstrap(whole, 40, seed=3, want_rep=True)
    whole.close(); ctx.close()
    world = 4
    grp = core.LocalGroup(world)
    outs, errs, refreshed = [None] * world, [None] * world, [None] * world
    c0 = ob.Context(0)
    w2 = ob.Design.pack(c0, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"] * 2.0 + 1.0, d["weights"], d["group"])
    ref2 = w2.download()
    w2.close(); c0.close()

    def work(r):
        try:
            c = ob.Context(0)
            c.init_local(grp, r)
            full = obd.pack_replicated(c, d, r, world)
            b, e = obd.shard_range(r, world, 40)
            outs[r] = (full.download(), ob.bootstrap(full, 40, seed=3, rep_begin=b, rep_end=e, skip_reduce=True))
            # the frame-row map travelled with the rows: an outcome refresh on the gathered design equals a re-pack
            full.update_outcome(d["outcome"] * 2.0 + 1.0)
            refreshed[r] = full.download()
            full.close(); c.close()
        except Exception as ex:  # noqa: BLE001
            errs[r] = ex
    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert all(e is None for e in errs), errs
    for r, (mats, part) in enumerate(outs):
        for a, b_ in zip(mats, ref):
            if weighted or not np.all(np.isnan(b_)):
                assert np.array_equal(a, b_, equal_nan=True)
        b, e = obd.shard_range(r, world, 40)
        assert _same(part["rep_stats"], ref_run["rep_stats"][b:e])
        assert _same(part["point_stats"], ref_run["point_stats"])
        for a, b_ in zip(refreshed[r], ref2):
            if weighted or not np.all(np.isnan(b_)):
                assert np.array_equal(a, b_, equal_nan=True)


def test_chunked_generator_shards_match_the_whole_frame():
    """synth.make_wage_rows (the generator bench.py uses in mode N: every rank produces only its rows, chunk by chunk):
    the ranks' rows, concatenated per group in rank order, are exactly the world = 1 frame."""
    from oaxaca_blinder_rs_b200 import synth
    kw = dict(n_cont=3, cat_levels=(4,), weights=True, chunk=1 << 14)
    full = synth.make_wage_rows(150_001, **kw)
    for world in (2, 8):
        parts = [synth.make_wage_rows(150_001, rank=r, world=world, **kw) for r in range(world)]
        assert sum(p["n"] for p in parts) == full["n"] == 150_001
        for g in (0, 1):
            sel = full["group"] == g
            for key in ("outcome", "weights"):
                np.testing.assert_array_equal(np.concatenate([p[key][p["group"] == g] for p in parts]), full[key][sel])
            np.testing.assert_array_equal(np.concatenate([p["cont"][2][p["group"] == g] for p in parts]), full["cont"][2][sel])
            np.testing.assert_array_equal(np.concatenate([p["cat_codes"][0][p["group"] == g] for p in parts]), full["cat_codes"][0][sel])
        assert all(p["n_a_global"] == full["n_a_global"] and p["n_b_global"] == full["n_b_global"] for p in parts)


def test_frame_slices_tile_the_frame():
    from oaxaca_blinder_rs_b200 import distributed as obd, synth
    d = synth.make_wage(10_007, 2, cat_levels=(3,), weights=True, seed=3)
    for world in (1, 2, 3, 8):
        sl = [obd.frame_slice(d, r, world) for r in range(world)]
        assert sum(s["n"] for s in sl) == d["n"]
        np.testing.assert_array_equal(np.concatenate([s["outcome"] for s in sl]), d["outcome"])
        np.testing.assert_array_equal(np.concatenate([s["group"] for s in sl]), d["group"])
        np.testing.assert_array_equal(np.concatenate([s["cat_codes"][0] for s in sl]), d["cat_codes"][0])


@pytest.mark.gpu
@pytest.mark.parametrize("world,sort_by_group", [(2, False), (4, False), (4, True), (8, False)])
def test_redistributed_slices_equal_row_shards(world, sort_by_group):
    """ob_design_redistribute_rows: frame slices -> row shards over the communicator.  The shard equals packing exactly
    the plan's rows (download bit-identical), and the sharded bootstrap equals one GPU -- also for a frame sorted by
    group, where whole slices change hands."""
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import core, synth, distributed as obd
    d = synth.make_wage(50_003, 3, cat_levels=(3,), weights=True, seed=13)
    if sort_by_group:
        order = np.argsort(d["group"], kind="stable")
        for k in ("outcome", "weights", "group"):
            d[k] = d[k][order]
        d["cont"] = [c[order] for c in d["cont"]]
        d["cat_codes"] = [c[order] for c in d["cat_codes"]]
    norm = [ob.NormVar(m, i) for m, i in synth.norm_spec(d)]
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    one = ob.bootstrap(des, 150, ref_kind=ob.REF_POOLED, norm=norm, seed=41, want_rep=True)
    des.close(); ctx.close()
    grp = core.LocalGroup(world)
    outs, errs = [None] * world, [None] * world

    def work(r):
        try:
            c = ob.Context(0)
            c.init_local(grp, r)
            shard = obd.pack_row_shard_from_slice(c, d, r, world)
            direct = obd.pack_row_shard(c, d, r, world)
            assert (shard.world, shard.rank, shard.n_a_global, shard.n_b_global) == (world, r, direct.n_a_global, direct.n_b_global)
            same = all(np.array_equal(a, b_, equal_nan=True) for a, b_ in zip(shard.download(), direct.download()))
            direct.close()
            outs[r] = (same, ob.bootstrap(shard, 150, ref_kind=ob.REF_POOLED, norm=norm, seed=41, want_rep=True))
            # outcome refresh on the shard (whole-frame y on every rank) == a shard packed from the new outcome
            y2 = d["outcome"] * 0.5 - 3.0
            shard.update_outcome(y2)
            d2 = dict(d, outcome=y2)
            direct2 = obd.pack_row_shard(c, d2, r, world)
            assert all(np.array_equal(a, b_, equal_nan=True) for a, b_ in zip(shard.download(), direct2.download()))
            direct2.close()
            shard.close(); c.close()
        except Exception as ex:  # noqa: BLE001
            errs[r] = ex
    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert all(e is None for e in errs), errs
    for same, o in outs:
        assert same
        for k in ("point_stats", "rep_stats", "std_err", "ci_lower", "ci_upper", "p_value"):
            assert _same(o[k], one[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("world,n,sort_by_group", [(2, 50_003, False), (4, 700_001, False), (4, 50_003, True), (8, 1_300_000, False),
                                                  (2, 600_000, True)])
def test_async_row_shard_pack_equals_row_shards(world, n, sort_by_group):
    """ob_design_pack_row_shard_async: rows written straight into the shard while the slice uploads, one exchange of the
    rows other ranks own, split Gram launches.  Shard (download) and sharded bootstrap equal the direct row shard / one
    GPU bit for bit -- small slices (one chunk), large ones (two chunks: the first Gram launch runs before the exchange)
    and a frame sorted by group (whole slices change hands; a rank may keep nothing of its own slice)."""
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import core, synth, distributed as obd
    d = synth.make_wage(n, 3, cat_levels=(3,), weights=True, seed=19)
    if sort_by_group:
        order = np.argsort(d["group"], kind="stable")
        for k in ("outcome", "weights", "group"):
            d[k] = d[k][order]
        d["cont"] = [c[order] for c in d["cont"]]
        d["cat_codes"] = [c[order] for c in d["cat_codes"]]
    norm = [ob.NormVar(m, i) for m, i in synth.norm_spec(d)]
    reps = 140
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    one = ob.bootstrap(des, reps, ref_kind=ob.REF_WEIGHTED, norm=norm, seed=43, want_rep=True)
    des.close(); ctx.close()
    grp = core.LocalGroup(world)
    outs, errs = [None] * world, [None] * world

    def work(r):
        try:
            c = ob.Context(0)
            c.init_local(grp, r)
            sh = obd.pack_row_shard_async(c, d, r, world)          # straight into the bootstrap
            o = ob.bootstrap(sh, reps, ref_kind=ob.REF_WEIGHTED, norm=norm, seed=43, want_rep=True)
            direct = obd.pack_row_shard(c, d, r, world)
            same = all(np.array_equal(a, b_, equal_nan=True) for a, b_ in zip(sh.download(), direct.download()))
            sh.update_outcome(d["outcome"])                        # the frame-row map is global
            same = same and all(np.array_equal(a, b_, equal_nan=True) for a, b_ in zip(sh.download(), direct.download()))
            sh.close()
            sh2 = obd.pack_row_shard_async(c, d, r, world)         # completed by download() instead (exchange inside)
            same = same and all(np.array_equal(a, b_, equal_nan=True) for a, b_ in zip(sh2.download(), direct.download()))
            sh2.close(); direct.close(); c.close()
            outs[r] = (same, o)
        except Exception as ex:  # noqa: BLE001
            errs[r] = ex
    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=600)
    assert all(e is None for e in errs), errs
    for same, o in outs:
        assert same
        for k in ("point_stats", "rep_stats", "std_err", "ci_lower", "ci_upper", "p_value"):
            assert _same(o[k], one[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_rif_on_row_shards_is_bit_identical(world, orc):
    """decompose_quantile's RIF pre-step on a row-sharded design: radix-select histograms and leaf partial sums are
    all-reduced over the communicator; the RIF outcome of every shard, and the sharded bootstrap of a one-pass quantile
    sweep, equal the unsharded ones bit for bit (and the oracle's RIF within 1e-10)."""
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import core, synth, distributed as obd
    from helpers import relerr
    d = synth.make_wage(70_001, 3, cat_levels=(3,), weights=True, seed=23)
    taus = (0.1, 0.5, 0.9)
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    _, ya_raw, _, _, yb_raw, _ = des.download()
    des.apply_rif(0.9)
    _, ya9, _, _, yb9, _ = des.download()
    assert relerr(ya9, orc.rif(ya_raw, 0.9)) <= 1e-10 and relerr(yb9, orc.rif(yb_raw, 0.9)) <= 1e-10
    des.apply_rif_multi(taus)
    one = ob.bootstrap(des, 120, ref_kind=ob.REF_GROUP_B, seed=8, want_rep=True)
    des.close(); ctx.close()
    grp = core.LocalGroup(world)
    outs, errs = [None] * world, [None] * world

    def work(r):
        try:
            c = ob.Context(0)
            c.init_local(grp, r)
            sh = obd.pack_row_shard_async(c, d, r, world)
            sh.apply_rif(0.9)
            y9 = sh.download()
            sh.apply_rif_multi(taus)
            o = ob.bootstrap(sh, 120, ref_kind=ob.REF_GROUP_B, seed=8, want_rep=True)
            sh.close(); c.close()
            outs[r] = (y9[1], y9[4], o)
        except Exception as ex:  # noqa: BLE001
            errs[r] = ex
    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert all(e is None for e in errs), errs
    assert np.array_equal(np.concatenate([o[0] for o in outs]), ya9) and np.array_equal(np.concatenate([o[1] for o in outs]), yb9)
    for _, _, o in outs:
        for k in ("point_stats", "rep_stats", "std_err", "ci_lower", "ci_upper", "p_value"):
            assert _same(o[k], one[k]), k


@pytest.mark.gpu
def test_async_pack_marked_as_row_shard_while_in_flight():
    """Rows selected on the host (ob_row_shard_plan), uploaded with ob_design_pack_async and marked with
    ob_design_set_row_shard while the upload is still in flight: the sharded bootstrap overlaps it and equals one GPU."""
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import core, synth, distributed as obd
    d = synth.make_wage(900_000, 3, cat_levels=(3,), weights=False, seed=29)
    world, reps = 2, 130
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    one = ob.bootstrap(des, reps, seed=9, want_rep=True)
    des.close(); ctx.close()
    grp = core.LocalGroup(world)
    outs, errs = [None] * world, [None] * world

    def work(r):
        try:
            c = ob.Context(0)
            c.init_local(grp, r)
            loc = obd.shard_frame(d, r, world)
            sh = ob.Design.pack(c, loc["cont"], loc["cat_codes"], loc["cat_levels"], loc["outcome"], loc["weights"], loc["group"],
                                asynchronous=True)
            sh.set_row_shard(loc["n_a_global"], loc["n_b_global"], world, r)
            outs[r] = ob.bootstrap(sh, reps, seed=9, want_rep=True)
            sh.close(); c.close()
        except Exception as ex:  # noqa: BLE001
            errs[r] = ex
    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert all(e is None for e in errs), errs
    for o in outs:
        for k in ("point_stats", "rep_stats", "std_err", "ci_lower", "ci_upper"):
            assert _same(o[k], one[k]), k


@pytest.mark.gpu
def test_async_row_shard_pack_with_more_ranks_than_leaves():
    """Eight ranks, 700 rows: the row-shard plan leaves most ranks without rows of one group or of both; every rank
    still takes part in the exchange and the collectives, and the result equals one GPU."""
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import core, synth, distributed as obd
    d = synth.make_wage(700, 2, cat_levels=(), weights=False, seed=31)
    world, reps = 8, 70
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    one = ob.bootstrap(des, reps, seed=6, want_rep=True)
    des.close(); ctx.close()
    grp = core.LocalGroup(world)
    outs, errs = [None] * world, [None] * world

    def work(r):
        try:
            c = ob.Context(0)
            c.init_local(grp, r)
            sh = obd.pack_row_shard_async(c, d, r, world)
            outs[r] = ob.bootstrap(sh, reps, seed=6, want_rep=True)
            sh.close(); c.close()
        except Exception as ex:  # noqa: BLE001
            errs[r] = ex
    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert all(e is None for e in errs), errs
    for o in outs:
        for k in ("point_stats", "rep_stats", "std_err", "ci_lower", "ci_upper"):
            assert _same(o[k], one[k]), k

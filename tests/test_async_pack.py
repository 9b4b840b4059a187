"""ob_design_pack_async: the frame is uploaded and packed in row chunks on the copy stream while ob_bootstrap_run
already generates replicates and contracts the leaves that have arrived (two Gram launches instead of one).  The
design, every statistic and the deferred errors must equal the synchronous path bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _same(a, b):
    return np.array_equal(np.nan_to_num(np.asarray(a), nan=-7.0), np.nan_to_num(np.asarray(b), nan=-7.0))


KEYS = ("point_stats", "rep_stats", "rep_status", "rep_beta_a", "rep_beta_b", "std_err", "ci_lower", "ci_upper", "p_value",
        "t_stat", "xa_mean", "xb_mean", "beta_star", "residuals_b")


def _pinned_frame(ob, d):
    """The frame's columns copied into page-locked buffers (ob_host_alloc)."""
    keep = []

    def pin(a):
        b = ob.PinnedBuffer(a.shape, a.dtype)
        b.array[...] = a
        keep.append(b)
        return b.array
    fr = dict(cont=[pin(c) for c in d["cont"]], cat_codes=[pin(c) for c in d["cat_codes"]], cat_levels=d["cat_levels"],
              outcome=pin(d["outcome"]), weights=None if d["weights"] is None else pin(d["weights"]), group=pin(d["group"]))
    return fr, keep


@pytest.mark.parametrize("n,weighted,pinned", [(3_000, False, True), (300_000, True, True), (700_001, True, False),
                                               (1_200_000, False, True)])
def test_async_pack_equals_sync_pack(n, weighted, pinned):
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(n, 5, cat_levels=(4, 3), weights=weighted, seed=31)
    norm = [ob.NormVar(m, i) for m, i in synth.norm_spec(d)]
    fr, keep = _pinned_frame(ob, d) if pinned else (d, None)
    ctx = ob.Context(0)
    args = (fr["cont"], fr["cat_codes"], fr["cat_levels"], fr["outcome"], fr["weights"], fr["group"])
    sync = ob.Design.pack(ctx, *args)
    ref_mats = sync.download()
    ref = ob.bootstrap(sync, 200, ref_kind=ob.REF_POOLED, norm=norm, seed=17, want_rep=True)
    sync.close()
    # (1) straight into the bootstrap: the Gram contraction starts on the first chunk's leaves
    a1 = ob.Design.pack(ctx, *args, asynchronous=True)
    assert (a1.n_a, a1.n_b, a1.K) == (len(ref_mats[1]), len(ref_mats[4]), ref_mats[0].shape[1])    # shape known at once
    out = ob.bootstrap(a1, 200, ref_kind=ob.REF_POOLED, norm=norm, seed=17, want_rep=True)
    for k in KEYS:
        assert _same(out[k], ref[k]), k
    assert out["n_ok"] == ref["n_ok"] and out["total_gap"] == ref["total_gap"]
    again = ob.bootstrap(a1, 200, ref_kind=ob.REF_POOLED, norm=norm, seed=17, want_rep=True)     # now an ordinary resident design
    assert _same(again["rep_stats"], ref["rep_stats"])
    for m_a, m_s in zip(a1.download(), ref_mats):
        assert np.array_equal(m_a, m_s, equal_nan=True)
    a1.close()
    # (2) other entry points complete the pack first
    a2 = ob.Design.pack(ctx, *args, asynchronous=True)
    for m_a, m_s in zip(a2.download(), ref_mats):
        assert np.array_equal(m_a, m_s, equal_nan=True)
    a2.close()
    a3 = ob.Design.pack(ctx, *args, asynchronous=True)
    a3.wait()
    h2d, pk = a3.pack_timings()
    assert h2d > 0.0 and pk >= 0.0
    a3.close()
    # (3) destroyed while in flight
    a4 = ob.Design.pack(ctx, *args, asynchronous=True)
    a4.close()
    ctx.close()


def test_async_pack_reports_deferred_errors():
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import synth
    d = synth.make_wage(400_000, 3, cat_levels=(3,), weights=True, seed=2)
    ctx = ob.Context(0)
    w = d["weights"].copy()
    w[333_333] = -1.0                                   # ols.rs:60-66: negative weight
    bad = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], w, d["group"], asynchronous=True)
    with pytest.raises(ob.OaxacaError) as e:
        ob.bootstrap(bad, 16, seed=1)
    assert e.value.kind == "InvalidGroupVariable" and "negative" in str(e.value)
    with pytest.raises(ob.OaxacaError) as e:            # the design stays unusable
        bad.download()
    assert e.value.kind == "InvalidGroupVariable"
    bad.close()
    codes = d["cat_codes"][0].copy()
    codes[5] = 9
    bad = ob.Design.pack(ctx, d["cont"], [codes], d["cat_levels"], d["outcome"], d["weights"], d["group"], asynchronous=True)
    with pytest.raises(ob.OaxacaError) as e:
        bad.wait()
    assert e.value.kind == "InvalidArgument"
    bad.close()
    # the synchronous path reports the same errors at pack time
    with pytest.raises(ob.OaxacaError) as e:
        ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], w, d["group"])
    assert e.value.kind == "InvalidGroupVariable"
    ctx.close()


def test_async_pack_edge_shapes():
    """Tiny frames (one chunk, fewer rows than a pipeline stage), an empty group (the error arrives from the bootstrap,
    the in-flight design is destroyed cleanly) and an empty frame."""
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import synth
    ctx = ob.Context(0)
    for n in (40, 33, 129):
        d = synth.make_wage(n, 2, seed=n)
        args = (d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
        s = ob.Design.pack(ctx, *args)
        a = ob.Design.pack(ctx, *args, asynchronous=True)
        rs, ra = ob.bootstrap(s, 50, seed=2, want_rep=True), ob.bootstrap(a, 50, seed=2, want_rep=True)
        assert _same(rs["rep_stats"], ra["rep_stats"]) and _same(rs["std_err"], ra["std_err"])
        s.close(); a.close()
    d = synth.make_wage(5_000, 2, seed=1)
    g = np.zeros_like(d["group"])                     # everybody in group A
    a = ob.Design.pack(ctx, d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], g, asynchronous=True)
    assert (a.n_a, a.n_b) == (5_000, 0)
    with pytest.raises(ob.OaxacaError) as e:
        ob.bootstrap(a, 8, seed=1)
    assert e.value.kind == "InvalidGroupVariable"
    a.close()
    e0 = ob.Design.pack(ctx, [np.empty(0)], [], [], np.empty(0), None, np.empty(0, dtype=np.uint8), asynchronous=True)
    assert (e0.n_a, e0.n_b) == (0, 0)
    e0.close()
    ctx.close()

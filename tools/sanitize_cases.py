#!/usr/bin/env python
"""Small shapes through every kernel family, for compute-sanitizer (racecheck / memcheck / synccheck on the
hand-rolled mbarrier ring of the Gram kernel, the histogram atomics, the radix select, the collectives):

    compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py
    compute-sanitizer --tool synccheck python tools/sanitize_cases.py

One tool per gpurun call (B200_PROFILING.md).  Exits non-zero if a result differs between the paths compared."""
import os
import sys
import threading

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oaxaca_blinder_rs_b200 as ob                                    # noqa: E402
from oaxaca_blinder_rs_b200 import core, distributed as obd, synth    # noqa: E402


def same(a, b):
    return np.array_equal(np.nan_to_num(a, nan=-7.0), np.nan_to_num(b, nan=-7.0))


def main():
    d = synth.make_wage(6_000, 5, cat_levels=(4, 3), weights=True, seed=7)
    norm = [ob.NormVar(m, i) for m, i in synth.norm_spec(d)]
    args = (d["cont"], d["cat_codes"], d["cat_levels"], d["outcome"], d["weights"], d["group"])
    kw = dict(ref_kind=ob.REF_POOLED, norm=norm, seed=3, want_rep=True)
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, *args)
    one = ob.bootstrap(des, 150, **kw)                                  # native Philox stream: 2 panels, partly filled tail
    rng = np.random.default_rng(1)
    ia = rng.integers(0, des.n_a, size=(20, des.n_a), dtype=np.uint32)
    ib = rng.integers(0, des.n_b, size=(20, des.n_b), dtype=np.uint32)
    ob.bootstrap(des, 20, idx_a=ia, idx_b=ib, ref_kind=ob.REF_WEIGHTED, norm=norm, want_rep=True)    # histogram kernel
    ob.bootstrap(des, 150, count_bits=16, **kw)                         # uint16 multiplicities
    ob.bootstrap(des, 150, max_workspace_bytes=3_000_000, **kw)         # several panel batches
    des.apply_rif_multi((0.1, 0.5, 0.9))                                # radix select + re-layout + multi-outcome solve
    ob.bootstrap(des, 40, **kw)
    des.update_outcome(d["outcome"])
    des.close()
    a = ob.Design.pack(ctx, *args, asynchronous=True)                   # chunked pack on the copy stream
    out = ob.bootstrap(a, 150, **kw)
    assert same(out["rep_stats"], one["rep_stats"]), "asynchronous pack differs"
    a.close()
    big = synth.make_wage(300_000, 3, seed=2)                           # two chunks, two Gram launches
    a = ob.Design.pack(ctx, big["cont"], big["cat_codes"], big["cat_levels"], big["outcome"], big["weights"], big["group"], asynchronous=True)
    ob.bootstrap(a, 16, seed=1)
    a.close()
    ctx.close()

    world = 2
    grp = core.LocalGroup(world)
    outs, errs = [None] * world, [None] * world

    def work(r):
        try:
            c = ob.Context(0)
            c.init_local(grp, r)
            sh = obd.pack_row_shard_from_slice(c, d, r, world)          # redistribute + mode N collectives
            o1 = ob.bootstrap(sh, 150, max_workspace_bytes=2_000_000, **kw)
            sh.close()
            full = obd.pack_replicated(c, d, r, world)                  # all-gather rows + mode R inside the library
            o2 = ob.bootstrap(full, 150, shard_replicates=True, **kw)
            full.close(); c.close()
            outs[r] = (o1, o2)
        except Exception as e:  # noqa: BLE001
            errs[r] = e
    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert all(e is None for e in errs), errs
    for o1, o2 in outs:
        assert same(o1["rep_stats"], one["rep_stats"]) and same(o2["rep_stats"], one["rep_stats"]), "sharded run differs"
        assert same(o1["std_err"], one["std_err"]) and same(o2["std_err"], one["std_err"])
    print("sanitize_cases: all paths ran and agree")


if __name__ == "__main__":
    main()

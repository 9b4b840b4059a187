#!/usr/bin/env python
"""Times the Heckman two-step bootstrap on the GPU and, on a bounded number of replicates, the oracle's
restatement of the reference on the host cores.  python tools/heckman_probe.py [n] [n_x] [reps] -> one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oaxaca_blinder_rs_b200 as ob                      # noqa: E402
from test_gpu_heckman import dense, make_selection_frame  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    n_x = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    fr = make_selection_frame(n, n_x, seed=1)
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, fr["cont"], [fr["cat"]], [3], fr["y"], None, fr["group"])
    des.attach_selection(fr["s"], fr["z"])
    ob.bootstrap(des, reps, ref_kind=1, seed=1)
    best, out = None, None
    for _ in range(3):
        t0 = time.perf_counter()
        out = ob.bootstrap(des, reps, ref_kind=1, seed=1)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    des.close(); ctx.close()
    line = {"workload": f"heckman n={n}, K={n_x + 3}, K1=3, B={reps}", "gpu_seconds": best, "gpu_reps_per_s": reps / best,
            "n_ok": out["n_ok"], "stage_ms": out["timings_ms"], "gpu_launches": out["gpu_launches"]}
    if os.environ.get("HECKMAN_PROBE_CPU", "1") == "1":
        from oracle import pyoracle as orc
        (Xa, ya, Za, sa), (Xb, yb, Zb, sb) = dense(fr)
        threads = os.cpu_count() or 1
        r_cpu = max(threads, 8)
        ia, ib = orc.index_stream(1, r_cpu, 0, len(ya)), orc.index_stream(1, r_cpu, 1, len(yb))
        t0 = time.perf_counter()
        orc.heckman_run(1, Xa, ya, Za, sa, Xb, yb, Zb, sb, r_cpu, ia, ib, nthreads=threads, precise=False)
        dt = time.perf_counter() - t0
        line["cpu_port"] = {"reps_per_s": (r_cpu + 1) / dt, "threads": threads, "sample": f"{r_cpu} replicates + point pass, {dt:.1f} s"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()

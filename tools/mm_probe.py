#!/usr/bin/env python
"""Timing probe of the Machado-Mata path (ob_mm_run) on the GPU, with the oracle port on the host cores beside it on a
bounded sample of the same regressions.  Usage: python tools/mm_probe.py [n] [n_x] [sims] [reps] [out.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    n_x = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    sims = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    out_path = sys.argv[5] if len(sys.argv) > 5 else None
    import oaxaca_blinder_rs_b200 as ob
    from test_gpu_mm import dense, make_frame
    fr = make_frame(n, n_x, seed=1)
    ctx = ob.Context(0)
    des = ob.Design.pack(ctx, fr["cont"], [fr["cat"]], [3], fr["y"], None, fr["group"])
    q = [0.1, 0.25, 0.5, 0.75, 0.9]
    ob.machado_mata(des, q, simulations=min(sims, 16), reps=1, seed=1)          # warm-up
    runs = []
    for it in range(2):
        t0 = time.perf_counter()
        r = ob.machado_mata(des, q, simulations=sims, reps=reps, seed=2 + it)
        runs.append(dict(wall_s=time.perf_counter() - t0, ms=r["timings_ms"], qr=r["qr"], launches=r["gpu_launches"]))
    K = des.K
    best = min(runs, key=lambda x: x["wall_s"])
    nprob = best["qr"]["total"]
    rec = dict(workload=f"machado-mata n={n} K={K} sims={sims} reps={reps}", n_a=des.n_a, n_b=des.n_b, K=K, problems=nprob,
               gpu_runs=runs, gpu_regressions_per_s=nprob / best["ms"]["qr"] * 1e3,
               gpu_passes_per_s=(reps + 1) / best["wall_s"], point_stats=r["point_stats"].tolist(), std_err=r["std_err"].tolist(),
               mean_ipm_iterations=best["qr"]["iterations"] / max(nprob, 1))
    des.close(); ctx.close()
    # CPU: the oracle port on a bounded sample of regressions of group A (OpenMP over regressions is the pass loop's job;
    # here one thread per regression, all cores)
    from oracle import pyoracle as orc
    (Xa, ya), _ = dense(fr)
    rng = np.random.default_rng(0)
    taus = rng.uniform(0.01, 0.99, size=8)
    t0 = time.perf_counter()
    for t in taus:
        orc.qr(Xa, ya, float(t))
    cpu_s = (time.perf_counter() - t0) / len(taus)
    rec.update(cpu_oracle_s_per_regression_1thread=cpu_s, cpu_cores=os.cpu_count(),
               cpu_regressions_per_s_all_cores=os.cpu_count() / cpu_s)
    print(json.dumps(rec))
    if out_path:
        with open(out_path, "w") as f:
            json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main()

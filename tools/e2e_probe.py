"""Breaks one end-to-end step (pinned host columns -> pack -> bootstrap -> results) into its parts."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oaxaca_blinder_rs_b200 as ob
from oaxaca_blinder_rs_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = synth.make_wage(n, 44, cat_levels=(4, 4), weights=True)
ctx = ob.Context(0)
def pin(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:1]).dtype, pin_memory=True); t.numpy()[...] = a; return t
P = dict(cont=[pin(c) for c in d["cont"]], cat=[pin(c) for c in d["cat_codes"]], y=pin(d["outcome"]), w=pin(d["weights"]), g=pin(d["group"]))
norm = [ob.NormVar(m, i) for m, i in synth.norm_spec(d)]
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    des = ob.Design.pack(ctx, [t.numpy() for t in P["cont"]], [t.numpy() for t in P["cat"]], d["cat_levels"], P["y"].numpy(), P["w"].numpy(), P["g"].numpy())
    t1 = time.perf_counter()
    out = ob.bootstrap(des, 2000, norm=norm, seed=1)
    t2 = time.perf_counter()
    des.close(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("step %d: pack %.1f ms | bootstrap wall %.1f (device %.1f, gram %.1f) | close %.1f | total %.1f" %
          (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, out["timings_ms"]["total"], out["timings_ms"]["gram"], (t3 - t2) * 1e3, (t3 - t0) * 1e3), flush=True)

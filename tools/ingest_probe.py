"""Ingest -> pack at config-3 shape: device-side cleaning/coding (ob_ingest_begin/finish) against
the same steps done on the host with numpy followed by the ordinary pack.  Frame: n rows, 44 continuous predictors,
two dictionary-coded categoricals (unsorted dictionaries), weights, ~1 % of the rows carry a null somewhere.
Prints one JSON line."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import oaxaca_blinder_rs_b200 as ob                      # noqa: E402
from oaxaca_blinder_rs_b200 import core, synth          # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = synth.make_wage(n, 44, cat_levels=(4, 4), weights=True)
rng = np.random.default_rng(5)
dicts = [["manu", "tech", "agri", "serv"], ["west", "north", "south", "east"]]          # arbitrary dictionary order
perm = [np.array([sorted(dq).index(s) for s in dq]) for dq in dicts]                       # dict code -> sorted level
inv = [np.argsort(p) for p in perm]
cat_raw = [inv[q][d["cat_codes"][q]].astype(np.int32) for q in range(2)]                   # codes into the unsorted dictionaries
group_dict = ["M", "F"]
group_raw = d["group"].astype(np.int32)
cont = [c.copy() for c in d["cont"]]
y = d["outcome"].copy()
for col in cont[:8] + [y]:
    col[rng.choice(n, n // 1000, replace=False)] = np.nan
cat_raw[0][rng.choice(n, n // 1000, replace=False)] = -1

ctx = ob.Context(0)


def host_path():
    ok = ~np.isnan(y)
    for c in cont:
        ok &= ~np.isnan(c)
    ok &= (cat_raw[0] >= 0) & (cat_raw[1] >= 0) & (group_raw >= 0)
    keep = np.flatnonzero(ok)
    codes = []
    for q in range(2):
        present = np.unique(cat_raw[q][keep])
        lv = sorted(dicts[q][i] for i in present)
        remap = np.full(len(dicts[q]), -1, dtype=np.int32)
        for i in present:
            remap[i] = lv.index(dicts[q][i])
        codes.append(remap[cat_raw[q][keep]])
    g = np.where(group_raw[keep] == 0, 0, 1).astype(np.uint8)
    return ob.Design.pack(ctx, [c[keep] for c in cont], codes, [4, 4], y[keep], d["weights"][keep], g)


def device_path():
    des, meta = core.ingest(ctx, cont, [(cat_raw[0], dicts[0]), (cat_raw[1], dicts[1])], y, d["weights"],
                            (group_raw, group_dict), reference_group="F")
    return des


def pinned_copy(a):
    import torch
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:1]).dtype, pin_memory=True)
    t.numpy()[...] = a
    return t


pins = [pinned_copy(c) for c in cont] + [pinned_copy(c) for c in cat_raw] + [pinned_copy(y), pinned_copy(d["weights"]), pinned_copy(group_raw)]


def device_path_pinned():
    pc = [t.numpy() for t in pins[:44]]
    des, meta = core.ingest(ctx, pc, [(pins[44].numpy(), dicts[0]), (pins[45].numpy(), dicts[1])], pins[46].numpy(),
                            pins[47].numpy(), (pins[48].numpy(), group_dict), reference_group="F")
    return des


out = {}
for name, fn in (("host_clean_then_pack", host_path), ("device_ingest", device_path), ("device_ingest_pinned", device_path_pinned)):
    fn().close()                                   # warm-up (memory pools)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        des = fn()
        ts.append(time.perf_counter() - t0)
        shape = (des.n_a, des.n_b, des.K)
        mats = des.download() if n <= 2_000_000 else None
        des.close()
    out[name] = {"seconds": ts, "best": min(ts), "shape": shape}
    out.setdefault("_mats", []).append(mats)
mats = out.pop("_mats")
if mats[0] is not None:
    out["identical_design"] = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(mats[0], mats[1])) and \
        all(np.array_equal(a, b, equal_nan=True) for a, b in zip(mats[0], mats[2]))
out["n"] = n
out["speedup"] = out["host_clean_then_pack"]["best"] / out["device_ingest"]["best"]
out["speedup_pinned"] = out["host_clean_then_pack"]["best"] / out["device_ingest_pinned"]["best"]
print(json.dumps(out))

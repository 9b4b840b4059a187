"""Replicate sharding plumbing (oaxaca_blinder_rs_b200/distributed.py) on CPU with gloo, world_size 2 and 3:
each rank computes its shard of replicates (here with the oracle -- tests may), the package all-gathers the rows,
and the result must equal the unsharded run bit for bit, including uneven shards and failed replicates."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem():
    rng = np.random.default_rng(4)
    n = 80
    edu = rng.normal(13, 2, n)
    rare = np.zeros(n); rare[[3, 50]] = 1.0
    grp = np.arange(n) >= 40
    y = 1 + 0.5 * edu + rare + rng.normal(0, 1, n) + grp
    X = np.c_[np.ones(n), edu, rare]
    return X[grp], y[grp], X[~grp], y[~grp]


def _worker(rank, world, port, reps, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oaxaca_blinder_rs_b200 import distributed as obd
    from oracle import pyoracle as orc
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    Xa, ya, Xb, yb = _problem()
    spec = orc.Spec(K=3, n_cont=1)
    b, e = obd.shard_range(rank, world, reps)
    ia, ib = orc.index_stream(9, reps, 0, len(ya)), orc.index_stream(9, reps, 1, len(yb))
    S = spec.n_stats
    if e > b:
        part = orc.run(spec, Xa, ya, None, Xb, yb, None, e - b, ia[b:e], ib[b:e])
        ls, lst = part["rep_stats"], part["rep_status"]
    else:
        ls, lst = np.empty((0, S)), np.empty(0, dtype=np.int32)
    stats, status = obd.gather_replicates(ls, lst, reps, S)
    q.put((rank, stats, status))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,reps", [(2, 37), (3, 8), (2, 1)])
def test_gather_matches_unsharded(orc, world, reps):
    from oaxaca_blinder_rs_b200 import distributed as obd
    # shard ranges tile [0, reps) contiguously and evenly
    rng_ = [obd.shard_range(r, world, reps) for r in range(world)]
    assert rng_[0][0] == 0 and rng_[-1][1] == reps and all(a[1] == b[0] for a, b in zip(rng_, rng_[1:]))
    assert max(e - b for b, e in rng_) - min(e - b for b, e in rng_) <= 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, reps, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    Xa, ya, Xb, yb = _problem()
    spec = orc.Spec(K=3, n_cont=1)
    ia, ib = orc.index_stream(9, reps, 0, len(ya)), orc.index_stream(9, reps, 1, len(yb))
    full = orc.run(spec, Xa, ya, None, Xb, yb, None, reps, ia, ib)
    for rank, stats, status in got:
        np.testing.assert_array_equal(status, full["rep_status"])
        np.testing.assert_array_equal(np.nan_to_num(stats, nan=-1.0), np.nan_to_num(full["rep_stats"], nan=-1.0))
    red = orc.reduce(np.nan_to_num(got[0][1]), got[0][2], full["point"]["stats"])
    np.testing.assert_array_equal(red["se"], full["se"])
    assert red["n_ok"] == full["n_ok"]

"""Row sharding over the library's own NCCL communicator (one process per GPU): needs >= 2 GPUs, skipped otherwise.
Every rank must return statistics BIT-IDENTICAL to an unsharded run of the same frame on one GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oaxaca_blinder_rs_b200 as ob
    from oaxaca_blinder_rs_b200 import synth, distributed as obd
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    full = synth.make_wage_rows(400_000, 6, cat_levels=(4,), weights=True, chunk=1 << 16)
    norm = [ob.NormVar(m, i) for m, i in synth.norm_spec(full)]
    reps = 200
    ctx = ob.Context(rank)
    ctx.init_nccl(rank, world)
    loc = synth.make_wage_rows(400_000, 6, cat_levels=(4,), weights=True, chunk=1 << 16, rank=rank, world=world)
    des = obd.pack_row_shard(ctx, loc, rank, world)
    out = ob.bootstrap(des, reps, ref_kind=ob.REF_WEIGHTED, norm=norm, seed=5, want_rep=True, max_workspace_bytes=80_000_000)
    # frame slices -> row shards over NCCL send/recv: the same shard, bit for bit
    re = obd.pack_row_shard_from_slice(ctx, full, rank, world)
    for a, b_ in zip(re.download(), des.download()):
        assert np.array_equal(a, b_, equal_nan=True)
    re.close()
    # the asynchronous variant (rows straight into the shard, one exchange, overlapped with the bootstrap) over NCCL
    asy = obd.pack_row_shard_async(ctx, full, rank, world)
    out_async = ob.bootstrap(asy, reps, ref_kind=ob.REF_WEIGHTED, norm=norm, seed=5, want_rep=True, max_workspace_bytes=80_000_000)
    for a, b_ in zip(asy.download(), des.download()):
        assert np.array_equal(a, b_, equal_nan=True)
    for k in ("point_stats", "rep_stats", "std_err"):
        assert np.array_equal(np.nan_to_num(out_async[k], nan=-7.0), np.nan_to_num(out[k], nan=-7.0)), k
    asy.close()
    des.close()
    # mode R upload over the same communicator: frame slices packed per rank, full design gathered over NVLink
    rep = obd.pack_replicated(ctx, full, rank, world)
    gathered = rep.download()
    # mode R inside the library: replicate shards, statistics all-gathered device to device over NCCL
    mode_r = ob.bootstrap(rep, reps, ref_kind=ob.REF_WEIGHTED, norm=norm, seed=5, want_rep=True, shard_replicates=True)
    rep.close()
    ctx.comm_destroy()
    one = None
    if rank == 0:
        des1 = ob.Design.pack(ctx, full["cont"], full["cat_codes"], full["cat_levels"], full["outcome"], full["weights"], full["group"])
        one = ob.bootstrap(des1, reps, ref_kind=ob.REF_WEIGHTED, norm=norm, seed=5, want_rep=True)
        whole = des1.download()
        des1.close()
        for a, b_ in zip(gathered, whole):
            assert np.array_equal(a, b_, equal_nan=True)
    keys = ("point_stats", "rep_stats", "std_err", "ci_lower", "ci_upper", "p_value")
    q.put((rank, {k: out[k] for k in keys}, None if one is None else {k: one[k] for k in keys}, out["timings_ms"],
           {k: mode_r[k] for k in keys}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_nccl_row_sharding_matches_one_gpu_bit_for_bit():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    one = got[0][2]
    for rank, out, _, tm, mode_r in got:
        for k, v in out.items():
            assert np.array_equal(np.nan_to_num(v, nan=-7.0), np.nan_to_num(one[k], nan=-7.0)), (rank, k)
            assert np.array_equal(np.nan_to_num(mode_r[k], nan=-7.0), np.nan_to_num(one[k], nan=-7.0)), ("mode R", rank, k)
        assert tm["comm"] > 0.0

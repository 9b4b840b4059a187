"""Multi-GPU plumbing; one process per GPU.

Mode R, replicate sharding: every rank holds the full packed design, rank r computes a contiguous range of
global replicate ids with the counter-based stream keyed by global id, one all-gather of the
[reps_r x S] statistics block, then every rank reduces the identical gathered array.  With a communicator attached
to the context (Context.init_nccl) all of that happens INSIDE the library (ob_boot_opts.shard_replicates: device-to-
device all-gather over NVLink, no host hop, no torch); without one, gather_replicates() carries the rows through
torch.distributed (gloo in the CPU tests) and ob_reduce_stats reduces them.

Mode N, row sharding (n too large for one HBM): shard_frame() selects the rows of each group that
ob_row_shard_plan assigns to this rank; the packed shard is marked with set_row_shard and the library itself
exchanges column sums and per-rank Gram sums over its own NCCL communicator (Context.init_nccl).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(rank: int, world: int, reps: int) -> Tuple[int, int]:
    """Contiguous, balanced replicate range [begin, end) of `rank`; the first reps % world ranks get one more."""
    base, extra = divmod(reps, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_replicates(local_stats: np.ndarray, local_status: np.ndarray, reps: int, S: int, group=None, device=None):
    """All-gather the per-rank replicate rows into global replicate order.  Returns (stats [reps,S], status [reps])
    identical on every rank.  Uneven shards are padded to the largest shard for the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    b, e = shard_range(rank, world, reps)
    assert local_stats.shape == (e - b, S) and local_status.shape == (e - b,)
    cap = -(-reps // world) if reps else 0
    buf = torch.zeros((max(cap, 1), S + 1), dtype=torch.float64)
    if e > b:
        buf[: e - b, :S] = torch.from_numpy(np.ascontiguousarray(local_stats))
        buf[: e - b, S] = torch.from_numpy(local_status.astype(np.float64))
    if device is not None:
        buf = buf.to(device)
    out = torch.empty((world * buf.shape[0], S + 1), dtype=torch.float64, device=buf.device)   # concatenated form
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.cpu().numpy().reshape(world, buf.shape[0], S + 1)
    stats = np.empty((reps, S))
    status = np.empty(reps, dtype=np.int32)
    for r in range(world):
        rb, re = shard_range(r, world, reps)
        stats[rb:re] = out[r, : re - rb, :S]
        status[rb:re] = out[r, : re - rb, S].astype(np.int32)
    return stats, status


def bootstrap_sharded(design, reps: int, group=None, device=None, **kw) -> dict:
    """ob_bootstrap_run on this rank's replicate shard + all-gather + reduction (same result on all ranks).
    A context with a communicator does it all inside the library (shard_replicates); the host-gather path below is
    for contexts without one."""
    from . import core
    if design.ctx.comm_world > 1 and design.world == 1:
        kw.setdefault("want_residuals", design.ctx.comm_rank == 0)
        return core.bootstrap(design, reps, shard_replicates=True, **kw)
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    b, e = shard_range(rank, world, reps)
    kw.setdefault("want_residuals", rank == 0)     # OaxacaResults.residuals (builder.rs:946) is fetched once, on rank 0
    part = core.bootstrap(design, reps, rep_begin=b, rep_end=e, skip_reduce=True, **kw) if e > b else \
        core.bootstrap(design, 0, skip_reduce=True, **{k: v for k, v in kw.items() if k not in ("idx_a", "idx_b")})
    S = part["S"]
    local_stats = part["rep_stats"] if e > b else np.empty((0, S))
    local_status = part["rep_status"] if e > b else np.empty(0, dtype=np.int32)
    stats, status = gather_replicates(local_stats, local_status, reps, S, group=group, device=device)
    red = core.reduce_stats(design.ctx, stats, status, part["point_stats"]) if reps else \
        dict(std_err=np.full(S, np.nan), p_value=np.full(S, np.nan), ci_lower=np.full(S, np.nan),
             ci_upper=np.full(S, np.nan), t_stat=np.zeros(S), n_ok=0)
    out = dict(part)
    out.update(red)
    out["rep_stats"], out["rep_status"] = stats, status
    return out


def shard_frame(d: dict, rank: int, world: int) -> dict:
    """Rows of the frame `d` (synth.make_wage layout) that rank holds under row sharding: for each group the
    contiguous range ob_row_shard_plan gives, in frame order.  Adds n_a_global / n_b_global."""
    from . import core
    grp = d["group"]
    pos_a = np.cumsum(grp == 0) - 1          # position of each row inside its group (frame order)
    pos_b = np.cumsum(grp == 1) - 1
    na, nb = int((grp == 0).sum()), int((grp == 1).sum())
    a0, a1 = core.row_shard_plan(na, world, rank)
    b0, b1 = core.row_shard_plan(nb, world, rank)
    keep = ((grp == 0) & (pos_a >= a0) & (pos_a < a1)) | ((grp == 1) & (pos_b >= b0) & (pos_b < b1))
    out = dict(n=int(keep.sum()), cont=[c[keep] for c in d["cont"]], cat_codes=[c[keep] for c in d["cat_codes"]],
               cat_levels=list(d["cat_levels"]), outcome=d["outcome"][keep],
               weights=None if d["weights"] is None else d["weights"][keep], group=grp[keep],
               n_a_global=na, n_b_global=nb)
    return out


def pack_row_shard(ctx, d: dict, rank: int, world: int):
    """Packs this rank's rows and marks the design as a row shard (mode N)."""
    from . import core
    loc = d if "n_a_global" in d else shard_frame(d, rank, world)
    des = core.Design.pack(ctx, loc["cont"], loc["cat_codes"], loc["cat_levels"], loc["outcome"], loc["weights"], loc["group"])
    des.set_row_shard(loc["n_a_global"], loc["n_b_global"], world, rank)
    return des


def frame_slice(d: dict, rank: int, world: int) -> dict:
    """Contiguous rows [n r / world, n (r+1) / world) of the frame (views, no copy)."""
    n = d["n"]
    lo, hi = n * rank // world, n * (rank + 1) // world
    return dict(n=hi - lo, cont=[c[lo:hi] for c in d["cont"]], cat_codes=[c[lo:hi] for c in d["cat_codes"]],
                cat_levels=list(d["cat_levels"]), outcome=d["outcome"][lo:hi],
                weights=None if d["weights"] is None else d["weights"][lo:hi], group=d["group"][lo:hi])


def pack_replicated(ctx, d: dict, rank: int, world: int):
    """Mode R upload: this rank sends only its frame slice over PCIe and packs it; the full design is assembled on
    every GPU by ob_design_allgather_rows (needs Context.init_nccl / init_local).  Equals Design.pack(whole frame)."""
    from . import core
    sl = frame_slice(d, rank, world)
    loc = core.Design.pack(ctx, sl["cont"], sl["cat_codes"], sl["cat_levels"], sl["outcome"], sl["weights"], sl["group"])
    full = loc.allgather_rows()
    loc.close()
    return full


def pack_row_shard_from_slice(ctx, d: dict, rank: int, world: int):
    """Mode N upload without host-side row selection: this rank uploads and packs its contiguous frame slice, then
    ob_design_redistribute_rows re-cuts the groups' rows along the row-shard plan over the communicator.  Equals
    pack_row_shard (bit for bit) while the host touches only its 1/world of the frame."""
    from . import core
    sl = frame_slice(d, rank, world)
    loc = core.Design.pack(ctx, sl["cont"], sl["cat_codes"], sl["cat_levels"], sl["outcome"], sl["weights"], sl["group"])
    shard = loc.redistribute_rows()
    loc.close()
    return shard


def pack_row_shard_async(ctx, d: dict, rank: int, world: int):
    """ob_design_pack_row_shard_async on this rank's frame slice: like pack_row_shard_from_slice, but rows go straight to
    their place in the shard while the slice uploads, and the next bootstrap() overlaps upload, pack and the exchange
    of the few rows other ranks own with its first kernels.  Collective."""
    from . import core
    sl = frame_slice(d, rank, world)
    return core.Design.pack(ctx, sl["cont"], sl["cat_codes"], sl["cat_levels"], sl["outcome"], sl["weights"], sl["group"], row_shard=True)

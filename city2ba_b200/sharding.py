"""Camera sharding across ranks: one process per GPU, rank r owns the contiguous camera
range [r*C/N, (r+1)*C/N); mesh, BVH and points are replicated; the only exchange is one
all-gather of per-rank observation counts, from which every rank knows where its slab starts in
the global CSR.  The collective goes through torch.distributed (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import numpy as np

from .generate import VisGraph


def camera_range(num_cameras: int, rank: int, world: int):
    return (rank * num_cameras) // world, ((rank + 1) * num_cameras) // world


def exchange_counts(local_obs: int, group=None, device=None):
    """all-gather of the per-rank observation counts -> (counts[world], my global offset)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = torch.tensor([int(local_obs)], dtype=torch.int64, device=device)
    allc = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allc, mine, group=group)
    counts = allc.cpu().numpy()
    return counts, int(counts[:rank].sum())


def visibility_graph_sharded(scene, cameras, points, max_dist, *, group=None, device=None,
                             compute=None, gather_to=None, **kw):
    """Each rank computes the graph of its camera range.  Returns (local VisGraph, info) where
    info = {range, counts, obs_offset}; with gather_to=r the full graph is also assembled on rank r
    (info["global"]) — the BAL assembly step of the reference's host code."""
    import torch
    import torch.distributed as dist
    from .generate import _cam_array, visibility_graph
    compute = compute or visibility_graph
    cams = _cam_array(cameras)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = camera_range(len(cams), rank, world)
    local = compute(scene, cams[lo:hi], points, max_dist, **kw)
    counts, off = exchange_counts(local.num_observations, group, device)
    info = {"range": (lo, hi), "counts": counts, "obs_offset": off}
    if gather_to is not None:
        parts = [None] * world if rank == gather_to else None
        dist.gather_object((np.asarray(local.offsets), np.asarray(local.point_idx), np.asarray(local.uv)),
                           parts, dst=gather_to, group=group)
        if rank == gather_to:
            offs, base = [np.zeros(1, np.uint64)], 0
            for o, _, _ in parts:
                offs.append(o[1:].astype(np.uint64) + np.uint64(base))
                base += int(o[-1])
            info["global"] = VisGraph(np.concatenate(offs), np.concatenate([p[1] for p in parts]),
                                      np.concatenate([p[2].reshape(-1, 2) for p in parts]))
    return local, info

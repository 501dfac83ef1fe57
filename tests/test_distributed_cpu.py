"""world_size-2 gloo test of the multi-rank host logic (camera ranges, the count all-gather, CSR
assembly).  No GPU here: the per-rank compute is injected and played by the oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _oracle_compute(scene, cams, pts, max_dist, **kw):
    from oracle import oracle as o
    from city2ba_b200.generate import VisGraph
    v = o.visibility_graph(scene[0], scene[1], cams, pts, max_dist)
    return VisGraph(v.offsets, v.point_idx, v.uv)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as o
    from city2ba_b200.sharding import camera_range, visibility_graph_sharded
    cams, pts = o.grid_cameras(3, 2), o.grid_points(6, 2)
    mesh = o.city_mesh(2)
    local, info = visibility_graph_sharded(mesh, cams, pts, 10.0, compute=_oracle_compute, gather_to=0)
    lo, hi = camera_range(len(cams), rank, world)
    assert info["range"] == (lo, hi) and len(local) == hi - lo
    assert int(info["counts"][rank]) == local.num_observations
    assert info["obs_offset"] == int(info["counts"][:rank].sum())
    if rank == 0:
        full = o.visibility_graph(mesh[0], mesh[1], cams, pts, 10.0)
        g = info["global"]
        ok = (np.array_equal(g.offsets, full.offsets) and np.array_equal(g.point_idx, full.point_idx)
              and np.array_equal(g.uv, full.uv) and int(info["counts"].sum()) == full.n_obs)
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_assembles_the_single_rank_graph():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_camera_ranges_partition():
    from city2ba_b200.sharding import camera_range
    for C in (0, 1, 7, 800, 99840):
        for world in (1, 2, 3, 4, 8):
            r = [camera_range(C, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == C
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1

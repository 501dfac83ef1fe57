"""The multi-GPU entry of the library (c2b_init_multi / c2b_visibility_graph_multi): one process,
one ctx + worker thread per GPU, camera ranges, the points all-gathered from per-GPU shards and ONE NCCL
all-gather of the per-GPU observation counts, every GPU copying its slab into the one host CSR.  The result
must equal the single-GPU graph and the oracle's, bit for bit.  Runs with every GPU count the box offers
(G = 1 exercises the same phases without NCCL, so the path is covered on a one-GPU box too)."""
import numpy as np
import pytest

from conftest import assert_same_graph

pytestmark = pytest.mark.gpu


def gpu_counts():
    import torch
    n = torch.cuda.device_count()
    return [g for g in (1, 2, 3, 4, 8) if g <= n]


@pytest.fixture(scope="module")
def cfg2(orc):
    return orc.grid_cameras(10, 4), orc.grid_points(10, 4), *orc.city_mesh(4)


@pytest.mark.parametrize("G", [1, 2, 3, 4, 8])
def test_multi_equals_oracle_and_single(c2b, ctx, orc, cfg2, G):
    if G not in gpu_counts():
        pytest.skip(f"needs {G} GPUs")
    cams, pts, xyz, tri = cfg2
    ref = orc.visibility_graph(xyz, tri, cams, pts, 10.0)
    m = c2b.MultiContext(G)
    try:
        scene = c2b.MultiScene(xyz, tri, m)
        for occlusion, sc in (("mesh", scene), ("analytic", None), ("none", None)):
            g = c2b.visibility_graph_multi(m, sc, cams, pts, 10.0, occlusion=occlusion)
            single = c2b.visibility_graph(c2b.Scene(xyz, tri, ctx=ctx) if sc else None, cams, pts, 10.0,
                                          occlusion=occlusion, ctx=ctx)
            assert np.array_equal(g.offsets, single.offsets) and np.array_equal(g.point_idx, single.point_idx)
            assert np.array_equal(g.uv, single.uv)
            if occlusion == "mesh":
                assert_same_graph(g, ref, f"multi G={G}")
            mu = g.stats["multi"]
            # contiguous camera ranges (src/generate.rs:435: par_iter keeps camera order), slabs back to back
            assert mu["cam_begin"][0] == 0 and mu["cam_end"][-1] == len(cams)
            assert all(mu["cam_end"][k] == mu["cam_begin"][k + 1] for k in range(G - 1))
            assert mu["obs_base"] == list(np.concatenate([[0], np.cumsum(mu["n_obs"])[:-1]]))
            assert sum(mu["n_obs"]) == g.num_observations
        # ragged and empty inputs: fewer cameras than GPUs, no cameras, no points
        few = c2b.visibility_graph_multi(m, scene, cams[:max(1, G - 1)], pts, 10.0)
        assert_same_graph(few, orc.visibility_graph(xyz, tri, cams[:max(1, G - 1)], pts, 10.0), "fewer cameras than GPUs")
        none = c2b.visibility_graph_multi(m, scene, cams[:0], pts, 10.0)
        assert len(none.offsets) == 1 and none.num_observations == 0
        nop = c2b.visibility_graph_multi(m, scene, cams, pts[:0], 10.0)
        assert nop.num_observations == 0 and len(nop.offsets) == len(cams) + 1
        scene.close()
    finally:
        m.close()


def test_multi_full_size_hash_equals_single(c2b, ctx):
    """cfg3 (9,792 x 998,784) over all GPUs of the box: the one host CSR hashes like the single-GPU one"""
    import bench
    import torch
    G = max(gpu_counts())
    cams, pts, xyz, tri = bench.build_workload("cfg3")
    single = c2b.visibility_graph(c2b.Scene(xyz, tri, ctx=ctx), cams, pts, bench.MAX_DIST, ctx=ctx)
    m = c2b.MultiContext(G)
    try:
        scene = c2b.MultiScene(xyz, tri, m)
        want = bench.result_hash(single.offsets, single.point_idx, single.uv)
        assert want == bench.expected_hash("cfg3")
        ranges = set()
        for call in range(4):   # later calls may move the range boundaries (measured shares): the CSR must not change
            g = c2b.visibility_graph_multi(m, scene, cams, pts, bench.MAX_DIST)
            assert bench.result_hash(g.offsets, g.point_idx, g.uv) == want, f"call {call}"
            ranges.add(tuple(g.stats["multi"]["cam_end"]))
        assert np.array_equal(g.point_idx, single.point_idx) and np.array_equal(g.uv, single.uv)
        print(f"\n  camera range ends over 4 calls on {G} GPU(s): {sorted(ranges)}")
    finally:
        m.close()
    assert torch.cuda.device_count() >= G


def test_multi_errors(c2b):
    with pytest.raises(c2b.C2BError):
        c2b.MultiContext(0)
    with pytest.raises(c2b.C2BError):
        c2b.MultiContext(2, devices=[0, 0])

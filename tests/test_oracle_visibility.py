"""Oracle self-consistency and golden fixtures (CPU only)."""
import os

import numpy as np

from conftest import GOLDEN, points_on_mesh, procedural_scene, random_cameras


def test_lattice_counts_closed_form(orc):
    # C = 4*cpb*n(n+1), P = 12*ppb*n(n+1)  (loops at src/synthetic.rs:179-258)
    for cpb, ppb, n in [(10, 10, 4), (9, 306, 16), (6, 200, 64), (3, 5, 1)]:
        assert orc.lib().orc_grid_num_cameras(cpb, n) == 4 * cpb * n * (n + 1)
        assert orc.lib().orc_grid_num_points(ppb, n) == 12 * ppb * n * (n + 1)
    assert orc.grid_cameras(10, 4).shape == (800, 15)
    assert orc.grid_points(10, 4).shape == (2400, 3)


def test_lattice_first_cameras(orc):
    # first pushes at bx=by=i=0: x-street yaw -90 then +90, z-street yaw 180 then identity
    cams = orc.grid_cameras(10, 4)
    for k in range(4):
        assert np.allclose(orc.center(cams[k]), [0, 1, 0], atol=1e-12)
    fwd = [-c[:9].reshape(3, 3).T[2] for c in cams[:4]]  # viewing direction = -(third row of R)
    assert np.allclose(fwd[0], [-1, 0, 0], atol=1e-12) and np.allclose(fwd[1], [1, 0, 0], atol=1e-12)
    assert np.allclose(fwd[2], [0, 0, 1], atol=1e-12) and np.allclose(fwd[3], [0, 0, -1], atol=1e-12)


def test_golden_cfg2(orc):
    g = np.load(os.path.join(GOLDEN, "cfg2_blocks4.npz"))
    cams, pts = orc.grid_cameras(10, 4), orc.grid_points(10, 4)
    xyz, tri = orc.city_mesh(4)
    v = orc.visibility_graph(xyz, tri, cams, pts, 10.0, want_flags=True)
    assert np.array_equal(v.offsets, g["mesh_offsets"])
    assert np.array_equal(v.point_idx, g["mesh_idx"])
    assert np.array_equal(v.uv, g["mesh_uv"])
    assert np.array_equal(v.cand_occluded, g["cand_occluded"])
    assert np.array_equal(v.cand_flags, g["cand_flags"])
    assert [v.n_flag_edge, v.n_flag_graze, v.n_flag_endpoint, v.n_flag_cull] == list(g["flag_counts"])
    a = orc.synthetic_visibility(cams, pts, 10.0, True)
    assert np.array_equal(a.offsets, g["analytic_offsets"]) and np.array_equal(a.point_idx, g["analytic_idx"])
    # flags without want_flags are identical decisions (early exit must not change the answer)
    v2 = orc.visibility_graph(xyz, tri, cams, pts, 10.0)
    assert np.array_equal(v2.point_idx, v.point_idx) and np.array_equal(v2.offsets, v.offsets)


def test_golden_cfg1_scene(orc):
    g = np.load(os.path.join(GOLDEN, "cfg1_scene.npz"))
    v = orc.visibility_graph(g["xyz"], g["tri"], g["cams"], g["pts"], 100.0)
    assert np.array_equal(v.offsets, g["offsets"]) and np.array_equal(v.point_idx, g["idx"])
    assert np.array_equal(v.uv, g["uv"])
    # fixture is regenerated identically by the committed script's inputs
    xyz, tri = procedural_scene(0)
    assert np.array_equal(xyz, g["xyz"]) and np.array_equal(tri, g["tri"])


def test_cpu_ref_bvh_equals_bruteforce(orc):
    """the multithreaded CPU arm (BVH) must reproduce the brute-force oracle exactly"""
    rng = np.random.default_rng(5)
    xyz, tri = procedural_scene(3)
    cams = random_cameras(rng, 40)
    pts = points_on_mesh(rng, xyz, tri, 600)
    a = orc.visibility_graph(xyz, tri, cams, pts, 60.0)
    b, threads = orc.ref_visibility_graph(xyz, tri, cams, pts, 60.0)
    assert threads >= 1
    assert np.array_equal(a.offsets, b.offsets) and np.array_equal(a.point_idx, b.point_idx)
    assert np.array_equal(a.uv, b.uv)
    cams, pts = orc.grid_cameras(10, 4), orc.grid_points(10, 4)
    xyz, tri = orc.city_mesh(4)
    a = orc.visibility_graph(xyz, tri, cams, pts, 10.0)
    b, _ = orc.ref_visibility_graph(xyz, tri, cams, pts, 10.0, n_threads=2)
    assert np.array_equal(a.offsets, b.offsets) and np.array_equal(a.point_idx, b.point_idx)


def test_random_soup_bvh_equals_bruteforce(orc):
    rng = np.random.default_rng(11)
    xyz = rng.uniform(-10, 10, (300, 3)).astype(np.float32)
    tri = rng.integers(0, 300, (500, 3)).astype(np.uint32)
    cams = random_cameras(rng, 30, center=(0, 0, 0), spread=8.0)
    pts = rng.uniform(-10, 10, (500, 3))
    a = orc.visibility_graph(xyz, tri, cams, pts, 25.0)
    b, _ = orc.ref_visibility_graph(xyz, tri, cams, pts, 25.0)
    assert a.n_candidates > 100 and 0 < a.n_obs < a.n_candidates
    assert np.array_equal(a.offsets, b.offsets) and np.array_equal(a.point_idx, b.point_idx)


def test_ordering_and_frustum_properties(orc):
    cams, pts = orc.grid_cameras(10, 4), orc.grid_points(10, 4)
    xyz, tri = orc.city_mesh(4)
    v = orc.visibility_graph(xyz, tri, cams, pts, 10.0)
    for c in range(0, 800, 37):
        idx = v.point_idx[int(v.offsets[c]):int(v.offsets[c + 1])].astype(np.int64)
        assert np.all(np.diff(idx) > 0)  # ascending point index (src/generate.rs:446)
    assert np.all(np.abs(v.uv) <= 1.0)
    assert orc.total_reprojection_error(cams, pts, v.offsets, v.point_idx, v.uv, 2.0) == 0.0


def test_ray_predicate_basics(orc):
    tri = ([0, 0, -5], [1, 0, -5], [0, 1, -5])
    ray = np.array([0.2, 0.2, 0, 0, 0, -1, 10], np.float32)
    assert orc.ray_triangle(ray, *tri) == 1
    ray[6] = 4.0  # stops short
    assert orc.ray_triangle(ray, *tri) == 0
    ray[6] = 5.0  # t == tfar is a hit (T <= tfar*det)
    assert orc.ray_triangle(ray, *tri) == 1
    ray[:3] = (2, 2, 0)
    ray[6] = 10
    assert orc.ray_triangle(ray, *tri) == 0
    # shared edge: exactly one of the two triangles of a quad must be hit (watertight), never zero
    quad_a = ([0, 0, -5], [1, 0, -5], [1, 1, -5])
    quad_b = ([0, 0, -5], [1, 1, -5], [0, 1, -5])
    ray = np.array([0.5, 0.5, 0, 0, 0, -1, 10], np.float32)
    assert orc.ray_triangle(ray, *quad_a) + orc.ray_triangle(ray, *quad_b) >= 1
    # NaN direction: never occluded (Embree leaves invalid rays untouched)
    ray = np.array([0.2, 0.2, 0, np.nan, np.nan, np.nan, 10], np.float32)
    assert orc.ray_triangle(ray, *tri) == 0


def test_make_ray_matches_reference_formula(orc):
    c, p = np.array([1.0, 2.0, 3.0]), np.array([4.0, -2.0, 9.5])
    r = orc.make_ray(c, p)
    d = p - c
    n = np.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
    assert np.array_equal(r[:3], c.astype(np.float32))
    assert np.array_equal(r[3:6], (d * (1.0 / n)).astype(np.float32))
    assert r[6] == np.float32(np.float32(n) - np.float32(1e-6))  # src/generate.rs:464


def test_hits_building_quirk(orc):
    # a view segment crossing a building is blocked; along a street it is not (src/synthetic.rs:100-124)
    assert orc.hits_building([0, 1, 10], [20, 1, 10], 20.0, 1.0) == 1
    assert orc.hits_building([0, 1, 0], [19, 1, 0], 20.0, 1.0) == 0


import pytest  # noqa: E402


@pytest.mark.parametrize("name", ["cfg1_test_scene.npz", "box_obj.npz"])
def test_golden_reference_meshes(orc, name):
    """BASELINE config 1 on the reference's own test_scene.obj (and tests/box.obj): the committed
    arrays were derived from the OBJ files by tests/golden/make_golden_obj.py; the oracle must
    reproduce the stored graph, and its seeded point generator the stored points"""
    import os
    import numpy as np
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, name))
    md = float(g["max_dist"])
    v = orc.visibility_graph(g["xyz"], g["tri"], g["cams"], g["pts"], md)
    assert np.array_equal(v.offsets, g["offsets"]) and np.array_equal(v.point_idx, g["idx"])
    assert np.array_equal(v.uv, g["uv"]) and np.array_equal(v.cand_occluded, g["cand_occluded"])
    seed = {"cfg1_test_scene.npz": 20261017, "box_obj.npz": 7}[name]
    assert np.array_equal(orc.generate_world_points_uniform(g["xyz"], g["tri"], g["cams"], len(g["pts"]), md, seed), g["pts"])
    # every world point lies ON a triangle, so every ray is an end-point case (SURVEY finding 3)
    assert g["flag_counts"][2] >= 0.99 * len(g["cand_idx"])


def test_moeller_trumbore_ab_count(orc, capsys):
    """A/B against a restatement of Embree's DEFAULT (non-watertight Moeller-Trumbore) triangle test, the
    intersector the reference's scene actually uses: every cast ray on which it and the oracle's watertight
    predicate disagree must be one the flagged-epsilon protocol already reports (edge / grazing / end point);
    the counts go into DESIGN.md.  Both sides are restatements — parity with the Embree binary stays unpinned."""
    cases = []
    cams, pts = orc.grid_cameras(10, 4), orc.grid_points(10, 4)
    xyz, tri = orc.city_mesh(4)
    cases.append(("cfg2 lattice", xyz, tri, cams, pts, 10.0))
    for name in ("cfg1_test_scene.npz", "box_obj.npz"):
        g = np.load(os.path.join(GOLDEN, name))
        cases.append((name, g["xyz"], g["tri"], g["cams"], g["pts"], float(g["max_dist"])))
    rng = np.random.default_rng(17)
    xyz, tri = procedural_scene(2)
    cases.append(("procedural", xyz, tri, random_cameras(rng, 60), points_on_mesh(rng, xyz, tri, 3000), 30.0))
    for name, xyz, tri, cams, pts, md in cases:
        v = orc.visibility_graph(xyz, tri, cams, pts, md, want_flags=True)
        mt = orc.occluded_mt(xyz, tri, cams, pts, v)
        differ = mt != v.cand_occluded
        ray_flags = v.cand_flags & 7          # edge | graze | end point
        unflagged = int(np.count_nonzero(differ & (ray_flags == 0)))
        with capsys.disabled():
            print(f"\n  MT A/B {name}: {v.n_candidates} rays, {int(differ.sum())} disagree "
                  f"({int(np.count_nonzero(differ & ((ray_flags & 4) != 0)))} end point, "
                  f"{int(np.count_nonzero(differ & ((ray_flags & 1) != 0)))} edge, "
                  f"{int(np.count_nonzero(differ & ((ray_flags & 2) != 0)))} grazing), {unflagged} unflagged")
        assert unflagged == 0, f"{name}: the two predicates disagree on {unflagged} rays that carry no flag"
        assert v.n_candidates > 0

"""Input generation on the GPU: world points on the mesh,
generate_world_points_uniform (src/generate.rs:356-420), against the oracle's identical Philox
candidate stream (exact) and against the reference's distribution (area-weighted, uniform in the
triangle, within max_dist of a camera)."""
import numpy as np
import pytest

from conftest import procedural_scene, random_cameras

pytestmark = pytest.mark.gpu


def test_points_match_oracle_exactly(c2b, ctx, orc):
    from city2ba_b200.generate import generate_world_points_uniform
    rng = np.random.default_rng(5)
    xyz, tri = procedural_scene(1)
    cams = random_cameras(rng, 40, spread=12.0)
    for seed, n, md in ((1, 5000, 6.0), (2, 300, 100.0), (3, 2000, 4.0)):
        got = generate_world_points_uniform(xyz, tri, cams, n, md, seed=seed, ctx=ctx)
        want = orc.generate_world_points_uniform(xyz, tri, cams, n, md, seed)
        assert got.shape == (n, 3) and np.array_equal(got, want)
    # more than 10 n rejections before n acceptances: both sides give up (the reference panics)
    assert len(orc.generate_world_points_uniform(xyz, tri, cams, 2000, 2.5, 3)) == 0
    with pytest.raises(c2b.C2BError, match="Failed to generate enough points"):
        generate_world_points_uniform(xyz, tri, cams, 2000, 2.5, seed=3, ctx=ctx)
    again = generate_world_points_uniform(xyz, tri, cams, 5000, 6.0, seed=1, ctx=ctx)
    other = generate_world_points_uniform(xyz, tri, cams, 5000, 6.0, seed=7, ctx=ctx)
    assert np.array_equal(again, orc.generate_world_points_uniform(xyz, tri, cams, 5000, 6.0, 1))
    assert not np.array_equal(other, again)


def test_points_distribution_and_predicates(c2b, ctx, orc):
    """every point lies on the mesh and within max_dist of a camera; with max_dist = inf the share of
    points per triangle follows the triangle areas (chi-square well inside its 5-sigma band)"""
    from city2ba_b200.generate import generate_world_points_uniform
    rng = np.random.default_rng(6)
    xyz, tri = procedural_scene(2)
    cams = random_cameras(rng, 25, spread=10.0)
    pts = generate_world_points_uniform(xyz, tri, cams, 20000, 5.0, seed=11, ctx=ctx)
    cen = np.array([orc.center(c) for c in cams])
    d = np.linalg.norm(pts[:, None, :] - cen[None], axis=2).min(axis=1)
    assert np.all(d <= 5.0)
    valid = tri[(tri[:, 0] != tri[:, 1]) & (tri[:, 1] != tri[:, 2]) & (tri[:, 0] != tri[:, 2])]
    a, b, c = (xyz[valid[:, k]].astype(np.float64) for k in range(3))
    # distance of every point to its nearest triangle plane-and-inside test (barycentric), tolerance 1e-9
    n = np.cross(b - a, c - a)
    area = 0.5 * np.linalg.norm(n, axis=1)
    allp = generate_world_points_uniform(xyz, tri, cams, 60000, float("inf"), seed=12, ctx=ctx)
    unit = n / np.linalg.norm(n, axis=1, keepdims=True)
    owner = np.full(len(allp), -1)
    for t in range(len(valid)):
        rel = allp - a[t]
        off = np.abs(rel @ unit[t])
        # barycentric coordinates
        e1, e2 = b[t] - a[t], c[t] - a[t]
        d11, d12, d22 = e1 @ e1, e1 @ e2, e2 @ e2
        h1, h2 = rel @ e1, rel @ e2
        den = d11 * d22 - d12 * d12
        v, w = (d22 * h1 - d12 * h2) / den, (d11 * h2 - d12 * h1) / den
        inside = (off < 1e-6) & (v > -1e-9) & (w > -1e-9) & (v + w < 1 + 1e-9)
        owner[inside & (owner < 0)] = t
    assert np.all(owner >= 0)                       # every point lies on some triangle
    big = area > 0.02 * area.sum()                  # triangles with enough expected hits for the statistic
    cnt = np.bincount(owner, minlength=len(valid))[big]
    exp = len(allp) * area[big] / area.sum()
    # coplanar neighbours can swap ownership in this reconstruction, so compare with a generous band
    assert np.all(np.abs(cnt - exp) < 6 * np.sqrt(exp) + 0.02 * exp)


def test_points_error_paths(c2b, ctx):
    from city2ba_b200.generate import generate_world_points_uniform
    xyz, tri = procedural_scene(0)
    cams = random_cameras(np.random.default_rng(0), 3, spread=1.0)
    with pytest.raises(c2b.C2BError, match="0 cameras"):
        generate_world_points_uniform(xyz, tri, cams[:0], 10, 5.0, seed=1, ctx=ctx)
    far = cams.copy()
    far[:, 9:12] += 1e6                              # cameras nowhere near the mesh
    with pytest.raises(c2b.C2BError, match="Failed to generate enough points"):
        generate_world_points_uniform(xyz, tri, far, 100, 1.0, seed=1, ctx=ctx)


def test_camera_generators_python_mirror(c2b, ctx, tmp_path):
    """src/generate.rs:109-280 through the Python mirror on the small OBJ scene (the C++ mirror's counterpart is
    tests/cpp/test_host.cpp::generate_cameras): path cameras sit on the poly-line and look along it, path-step
    cameras are evenly spaced, Poisson cameras stand `height` above the tallest surface below them"""
    from city2ba_b200 import generate
    from test_host_mirror import _write_obj
    obj = str(tmp_path / "scene.obj")
    _write_obj(obj)
    models = generate.load_obj(obj)
    path, models = models[3], models[:3]
    scene = c2b.Scene(*generate.concat_models(models), ctx=ctx)
    assert scene.num_triangles == 2 + 12 + 3
    cams = generate.generate_cameras_path(scene, path, 40, seed=3)
    assert cams.shape == (40, 15)
    for rec in cams:
        cam = c2b.SnavelyCamera(record=rec)
        p = cam.center()
        assert abs(p[1] - 0.5) < 1e-12 and abs(p[2] - 3.0) < 1e-12 and -3.0 - 1e-12 <= p[0] <= 2.0 + 1e-12
        ahead = cam.project_world(p + np.array([1.0, 0.0, 0.0]))
        assert np.allclose(ahead, [0.0, 0.0, -1.0], atol=1e-9)          # +x along the path is straight ahead (-z)
    stepped = generate.generate_cameras_path_step(scene, path, 20, 0.2)
    xs = [c2b.SnavelyCamera(record=r).center()[0] for r in stepped]
    assert np.allclose(xs, -3.0 + 0.2 * np.arange(20), atol=1e-9)
    with pytest.raises(AssertionError):
        generate.generate_cameras_path_step(scene, path, 100, 0.2)        # 20 > 5: the reference's assert!
    poisson = generate.generate_cameras_poisson(scene, 100, 1.0, 10.0, seed=21)
    assert 30 < len(poisson) <= 200
    on_roof = 0
    for rec in poisson:
        p = c2b.SnavelyCamera(record=rec).center()
        over_cube = abs(p[0]) < 1.0 and abs(p[2]) < 1.0
        on_roof += over_cube
        if over_cube or abs(p[0]) > 1.6 or abs(p[2]) > 1.1:
            assert abs(p[1] - (3.0 if over_cube else 1.0)) < 1e-5
    assert on_roof > 0
    assert all(c2b.SnavelyCamera(record=r).center()[2] < 0.0 for r in generate.generate_cameras_poisson(scene, 100, 1.0, 0.0, seed=22))
    with_intr = generate.modify_intrinsics(poisson, (1.0, -0.1, 0.0), (2.0, 0.1, 0.0), seed=1)
    assert ((with_intr[:, 12] >= 1.0) & (with_intr[:, 12] < 2.0)).all() and (with_intr[:, 14] == 0.0).all()

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def c2b():
    import city2ba_b200
    return city2ba_b200


@pytest.fixture(scope="session")
def ctx(c2b):
    return c2b.context(0)


def procedural_scene(seed=0):
    """A small OBJ-like scene in the spirit of the reference's test scene: a ground plane, a box,
    a cone and a cylinder, polygons fan-triangulated the way tobj does (a,b,c),(a,c,d),...
    plus two degenerate `l`-style index triples.  Returns (xyz f32 (nv,3), tri u32 (nt,3))."""
    rng = np.random.default_rng(seed)
    verts, tris = [], []

    def add(vs, faces):
        base = len(verts)
        verts.extend(vs)
        for f in faces:
            for k in range(1, len(f) - 1):
                tris.append((base + f[0], base + f[k], base + f[k + 1]))

    s = 40.0
    add([(-s, 0, -s), (s, 0, -s), (s, 0, s), (-s, 0, s)], [(0, 1, 2, 3)])
    cx, cz, h = 5.0, -3.0, 2.0
    add([(cx - 1, 0, cz - 1), (cx + 1, 0, cz - 1), (cx + 1, 0, cz + 1), (cx - 1, 0, cz + 1),
         (cx - 1, h, cz - 1), (cx + 1, h, cz - 1), (cx + 1, h, cz + 1), (cx - 1, h, cz + 1)],
        [(0, 1, 2, 3), (4, 7, 6, 5), (0, 4, 5, 1), (1, 5, 6, 2), (2, 6, 7, 3), (3, 7, 4, 0)])
    n = 16
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    ring = [(-6 + 1.5 * np.cos(a), 0.0, 4 + 1.5 * np.sin(a)) for a in ang]
    add(ring + [(-6.0, 3.0, 4.0)], [tuple(range(n))] + [(i, (i + 1) % n, n) for i in range(n)])
    lo = [(8 + np.cos(a), 0.0, 9 + np.sin(a)) for a in ang]
    hi = [(8 + np.cos(a), 2.5, 9 + np.sin(a)) for a in ang]
    add(lo + hi, [tuple(range(n)), tuple(range(n, 2 * n))] +
        [(i, (i + 1) % n, n + (i + 1) % n, n + i) for i in range(n)])
    xyz = np.array(verts, dtype=np.float32) + rng.normal(0, 1e-3, (len(verts), 3)).astype(np.float32)
    tri = np.array(tris + [(0, 1, 1), (2, 2, 3)], dtype=np.uint32)
    return xyz, tri


def random_cameras(rng, n, center=(0, 1, 0), spread=15.0):
    """n SnavelyCamera records with random yaw/pitch, positions around `center`."""
    from oracle import oracle as o
    cams = np.empty((n, 15))
    for i in range(n):
        pos = np.array(center) + rng.uniform(-spread, spread, 3) * np.array([1, 0.1, 1])
        R = o.from_axis_angle([0, 1, 0], rng.uniform(0, 2 * np.pi))
        Rx = o.from_axis_angle([1, 0, 0], rng.uniform(-0.4, 0.4))
        M = (R.reshape(3, 3).T @ Rx.reshape(3, 3).T).T.reshape(9)
        cams[i] = o.from_position_direction(pos, M)
        cams[i, 12:15] = (rng.uniform(0.8, 1.5), rng.uniform(-0.05, 0.05), rng.uniform(-0.01, 0.01))
    return cams


def points_on_mesh(rng, xyz, tri, n):
    """area-weighted uniform samples on the triangles (src/generate.rs:370-408 style)."""
    t = tri[(tri[:, 0] != tri[:, 1]) & (tri[:, 1] != tri[:, 2]) & (tri[:, 0] != tri[:, 2])]
    a, b, c = (xyz[t[:, k]].astype(np.float64) for k in range(3))
    area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)
    pick = rng.choice(len(t), size=n, p=area / area.sum())
    r1, r2 = rng.uniform(size=n), rng.uniform(size=n)
    flip = r1 + r2 > 1
    r1[flip], r2[flip] = 1 - r1[flip], 1 - r2[flip]
    return a[pick] + r1[:, None] * (b[pick] - a[pick]) + r2[:, None] * (c[pick] - a[pick])


def bits_equal(a, b):
    """f64 arrays equal BIT FOR BIT (np.array_equal calls +0.0 and -0.0 equal and a NaN unequal to itself; the
    synthetic cities produce v = +-0 for every observation, so the sign of zero is part of the result)"""
    a = np.ascontiguousarray(a, np.float64).reshape(-1)
    b = np.ascontiguousarray(b, np.float64).reshape(-1)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def assert_same_graph(g, ref, what=""):
    assert np.array_equal(np.asarray(g.offsets, np.uint64), ref.offsets), f"{what}: CSR offsets differ"
    assert np.array_equal(np.asarray(g.point_idx, np.uint64), ref.point_idx), f"{what}: indices differ"
    assert np.array_equal(np.asarray(g.uv), ref.uv), f"{what}: projections differ (bit-exact expected)"
    assert bits_equal(g.uv, ref.uv), f"{what}: projections differ in the sign of a zero"

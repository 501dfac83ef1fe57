"""Golden fixtures on the reference's OWN test meshes (BASELINE config 1:
`city2ba generate test_scene.obj --cameras 100 --points 200`, and tests/box.obj of tests/main.rs).

    python tests/golden/make_golden_obj.py        # needs /root/reference (build container only)

The OBJ files are read where they lie under /root/reference; only derived arrays are stored
(vertex / index arrays after tobj-style triangulation, the sampled cameras and points, and the
oracle's visibility graph), so the GPU box — which has no /root/reference — can run the parity
test.  The reference's sampling is unseeded (thread_rng), so cameras and points are drawn here by
seeded restatements of its procedures:
  * mesh: tobj 0.1.12 rules — quads (a,b,c),(a,c,d), n-gons as a fan from vertex 0, `l` records as
    index pairs appended to the index list (they become degenerate triples, src/generate.rs:74-105);
  * cameras: generate_cameras_poisson (src/generate.rs:217-280) with uniform (x, z) samples in place
    of the Poisson-disk set: a ray straight down from above the scene bounds, camera `height` above
    the hit, kept if pt[2] < lower_y + ground (the reference compares z, not y — kept as is), random
    yaw about y;
  * points: generate_world_points_uniform (src/generate.rs:356-420) through the oracle's seeded stream.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as o  # noqa: E402

REF = "/root/reference"


def load_obj(path):
    """(xyz f32 (nv,3), tri u32 (nt,3)) with tobj's triangulation; all objects concatenated."""
    v, idx = [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            v.append([float(x) for x in t[1:4]])
        elif t[0] == "f":
            f = [int(x.split("/")[0]) - 1 for x in t[1:]]
            for k in range(1, len(f) - 1):
                idx += [f[0], f[k], f[k + 1]]
        elif t[0] == "l":
            idx += [int(x.split("/")[0]) - 1 for x in t[1:3]]
    idx = idx[: len(idx) // 3 * 3]                      # num_tri = indices.len() / 3
    return np.array(v, np.float32), np.array(idx, np.uint32).reshape(-1, 3)


def closest_down(xyz, tri, org):
    """distance of the first hit of a ray from `org` straight down, brute force in f64; inf = miss"""
    best = np.inf
    d = np.array([0.0, -1.0, 0.0])
    for a, b, c in tri:
        if a == b or b == c or a == c:
            continue
        A, B, C = (xyz[k].astype(np.float64) for k in (a, b, c))
        e1, e2 = B - A, C - A
        p = np.cross(d, e2)
        det = e1 @ p
        if abs(det) < 1e-14:
            continue
        s = org - A
        u = (s @ p) / det
        q = np.cross(s, e1)
        w = (d @ q) / det
        t = (e2 @ q) / det
        if u >= 0 and w >= 0 and u + w <= 1 and t > 0:
            best = min(best, t)
    return best


def cameras_like_poisson(xyz, tri, num, height, ground, rng):
    valid = tri[(tri[:, 0] != tri[:, 1]) & (tri[:, 1] != tri[:, 2]) & (tri[:, 0] != tri[:, 2])]
    lo, hi = xyz[valid.ravel()].min(0).astype(np.float64), xyz[valid.ravel()].max(0).astype(np.float64)
    start = np.array([hi[0], hi[1] + 0.1, hi[2]])
    delta = np.array([hi[0] - lo[0], 0.0, hi[2] - lo[2]])
    cams = []
    for s in rng.uniform(size=(2 * num, 2)):
        origin = start - delta * np.array([s[0], 0.0, s[1]])
        t = closest_down(xyz, tri, origin.astype(np.float32).astype(np.float64))
        if not np.isfinite(t):
            continue
        pt = origin + np.array([0.0, -1.0, 0.0]) * float(np.float32(t)) + np.array([0.0, height, 0.0])
        if pt[2] < lo[1] + ground:
            cams.append(o.from_position_direction(pt, o.from_angle_y(rng.uniform(0.0, 2.0 * np.pi))))
    return np.array(cams).reshape(-1, 15)


def make(name, path, num_cameras, num_points, max_dist, ground, seed):
    rng = np.random.default_rng(seed)
    xyz, tri = load_obj(path)
    cams = cameras_like_poisson(xyz, tri, num_cameras, 1.0, ground, rng)
    pts = o.generate_world_points_uniform(xyz, tri, cams, num_points, max_dist, seed)
    assert len(pts) == num_points and len(cams) > 0
    v = o.visibility_graph(xyz, tri, cams, pts, max_dist, want_flags=True)
    np.savez_compressed(os.path.join(HERE, name), xyz=xyz, tri=tri, cams=cams, pts=pts, max_dist=max_dist,
                        offsets=v.offsets, idx=v.point_idx.astype(np.uint32), uv=v.uv,
                        cand_offsets=v.cand_offsets, cand_idx=v.cand_point.astype(np.uint32),
                        cand_occluded=v.cand_occluded,
                        flag_counts=np.array([v.n_flag_edge, v.n_flag_graze, v.n_flag_endpoint, v.n_flag_cull]))
    print(name, "vertices", len(xyz), "index triples", len(tri), "cameras", len(cams), "points", len(pts),
          "rays", v.n_candidates, "observations", v.n_obs,
          "flags edge/graze/end/cull", v.n_flag_edge, v.n_flag_graze, v.n_flag_endpoint, v.n_flag_cull)


if __name__ == "__main__":
    # BASELINE config 1 (defaults of src/bin/city2ba.rs: max-dist 100, ground 0, height 1)
    make("cfg1_test_scene.npz", os.path.join(REF, "test_scene.obj"), 100, 200, 100.0, 0.0, 20261017)
    # tests/main.rs:86-128 runs generate on tests/box.obj (bbox +-4); --ground -1.0 variant keeps nothing
    # above z = lower_y - 1, so the default ground is used here with a larger camera request
    make("box_obj.npz", os.path.join(REF, "tests", "box.obj"), 60, 300, 100.0, 0.0, 7)

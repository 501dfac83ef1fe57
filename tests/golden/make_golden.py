"""Generates the golden fixtures in this directory from the CPU oracle.

    python tests/golden/make_golden.py

The reference (Rust + Embree) cannot be run in this image and ships no golden visibility
vectors, so these fixtures pin the ORACLE's behaviour (and through it the CUDA path), not the
reference binary's: see DESIGN.md "parity".  Re-run only when the oracle's stated predicate
changes; the diff of the .npz files is then part of the review.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import oracle as o  # noqa: E402
from conftest import points_on_mesh, procedural_scene, random_cameras  # noqa: E402


def save(name, **kw):
    np.savez_compressed(os.path.join(HERE, name), **kw)
    print(name, {k: getattr(v, "shape", v) for k, v in kw.items()})


def main():
    # cfg2: `city2ba synthetic --blocks 4` lattice (800 cameras, 2400 points), max_dist 10
    cams, pts = o.grid_cameras(10, 4), o.grid_points(10, 4)
    xyz, tri = o.city_mesh(4)
    m = o.visibility_graph(xyz, tri, cams, pts, 10.0, want_flags=True)
    a = o.synthetic_visibility(cams, pts, 10.0, True)
    g = o.visibility_graph(xyz, tri, cams, pts, 10.0, endpoint_guard_rel=True)
    save("cfg2_blocks4.npz", mesh_offsets=m.offsets, mesh_idx=m.point_idx.astype(np.uint32),
         mesh_uv=m.uv, cand_offsets=m.cand_offsets, cand_idx=m.cand_point.astype(np.uint32),
         cand_occluded=m.cand_occluded, cand_flags=m.cand_flags,
         flag_counts=np.array([m.n_flag_edge, m.n_flag_graze, m.n_flag_endpoint, m.n_flag_cull]),
         analytic_offsets=a.offsets, analytic_idx=a.point_idx.astype(np.uint32),
         guard_offsets=g.offsets, guard_idx=g.point_idx.astype(np.uint32))
    # cfg1-shaped: OBJ-like scene, 60 cameras, 400 points on the mesh, max_dist 100
    rng = np.random.default_rng(1234)
    xyz, tri = procedural_scene(0)
    cams = random_cameras(rng, 60)
    pts = points_on_mesh(rng, xyz, tri, 400)
    v = o.visibility_graph(xyz, tri, cams, pts, 100.0, want_flags=True)
    save("cfg1_scene.npz", xyz=xyz, tri=tri, cams=cams, pts=pts, offsets=v.offsets,
         idx=v.point_idx.astype(np.uint32), uv=v.uv, cand_offsets=v.cand_offsets,
         cand_idx=v.cand_point.astype(np.uint32), cand_occluded=v.cand_occluded,
         flag_counts=np.array([v.n_flag_edge, v.n_flag_graze, v.n_flag_endpoint, v.n_flag_cull]))
    # noise: deterministic drift (std 0) and a seeded Gaussian pass on the cfg2 lattice
    cams, pts = o.grid_cameras(10, 4), o.grid_points(10, 4)
    dc, dp = o.add_drift_normalized(cams, pts, 0.001, 0.0, 0.0, 1)
    nc, npts, nuv = o.add_noise(cams, pts, m.uv, 0.01, 0.0001, 0.01, 0.001, 42)
    save("cfg2_noise.npz", drift_cams=dc, drift_pts=dp, noise_cams=nc, noise_pts=npts, noise_uv=nuv,
         philox=np.array([o.philox4x32_10([0, 0, 0, 0], [0, 0]),
                          o.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2),
                          o.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                                          [0xa4093822, 0x299f31d0])]))


if __name__ == "__main__":
    main()

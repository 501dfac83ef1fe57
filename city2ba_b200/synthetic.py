"""Host mirror of city2ba::synthetic (reference: src/synthetic.rs).

The lattices come from the C ABI's host generators; visibility runs on the GPU.  The reference's
`synthetic_grid` decides occlusion with an analytic 2-D wall test (src/synthetic.rs:52-124,
`occlusion="analytic"`, the default here for drop-in parity); `occlusion="mesh"` casts rays at
the city-block box mesh instead (the north-star configuration).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import CAM_STRIDE, check, lib
from .baproblem import BAProblem
from .generate import Scene, visibility_graph

_pd = C.POINTER(C.c_double)


def grid_cameras(num_cameras_per_block, num_blocks, block_length, camera_height):
    """src/synthetic.rs:178-210 -> (C,15) camera records"""
    n = int(lib().c2b_grid_num_cameras(num_cameras_per_block, num_blocks))
    out = np.empty((n, CAM_STRIDE))
    check(lib().c2b_grid_cameras(num_cameras_per_block, num_blocks, block_length, camera_height,
                                 out.ctypes.data_as(_pd)))
    return out


def grid_points(num_points_per_block, num_blocks, block_length, block_inset, point_height):
    """src/synthetic.rs:213-258 -> (P,3)"""
    n = int(lib().c2b_grid_num_points(num_points_per_block, num_blocks))
    out = np.empty((n, 3))
    check(lib().c2b_grid_points(num_points_per_block, num_blocks, block_length, block_inset,
                                point_height, out.ctypes.data_as(_pd)))
    return out


def city_mesh(num_blocks, block_length=20.0, block_inset=1.0, height=10.0):
    """One box per city block (8 vertices, 12 triangles): the synthetic triangle scene."""
    nb = num_blocks * num_blocks
    xyz = np.empty((8 * nb, 3), np.float32)
    tri = np.empty((12 * nb, 3), np.uint32)
    check(lib().c2b_city_mesh(num_blocks, block_length, block_inset, height,
                              xyz.ctypes.data_as(C.POINTER(C.c_float)),
                              tri.ctypes.data_as(C.POINTER(C.c_uint32))))
    return xyz, tri


def city_mesh_tessellated(num_blocks, k, block_length=20.0, block_inset=1.0, height=10.0, amplitude=0.05):
    """The city-block boxes with every wall and roof tessellated into k x k quads (2 k^2 triangles per
    face, 10 k^2 per block) and the INTERIOR vertices of each face displaced along the face normal by
    a deterministic hash of their lattice position (borders stay put, so the boxes remain closed).
    This is BASELINE config 5's "large OBJ city mesh": 64 blocks x k = 11 gives 4.96 M triangles.
    A new artefact (the reference has no synthetic mesh); plain numpy, no RNG."""
    n, L, ins, H = int(num_blocks), float(block_length), float(block_inset), float(height)
    g = np.arange(k + 1, dtype=np.float64) / k
    a, b = np.meshgrid(g, g, indexing="ij")                      # face parameters in [0, 1]^2
    interior = ((a > 0) & (a < 1) & (b > 0) & (b < 1)).astype(np.float64)
    ia, ib = np.meshgrid(np.arange(k + 1), np.arange(k + 1), indexing="ij")
    vid = np.arange((k + 1) * (k + 1)).reshape(k + 1, k + 1)
    q00, q10, q11, q01 = vid[:-1, :-1].ravel(), vid[1:, :-1].ravel(), vid[1:, 1:].ravel(), vid[:-1, 1:].ravel()
    face_tri = np.concatenate([np.stack([q00, q10, q11], 1), np.stack([q00, q11, q01], 1)])
    bx, bz = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    bx, bz = bx.ravel()[:, None, None], bz.ravel()[:, None, None]  # (blocks, 1, 1)
    x0, x1 = bx * L + ins, (bx + 1) * L - ins
    z0, z1 = bz * L + ins, (bz + 1) * L - ins
    A, B = a[None], b[None]

    def disp(face):
        # integer hash of (block, face, lattice position) -> [-1, 1]
        h = (bx * 73856093) ^ (bz * 19349663) ^ (face * 83492791) ^ (ia[None] * 2654435761) ^ (ib[None] * 40503)
        h = (h ^ (h >> 13)) * 1274126177 & 0xffffffff
        return ((h & 0xffff) / 32767.5 - 1.0) * amplitude * interior[None]

    faces = []
    one = np.ones_like(A)
    faces.append((x0 + A * (x1 - x0), B * H * one, z0 * one - disp(0)))            # z lo wall
    faces.append((x0 + A * (x1 - x0), B * H * one, z1 * one + disp(1)))            # z hi wall
    faces.append((x0 * one - disp(2), B * H * one, z0 + A * (z1 - z0)))            # x lo wall
    faces.append((x1 * one + disp(3), B * H * one, z0 + A * (z1 - z0)))            # x hi wall
    faces.append((x0 + A * (x1 - x0), H * one + disp(4), z0 + B * (z1 - z0)))      # roof
    nv_face = (k + 1) * (k + 1)
    xyz = np.empty((n * n, 5, nv_face, 3), np.float32)
    for f, (X, Y, Z) in enumerate(faces):
        xyz[:, f, :, 0] = np.broadcast_to(X, (n * n, k + 1, k + 1)).reshape(n * n, -1)
        xyz[:, f, :, 1] = np.broadcast_to(Y, (n * n, k + 1, k + 1)).reshape(n * n, -1)
        xyz[:, f, :, 2] = np.broadcast_to(Z, (n * n, k + 1, k + 1)).reshape(n * n, -1)
    base = (np.arange(n * n * 5, dtype=np.int64) * nv_face)[:, None, None]
    tri = (face_tri[None].astype(np.int64) + base).reshape(-1, 3).astype(np.uint32)
    return xyz.reshape(-1, 3), tri


def synthetic_grid(num_cameras_per_block, num_points_per_block, num_blocks, block_length,
                   block_inset, camera_height, point_height, max_dist, verbose=False, *,
                   occlusion="analytic", building_height=10.0, cull=True, cull_mode="grid",
                   ctx=None) -> BAProblem:
    """src/synthetic.rs:163-300"""
    assert block_inset * 2.0 < block_length, (
        "Block inset ({}) must be less than half the block length ({}), to not violate physical "
        "constraints.".format(block_inset, block_length))
    cams = grid_cameras(num_cameras_per_block, num_blocks, block_length, camera_height)
    pts = grid_points(num_points_per_block, num_blocks, block_length, block_inset, point_height)
    scene = None
    if occlusion == "mesh":
        scene = Scene(*city_mesh(num_blocks, block_length, block_inset, building_height), ctx=ctx)
    vis = visibility_graph(scene, cams, pts, max_dist, verbose, occlusion=occlusion,
                           block_length=block_length, block_inset=block_inset,
                           cull_mode=cull_mode, ctx=ctx)
    ba = BAProblem.from_visibility(cams, pts, vis)
    return ba.cull() if cull else ba


def line_cameras(num_cameras, length, camera_height):
    out = np.empty((num_cameras, CAM_STRIDE))
    check(lib().c2b_line_cameras(num_cameras, length, camera_height, out.ctypes.data_as(_pd)))
    return out


def line_points(num_points, length, point_offset, point_height):
    out = np.empty((num_points, 3))
    check(lib().c2b_line_points(num_points, length, point_offset, point_height,
                                out.ctypes.data_as(_pd)))
    return out


def synthetic_line(num_cameras, num_points, length, point_offset, camera_height, point_height,
                   max_dist, verbose=False, *, cull=True, ctx=None) -> BAProblem:
    """src/synthetic.rs:313-381 (no occlusion test)"""
    cams = line_cameras(num_cameras, length, camera_height)
    pts = line_points(num_points, length, point_offset, point_height)
    vis = visibility_graph(None, cams, pts, max_dist, verbose, occlusion="none", ctx=ctx)
    ba = BAProblem.from_visibility(cams, pts, vis)
    return ba.cull() if cull else ba

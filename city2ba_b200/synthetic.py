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

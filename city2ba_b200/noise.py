"""Host mirror of city2ba::noise for the elementwise passes (reference: src/noise.rs:35-177).

The reference draws from rand's thread_rng(), which cannot be seeded; here every function takes
a `seed` for the Philox4x32-10 stream (default: a fresh 64-bit seed from os.urandom, i.e. the
reference's "different every run" behaviour).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._lib import check, context, lib
from .baproblem import BAProblem
from .generate import VisGraph

_pd = C.POINTER(C.c_double)


def _seed(seed):
    return int.from_bytes(os.urandom(8), "little") if seed is None else int(seed) & (2 ** 64 - 1)


def _p(a):
    return a.ctypes.data_as(_pd)


def add_drift(bal: BAProblem, strength, angle_strength, std, dir, seed=None, ctx=None) -> BAProblem:
    """src/noise.rs:68-116"""
    ctx = ctx or context()
    cams, pts = bal.cameras.copy(), bal.points.copy()
    d = np.ascontiguousarray(dir, np.float64)
    check(lib().c2b_add_drift(ctx.handle, _p(cams), len(cams), _p(pts), len(pts), float(strength),
                              float(angle_strength), float(std), _p(d), _seed(seed)))
    return BAProblem(cams, pts, bal.vis_graph)


def add_drift_normalized(bal: BAProblem, strength, angle_strength, std, seed=None, ctx=None) -> BAProblem:
    """src/noise.rs:47-56"""
    ctx = ctx or context()
    cams, pts = bal.cameras.copy(), bal.points.copy()
    check(lib().c2b_add_drift_normalized(ctx.handle, _p(cams), len(cams), _p(pts), len(pts),
                                         float(strength), float(angle_strength), float(std),
                                         _seed(seed)))
    return BAProblem(cams, pts, bal.vis_graph)


def add_noise(bal: BAProblem, translation_std, rotation_std, point_std, observations_std,
              seed=None, ctx=None) -> BAProblem:
    """src/noise.rs:119-177"""
    ctx = ctx or context()
    cams, pts = bal.cameras.copy(), bal.points.copy()
    g = bal.vis_graph
    uv = np.ascontiguousarray(g.uv, np.float64).copy()
    check(lib().c2b_add_noise(ctx.handle, _p(cams), len(cams), _p(pts), len(pts), _p(uv), len(uv),
                              float(translation_std), float(rotation_std), float(point_std),
                              float(observations_std), _seed(seed)))
    return BAProblem(cams, pts, VisGraph(g.offsets, g.point_idx, uv))


def add_sin_noise(bal: BAProblem, dir, noise_dir, strength, frequency, ctx=None) -> BAProblem:
    """src/noise.rs:388-416"""
    ctx = ctx or context()
    cams, pts = bal.cameras.copy(), bal.points.copy()
    d = np.ascontiguousarray(dir, np.float64)
    nd = np.ascontiguousarray(noise_dir, np.float64)
    check(lib().c2b_add_sin_noise(ctx.handle, _p(cams), len(cams), _p(pts), len(pts), _p(d), _p(nd),
                                  float(strength), float(frequency)))
    return BAProblem(cams, pts, bal.vis_graph)


def last_timing(ctx=None):
    """device milliseconds of the last noise call: (upload, statistics + kernels, download)"""
    ctx = ctx or context()
    ms = (C.c_float * 3)()
    check(lib().c2b_noise_timing(ctx.handle, ms))
    return tuple(float(x) for x in ms)

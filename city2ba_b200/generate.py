"""Host mirror of city2ba::generate for the hot path (reference: src/generate.rs).

`Scene` stands where `embree_rs::CommittedScene` stood (src/bin/city2ba.rs:515-521) and
`visibility_graph` keeps the reference's signature and result order
(src/generate.rs:424-481); the work happens in libcity2ba_cuda.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import CAM_STRIDE, Obs, VisOptions, check, context, lib


def _cam_array(cameras) -> np.ndarray:
    """Accept an (C,15) array or a sequence of SnavelyCamera."""
    if isinstance(cameras, np.ndarray):
        a = np.ascontiguousarray(cameras, dtype=np.float64)
    elif len(cameras) and hasattr(cameras[0], "to_record"):
        a = np.stack([c.to_record() for c in cameras]) if len(cameras) else np.zeros((0, CAM_STRIDE))
    else:
        a = np.ascontiguousarray(cameras, dtype=np.float64)
    return a.reshape(-1, CAM_STRIDE)


def _pts_array(points) -> np.ndarray:
    return np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)


class Scene:
    """Triangle scene committed to a GPU LBVH (replaces Scene::new + model_to_geometry + commit)."""

    def __init__(self, xyz, tri, ctx: _lib.Context | None = None):
        self.ctx = ctx or context()
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        tri = np.ascontiguousarray(tri, dtype=np.uint32).reshape(-1, 3)
        self._h = C.c_void_p()
        check(lib().c2b_scene_create(
            self.ctx.handle, xyz.ctypes.data_as(C.POINTER(C.c_float)), xyz.shape[0],
            tri.ctypes.data_as(C.POINTER(C.c_uint32)), tri.shape[0], C.byref(self._h)))

    @classmethod
    def from_models(cls, models, ctx=None):
        """models: iterable of (positions f32 (nv,3), indices u32 (nt,3)) per OBJ object —
        the loop over `model_to_geometry` + `attach_geometry` (src/bin/city2ba.rs:517-520)."""
        xyz, tri, base = [], [], 0
        for pos, idx in models:
            pos = np.asarray(pos, dtype=np.float32).reshape(-1, 3)
            idx = np.asarray(idx, dtype=np.uint32).reshape(-1, 3)
            xyz.append(pos)
            tri.append(idx + np.uint32(base))
            base += pos.shape[0]
        if not xyz:
            return cls(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32), ctx)
        return cls(np.concatenate(xyz), np.concatenate(tri), ctx)

    @property
    def handle(self):
        return self._h

    @property
    def num_triangles(self) -> int:
        return int(lib().c2b_scene_num_triangles(self._h))

    @property
    def num_nodes(self) -> int:
        return int(lib().c2b_scene_num_nodes(self._h))

    def bounds(self):
        """(lower xyz, upper xyz) as f32 — CommittedScene::bounds(), src/generate.rs:237."""
        lo = np.zeros(3, np.float32)
        hi = np.zeros(3, np.float32)
        check(lib().c2b_scene_bounds(self._h, lo.ctypes.data_as(C.POINTER(C.c_float)),
                                     hi.ctypes.data_as(C.POINTER(C.c_float))))
        return lo, hi

    def occluded(self, org, dirs, tfar) -> np.ndarray:
        """Embree-shaped any-hit on explicit f32 rays: returns a bool array (True = occluded)."""
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        tfar = np.ascontiguousarray(tfar, np.float32).reshape(-1)
        n = org.shape[0]
        rays = np.zeros(n, dtype=np.dtype([
            ("org", np.float32, 3), ("tnear", np.float32), ("dir", np.float32, 3),
            ("time", np.float32), ("tfar", np.float32), ("mask", np.uint32), ("id", np.uint32),
            ("flags", np.uint32)]))
        assert rays.dtype.itemsize == 48
        rays["org"], rays["dir"], rays["tfar"] = org, dirs, tfar
        check(lib().c2b_occluded(self.ctx.handle, self._h,
                                 rays.ctypes.data_as(C.POINTER(_lib.Ray48)), n))
        return np.isneginf(rays["tfar"])

    def intersect(self, org, dirs, tfar=np.inf):
        """Embree-shaped closest hit on explicit f32 rays: (hit bool array, distance array; inf = miss)."""
        org = np.ascontiguousarray(org, np.float32).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = org.shape[0]
        rays = np.zeros(n, dtype=np.dtype([
            ("org", np.float32, 3), ("tnear", np.float32), ("dir", np.float32, 3),
            ("time", np.float32), ("tfar", np.float32), ("mask", np.uint32), ("id", np.uint32),
            ("flags", np.uint32)]))
        rays["org"], rays["dir"] = org, dirs
        rays["tfar"] = np.broadcast_to(np.asarray(tfar, np.float32), (n,))
        check(lib().c2b_intersect(self.ctx.handle, self._h, rays.ctypes.data_as(C.POINTER(_lib.Ray48)), n))
        hit = rays["flags"] != 0
        return hit, np.where(hit, rays["tfar"], np.float32(np.inf))

    def intersect1(self, org, direction):
        """Closest hit of one ray: (hit, distance) — scene.intersect, src/generate.rs:253-262."""
        o = np.ascontiguousarray(org, np.float32)
        d = np.ascontiguousarray(direction, np.float32)
        hit, t = C.c_int(0), C.c_float(0)
        check(lib().c2b_intersect1(self.ctx.handle, self._h, o.ctypes.data_as(C.POINTER(C.c_float)),
                                   d.ctypes.data_as(C.POINTER(C.c_float)), C.byref(hit), C.byref(t)))
        return bool(hit.value), float(t.value)

    def close(self):
        if self._h:
            lib().c2b_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class VisGraph:
    """CSR visibility graph.  Behaves like the reference's Vec<Vec<(usize,(f64,f64))>>:
    len(g) == cameras, g[i] -> list of (point_index, (u, v)) in ascending point index."""

    def __init__(self, offsets, point_idx, uv, stats=None):
        self.offsets = np.asarray(offsets, dtype=np.uint64)
        self.point_idx = np.asarray(point_idx, dtype=np.uint64)
        self.uv = np.asarray(uv, dtype=np.float64).reshape(-1, 2)
        self.stats = stats or {}

    def __len__(self):
        return len(self.offsets) - 1

    def __getitem__(self, i):
        a, b = int(self.offsets[i]), int(self.offsets[i + 1])
        return [(int(p), (float(u), float(v)))
                for p, (u, v) in zip(self.point_idx[a:b], self.uv[a:b])]

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    @property
    def num_observations(self) -> int:
        return int(self.offsets[-1])

    def counts(self):
        return np.diff(self.offsets.astype(np.int64))


def _options(cull_mode, occlusion, endpoint_guard_rel, count_traversal, block_length, block_inset):
    o = VisOptions()
    lib().c2b_vis_options_default(C.byref(o))
    o.cull_mode = {"grid": _lib.CULL_GRID, "exhaustive": _lib.CULL_EXHAUSTIVE}[cull_mode]
    o.occlusion = {"mesh": _lib.OCC_MESH, "none": _lib.OCC_NONE, "analytic": _lib.OCC_ANALYTIC}[occlusion]
    o.endpoint_guard_rel = int(bool(endpoint_guard_rel))
    o.count_traversal = int(bool(count_traversal))
    o.block_length = float(block_length)
    o.block_inset = float(block_inset)
    return o


def _stats(o: Obs) -> dict:
    return {k: getattr(o, k) for k, _ in Obs._fields_ if k not in ("offsets", "point_idx", "uv")}


def visibility_graph(scene, cameras, points, max_dist, verbose=False, *, cull_mode="grid",
                     occlusion="mesh", endpoint_guard_rel=False, count_traversal=False,
                     block_length=20.0, block_inset=1.0, ctx=None) -> VisGraph:
    """Compute the camera-point visibility graph (src/generate.rs:424-481).

    scene: Scene or None (only with occlusion != "mesh").  cameras: (C,15) records or
    SnavelyCamera list.  points: (P,3).  Returns a VisGraph (camera-major CSR).
    `verbose` is accepted for signature parity; there is no progress bar (the call is one
    kernel pipeline).
    """
    ctx = ctx or (scene.ctx if scene is not None else context())
    cams, pts = _cam_array(cameras), _pts_array(points)
    opt = _options(cull_mode, occlusion, endpoint_guard_rel, count_traversal, block_length, block_inset)
    out = Obs()
    check(lib().c2b_visibility_graph(
        ctx.handle, scene.handle if scene is not None else None, cams.ctypes.data, cams.shape[0],
        pts.ctypes.data, pts.shape[0], float(max_dist), C.byref(opt), C.byref(out)))
    Cn, O = int(out.n_cameras), int(out.n_obs)
    offsets = np.ctypeslib.as_array(out.offsets, shape=(Cn + 1,)).copy()
    if O:
        idx = np.ctypeslib.as_array(out.point_idx, shape=(O,)).copy()
        uv = np.ctypeslib.as_array(out.uv, shape=(2 * O,)).copy()
    else:
        idx, uv = np.zeros(0, np.uint32), np.zeros(0, np.float64)
    st = _stats(out)
    lib().c2b_obs_free(ctx.handle, C.byref(out))
    return VisGraph(offsets, idx, uv, st)


class ResidentProblem:
    """Keeps cameras / points / result in HBM between calls (the three-call form of the ABI)."""

    def __init__(self, ctx=None):
        self.ctx = ctx or context()

    def upload_points(self, points):
        p = _pts_array(points)
        check(lib().c2b_upload_points(self.ctx.handle, p.ctypes.data, p.shape[0]))

    def upload_cameras(self, cameras):
        c = _cam_array(cameras)
        check(lib().c2b_upload_cameras(self.ctx.handle, c.ctypes.data, c.shape[0]))

    def upload_points_ptr(self, ptr: int, n: int):
        check(lib().c2b_upload_points(self.ctx.handle, ptr, n))

    def upload_cameras_ptr(self, ptr: int, n: int):
        check(lib().c2b_upload_cameras(self.ctx.handle, ptr, n))

    def run(self, scene, max_dist, **kw) -> dict:
        opt = _options(kw.get("cull_mode", "grid"), kw.get("occlusion", "mesh"),
                       kw.get("endpoint_guard_rel", False), kw.get("count_traversal", False),
                       kw.get("block_length", 20.0), kw.get("block_inset", 1.0))
        out = Obs()
        check(lib().c2b_visibility_graph_resident(
            self.ctx.handle, scene.handle if scene is not None else None, float(max_dist),
            C.byref(opt), C.byref(out)))
        return _stats(out)

    def download(self) -> VisGraph:
        out = Obs()
        check(lib().c2b_download_obs(self.ctx.handle, C.byref(out)))
        Cn, O = int(out.n_cameras), int(out.n_obs)
        offsets = np.ctypeslib.as_array(out.offsets, shape=(Cn + 1,)).copy()
        idx = np.ctypeslib.as_array(out.point_idx, shape=(O,)).copy() if O else np.zeros(0, np.uint32)
        uv = np.ctypeslib.as_array(out.uv, shape=(2 * O,)).copy() if O else np.zeros(0)
        return VisGraph(offsets, idx, uv, _stats(out))

    def download_raw(self) -> Obs:
        """Download into the ctx's pinned buffers without copying to numpy (bench e2e)."""
        out = Obs()
        check(lib().c2b_download_obs(self.ctx.handle, C.byref(out)))
        return out

    def reprojection_error(self, norm: float) -> float:
        v = C.c_double(0)
        check(lib().c2b_reprojection_error_resident(self.ctx.handle, float(norm), C.byref(v)))
        return v.value


def generate_world_points_uniform(xyz, tri, cameras, num_points, max_dist, seed=None, ctx=None) -> np.ndarray:
    """src/generate.rs:356-420 on the GPU: `num_points` points on the mesh (area-weighted triangle,
    uniform inside it) that lie within `max_dist` of some camera.  `xyz`/`tri` are all models'
    vertices / index triples concatenated.  The reference draws from thread_rng(); here the sample is a
    pure function of `seed` (default: a fresh one from os.urandom)."""
    import os
    ctx = ctx or context()
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    tri = np.ascontiguousarray(tri, dtype=np.uint32).reshape(-1, 3)
    cams = _cam_array(cameras)
    seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed) & (2 ** 64 - 1)
    out = np.empty((int(num_points), 3))
    n = C.c_uint64(0)
    check(lib().c2b_generate_world_points_uniform(
        ctx.handle, xyz.ctypes.data_as(C.POINTER(C.c_float)), xyz.shape[0],
        tri.ctypes.data_as(C.POINTER(C.c_uint32)), tri.shape[0], cams.ctypes.data_as(C.POINTER(C.c_double)),
        cams.shape[0], int(num_points), float(max_dist), seed, out.ctypes.data_as(C.POINTER(C.c_double)),
        C.byref(n)))
    return out[: n.value]

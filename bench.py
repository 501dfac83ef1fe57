#!/usr/bin/env python
"""bench.py — camera x point visibility tests/s on a synthetic city grid (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg4|cfg2]
                    [--cull-mode grid|exhaustive] [--impl reference]

One "step" = one pass of the hot path (c2b_visibility_graph: cull -> sort -> BVH traversal ->
compaction) over the whole workload; cameras are sharded across ranks (contiguous ranges, the
mesh, BVH and points replicated), so total work is fixed as N grows ("strong" scaling on the
named configuration).  Prints ONE JSON line on rank 0.

  value : C*P / t with inputs resident in HBM, t = device time of the step (CUDA events on the
          library's stream), max over ranks.
  e2e   : the same metric through the host-buffer C-ABI call (pinned host inputs copied H2D and
          the CSR result copied D2H inside the timed region, wall clock around the call).
  --impl reference : the reference's algorithm on the box's host cores (the oracle's OpenMP +
          CPU-BVH arm; the Rust/Embree reference cannot be built in this image), on a bounded
          camera sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (blocks, cameras_per_block, points_per_block)  — BASELINE.md section 4
    "cfg2": (4, 10, 10),
    "cfg3": (16, 9, 306),
    "cfg4": (64, 6, 200),
    # config 5: the same 64x64 city with every wall tessellated 11 x 11 and displaced (4.96 M triangles),
    # 49,920 lattice cameras, 4,992,000 points sampled on the mesh (area-weighted, seeded)
    "cfg5": (64, 3, 100),
    "cfg5s": (16, 3, 100),  # a 16x16-block cut of cfg5 (310 k triangles) for quick runs
}
TESS_K = 11
MAX_DIST = 10.0
BLOCK_LENGTH, BLOCK_INSET, CAM_H, PT_H, BUILDING_H = 20.0, 1.0, 1.0, 1.0, 10.0


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_traffic(stage, workload, cull_mode):
    """dram__bytes_read.sum + dram__bytes_write.sum of the stage's kernel, per launch, from the
    committed `ncu --set full` capture (profiles/traffic.json, written by profiles/summarize.py);
    None when no capture matches this workload."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))
        e = t[f"{workload}/{cull_mode}/{stage}"]
        return float(e["dram_bytes"]), f'{e["kernel"]}: {e["source"]}'
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(device), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                   "reasons": sorted(reasons), "samples": len(sm)}
        return out


def build_workload(name):
    from city2ba_b200 import synthetic
    n, cpb, ppb = WORKLOADS[name]
    cams = synthetic.grid_cameras(cpb, n, BLOCK_LENGTH, CAM_H)
    if name.startswith("cfg5"):
        xyz, tri = synthetic.city_mesh_tessellated(n, TESS_K, BLOCK_LENGTH, BLOCK_INSET, BUILDING_H)
        pts = sample_points_on_walls(xyz, tri, int(synthetic.lib().c2b_grid_num_points(ppb, n)))
        return cams, pts, xyz, tri
    pts = synthetic.grid_points(ppb, n, BLOCK_LENGTH, BLOCK_INSET, PT_H)
    xyz, tri = synthetic.city_mesh(n, BLOCK_LENGTH, BLOCK_INSET, BUILDING_H)
    return cams, pts, xyz, tri


def sample_points_on_walls(xyz, tri, n, seed=0xC17B2A):
    """n world points on the mesh, area-weighted over the triangles that reach below 3 m (what
    street-level cameras can see), uniform inside each triangle — the rule of src/generate.rs:370-408 with a
    seeded generator (the reference's thread_rng cannot be seeded)."""
    rng = np.random.default_rng(seed)
    a, b, c = (xyz[tri[:, k]].astype(np.float64) for k in range(3))
    low = np.minimum(np.minimum(a[:, 1], b[:, 1]), c[:, 1]) <= 3.0
    a, b, c = a[low], b[low], c[low]
    area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)
    cdf = np.cumsum(area)
    pick = np.minimum(np.searchsorted(cdf, rng.uniform(0, cdf[-1], n)), len(cdf) - 1)
    r1, r2 = rng.uniform(size=n), rng.uniform(size=n)
    flip = r1 + r2 > 1
    r1[flip], r2[flip] = 1 - r1[flip], 1 - r2[flip]
    return np.ascontiguousarray(a[pick] + r1[:, None] * (b[pick] - a[pick]) + r2[:, None] * (c[pick] - a[pick]))


def shard(C, rank, world):
    return (rank * C) // world, ((rank + 1) * C) // world


def cpu_arm(cams, pts, xyz, tri, budget_s, threads=0):
    """Times the oracle's multithreaded CPU arm on a bounded camera sample (every k-th camera).
    Returns (tests_per_s, cores, sample description, obs_per_s)."""
    from oracle import oracle as orc
    C, P = len(cams), len(pts)
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if threads <= 0:
        threads = ncores  # explicit: torchrun exports OMP_NUM_THREADS=1, which would serialise the arm
    n = min(C, max(ncores, 8))
    for _ in range(4):  # grow the sample until it fills about the budget
        idx = np.linspace(0, C - 1, n).astype(np.int64)
        t0 = time.perf_counter()
        v, used = orc.ref_visibility_graph(xyz, tri, cams[idx], pts, MAX_DIST, n_threads=threads)
        dt = time.perf_counter() - t0
        if dt >= 0.5 * budget_s or n >= C:
            break
        n = int(min(C, max(n + used, 0.9 * n * budget_s / max(dt, 1e-6))))
        n = max(used, (n // used) * used)
    return n * P / dt, used, f"{n} of {C} cameras (evenly spaced) x all {P} points, {dt:.1f} s", v.n_obs / dt


def run_reference(args):
    """--impl reference: the CPU arm, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    cams, pts, xyz, tri = build_workload(args.workload)
    C, P = len(cams), len(pts)
    budget = 8.0
    vals, obs = [], []
    cores, sample = 1, ""
    for s in range(args.warmup + args.steps):
        v, cores, sample, o = cpu_arm(cams, pts, xyz, tri, budget)
        if s >= args.warmup:
            vals.append(v)
            obs.append(o)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "camera-point visibility tests/sec", "value": value,
        "unit": "tests/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * C * P / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: synthetic {WORKLOADS[args.workload][0]}x"
                               f"{WORKLOADS[args.workload][0]}-block city{', walls tessellated' if args.workload.startswith('cfg5') else ''}, "
                               f"{C} cameras x {P} points, max_dist {MAX_DIST}", "cameras": C, "points": P,
                   "triangles": int(len(tri)), "note": "reference binary (Rust + Embree 3.8) cannot be "
                   "built here; this is the oracle's C restatement of its algorithm: OpenMP over "
                   "cameras, brute-force point loop, CPU BVH any-hit; ms_per_step extrapolates the "
                   "sample to the whole workload"},
        "observations_per_s": float(np.mean(obs)),
        "cpu_baseline": {"value": value, "unit": "tests/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--cull-mode", default="grid", choices=["grid", "exhaustive"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-exhaustive", action="store_true")
    ap.add_argument("--no-noise", action="store_true")
    ap.add_argument("--replicate-points", action="store_true",
                    help="N > 1: every rank uploads all points itself instead of 1/N + NCCL all-gather")
    args = ap.parse_args()
    if args.workload is None:
        # the metric's target is quoted on the 64x64-block city (cfg4); it fits one GPU
        args.workload = "cfg4"
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import city2ba_b200 as c2b
    from city2ba_b200 import _lib
    from city2ba_b200.generate import ResidentProblem, _options

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = c2b.context(local)
    L = _lib.lib()

    cams, pts, xyz, tri = build_workload(args.workload)
    C, P = len(cams), len(pts)
    c0, c1 = shard(C, rank, world)
    my_cams = torch.from_numpy(np.ascontiguousarray(cams[c0:c1])).pin_memory()
    pin_pts = torch.from_numpy(pts).pin_memory()
    Cr = c1 - c0
    scene = c2b.Scene(xyz, tri, ctx=ctx)
    rp = ResidentProblem(ctx)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def flush_l2():
        flush.zero_()
        torch.cuda.synchronize()

    # ---- device-resident arm ("value") -------------------------------------------------------
    def resident_arm(mode, steps, warmup, count=False):
        rp.upload_points_ptr(pin_pts.data_ptr(), P)
        rp.upload_cameras_ptr(my_cams.data_ptr(), Cr)
        torch.cuda.synchronize()
        st = None
        tot = {k: 0.0 for k in ("ms_total", "ms_prep", "ms_cull", "ms_sort", "ms_traverse", "ms_compact")}
        for s in range(warmup + steps):
            if s == warmup:
                barrier()
            flush_l2()
            st = rp.run(scene, MAX_DIST, cull_mode=mode, count_traversal=count)
            if s >= warmup:
                for k in tot:
                    tot[k] += st[k]
        barrier()
        return tot, st

    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = L.c2b_kernel_launches()
    tot, st = resident_arm(args.cull_mode, args.steps, args.warmup)
    launches = L.c2b_kernel_launches() - launches0
    launches_per_step = launches // (args.steps + args.warmup)
    t_dev = tot["ms_total"] / 1e3

    # ---- end-to-end arm: host buffers -> C-ABI call -> host CSR ------------------------------------
    opt = _options(args.cull_mode, "mesh", False, False, BLOCK_LENGTH, BLOCK_INSET)
    import ctypes as Ct
    out = _lib.Obs()
    e2e_t = 0.0
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.zeros(1, dtype=torch.int64, device=dev)
    # N > 1: every rank uploads 1/N of the points over its own PCIe link and the shards are
    # all-gathered over NVLink (NCCL) instead of N full uploads — the path's point exchange.
    # (world == 1: the plain host-pointer call.)
    shard_pts = world > 1 and not args.replicate_points
    if shard_pts:
        per = -(-P // world)
        lo_p, hi_p = min(P, rank * per), min(P, (rank + 1) * per)
        pin_shard = torch.zeros(per * 3, dtype=torch.float64).pin_memory()
        pin_shard[:(hi_p - lo_p) * 3] = torch.from_numpy(pts[lo_p:hi_p].reshape(-1))
        d_shard = torch.empty(per * 3, dtype=torch.float64, device=dev)
        d_full = torch.empty(world * per * 3, dtype=torch.float64, device=dev)
    h2d_pts = 0
    for s in range(args.warmup + args.steps):
        if s == args.warmup:
            barrier()
        flush_l2()
        t0 = time.perf_counter()
        if shard_pts:
            d_shard.copy_(pin_shard, non_blocking=True)
            dist.all_gather_into_tensor(d_full, d_shard)
            torch.cuda.synchronize()
            _lib.check(L.c2b_upload_points_device(ctx.handle, d_full.data_ptr(), P))
            h2d_pts = per * 24
            _lib.check(L.c2b_visibility_graph(ctx.handle, scene.handle, my_cams.data_ptr(), Cr,
                                              None, P, MAX_DIST, Ct.byref(opt), Ct.byref(out)))
        else:
            _lib.check(L.c2b_visibility_graph(ctx.handle, scene.handle, my_cams.data_ptr(), Cr,
                                              pin_pts.data_ptr(), P, MAX_DIST, Ct.byref(opt), Ct.byref(out)))
        if world > 1:
            # per-rank observation counts -> global CSR offsets
            mine[0] = int(out.n_obs)
            dist.all_gather_into_tensor(counts, mine)
            counts_host = counts.cpu()
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            e2e_t += dt
    barrier()
    clocks = sampler.stop() if sampler else None
    h2d, d2h = int(out.h2d_bytes) + h2d_pts, int(out.d2h_bytes)
    n_obs_local = int(out.n_obs)
    e2e_stage = {k: float(getattr(out, k)) for k in ("ms_h2d", "ms_prep", "ms_cull", "ms_sort", "ms_traverse", "ms_compact", "ms_d2h")}

    # ---- noise pass on the generated problem (config 2 / 5: drift + Gaussian noise), rank 0 only ------
    noise = None
    if not args.no_noise and rank == 0:
        O = int(out.n_obs)
        pd_ = Ct.POINTER(Ct.c_double)
        # pinned host arrays (what a host program that cares about transfer time would hand over)
        t_uv = torch.from_numpy(np.ctypeslib.as_array(out.uv, shape=(2 * O,)).copy() if O else np.zeros(0)).pin_memory()
        t_cam = torch.from_numpy(np.ascontiguousarray(cams[c0:c1]).copy()).pin_memory()
        t_pts = torch.from_numpy(pts.copy()).pin_memory()
        uv, ncam, npts = t_uv.numpy(), t_cam.numpy(), t_pts.numpy()
        ms3 = (Ct.c_float * 3)()
        noise = {}
        for name, call, nbytes in (
            ("add_drift_normalized", lambda: L.c2b_add_drift_normalized(
                ctx.handle, ncam.ctypes.data_as(pd_), len(ncam), npts.ctypes.data_as(pd_), len(npts),
                0.001, 0.0, 0.0, 1), 2 * (120 * len(ncam) + 24 * len(npts)) + 3 * 24 * (len(ncam) + len(npts))),
            ("add_noise", lambda: L.c2b_add_noise(
                ctx.handle, ncam.ctypes.data_as(pd_), len(ncam), npts.ctypes.data_as(pd_), len(npts),
                uv.ctypes.data_as(pd_), O, 0.0, 0.0001, 0.01, 0.001, 42),
             2 * (120 * len(ncam) + 24 * len(npts) + 16 * O) + 2 * 24 * (len(ncam) + len(npts))),
        ):
            best_wall, kern = 1e9, 0.0
            for _ in range(3):
                t0 = time.perf_counter()
                _lib.check(call())
                best_wall = min(best_wall, time.perf_counter() - t0)
                _lib.check(L.c2b_noise_timing(ctx.handle, ms3))
                kern = float(ms3[1])
            noise[name] = {"elements": len(ncam) + len(npts) + (O if name == "add_noise" else 0),
                           "ms_host_to_host": 1e3 * best_wall, "ms_kernels": kern,
                           "algorithmic_bytes": nbytes, "kernel_GBps": nbytes / (kern * 1e-3) / 1e9 if kern > 0 else None,
                           "note": "pinned host arrays in and out (H2D + D2H inside ms_host_to_host); "
                                   "ms_kernels = statistics reductions + elementwise kernels (CUDA events)"}

    # ---- traversal counters (one untimed instrumented run) and the exhaustive arm ------------------
    _, stc = resident_arm(args.cull_mode, 1, 0, count=True)
    ex = None
    if not args.no_exhaustive and args.cull_mode == "grid":
        ex_steps = 2 if C * P > 2e11 else 5
        ex_tot, ex_st = resident_arm("exhaustive", ex_steps, 1)
        ex = (ex_tot, ex_st, ex_steps)

    # ---- reduce over ranks (max time, summed work) -----------------------------------------------------
    def rmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def rsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    t_dev_max = rmax(t_dev)
    e2e_max = rmax(e2e_t)
    n_obs = rsum(float(n_obs_local))
    n_cand = rsum(float(st["n_candidates"]))
    pairs_eval = rsum(float(st["pairs_evaluated"]))
    h2d_all, d2h_all = rsum(float(h2d)), rsum(float(d2h))
    stage_ms = {k: rmax(tot[k] / args.steps) for k in tot}
    nodes = rsum(float(stc["nodes_visited"]))
    tris_t = rsum(float(stc["tris_tested"]))
    if ex is not None:
        ex_t = rmax(ex[0]["ms_total"] / 1e3 / ex[2])
        ex_cull = rmax(ex[0]["ms_cull"] / ex[2])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_kind = measured_peaks()
    K = args.steps
    value = C * P * K / t_dev_max
    e2e_value = C * P * K / e2e_max
    # algorithmic bytes per stage (DESIGN.md "roofline bookkeeping"), whole job, per step
    n_words = (n_cand + 31) // 32
    nodes_per_cam_sum = nodes / max(1.0, n_cand / 32.0) * C if n_cand else 0.0  # ~ one list per camera
    if args.cull_mode == "grid":
        # fused grid schedule: plan -> one fused pass per camera -> sort + write
        b_cull = 144.0 * C + 4.0 * C * 2 + 4.0 * nodes_per_cam_sum  # plan: camera + centre, count + offset;
        #                                                            leaf lists written (4 B per entry)
        b_trav = (24.0 * pairs_eval + 4.0 * n_cand + 144.0 * C      # grid-ordered point per evaluated pair,
                  + 36.0 * nodes + 48.0 * tris_t + 4.0 * n_obs)     # index per candidate, camera; list entry (4)
        #            + leaf box (32) per entry examined, 48 B record per warp-triangle test, scratch index out
        b_sort = 0.0
        b_comp = 4.0 * n_obs + 24.0 * n_obs + 20.0 * n_obs + 128.0 * C
        #        scratch index read, point gather, CSR record written, camera record + offsets
    else:
        # exhaustive schedule: every pair -> pool (key, uv) -> radix sort -> ordered traversal -> compaction
        b_cull = 24.0 * pairs_eval / 64.0 + 144.0 * C + 28.0 * n_cand  # a 64-camera tile reads each point once
        b_trav = 56.0 * n_cand + 32.0 * nodes + 48.0 * tris_t + 4.0 * n_words
        b_sort = 24.0 * n_cand * max(1, -(-(int(np.ceil(np.log2(max(C // world, 2)))) + int(np.ceil(np.log2(P)))) // 8))
        b_comp = 12.0 * n_cand + 8.0 * n_words + 24.0 * n_obs * 2 + 8.0 * (C + 1)
    stages = {
        "cull": (b_cull, stage_ms["ms_cull"]), "sort": (b_sort, stage_ms["ms_sort"]),
        "traverse": (b_trav, stage_ms["ms_traverse"]), "compact": (b_comp, stage_ms["ms_compact"]),
    }
    dom = max(stages, key=lambda k: stages[k][1])
    ach = stages[dom][0] / (stages[dom][1] * 1e-3) / 1e9 if stages[dom][1] > 0 else 0.0
    kernel_of = ({"cull": "k_cam_plan + k_cam_trilist", "traverse": "k_visibility_fused (cull + ray build + occlusion)",
                  "compact": "k_sort_write", "sort": "-"} if args.cull_mode == "grid" else
                 {"cull": "k_cull_exhaustive", "sort": "k_rs_scatter (radix sort)", "traverse": "k_traverse",
                  "compact": "k_compact_write"})
    traffic, traffic_src = ncu_traffic(dom, args.workload, args.cull_mode)
    line = {
        "metric": "camera-point visibility tests/sec", "value": value, "unit": "tests/s",
        "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev_max / K,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": f"{args.workload}: synthetic {WORKLOADS[args.workload][0]}x"
                        f"{WORKLOADS[args.workload][0]}-block city{', walls tessellated' if args.workload.startswith('cfg5') else ''}, "
                        f"{C} cameras x {P} points, max_dist {MAX_DIST}",
            "cameras": C, "points": P, "triangles": int(len(tri)), "bvh_nodes": scene.num_nodes,
            "cull_mode": args.cull_mode, "parallelism": f"camera ranges over {world} GPU(s), mesh/BVH/points replicated"
                           + ("; e2e: points uploaded 1/N per rank + NCCL all-gather" if shard_pts else ""),
            "l2": "256 MB buffer written between timed steps (L2 flush)",
            "candidates": int(n_cand), "observations": int(n_obs),
            "pairs_evaluated_per_step": int(pairs_eval),
        },
        "observations_per_s": n_obs * K / t_dev_max,
        "evaluated_pairs_per_s": pairs_eval * K / t_dev_max,
        "e2e": {"value": e2e_value, "unit": "tests/s", "h2d_bytes_per_step": int(h2d_all),
                "d2h_bytes_per_step": int(d2h_all), "ms_per_step": 1e3 * e2e_max / K,
                "observations_per_s": n_obs * K / e2e_max, "stage_ms_last_step_rank0": e2e_stage},
        "gpu_launches": int(launches_per_step * K),
        "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
        "roofline": {
            "bound": "hbm", "stage": dom, "kernel": kernel_of[dom], "achieved": ach, "peak": peak * world, "unit": "GB/s",
            "frac": ach / (peak * world), "traffic": traffic, "traffic_source": traffic_src, "peak_kind": f"of {peak_kind} (MEASURED_PEAKS.json hbm_gbs x n_gpus)",
            "algorithmic_bytes_per_step": stages[dom][0],
            "all_stages_GBps": {k: (v[0] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else None) for k, v in stages.items()},
            "warp_node_visits": int(nodes), "warp_triangle_tests": int(tris_t),
            # the whole pass by BASELINE.md section 5's single formula (node visits and triangle tests are
            # per 32-ray packet here: one fetch serves the packet)
            "pipeline": (lambda b: {"algorithmic_bytes_per_step": b, "achieved": b / (t_dev_max / K) / 1e9,
                                    "frac": b / (t_dev_max / K) / 1e9 / (peak * world),
                                    "formula": "24*pairs_evaluated + 120*C + 64*node_visits + 48*triangle_tests + 24*O + 8*(C+1)"})(
                24.0 * pairs_eval + 120.0 * C + 64.0 * nodes + 48.0 * tris_t + 24.0 * n_obs + 8.0 * (C + 1)),
        },
        "clocks": clocks,
    }
    if noise is not None:
        line["noise"] = noise
    if ex is not None:
        line["exhaustive"] = {"value": C * P / ex_t, "unit": "tests/s", "ms_per_step": 1e3 * ex_t,
                              "cull_ms": ex_cull, "steps": ex[2],
                              "note": "every camera x point pair tested on the GPU (the reference's loop)"}
        # secondary bound: the exhaustive cull's hot loop issues 7 FP64-pipe instructions per
        # pair (3 DADD + DMUL + 2 DFMA + DSETP); the probe is a loop of independent DFMAs on the same device
        import ctypes as _C
        rate = _C.c_double(0.0)
        if _lib.lib().c2b_probe_fp64(ctx.handle, _C.byref(rate)) == 0 and rate.value > 0 and ex_cull > 0:
            per_rank_pairs = C * P / world
            line["exhaustive"]["fp64_probe_dfma_per_s"] = rate.value
            line["exhaustive"]["fp64_pipe_frac"] = 7.0 * per_rank_pairs / (ex_cull * 1e-3) / rate.value
    if not args.no_cpu_baseline and world == 1:
        from oracle import oracle as orc
        orc.build()
        v, cores, sample, _ = cpu_arm(cams, pts, xyz, tri, 12.0)
        line["cpu_baseline"] = {"value": v, "unit": "tests/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

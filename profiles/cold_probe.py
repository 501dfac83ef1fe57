"""Where a one-shot caller's time goes: a fresh process, no torch — c2b_init, scene builds, the first and the
following c2b_visibility_graph calls with pageable inputs (what `city2ba generate` or a Rust caller does).

    python profiles/cold_probe.py [--workload cfg4]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import city2ba_b200 as c2b  # noqa: E402
from city2ba_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg4")
    a = ap.parse_args()
    cams, pts, xyz, tri = bench.build_workload(a.workload)
    out = {"workload": a.workload}
    t0 = time.perf_counter()
    ctx = _lib.Context(0)
    out["c2b_init_ms (creates the CUDA context)"] = 1e3 * (time.perf_counter() - t0)
    for k in range(3):
        t0 = time.perf_counter()
        scene = c2b.Scene(xyz, tri, ctx=ctx)
        out[f"scene_build_{k}_ms"] = 1e3 * (time.perf_counter() - t0)
    for k in range(3):
        t0 = time.perf_counter()
        g = c2b.visibility_graph(scene, cams, pts, bench.MAX_DIST, ctx=ctx)
        out[f"visibility_graph_{k}_ms (incl. numpy copies of the result)"] = 1e3 * (time.perf_counter() - t0)
        out[f"visibility_graph_{k}_stage_ms"] = {s: round(float(g.stats[s]), 3) for s in
                                                 ("ms_h2d", "ms_prep", "ms_cull", "ms_traverse", "ms_compact", "ms_d2h")}
    t0 = time.perf_counter()
    ctx2 = _lib.Context(0)
    out["second_ctx_init_ms"] = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    scene2 = c2b.Scene(xyz, tri, ctx=ctx2)
    out["second_ctx_scene_build_ms"] = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    c2b.visibility_graph(scene2, cams, pts, bench.MAX_DIST, ctx=ctx2)
    out["second_ctx_first_call_ms"] = 1e3 * (time.perf_counter() - t0)
    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in out.items()}))


if __name__ == "__main__":
    main()

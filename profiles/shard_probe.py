"""One camera shard of a bench workload on ONE GPU, without torch.distributed (so that it can run under
ncu, which must never wrap a multi-rank command): resident passes of shard r of N, stage times printed.

    python profiles/shard_probe.py --workload cfg4 --shard 0/8 [--steps 5]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import city2ba_b200 as c2b  # noqa: E402
from city2ba_b200.generate import ResidentProblem  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg4")
    ap.add_argument("--shard", default="0/8")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--max-dist", type=float, default=None, help="override the workload's max_dist (startup-cost probe)")
    ap.add_argument("--noise", type=int, default=0, help="then this many resident add_noise passes (for ncu -k regex:k_noise)")
    ap.add_argument("--drop-grid", action="store_true", help="rebuild the point grid in every pass")
    a = ap.parse_args()
    md = bench.MAX_DIST if a.max_dist is None else a.max_dist
    r, n = (int(x) for x in a.shard.split("/"))
    ctx = c2b.context(0)
    cams, pts, xyz, tri = bench.build_workload(a.workload)
    c0, c1 = bench.shard(len(cams), r, n)
    my = torch.from_numpy(np.ascontiguousarray(cams[c0:c1])).pin_memory()
    pp = torch.from_numpy(pts).pin_memory()
    scene = c2b.Scene(xyz, tri, ctx=ctx)
    rp = ResidentProblem(ctx)
    rp.upload_points_ptr(pp.data_ptr(), len(pts))
    rp.upload_cameras_ptr(my.data_ptr(), c1 - c0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    tot = {}
    for s in range(3 + a.steps):
        flush.zero_()
        torch.cuda.synchronize()
        if a.drop_grid:
            c2b._lib.check(c2b._lib.lib().c2b_drop_point_grid(ctx.handle))
        st = rp.run(scene, md, cull_mode="grid", count_traversal=False)
        if s >= 3:
            for k in ("ms_total", "ms_cull", "ms_traverse", "ms_compact"):
                tot[k] = tot.get(k, 0.0) + st[k] / a.steps
    for k in range(a.noise):
        flush.zero_()
        torch.cuda.synchronize()
        rp.add_noise(0.0, 0.0001, 0.01, 0.001, seed=42 + k)
        print("add_noise resident: upload / kernels / download ms", c2b.noise.last_timing(ctx))
    print(json.dumps({"shard": a.shard, "max_dist": md, "cameras": c1 - c0, "pairs": int(st["pairs_evaluated"]), "rays": int(st["n_candidates"]), **{k: round(v, 4) for k, v in tot.items()}}))


if __name__ == "__main__":
    main()

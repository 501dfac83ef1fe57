"""BASELINE config 5 as the reference would meet it — "large OBJ city mesh (~5M triangles)" on disk: writes the
tessellated city of bench.py's cfg5 as a Wavefront OBJ (one object per city block), then times

  * tobj-rule OBJ ingest alone (tests/cpp/test_host.cpp `obj` mode = include/city2ba.hpp tobj::load_obj), and
  * the whole command line:  city2ba generate city.obj out.bbal --cameras 50000 --points 5000000 --max-dist 10
    [--gpus N], then  city2ba noise out.bbal noised.bbal --drift-strength 0.001 --rotation-std 0.0001 ...

    python profiles/obj_ingest_probe.py [--blocks 64] [--out /tmp/c2b_city]
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def write_obj(path, xyz, tri, n_blocks):
    per_block_v = len(xyz) // (n_blocks * n_blocks)
    per_block_t = len(tri) // (n_blocks * n_blocks)
    with open(path, "w") as f:
        f.write("# city2ba_b200 cfg5: tessellated city blocks\n")
        for b in range(n_blocks * n_blocks):
            v = xyz[b * per_block_v:(b + 1) * per_block_v]
            t = tri[b * per_block_t:(b + 1) * per_block_t].astype(np.int64) + 1   # OBJ indices are global, 1-based
            f.write(f"o block{b}\n")
            f.write("".join("v %.9g %.9g %.9g\n" % (x, y, z) for x, y, z in v))
            f.write("".join("f %d %d %d\n" % (a, b_, c) for a, b_, c in t))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=64)
    ap.add_argument("--out", default="/tmp/c2b_city")
    ap.add_argument("--cameras", type=int, default=50000)
    ap.add_argument("--points", type=int, default=5000000)
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    obj = os.path.join(a.out, "city.obj")
    t0 = time.perf_counter()
    xyz, tri = bench.city_mesh_tessellated(a.blocks, bench.TESS_K)
    write_obj(obj, xyz, tri, a.blocks)
    res = {"triangles": int(len(tri)), "vertices": int(len(xyz)), "obj_bytes": os.path.getsize(obj),
           "write_obj_s": round(time.perf_counter() - t0, 2)}
    exe = os.path.join(a.out, "test_host")
    so_dir = os.path.join(ROOT, "city2ba_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_host.cpp"), "-o", exe, "-L", so_dir,
                           "-lcity2ba_cuda", f"-Wl,-rpath,{so_dir}"])
    t0 = time.perf_counter()
    r = subprocess.run([exe, "obj", obj], capture_output=True, text=True)
    res["load_obj_s"] = round(time.perf_counter() - t0, 2)
    res["models"] = len(r.stdout.splitlines())
    import __graft_entry__ as entry
    cli = entry.build_cli()
    out = os.path.join(a.out, "city.bbal")
    t0 = time.perf_counter()
    r = subprocess.run([cli, "generate", obj, out, "--cameras", str(a.cameras), "--points", str(a.points),
                        "--max-dist", "10", "--ground", "1000", "--height", "1", "--seed", "7", "--gpus", str(a.gpus)],
                       capture_output=True, text=True, timeout=1500)
    res["gpus"] = a.gpus
    res["generate_s"] = round(time.perf_counter() - t0, 2)
    res["generate_rc"] = r.returncode
    res["generate_stdout"] = r.stdout.strip().splitlines()
    res["generate_stderr_tail"] = r.stderr.strip()[-400:]
    if os.path.exists(out):
        res["bbal_bytes"] = os.path.getsize(out)
        # BASELINE config 5 ends with "plus full noise pass": drift + Gaussian camera / point / observation noise
        noised = os.path.join(a.out, "city_noised.bbal")
        t0 = time.perf_counter()
        r = subprocess.run([cli, "noise", out, noised, "--drift-strength", "0.001", "--rotation-std", "0.0001",
                            "--point-std", "0.01", "--observation-std", "0.001", "--seed", "7"],
                           capture_output=True, text=True, timeout=600)
        res["noise_s"] = round(time.perf_counter() - t0, 2)
        res["noise_rc"] = r.returncode
        res["noise_stdout"] = r.stdout.strip().splitlines()
    print(json.dumps(res))


if __name__ == "__main__":
    main()

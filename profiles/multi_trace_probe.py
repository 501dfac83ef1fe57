import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench, city2ba_b200 as c2b
G = torch.cuda.device_count()
cams, pts, xyz, tri = bench.build_workload("cfg4")
m = c2b.MultiContext(G)
sc = c2b.MultiScene(xyz, tri, m)
pin_c = torch.from_numpy(cams).pin_memory(); pin_p = torch.from_numpy(pts).pin_memory()
import ctypes as Ct
from city2ba_b200 import _lib
from city2ba_b200.generate import _options
L = _lib.lib(); opt = _options("grid", "mesh", False, False, 20.0, 1.0); out = _lib.Obs()
for name, c, p in (("pinned", pin_c.data_ptr(), pin_p.data_ptr()), ("pageable", cams.ctypes.data, pts.ctypes.data)):
    for k in range(3):
        print(f"--- {name} call {k}", file=sys.stderr, flush=True)
        t0 = time.perf_counter()
        _lib.check(L.c2b_visibility_graph_multi(m.handle, sc.handle, c, len(cams), p, len(pts), 10.0, Ct.byref(opt), Ct.byref(out), None))
        print(f"--- {name} call {k}: {1e3*(time.perf_counter()-t0):.2f} ms", file=sys.stderr, flush=True)

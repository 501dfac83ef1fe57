"""Two-rank NCCL run of the camera-sharded path on real GPUs (skipped with fewer than 2 GPUs):
the graph assembled from two camera ranges equals the single-GPU graph, bit for bit."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SCRIPT = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
import city2ba_b200 as c2b
from city2ba_b200 import synthetic
from city2ba_b200.sharding import visibility_graph_sharded
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = c2b.context(local)
cams = synthetic.grid_cameras(10, 4, 20.0, 1.0)
pts = synthetic.grid_points(10, 4, 20.0, 1.0, 1.0)
scene = c2b.Scene(*synthetic.city_mesh(4), ctx=ctx)
local_g, info = visibility_graph_sharded(scene, cams, pts, 10.0, device=torch.device("cuda", local),
                                         gather_to=0, ctx=ctx)
if dist.get_rank() == 0:
    full = c2b.visibility_graph(scene, cams, pts, 10.0, ctx=ctx)
    g = info["global"]
    assert np.array_equal(g.offsets, full.offsets) and np.array_equal(g.point_idx, full.point_idx)
    assert np.array_equal(g.uv, full.uv)
    assert int(info["counts"].sum()) == full.num_observations
    print("MULTI_OK", info["counts"].tolist())
dist.barrier()
dist.destroy_process_group()
""" % ROOT


def test_two_gpu_sharded_graph_equals_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "multi.py"
    script.write_text(SCRIPT)
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
        capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MULTI_OK" in out.stdout

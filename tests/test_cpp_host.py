"""The C++ host mirror (include/city2ba.hpp) against the reference's own tests, restated in
tests/cpp/test_host.cpp, and against the Python mirror (same C ABI underneath)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def test_host(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cpp") / "test_host")
    so_dir = os.path.join(ROOT, "city2ba_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_host.cpp"), "-o", exe, "-L", so_dir,
                           "-lcity2ba_cuda", f"-Wl,-rpath,{so_dir}"])
    return exe


def test_cpp_host_camera_graph_and_io(test_host, tmp_path):
    """src/baproblem.rs:64-75,227-249 (inline unit tests), cull on a hand-made graph, BAL round trips"""
    r = subprocess.run([test_host, "cpu"], capture_output=True, text=True, env={**os.environ, "TMPDIR": str(tmp_path)})
    assert r.returncode == 0, r.stdout + r.stderr


def test_python_mirror_lcc_matches_reference_quirk(c2b):
    """the same hand-made problem as tests/cpp/test_host.cpp::graph_and_io"""
    cams = np.zeros((3, 15))
    cams[:, [0, 4, 8, 12]] = 1.0
    pts = np.array([[0.5 * i, 0.1 * i, -3.0 - 0.2 * i] for i in range(6)])
    idx = np.tile(np.arange(5, dtype=np.uint64), 2)
    g = c2b.VisGraph(np.array([0, 5, 10, 10], np.uint64), idx, np.zeros((10, 2)))
    ba = c2b.BAProblem(cams, pts, g)
    lcc = ba.largest_connected_component()
    assert (lcc.num_cameras(), lcc.num_points(), lcc.vis_graph.num_observations) == (2, 5, 8)
    culled = ba.cull()
    assert (culled.num_cameras(), culled.num_points(), culled.vis_graph.num_observations) == (2, 4, 8)


@pytest.mark.gpu
def test_cpp_host_library_properties(test_host, tmp_path, c2b, ctx):
    """tests/main.rs:130-201 through the C++ mirror; the culled synthetic grid it writes must equal the
    Python mirror's (both drive the same library)"""
    out = str(tmp_path / "grid.bbal")
    r = subprocess.run([test_host, "gpu", out], capture_output=True, text=True, env={**os.environ, "TMPDIR": str(tmp_path)})
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Bundle Adjustment Problem with" in r.stdout
    from city2ba_b200 import synthetic
    ba = synthetic.synthetic_grid(10, 20, 3, 5.0, 1.0, 1.0, 1.0, 10.0, False, ctx=ctx)
    back = c2b.BAProblem.from_file(out)
    assert back.num_cameras() == ba.num_cameras() and back.num_points() == ba.num_points()
    assert np.array_equal(back.vis_graph.offsets, ba.vis_graph.offsets)
    assert np.array_equal(back.vis_graph.point_idx, ba.vis_graph.point_idx)
    assert np.array_equal(back.vis_graph.uv, ba.vis_graph.uv)
    assert np.allclose(back.points, ba.points, rtol=0, atol=0)

"""The command line (city2ba_b200/cli/city2ba.cpp) against the reference's own CLI tests
(tests/main.rs:11-128: exit status + stdout substrings) and against BASELINE configs 1 and 2, which
are literal command lines.  The reference's tests read tests/box.obj (Cube + Plane + a poly-line object
`BezierCurve`, bbox +-4); the GPU box has no /root/reference, so a scene of the same structure is
written here."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

CLI = os.path.join(ROOT, "city2ba_b200", "bin", "city2ba")


@pytest.fixture(scope="module", autouse=True)
def _cli_binary():
    """built by __graft_entry__.build(); rebuilt here (g++ only) if it is missing or older than its sources"""
    import sys
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    assert os.path.exists(entry.SO), "libcity2ba_cuda.so is missing: run __graft_entry__.build()"
    entry.build_cli()


def run(*args):
    return subprocess.run([CLI, *map(str, args)], capture_output=True, text=True, timeout=600)


@pytest.fixture(scope="module")
def box_obj(tmp_path_factory):
    """quads only, like tests/box.obj: a cube (y in [0, 2]), a ground plane and a 49-segment curve
    (about 16 long, so that 100 cameras at --step-size 0.1 fit, as on the reference's BezierCurve)"""
    path = tmp_path_factory.mktemp("obj") / "box.obj"
    lines = ["# box scene", "mtllib box.mtl", "o Cube"]
    for k in range(8):
        lines.append(f"v {1 if k & 1 else -1} {2 if k & 2 else 0} {1 if k & 4 else -1}")
    lines += ["vn 0 1 0", "usemtl None", "s off",
              "f 1//1 2//1 4//1 3//1", "f 5//1 7//1 8//1 6//1", "f 1//1 5//1 6//1 2//1",
              "f 3//1 4//1 8//1 7//1", "f 1//1 3//1 7//1 5//1", "f 2//1 6//1 8//1 4//1", "o BezierCurve"]
    t = np.linspace(0.0, 1.0, 50)
    for s in t:
        lines.append(f"v {-3.5 + 7.0 * s:.6f} 0.5 {2.6 + 1.2 * np.sin(6.0 * np.pi * s):.6f}")
    lines += [f"l {9 + k} {10 + k}" for k in range(49)]
    lines += ["o Plane", "v -4 0 -4", "v 4 0 -4", "v 4 0 4", "v -4 0 4", "s off", "f 59 60 61 62"]
    path.write_text("\n".join(lines) + "\n")
    return str(path)


def test_cli_is_built_and_reports_usage():
    assert os.path.exists(CLI), "run __graft_entry__.build()"
    assert run().returncode == 2                                   # no sub-command (clap: exit status 2)
    r = run("synthetic")                                           # missing <OUTPUT>
    assert r.returncode == 2 and "required arguments" in r.stderr
    r = run("synthetic", "--no-such-flag", "1", "/tmp/x.bal")
    assert r.returncode == 2 and "wasn't expected" in r.stderr
    r = run("generate", "a.obj", "b.bal", "--path", "p", "--ground", "-1.0")  # conflicts_with, :104
    assert r.returncode == 2 and "cannot be used with" in r.stderr
    r = run("generate", "/nonexistent/scene.obj", "/tmp/x.bal")
    assert r.returncode == 1 and "Could not open file" in r.stderr  # src/bin/city2ba.rs:481-485
    # values are parsed, and input files opened, before the GPU is touched: the same failures with or without one
    r = run("noise", "/nonexistent/in.bbal", "/tmp/x.bbal")
    assert r.returncode == 1 and "IOError" in r.stderr               # BAProblem::from_file, src/bin/city2ba.rs:281
    r = run("synthetic", "--blocks", "many", "/tmp/x.bal")
    assert r.returncode == 2 and "Invalid value for '--blocks" in r.stderr
    r = run("noise", "a.bbal", "b.bbal", "--drift-std=abc")
    assert r.returncode == 2 and "invalid float literal" in r.stderr
    r = run("generate", "a.obj", "b.bal", "--intrinsics-start", "1,2")  # parse_vec3 unwraps a missing component, :25-31
    assert r.returncode == 101 and "panicked" in r.stderr


def test_cli_has_no_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run("synthetic", tmp_path / "blocks.bbal")
    assert r.returncode == 1 and "no CPU fallback" in r.stderr
    assert not (tmp_path / "blocks.bbal").exists()


@pytest.mark.gpu
@pytest.mark.parametrize("ext", ["bbal", "bal"])
def test_synthetic_blocks(tmp_path, ext, c2b, ctx):
    """tests/main.rs:11-35 (synthetic_blocks, test_bal_output) + the file equals the Python mirror's"""
    out = tmp_path / f"blocks.{ext}"
    r = run("synthetic", out)
    assert r.returncode == 0, r.stderr
    assert "Bundle Adjustment Problem" in r.stdout
    from city2ba_b200 import synthetic
    ba = synthetic.synthetic_grid(10, 10, 5, 20.0, 1.0, 1.0, 1.0, 10.0, False, ctx=ctx)   # the defaults, :117-147
    assert r.stdout.strip() == (f"Bundle Adjustment Problem with {ba.num_cameras()} cameras, {ba.num_points()} points, "
                                f"and {ba.vis_graph.num_observations} observations")
    back = c2b.BAProblem.from_file(str(out))
    assert np.array_equal(back.vis_graph.offsets, ba.vis_graph.offsets)
    assert np.array_equal(back.vis_graph.point_idx, ba.vis_graph.point_idx)
    assert np.array_equal(back.vis_graph.uv, ba.vis_graph.uv)        # text BAL round-trips f64 exactly
    assert np.array_equal(back.points, ba.points)


@pytest.mark.gpu
def test_noise_blocks(tmp_path):
    """tests/main.rs:37-63"""
    bbal = tmp_path / "blocks.bbal"
    assert run("synthetic", bbal).returncode == 0
    r = run("noise", bbal, tmp_path / "blocks_noised.bbal", "--drift-strength", "0.00001", "--mismatch-chance", "0.00001")
    assert r.returncode == 0, r.stderr
    assert "Initial error" in r.stdout and "Final error" in r.stdout


@pytest.mark.gpu
def test_baseline_config2_commands(tmp_path, c2b, ctx, orc):
    """BASELINE config 2, literally: `city2ba synthetic --blocks 4` then
    `noise --drift-strength 0.001 --rotation-std 0.0001`.  The drift part is deterministic (--drift-std 0),
    the Gaussian part is pinned by --seed: the output must equal the oracle's on the same stream."""
    bal, noised = tmp_path / "b4.bbal", tmp_path / "b4_noised.bbal"
    r = run("synthetic", "--blocks", "4", bal)
    assert r.returncode == 0, r.stderr
    ba = c2b.BAProblem.from_file(str(bal))
    assert ba.num_cameras() <= 800 and ba.num_points() <= 2400 and ba.num_cameras() > 700
    r = run("noise", bal, noised, "--drift-strength", "0.001", "--rotation-std", "0.0001", "--seed", "42")
    assert r.returncode == 0, r.stderr
    out = c2b.BAProblem.from_file(str(noised))
    lines = r.stdout.splitlines()
    assert lines[0].startswith("Initial error: ") and lines[0].endswith("(L2)")
    assert lines[1] == (f"BA Problem with {out.num_cameras()} cameras, {out.num_points()} points, "
                        f"{out.vis_graph.num_observations} correspondences")
    # the same sequence through the oracle (seed + 4 for the drift, seed + 5 for add_noise, as the CLI numbers them)
    cams, pts = orc.add_drift_normalized(ba.cameras, ba.points, 0.001, 0.0, 0.0, 42 + 4)
    cams, pts, uv = orc.add_noise(cams, pts, ba.vis_graph.uv, 0.0, 0.0001, 0.0, 0.0, 42 + 5)
    # cameras pass through the 9-parameter Rodrigues form in the file: compare projections of the points
    assert np.allclose(out.points, pts, rtol=1e-9, atol=1e-12)
    assert np.array_equal(out.vis_graph.uv, uv)                      # observation-std 0: untouched
    for c in range(0, out.num_cameras(), 37):
        cen = orc.center(cams[c])
        assert np.allclose(orc.center(out.cameras[c]), cen, rtol=1e-8, atol=1e-9)
    err = float(lines[2].split()[2])
    assert err > float(lines[0].split()[2])                          # Final error > Initial error


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [("--path", "BezierCurve"), ("--path", "BezierCurve", "--step-size", "0.1"),
                                   ("--ground", "-1.0")])
def test_from_box(tmp_path, box_obj, extra, c2b):
    """tests/main.rs:65-128 (from_box_path, from_box_path_step, from_box_path_ground)"""
    out = tmp_path / "box.bal"
    r = run("generate", box_obj, out, "--cameras", "100", "--points", "100", *extra, "--seed", "3")
    if "--ground" in extra:
        # lower_y + ground = -1: only cameras with z < -1 survive the filter of src/generate.rs:264
        assert r.returncode in (0, 1), r.stderr
        if r.returncode == 1:
            assert "EmptyProblem" in r.stderr
            return
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Total reprojection error" in r.stdout
    assert "Generated 100 world points" in r.stdout
    ba = c2b.BAProblem.from_file(str(out))
    assert ba.num_cameras() > 0 and ba.total_reprojection_error(1.0) < 1e-6


@pytest.mark.gpu
def test_baseline_config1_command(tmp_path, box_obj, c2b, ctx):
    """BASELINE config 1's command line shape: `generate scene.obj --cameras 100 --points 200`; the graph in the
    file must be what the library computes for the cameras and points in the file (--no-lcc keeps them all)."""
    out = tmp_path / "scene.bbal"
    r = run("generate", box_obj, out, "--cameras", "100", "--points", "200", "--ground", "10", "--no-lcc", "--seed", "5")
    assert r.returncode == 0, r.stdout + r.stderr
    ba = c2b.BAProblem.from_file(str(out))
    assert ba.num_points() == 200 and ba.num_cameras() > 20
    assert f"Computed visibility graph with {ba.vis_graph.num_observations} edges" in r.stdout
    # repeatable with the same seed
    out2 = tmp_path / "scene2.bbal"
    assert run("generate", box_obj, out2, "--cameras", "100", "--points", "200", "--ground", "10", "--no-lcc",
               "--seed", "5").returncode == 0
    assert open(out, "rb").read() == open(out2, "rb").read()


@pytest.mark.gpu
def test_gpus_predicate_and_mesh_occlusion_flags(tmp_path, box_obj, c2b):
    """the flags this build adds to the reference's command line: --gpus N (same output file as one GPU),
    --predicate mt / --compare-predicates, synthetic --occlusion mesh"""
    import torch
    base = ("generate", box_obj, "--cameras", "100", "--points", "200", "--ground", "10", "--no-lcc", "--seed", "5")
    one = tmp_path / "one.bbal"
    r = run(base[0], base[1], one, *base[2:], "--compare-predicates")
    assert r.returncode == 0, r.stdout + r.stderr
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("Occlusion predicate watertight: ")]
    assert len(line) == 1 and "; mt: " in line[0]
    n_water = int(line[0].split()[3])
    n_mt = int(line[0].split()[-2])
    assert f"Computed visibility graph with {n_water} edges" in r.stdout
    mt = tmp_path / "mt.bbal"
    r = run(base[0], base[1], mt, *base[2:], "--predicate", "mt")
    assert r.returncode == 0 and f"Computed visibility graph with {n_mt} edges" in r.stdout
    assert run(base[0], base[1], mt, *base[2:], "--predicate", "embree").returncode == 2
    G = min(torch.cuda.device_count(), 4)
    many = tmp_path / "many.bbal"
    r = run(base[0], base[1], many, *base[2:], "--gpus", G)
    assert r.returncode == 0, r.stdout + r.stderr
    assert open(one, "rb").read() == open(many, "rb").read()
    assert run(base[0], base[1], many, *base[2:], "--gpus", "0").returncode == 2
    # synthetic: analytic (reference) and mesh occlusion, one GPU and all of them
    a1, aG, m1 = tmp_path / "a1.bbal", tmp_path / "aG.bbal", tmp_path / "m1.bbal"
    assert run("synthetic", a1, "--blocks", "3").returncode == 0
    assert run("synthetic", aG, "--blocks", "3", "--gpus", G).returncode == 0
    assert open(a1, "rb").read() == open(aG, "rb").read()
    r = run("synthetic", m1, "--blocks", "3", "--occlusion", "mesh")
    assert r.returncode == 0 and "Bundle Adjustment Problem" in r.stdout
    assert c2b.BAProblem.from_file(str(m1)).num_cameras() > 0
    assert run("synthetic", m1, "--occlusion", "raytrace").returncode == 2

"""Host-side mirror of the reference's library surface (no GPU): lattices, mesh, BAProblem graph
ops, BAL I/O."""
import numpy as np

from conftest import procedural_scene  # noqa: F401


def test_lattices_and_mesh_match_oracle(orc, c2b):
    from city2ba_b200 import synthetic
    for cpb, ppb, n in [(10, 10, 4), (3, 7, 2), (1, 1, 1)]:
        assert np.array_equal(synthetic.grid_cameras(cpb, n, 20.0, 1.0), orc.grid_cameras(cpb, n))
        assert np.array_equal(synthetic.grid_points(ppb, n, 20.0, 1.0, 1.0), orc.grid_points(ppb, n))
        x, t = synthetic.city_mesh(n)
        ox, ot = orc.city_mesh(n)
        assert np.array_equal(x, ox) and np.array_equal(t, ot)
    assert np.array_equal(synthetic.grid_cameras(10, 3, 5.0, 1.0), orc.grid_cameras(10, 3, 5.0, 1.0))
    assert np.array_equal(synthetic.line_cameras(30, 10.0, 1.0), orc.line_cameras(30, 10.0, 1.0))
    assert np.array_equal(synthetic.line_points(40, 10.0, 1.0, 1.0), orc.line_points(40, 10.0, 1.0, 1.0))


def _problem(orc, c2b):
    cams, pts = orc.grid_cameras(10, 3, 5.0, 1.0), orc.grid_points(20, 3, 5.0, 1.0, 1.0)
    v = orc.synthetic_visibility(cams, pts, 10.0, True, 5.0, 1.0)
    g = c2b.VisGraph(v.offsets, v.point_idx, v.uv)
    return c2b.BAProblem.from_visibility(cams, pts, g)


def test_baproblem_cull_invariants(orc, c2b):
    ba = _problem(orc, c2b)
    culled = ba.cull()
    g = culled.vis_graph
    assert culled.num_cameras() <= ba.num_cameras() and culled.num_points() <= ba.num_points()
    assert np.all(g.counts() > 3)                      # cameras see >= 4 points (src/baproblem.rs:432)
    pc = np.bincount(g.point_idx.astype(np.int64), minlength=culled.num_points())
    assert np.all(pc > 1)                              # points seen at least twice (:448)
    again = culled.cull()                              # fixed point (:542)
    assert again.num_cameras() == culled.num_cameras() and again.num_points() == culled.num_points()
    assert culled.total_reprojection_error(2.0) < 1e-12
    assert str(culled).startswith("Bundle Adjustment Problem with ")


def test_baproblem_list_of_lists_round_trip(orc, c2b):
    ba = _problem(orc, c2b)
    lol = [ba.vis_graph[i] for i in range(5)] + [[] for _ in range(ba.num_cameras() - 5)]
    ba2 = c2b.BAProblem.from_visibility(ba.cameras, ba.points, lol)
    assert ba2.vis_graph[3] == ba.vis_graph[3]
    assert ba2.num_observations() == sum(len(x) for x in lol)


def test_bal_binary_and_text_round_trip(orc, c2b, tmp_path):
    ba = _problem(orc, c2b).cull()
    for ext in ("bbal", "bal"):
        path = tmp_path / f"p.{ext}"
        ba.write(path)
        back = c2b.BAProblem.from_file(path)
        assert back.num_cameras() == ba.num_cameras() and back.num_points() == ba.num_points()
        assert np.array_equal(back.vis_graph.offsets, ba.vis_graph.offsets)
        assert np.array_equal(back.vis_graph.point_idx, ba.vis_graph.point_idx)
        assert np.array_equal(back.vis_graph.uv, ba.vis_graph.uv)      # shortest round-trip text
        assert np.array_equal(back.points, ba.points)
        assert np.allclose(back.cameras, ba.cameras, atol=1e-12)        # through Rodrigues vectors
    head = open(tmp_path / "p.bal").readline().split()
    assert [int(x) for x in head] == [ba.num_cameras(), ba.num_points(), ba.num_observations()]
    raw = open(tmp_path / "p.bbal", "rb").read()
    assert int.from_bytes(raw[:8], "big") == ba.num_cameras()           # big-endian (src/baproblem.rs:738)


def test_text_format_never_uses_exponent(c2b, tmp_path):
    cams = np.zeros((1, 15))
    cams[0, [0, 4, 8, 12]] = 1.0
    g = c2b.VisGraph(np.array([0, 1], np.uint64), np.array([0], np.uint64), np.array([[1e-7, 0.30000000000000004]]))
    ba = c2b.BAProblem(cams, np.array([[1e21, -2.5e-9, 1.0]]), g)
    ba.write(tmp_path / "x.bal")
    txt = open(tmp_path / "x.bal").read()
    assert "e" not in txt.lower() and "0.0000001" in txt and "0.30000000000000004" in txt

"""Pins the oracle's camera math to the reference's own known-answer tests
(src/baproblem.rs:64-75 rodrigues_idempotent, :227-234 test_project_world, :236-242 test_project,
:244-249 test_project_isomorphic) and the product's host mirror to the oracle."""
import numpy as np
import pytest


@pytest.mark.parametrize("v", [[1.0, 2.0, 3.0], [0.0, 0.0, 0.0], [-1.2, 0.0, 1.7]])
def test_rodrigues_idempotent(orc, v):  # src/baproblem.rs:64-75, tolerance 1e-10 as there
    v_ = orc.to_rodrigues(orc.from_rodrigues(v))
    assert np.linalg.norm(v_ - np.array(v)) < 1e-10


def test_project_world(orc):  # src/baproblem.rs:227-234
    c = orc.from_vec([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0])
    pc = orc.project_world(c, [0.0, 0.0, -1.0])
    assert pc[2] < 0.0
    assert pc[0] == 0.0 and pc[1] == 0.0


def test_project(orc):  # src/baproblem.rs:236-242
    c = orc.from_vec([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0])
    uv = orc.project(c, orc.project_world(c, [0.0, 0.0, -1.0]))
    assert uv[0] == 0.0 and uv[1] == 0.0


def test_project_isomorphic(orc):  # src/baproblem.rs:244-249, tolerance 1e-8 as there
    p = np.array([1.0, 3.0, -1.0])
    c = orc.from_vec([3.0, 5.0, -2.0, 0.5, -0.2, 0.1, 1.0, 0.0, 0.0])
    assert np.allclose(orc.to_world(c, orc.project_world(c, p)), p, atol=1e-8, rtol=0)


def test_center_is_minus_Rt_t(orc):
    rng = np.random.default_rng(0)
    for _ in range(20):
        c = orc.from_vec(np.concatenate([rng.normal(size=3), rng.normal(size=3) * 5, [1, 0, 0]]))
        R = c[:9].reshape(3, 3).T
        assert np.allclose(orc.center(c), -R.T @ c[9:12], atol=1e-12)
        # from_position_direction(center, R) reproduces t (src/baproblem.rs:153-159)
        assert np.allclose(orc.from_position_direction(orc.center(c), c[:9])[9:12], c[9:12], atol=1e-12)


def test_transform_moves_center_with_old_rotation(orc):
    # src/baproblem.rs:165-171: loc' = -R_old (center + delta); R' = R * dR
    rng = np.random.default_rng(1)
    c = orc.from_vec(np.concatenate([rng.normal(size=3), rng.normal(size=3), [1.2, 0.01, 0.0]]))
    dR = orc.from_axis_angle([0, 0, 1], 0.3)
    d = np.array([0.5, -1.0, 2.0])
    t = orc.transform(c, dR, d)
    R = c[:9].reshape(3, 3).T
    assert np.allclose(t[9:12], -R @ (orc.center(c) + d), atol=1e-12)
    assert np.allclose(t[:9].reshape(3, 3).T, R @ dR.reshape(3, 3).T, atol=1e-12)
    assert np.array_equal(t[12:], c[12:])


def test_host_mirror_camera_matches_oracle(orc, c2b):
    """city2ba_b200.SnavelyCamera (product host code) against the oracle, bit for bit."""
    rng = np.random.default_rng(2)
    for _ in range(50):
        v9 = np.concatenate([rng.normal(size=3), rng.normal(size=3) * 3, [rng.uniform(0.5, 2), 0.01, -0.002]])
        cam = c2b.SnavelyCamera.from_vec(v9)
        oc = orc.from_vec(v9)
        assert np.allclose(cam.rec, oc, rtol=0, atol=1e-15)
        cam = c2b.SnavelyCamera(record=oc)
        p = rng.normal(size=3) * 4
        assert np.array_equal(cam.center(), orc.center(oc))
        assert np.array_equal(cam.project_world(p), orc.project_world(oc, p))
        assert np.array_equal(cam.project(cam.project_world(p)), orc.project(oc, orc.project_world(oc, p)))
        dR = orc.from_axis_angle([1, 0, 0], 0.1)
        assert np.array_equal(cam.transform(dR, [0.1, 0.2, 0.3]).rec, orc.transform(oc, dR, [0.1, 0.2, 0.3]))
        assert np.allclose(cam.to_vec(), orc.to_vec(oc), rtol=0, atol=1e-13)
        assert np.allclose(cam.to_world(cam.project_world(p)), p, atol=1e-9)


def test_host_mirror_known_answers(c2b):
    """the reference's four unit tests, run against the product's SnavelyCamera"""
    from city2ba_b200.baproblem import from_rodrigues, to_rodrigues
    for v in ([1.0, 2.0, 3.0], [0.0, 0.0, 0.0], [-1.2, 0.0, 1.7]):
        assert np.linalg.norm(to_rodrigues(from_rodrigues(v)) - np.array(v)) < 1e-10
    c = c2b.SnavelyCamera.from_vec([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0])
    pc = c.project_world([0.0, 0.0, -1.0])
    assert pc[2] < 0.0 and pc[0] == 0.0 and pc[1] == 0.0
    uv = c.project(pc)
    assert uv[0] == 0.0 and uv[1] == 0.0
    p = np.array([1.0, 3.0, -1.0])
    c = c2b.SnavelyCamera.from_vec([3.0, 5.0, -2.0, 0.5, -0.2, 0.1, 1.0, 0.0, 0.0])
    assert np.allclose(c.to_world(c.project_world(p)), p, atol=1e-8, rtol=0)

"""The noise draws of the oracle (oracle/c2b_oracle.c "the noise draws"; the CUDA path evaluates the same
operations): the table + polynomial evaluations are pinned against libm, and the three distributions the
reference specifies — a normalised Gaussian pair = uniform direction on the circle (src/noise.rs:159-163), a
normalised Gaussian triple = uniform point on the sphere (unit_random, :35-43), Normal(0, 1) — by moments and
Kolmogorov-Smirnov tests on Philox-generated words."""
import math

import numpy as np
from scipy import stats


def test_unit2_against_libm(orc):
    rng = np.random.default_rng(1)
    ws = np.concatenate([rng.integers(0, 2 ** 32, 20000, dtype=np.uint64),
                         [0, 1, 2 ** 24 - 1, 2 ** 24, 2 ** 31, 2 ** 32 - 1, 2 ** 30, 3 * 2 ** 30]])
    worst = 0.0
    for w in ws:
        c, s = orc.unit2(int(w))
        a = (2 * np.pi * np.longdouble(1)) * np.longdouble(int(w)) / np.longdouble(2.0 ** 32)   # 80-bit reference
        worst = max(worst, abs(float(np.longdouble(c) - np.cos(a))), abs(float(np.longdouble(s) - np.sin(a))))
        assert abs(c * c + s * s - 1.0) < 1e-15
    assert worst < 8e-16, worst   # table angles k * (2 pi / 256) are rounded doubles: a few 1e-16 absolute


def test_neg2ln40_against_libm(orc):
    rng = np.random.default_rng(2)
    ns = np.concatenate([rng.integers(0, 2 ** 40, 20000, dtype=np.uint64),
                         2 ** 40 - 1 - rng.integers(0, 1000, 200, dtype=np.uint64),   # u close to 1
                         rng.integers(0, 1000, 200, dtype=np.uint64),                  # u close to 2^-40
                         [0, 1, 2 ** 39, 2 ** 40 - 1, 2 ** 40 - 2]])
    for n in ns:
        u = (int(n) + 1) * 2.0 ** -40
        want = -2.0 * math.log(u)
        got = orc.neg2ln40(int(n))
        assert got >= 0.0
        assert abs(got - want) <= 4e-16 + 4e-16 * want, (int(n), got, want)
    assert orc.neg2ln40(2 ** 40 - 1) == 0.0  # u = 1: no negative radicand


def blocks(orc, n, seed=7, stream=5):
    # Philox words the way the kernels draw them: counter (index, 0, stream, 0), key = seed
    return np.array([orc.philox4x32_10([i, 0, stream, 0], [seed, 0]) for i in range(n)], dtype=np.uint32)


def test_distributions(orc):
    n = 100_000
    d = orc.noise_draws(blocks(orc, n))
    z = d["normal"]
    assert abs(z.mean()) < 4 / math.sqrt(n) and abs(z.std() - 1.0) < 0.01
    assert abs(stats.kurtosis(z)) < 0.06 and abs(stats.skew(z)) < 0.03
    assert stats.kstest(z, "norm").pvalue > 1e-3
    # circle: the angle is uniform, and what the reference draws — (nx, ny) / |(nx, ny)| of two normals — has
    # the same law: compare with that construction from an independent generator
    c = d["circle"]
    ang = np.arctan2(c[:, 1], c[:, 0])
    assert stats.kstest((ang + math.pi) / (2 * math.pi), "uniform").pvalue > 1e-3
    g = np.random.default_rng(3).normal(size=(n, 2))
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    assert stats.ks_2samp(ang, np.arctan2(g[:, 1], g[:, 0])).pvalue > 1e-3
    # sphere: unit length, each coordinate uniform on [-1, 1] (Archimedes), like a normalised Gaussian triple
    s = d["sphere"]
    assert np.allclose(np.linalg.norm(s, axis=1), 1.0, atol=1e-15)
    g3 = np.random.default_rng(4).normal(size=(n, 3))
    g3 /= np.linalg.norm(g3, axis=1, keepdims=True)
    for k in range(3):
        assert stats.kstest((s[:, k] + 1) / 2, "uniform").pvalue > 1e-3
        assert stats.ks_2samp(s[:, k], g3[:, k]).pvalue > 1e-3
    assert np.all(np.abs(s.mean(axis=0)) < 0.01)
    assert np.all(np.abs((s[:, [0, 0, 1]] * s[:, [1, 2, 2]]).mean(axis=0)) < 0.01)  # no axis correlation
    # the three draws of one block are independent of each other
    assert abs(np.corrcoef(z, ang)[0, 1]) < 0.01 and abs(np.corrcoef(z, s[:, 2])[0, 1]) < 0.01


def test_add_noise_moments_on_the_oracle(orc):
    """src/noise.rs:149,159-168 through orc_add_noise: |dp| ~ |N(0, sigma_p)|, |duv| ~ |N(0, sigma_o)|"""
    n = 50_000
    cams = np.zeros((1, 15))
    cams[0, [0, 4, 8, 12]] = 1.0
    _, p, uv = orc.add_noise(cams, np.zeros((n, 3)), np.zeros((n, 2)), 0.0, 0.0, 0.5, 0.25, 11)
    r, ro = np.linalg.norm(p, axis=1), np.linalg.norm(uv, axis=1)
    assert abs(r.mean() - 0.5 * math.sqrt(2 / math.pi)) < 6e-3 and abs(math.sqrt((r ** 2).mean()) - 0.5) < 6e-3
    assert abs(math.sqrt((ro ** 2).mean()) - 0.25) < 3e-3
    assert stats.kstest(r / 0.5, "halfnorm").pvalue > 1e-3 and stats.kstest(ro / 0.25, "halfnorm").pvalue > 1e-3

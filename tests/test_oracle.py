"""The oracle against what pins it: the reference's golden vectors
(gp/gp_test.go), its own two gradient modes, finite differences and mpmath."""
import math

import numpy as np
import pytest

from oracle import kernels as ok
from oracle.gp import GP, NotPositiveDefinite, mean_std
from tests import cases
from tests.golden_ref import DX, ELEMENTAL, EPS, PRODUCE, arr


def _noise(n):
    return ok.UniformNoise if n == "uniform" else ok.ConstantNoise(n)


@pytest.mark.parametrize("case", PRODUCE, ids=[c[0] for c in PRODUCE])
def test_produce_goldens(case):
    name, nstd, theta, x, y, z, mu, sigma = case
    g = GP(1, ok.Normal, ok.ConstantNoise(nstd), theta_simil=theta)
    g.absorb(arr(x).reshape(-1, 1), arr(y))
    m, s = g.produce(arr(z), clamp=True)
    assert len(m) == len(mu) and len(s) == len(sigma)
    assert np.all(np.abs(m - arr(mu)) <= 1e-6)
    assert np.all(np.abs(s - arr(sigma)) <= 1e-6)


@pytest.mark.parametrize("case", ELEMENTAL, ids=[c[0] for c in ELEMENTAL])
def test_elemental_goldens(case):
    name, n, x, ll = case
    x = arr(x)
    g = GP(1, ok.Normal, _noise(n))
    v = g.observe(x)
    dll = g.gradient()
    assert abs(v - ll) < 1e-6
    assert len(dll) == len(x)
    for j in range(len(x)):  # the reference's forward-difference check, gp/gp_test.go:242-252
        x0 = x[j]
        x[j] += DX
        vj = GP(1, ok.Normal, _noise(n)).observe(x)
        x[j] = x0
        assert abs(dll[j] - (vj - v) / DX) <= EPS
    P = g.Simil.ntheta + g.Noise.ntheta  # hyper-parameters only, gp/gp_test.go:254-267
    v2 = g.observe(x[:P].copy())
    assert abs(v2 - ll) < 1e-6
    assert len(g.gradient()) == P


@pytest.mark.parametrize("name", ["normal_uniform", "periodic", "matern52", "hyperpriors", "c3_ard3", "sum_times"])
def test_literal_equals_fast_gradient(name):
    X, y, logt = cases.synth(name, 9, seed=3)
    x = np.concatenate([logt, X.ravel(), y])
    g = cases.make_oracle_gp(name)
    g.observe(x)
    lit = g.gradient("literal")
    g.observe(x)
    fast = g.gradient("fast")
    assert cases.relerr(fast, lit) < 1e-11


def test_barebones_kat():
    """bootstrap KAT for config 1 (restatement-derived, mpmath-checked)."""
    d = np.loadtxt(cases.GOLDEN + "/barebones.csv", delimiter=",")
    m, s = mean_std(d[:, 1])
    g = cases.make_oracle_gp("barebones")
    g.X, g.Y = d[:, :1].copy(), (d[:, 1] - m) / s
    assert abs(g.observe(np.zeros(3)) - (-8.482987052156)) < 1e-10
    assert np.allclose(g.gradient(), [-3.2320901320, 6.9054771170, -0.4752473692], atol=1e-9)
    assert abs(g.observe(arr([-0.5, 0.3, 1.0])) - (-9.103987864150)) < 1e-10
    assert np.allclose(g.gradient(), [-0.1009289296, 2.3769085404, -7.5447626235], atol=1e-9)


def test_not_positive_definite():
    g = GP(1, ok.Normal, ok.ConstantNoise(0.0), theta_simil=[1.0])
    with pytest.raises(NotPositiveDefinite):
        g.absorb(arr([[0.0], [0.0]]), arr([1.0, 2.0]))  # duplicate input, no noise -> singular K


def test_bad_length_panics():
    g = GP(2, ok.ArdMatern32(2), ok.UniformNoise)
    with pytest.raises(ValueError):
        g.observe(np.zeros(3 + 1 + 4))  # 4 is not a multiple of NDim+1


def test_matern52_ships_with_unit_coefficient():
    """kernel/kernel.go:91: 5/3 is integer division in Go."""
    from oracle.dual import Dual
    d = 0.7
    v = ok.matern52_cov(Dual(1.0), Dual(d), Dual(0.0)).v
    assert abs(v - (1 + ok.SQRT5 * d + d * d) * math.exp(-ok.SQRT5 * d)) < 1e-16

"""Kernel configurations shared by the parity tests: the device-side expression
(gogp_b200.kernel) next to the oracle's restatement of the same reference code
(oracle.kernels).  The oracle side is test infrastructure only."""
import os

import numpy as np

from gogp_b200 import kernel as k
from oracle import kernels as ok

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def ard_normal_periodic(D):
    e = k.Param(0)
    for d in range(D):
        e = e * k.Normal.Of(l=1 + d, dim=d)
    return e * k.Periodic.Of(l=1 + D, p=2 + D, dim=0)


def ard_matern32(D):
    e = k.Param(0)
    for d in range(D):
        e = e * k.Matern32.Of(l=1 + d, dim=d)
    return e


class _Sum2:
    """(theta0*Normal(theta1) + theta2) * Matern32(theta3): exercises the host-side
    distribution of a product over a sum and a bare-parameter term."""
    ntheta = 4

    def observe(self, x):
        return (x[0] * ok.normal_cov(x[1], x[4], x[5]) + x[2]) * ok.matern32_cov(x[3], x[4], x[5])


# name -> (ndim, device simil, device noise, oracle simil, oracle noise)
CASES = {
    "normal_const": (1, k.Normal, k.ConstantNoise(0.1), ok.Normal, ok.ConstantNoise(0.1)),
    "normal_uniform": (1, k.Normal, k.UniformNoise, ok.Normal, ok.UniformNoise),
    "periodic": (1, k.Periodic, k.ConstantNoise(0.3), ok.Periodic, ok.ConstantNoise(0.3)),
    "matern32": (1, k.Matern32, k.ConstantNoise(0.2), ok.Matern32, ok.ConstantNoise(0.2)),
    "matern52": (1, k.Matern52, k.UniformNoise, ok.Matern52, ok.UniformNoise),
    "barebones": (1, k.Param(0) * k.Matern32.Of(l=1), 0.01 * k.UniformNoise,
                  ok.BarebonesSimil(), ok.ScaledUniformNoise(0.01)),
    "hyperpriors": (1, k.Param(0) * k.Matern52.Of(l=2) + k.Param(1) * k.Periodic.Of(l=3, p=(4, 10.0)),
                    0.01 * k.UniformNoise, ok.HyperpriorsSimil(), ok.ScaledUniformNoise(0.01)),
    "anynoise": (1, k.Param(0) * k.Matern52.Of(l=1), k.Const(1e-5).WithNTheta(1),
                 ok.AnynoiseSimil(), ok.AnynoiseNoise()),
    "warpedtime": (1, k.Param(0) * k.Matern52.Of(l=1), 0.01 * k.UniformNoise,
                   ok.AnynoiseSimil(), ok.ScaledUniformNoise(0.01)),
    "c2_rbf": (1, k.Param(0) * k.Normal.Of(l=1), k.UniformNoise, ok.ScaledNormal1D(), ok.UniformNoise),
    "c3_ard8": (8, ard_normal_periodic(8), k.UniformNoise, ok.ArdNormalTimesPeriodic(8), ok.UniformNoise),
    "c3_ard3": (3, ard_normal_periodic(3), k.UniformNoise, ok.ArdNormalTimesPeriodic(3), ok.UniformNoise),
    "c5_matern4": (4, ard_matern32(4), k.UniformNoise, ok.ArdMatern32(4), ok.UniformNoise),
    "events": (1, k.Param(0) * k.Matern52.Of(l=1) * k.Events([(1.0, 2.5, 0.3), (3.0, 6.0, 0.5)]), 0.01 * k.UniformNoise,
               ok.EventsSimil([(1.0, 2.5, 0.3), (3.0, 6.0, 0.5)]), ok.ScaledUniformNoise(0.01)),
    "sum_times": (1, (k.Param(0) * k.Normal.Of(l=1) + k.Param(2)) * k.Matern32.Of(l=3), k.UniformNoise,
                  _Sum2(), ok.UniformNoise),
}


def synth(name, N, seed=0, spread=None):
    """Seeded synthetic inputs in the style of SURVEY.md section 8(d): x ~ U(0, spread)^D,
    y = sum_d sin(x_d) + 0.1 N(0,1) normalised, log theta = small jitter around 0 with a
    noise level that keeps cond(K) moderate."""
    ndim = CASES[name][0]
    rng = np.random.default_rng(seed)
    if spread is None:
        spread = max(4.0, N / 50.0) if ndim == 1 else 4.0
    X = rng.uniform(0.0, spread, size=(N, ndim))
    y = np.sin(X).sum(axis=1) + 0.1 * rng.standard_normal(N)
    if N > 1:
        y = (y - y.mean()) / y.std(ddof=1)
    nts = CASES[name][1].NTheta()
    ntn = CASES[name][2].NTheta()
    logt = 0.1 * rng.standard_normal(nts + ntn)
    if name == "hyperpriors":
        logt[4] += np.log(0.3)  # period 10*theta4 ~ 3
    if ntn and name not in ("anynoise",):
        scale = 0.01 if name in ("barebones", "hyperpriors", "warpedtime", "events") else 1.0
        logt[nts] = 0.5 * np.log(0.01 / scale) + 0.05 * rng.standard_normal()  # noise variance ~1e-2
    return X, y, logt


def make_device_gp(name):
    from gogp_b200 import GP
    ndim, ds, dn, _, _ = CASES[name]
    return GP(NDim=ndim, Simil=ds, Noise=dn)


def make_oracle_gp(name):
    from oracle.gp import GP
    ndim, _, _, os_, on = CASES[name]
    return GP(ndim, os_, on)


def relerr(a, b, floor=1.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0

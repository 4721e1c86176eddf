"""Host-side mirror of the tutorial drivers (gogp_b200/tutorial.py; reference tutorial/tutorial.go,
tutorial/{anynoise,warpedtime,hyperpriors}/model/model.go) on a machine without a GPU: the priors' closed-form
gradients against finite differences, the host L-BFGS, and tutorial.Evaluate in the reference's OPTINP mode (inputs and
outputs optimised with the hyper-parameters) with the ORACLE standing in for the device GP behind the same interface."""
import io

import numpy as np
import pytest

from gogp_b200 import tutorial as T
from tests import cases


class OracleBackedGP:
    """gp.GP's surface (NDim, Simil, Noise, X, Y, Observe, Gradient, Produce) over oracle.gp.GP: test infrastructure."""

    def __init__(self, name):
        self.NDim, self.Simil, self.Noise = cases.CASES[name][0], cases.CASES[name][1], cases.CASES[name][2]
        self.og = cases.make_oracle_gp(name)
        self.X, self.Y = [], []

    def Observe(self, x):
        P = self.Simil.NTheta() + self.Noise.NTheta()
        x = np.ascontiguousarray(x, dtype=np.float64)
        if len(x) == P:
            self.og.X, self.og.Y = np.asarray(self.X, dtype=np.float64).reshape(-1, self.NDim), np.asarray(self.Y)
        v = self.og.observe(x)
        self.X, self.Y = self.og.X, self.og.Y
        return v

    def Gradient(self):
        return self.og.gradient("fast")

    def Produce(self, Z):
        mu, sigma = self.og.produce(np.asarray(Z, dtype=np.float64), clamp=True)
        return mu, sigma, None


@pytest.mark.parametrize("cls", [T.AnynoisePriors, T.WarpedtimePriors, T.HyperPriors])
def test_priors_gradients_against_finite_differences(cls):
    rng = np.random.default_rng(1)
    if cls is T.HyperPriors:
        x0 = 0.3 * rng.standard_normal(6)
    else:
        n = 7
        x0 = np.concatenate([0.3 * rng.standard_normal(3), np.sort(rng.uniform(0, 5, n)), rng.standard_normal(n)])
    p = cls()
    p.Observe(x0)                       # the first call memoises the initial outputs / steps
    x = x0 + 0.05 * rng.standard_normal(len(x0))
    p.Observe(x)
    g = np.asarray(p.Gradient())
    h = 1e-6
    fd = np.array([(p.Observe(x + h * e) - p.Observe(x - h * e)) / (2 * h) for e in np.eye(len(x))])
    assert np.max(np.abs(fd - g)) <= 1e-7 * max(1.0, np.max(np.abs(g)))


def test_host_lbfgs_maximises():
    ros = lambda x: (-((1 - x[0]) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2),
                     -np.array([-2 * (1 - x[0]) - 400 * x[0] * (x[1] - x[0] ** 2), 200 * (x[1] - x[0] ** 2)]))
    x, f, it = T.lbfgs_ascent(ros, np.array([-1.2, 1.0]), 500, 1e-8)
    assert np.max(np.abs(x - 1.0)) < 1e-6 and it < 200
    # a trial point that cannot be evaluated is a rejected step, not an error
    def walled(x):
        if x[0] > 2.0:
            raise RuntimeError("not positive definite")
        return -(x[0] - 1.9) ** 2, np.array([-2 * (x[0] - 1.9)])
    x, f, it = T.lbfgs_ascent(walled, np.array([-50.0]), 100, 1e-10)
    assert abs(x[0] - 1.9) < 1e-6


@pytest.mark.parametrize("name,model,priors", [("anynoise", T.AnyNoise, T.AnynoisePriors),
                                               ("warpedtime", T.WarpedTime, T.WarpedtimePriors)])
def test_evaluate_in_optinp_mode(name, model, priors):
    """tutorial/anynoise and tutorial/warpedtime: OPTINP = true, the model is gp.Model with the tutorial's priors and
    its gradient mask.  On the first 14 rows of the shipped data: one output row per window in the reference's format,
    the optimisation never lowers the objective, the masked arguments do not move."""
    text = "".join(open(cases.GOLDEN + "/%s.csv" % name).readlines()[:14])
    g = OracleBackedGP(name)
    m = model(g, priors())
    P = g.Simil.NTheta() + g.Noise.NTheta()
    out = io.StringIO()
    rows = T.Evaluate(g, m, np.zeros(P), text, out, iters=8, optinp=True, rng=np.random.default_rng(5))
    lines = out.getvalue().splitlines()
    assert len(rows) == 14 and len(lines) == 14 and all(len(l.split(",")) == 1 + 5 + P for l in lines)
    assert all(np.isfinite(r).all() for r in rows)
    assert all(r[5] >= r[4] - 1e-9 for r in rows)           # lml >= lml0 in every window
    X, Y = T.load(text)
    mean, std = T.mean_std(Y)
    if name == "anynoise":                                   # inputs are held, outputs inferred
        assert np.allclose(np.asarray(g.X).reshape(-1), X[:13, 0])
        assert not np.allclose(np.asarray(g.Y), (Y[:13] - mean) / std)
    else:                                                    # first and last input and all outputs are held
        gx = np.asarray(g.X).reshape(-1)
        assert gx[0] == X[0, 0] and gx[-1] == X[12, 0] and np.allclose(np.asarray(g.Y), (Y[:13] - mean) / std)
        assert not np.allclose(gx[1:-1], X[1:12, 0])


def test_events_flag_format():
    assert T.parse_events("1.:2.5:0.3,3:6:0.5") == [(1.0, 2.5, 0.3), (3.0, 6.0, 0.5)] and T.parse_events("") == []
    with pytest.raises(ValueError):
        T.parse_events("1:x:0.3")
    from gogp_b200 import kernel as k
    e = k.Param(0) * k.Matern52.Of(l=1) * k.Events(T.parse_events("1.:2.5:0.3,3:6:0.5"))
    assert e.NTheta() == 2 and len(e.events) == 2

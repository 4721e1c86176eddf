"""Independent check of the oracle's mathematics against scikit-learn's GaussianProcessRegressor
(third-party, unrelated to the reference): log marginal likelihood and its gradient with respect to the
LOG hyper-parameters, for the kernel families the reference's own tests never pin (gp/gp_test.go covers
only kernel.Normal with N <= 2).  This does not replace the reference goldens -- it shows that the
restated formulas (values, log-scale chain rule, the 0.5 tr((aa^T - K^-1) dK) gradient) are the standard
ones at sizes and kernels beyond them.

Parameter correspondence (sklearn theta is log of: ConstantKernel value, length scale[, periodicity],
WhiteKernel noise_level):
    theta0 * Normal(l)            <-> ConstantKernel(theta0) * RBF(l)
    theta0 * Matern32(l)          <-> ConstantKernel(theta0) * Matern(l, nu=1.5)
    theta0 * Matern52_textbook(l) <-> ConstantKernel(theta0) * Matern(l, nu=2.5)
    theta0 * Periodic(l, p)       <-> ConstantKernel(theta0) * ExpSineSquared(l, p)
    UniformNoise(s): variance s^2 <-> WhiteKernel(s^2): d/dlog(s) = 2 d/dlog(noise_level)
"""
import numpy as np
import pytest
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, ExpSineSquared, Matern, WhiteKernel

from oracle import kernels as ok
from oracle.gp import GP


class _Scaled:
    """theta0 * leaf(theta1[, theta2]) over 1-D inputs, in the reference's argument layout."""

    def __init__(self, cov, nleaf):
        self.cov, self.ntheta = cov, 1 + nleaf

    def observe(self, x):
        n = self.ntheta
        return x[0] * self.cov(*x[1:n], x[n], x[n + 1])


CASES = {
    "rbf": (_Scaled(ok.normal_cov, 1), lambda c, l: ConstantKernel(c) * RBF(l)),
    "matern32": (_Scaled(ok.matern32_cov, 1), lambda c, l: ConstantKernel(c) * Matern(l, nu=1.5)),
    "matern52_textbook": (_Scaled(ok.matern52_textbook_cov, 1), lambda c, l: ConstantKernel(c) * Matern(l, nu=2.5)),
    "periodic": (_Scaled(ok.periodic_cov, 2), lambda c, l, p: ConstantKernel(c) * ExpSineSquared(l, p)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_lml_and_gradient_agree_with_sklearn(name):
    simil, make = CASES[name]
    rng = np.random.default_rng(17)
    N = 60
    X = np.sort(rng.uniform(0.0, 6.0, size=(N, 1)), axis=0)
    y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(N)
    nleaf = simil.ntheta - 1
    nat = np.concatenate([[0.8], [1.3, 2.1][:nleaf], [0.2]])       # theta0, l[, p], noise std
    g = GP(1, simil, ok.UniformNoise)
    g.X, g.Y = X, y
    lml = g.observe(np.log(nat))
    grad = g.gradient()

    kern = make(*nat[:1 + nleaf]) + WhiteKernel(nat[-1] ** 2)
    gpr = GaussianProcessRegressor(kernel=kern, optimizer=None, alpha=0.0).fit(X, y)
    theta = gpr.kernel_.theta                                       # log of (c, l[, p], noise_level)
    ref, gref = gpr.log_marginal_likelihood(theta, eval_gradient=True)
    gref = np.array(gref, dtype=np.float64)
    gref[-1] *= 2.0                                                 # d/dlog(std) = 2 d/dlog(variance)
    assert np.allclose(np.exp(theta[:-1]), nat[:-1]) and np.isclose(np.exp(theta[-1]), nat[-1] ** 2)
    assert abs(lml - ref) <= 1e-9 * max(1.0, abs(ref)), (lml, ref)
    assert np.max(np.abs(grad - gref)) <= 1e-7 * max(1.0, np.max(np.abs(gref))), (grad, gref)


@pytest.mark.parametrize("name", ["rbf", "matern32"])
def test_predictive_moments_agree_with_sklearn(name):
    """gp.GP.Produce (gp/gp.go:258-360): posterior mean and the LATENT standard deviation.  sklearn's
    predictive variance at new points includes the WhiteKernel level, the reference's sigma does not."""
    simil, make = CASES[name]
    rng = np.random.default_rng(23)
    N, M = 50, 15
    X = rng.uniform(0.0, 5.0, size=(N, 1))
    y = np.cos(X[:, 0]) + 0.1 * rng.standard_normal(N)
    Z = rng.uniform(-0.5, 5.5, size=(M, 1))
    nat = np.array([1.2, 0.9, 0.15])
    g = GP(1, simil, ok.UniformNoise)
    g.X, g.Y = X, y
    g.observe(np.log(nat))
    mu, sigma = g.produce(Z, clamp=True)
    kern = make(nat[0], nat[1]) + WhiteKernel(nat[2] ** 2)
    gpr = GaussianProcessRegressor(kernel=kern, optimizer=None, alpha=0.0).fit(X, y)
    mref, sref = gpr.predict(Z, return_std=True)
    assert np.max(np.abs(mu - mref)) <= 1e-7 * max(1.0, np.max(np.abs(mref)))
    assert np.max(np.abs(sigma ** 2 - (sref ** 2 - nat[2] ** 2))) <= 1e-7

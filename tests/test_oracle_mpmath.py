"""50-digit mpmath evaluation of the same formulas, to check the ORACLE itself
(the Go reference cannot be run here; see oracle/__init__.py)."""
import mpmath as mp
import pytest

from tests import cases

mp.mp.dps = 50
S3, S5 = mp.mpf("1.7320508075688772"), mp.mpf("2.2360679774997900")  # the reference's literals


def normal(l, a, b):
    d = (a - b) / l
    return mp.e ** (-d * d / 2)


def periodic(l, p, a, b):
    d = mp.sin(mp.pi * abs(a - b) / p) / l
    return mp.e ** (-2 * d * d)


def m32(l, a, b):
    d = abs(a - b) / l
    return (1 + S3 * d) * mp.e ** (-S3 * d)


def m52(l, a, b):
    d = abs(a - b) / l
    return (1 + S5 * d + d * d) * mp.e ** (-S5 * d)


KERNELS = {
    "barebones": (lambda t, a, b: t[0] * m32(t[1], a[0], b[0]), lambda t: mp.mpf("0.01") * t[0] ** 2, 2),
    "hyperpriors": (lambda t, a, b: t[0] * m52(t[2], a[0], b[0]) + t[1] * periodic(t[3], 10 * t[4], a[0], b[0]),
                    lambda t: mp.mpf("0.01") * t[0] ** 2, 5),
    "c3_ard3": (lambda t, a, b: t[0] * normal(t[1], a[0], b[0]) * normal(t[2], a[1], b[1]) * normal(t[3], a[2], b[2])
                * periodic(t[4], t[5], a[0], b[0]), lambda t: t[0] ** 2, 6),
    "periodic": (lambda t, a, b: periodic(t[0], t[1], a[0], b[0]), lambda t: mp.mpf("0.3") ** 2, 2),
}


def mp_lml(name, logt, X, y):
    simil, noise, nts = KERNELS[name]
    th = [mp.e ** mp.mpf(v) for v in logt]
    ts, tn = th[:nts], th[nts:]
    n = len(y)
    Xm = [[mp.mpf(float(v)) for v in row] for row in X]
    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            K[i, j] = simil(ts, Xm[i], Xm[j]) + (noise(tn) if i == j else 0)
    yv = mp.matrix([mp.mpf(float(v)) for v in y])
    L = mp.cholesky(K)
    alpha = mp.cholesky_solve(K, yv)
    logdet = 2 * sum(mp.log(L[i, i]) for i in range(n))
    return -mp.mpf(n) / 2 * mp.log(2 * mp.pi) - logdet / 2 - (yv.T * alpha)[0] / 2


@pytest.mark.parametrize("name", sorted(KERNELS))
def test_lml_and_gradient_against_mpmath(name):
    X, y, logt = cases.synth(name, 6, seed=11)
    g = cases.make_oracle_gp(name)
    g.X, g.Y = X, y
    lml = g.observe(logt.copy())
    grad = g.gradient()
    ref = mp_lml(name, logt, X, y)
    assert abs(lml - float(ref)) < 1e-11 * max(1.0, abs(float(ref)))
    for p in range(len(logt)):
        def f(t, p=p):
            lt = [mp.mpf(v) for v in logt]
            lt[p] = t
            return mp_lml(name, lt, X, y)
        d = mp.diff(f, mp.mpf(logt[p]))
        assert abs(grad[p] - float(d)) < 1e-10 * max(1.0, abs(float(d)))

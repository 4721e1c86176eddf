"""Restatement of the reference kernel library (test infrastructure).

Every kernel is, like in the reference, an object with ``observe(x)`` taking the
flat argument vector ``[theta... | xa (NDim) | xb (NDim)]`` (similarity,
gp/gp.go:110-111) or ``[theta... | x (NDim)]`` (noise, gp/gp.go:134-135) and
``ntheta``.  Elements of ``x`` are ``oracle.dual.Dual`` so the same code yields
values and all partials, as the reference's AD tape does.
"""
import math

from . import dual as ad

SQRT3 = 1.7320508075688772  # kernel/kernel.go:51
SQRT5 = 2.2360679774997900  # kernel/kernel.go:52


# --- kernel/kernel.go ---------------------------------------------------------
def normal_cov(l, xa, xb):
    """kernel/kernel.go:23-26"""
    d = (xa - xb) / l
    return ad.exp(-d * d / 2)


def periodic_cov(l, p, xa, xb):
    """kernel/kernel.go:44-47"""
    d = ad.sin(math.pi * ad.fabs(xa - xb) / p) / l
    return ad.exp(-2 * d * d)


def matern32_cov(l, xa, xb):
    """kernel/kernel.go:70-73"""
    d = ad.fabs(xa - xb) / l
    return (1 + SQRT3 * d) * ad.exp(-SQRT3 * d)


def matern52_cov(l, xa, xb):
    """kernel/kernel.go:89-92.  ``5/3`` is an untyped integer constant division
    in Go and evaluates to 1 (kernel/ad/kernel.go:130 records ad.Value(1))."""
    d = ad.fabs(xa - xb) / l
    return (1 + SQRT5 * d + (5 // 3) * d * d) * ad.exp(-SQRT5 * d)


def matern52_textbook_cov(l, xa, xb):
    """Not in the reference: the textbook 5/3 coefficient, offered as its own leaf."""
    d = ad.fabs(xa - xb) / l
    return (1 + SQRT5 * d + (5.0 / 3.0) * d * d) * ad.exp(-SQRT5 * d)


class _Leaf1:
    ntheta = 1

    def __init__(self, cov):
        self.cov = cov

    def observe(self, x):  # kernel/kernel.go:15-17,58-60,77-79
        return self.cov(x[0], x[1], x[2])


class _Periodic:
    ntheta = 2

    def observe(self, x):  # kernel/kernel.go:36-38
        return periodic_cov(x[0], x[1], x[2], x[3])


Normal = _Leaf1(normal_cov)
Matern32 = _Leaf1(matern32_cov)
Matern52 = _Leaf1(matern52_cov)
Periodic = _Periodic()


# --- kernel/noise.go ------------------------------------------------------------
class ConstantNoise:
    """kernel/noise.go:21-34 -- fixed standard error, no parameters."""
    ntheta = 0

    def __init__(self, std):
        self.std = float(std)

    def observe(self, x):
        return ad.Dual(self.std * self.std)


class _UniformNoise:
    """kernel/noise.go:39-53 -- one parameter, the standard error."""
    ntheta = 1

    def observe(self, x):
        return x[0] * x[0]


UniformNoise = _UniformNoise()


# --- tutorial compositions ------------------------------------------------------
class BarebonesSimil:
    """tutorial/barebones/kernel/kernel.go:14-18"""
    ntheta = 2

    def observe(self, x):
        return x[0] * Matern32.observe(x[1:])


class ScaledUniformNoise:
    """tutorial/barebones/kernel/kernel.go:25-31 (Noise(0.01)) and the
    ``0.01 * UniformNoise`` of hyperpriors / warpedtime kernel.go:34-36,30-32."""
    ntheta = 1

    def __init__(self, c):
        self.c = float(c)

    def observe(self, x):
        return self.c * UniformNoise.observe(x)


class HyperpriorsSimil:
    """tutorial/hyperpriors/kernel/kernel.go:12-27"""
    ntheta = 5

    def observe(self, x):
        c1, c2, l1, l2, p, xa, xb = range(7)
        return (x[c1] * matern52_cov(x[l1], x[xa], x[xb])
                + x[c2] * periodic_cov(x[l2], 10 * x[p], x[xa], x[xb]))


class AnynoiseSimil:
    """tutorial/anynoise/kernel/kernel.go:12-23 == warpedtime kernel.go:12-23"""
    ntheta = 2

    def observe(self, x):
        return x[0] * matern52_cov(x[1], x[2], x[3])


class AnynoiseNoise:
    """tutorial/anynoise/kernel/kernel.go:31-35: constant 1e-5 *variance*, one
    declared-but-unused parameter."""
    ntheta = 1

    def observe(self, x):
        return ad.Dual(1e-5)


class EventsSimil:
    """tutorial/events/kernel/kernel.go:10-46: scaled Matern52, discounted when the pair straddles
    an event boundary (first matching event only)."""
    ntheta = 2

    def __init__(self, events):
        self.events = [tuple(float(v) for v in e) for e in events]

    def observe(self, x):
        import numpy as np
        k = x[0] * Matern52.observe(x[1:])
        xa, xb = np.asarray(x[2].v), np.asarray(x[3].v)
        lo, hi = np.minimum(xa, xb), np.maximum(xa, xb)
        factor = np.ones(np.broadcast(lo, hi).shape)
        done = np.zeros(factor.shape, dtype=bool)
        for (frm, to, disc) in self.events:
            hit = ((lo < frm) & (frm <= hi)) | ((lo < to) & (to <= hi))
            factor = np.where(hit & ~done, disc, factor)
            done |= hit
        return k * ad.Dual(factor)


# --- synthetic benchmark kernels (SURVEY.md section 8 d; not in the reference) ---
class ScaledNormal1D:
    """C2: theta0 * Normal(l)."""
    ntheta = 2

    def observe(self, x):
        return x[0] * normal_cov(x[1], x[2], x[3])


class ArdNormalTimesPeriodic:
    """C3: theta0 * prod_d Normal(l_d; dim d) * Periodic(l_p, p; dim 0), NDim = D.
    theta = [theta0, l_0..l_{D-1}, l_p, p]."""

    def __init__(self, ndim):
        self.ndim = ndim
        self.ntheta = ndim + 3

    def observe(self, x):
        D = self.ndim
        nt = self.ntheta
        xa = x[nt:nt + D]
        xb = x[nt + D:nt + 2 * D]
        k = x[0]
        for d in range(D):
            k = k * normal_cov(x[1 + d], xa[d], xb[d])
        return k * periodic_cov(x[1 + D], x[2 + D], xa[0], xb[0])


class ArdMatern32:
    """C5: theta0 * prod_d Matern32(l_d; dim d)."""

    def __init__(self, ndim):
        self.ndim = ndim
        self.ntheta = ndim + 1

    def observe(self, x):
        D = self.ndim
        nt = self.ntheta
        k = x[0]
        for d in range(D):
            k = k * matern32_cov(x[1 + d], x[nt + d], x[nt + D + d])
        return k

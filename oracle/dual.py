"""Vectorised forward-mode automatic differentiation (test infrastructure).

The reference obtains d k / d(arg) for every element of the kernel's argument
vector from infergo's reverse-mode tape (``model.Gradient(gp.Simil)``,
gp/gp.go:113,137; generated code in kernel/ad/kernel.go).  The oracle needs the
same numbers for arbitrary user kernels written as plain functions of an
argument list, without trusting the closed-form partials the CUDA path uses.
A dual number carrying a sparse {arg index: derivative array} map does that;
values broadcast over all (i, j) pairs at once.
"""
import numpy as np


class Dual:
    __slots__ = ("v", "d")

    def __init__(self, v, d=None):
        self.v = np.asarray(v, dtype=np.float64)
        self.d = d if d is not None else {}

    # -- helpers -----------------------------------------------------------
    @staticmethod
    def lift(x):
        return x if isinstance(x, Dual) else Dual(x)

    def _unary(self, v, dv):
        return Dual(v, {k: dv * g for k, g in self.d.items()})

    # -- arithmetic --------------------------------------------------------
    def __add__(self, o):
        o = Dual.lift(o)
        d = dict(self.d)
        for k, g in o.d.items():
            d[k] = d[k] + g if k in d else g
        return Dual(self.v + o.v, d)

    __radd__ = __add__

    def __neg__(self):
        return Dual(-self.v, {k: -g for k, g in self.d.items()})

    def __sub__(self, o):
        return self + (-Dual.lift(o))

    def __rsub__(self, o):
        return Dual.lift(o) + (-self)

    def __mul__(self, o):
        o = Dual.lift(o)
        d = {k: g * o.v for k, g in self.d.items()}
        for k, g in o.d.items():
            t = g * self.v
            d[k] = d[k] + t if k in d else t
        return Dual(self.v * o.v, d)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = Dual.lift(o)
        inv = 1.0 / o.v
        q = self.v * inv
        d = {k: g * inv for k, g in self.d.items()}
        for k, g in o.d.items():
            t = -g * q * inv
            d[k] = d[k] + t if k in d else t
        return Dual(q, d)

    def __rtruediv__(self, o):
        return Dual.lift(o) / self


def exp(x):
    x = Dual.lift(x)
    e = np.exp(x.v)
    return x._unary(e, e)


def sin(x):
    x = Dual.lift(x)
    return x._unary(np.sin(x.v), np.cos(x.v))


def fabs(x):
    # d|r|/dr = sign(r), 0 at r == 0: every partial that reaches this branch is
    # multiplied by d == 0 in the shipped kernels, so the convention at the kink
    # cannot show in a result.
    x = Dual.lift(x)
    return x._unary(np.abs(x.v), np.sign(x.v))

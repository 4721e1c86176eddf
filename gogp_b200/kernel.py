"""Host-side mirror of GoGP's kernel library (reference kernel/kernel.go,
kernel/noise.go), lowered to the device descriptor.

In the reference a kernel is Go code behind ``gp.Kernel{Observe, NTheta}``
(gp/gp.go:14-17) and is called, through an AD tape, once per matrix element.  A
user Simil written in the host language cannot run on the device, so here a
kernel is a small expression over the stock kernels that carries its own
descriptor: the names and ``NTheta()`` stay, ``Observe`` is replaced by
``Descriptor()``.

    from gogp_b200 import kernel as k
    simil = k.Param(0) * k.Matern32.Of(l=1)                  # tutorial/barebones
    noise = 0.01 * k.UniformNoise                            # kernel.go:27-29
    simil = (k.Param(0) * k.Matern52.Of(l=2)
             + k.Param(1) * k.Periodic.Of(l=3, p=(4, 10.0))) # tutorial/hyperpriors
"""
from . import _lib


def _pidx(p):
    """parameter reference: index or (index, constant multiplier)"""
    if isinstance(p, tuple):
        return int(p[0]), float(p[1])
    return int(p), 1.0


class Kernel:
    """An expression in postfix form + the number of parameters it declares."""

    def __init__(self, ops, ntheta=None, is_noise=False, events=None):
        self.ops = list(ops)
        self.events = events  # the Events table of tutorial/events' Simil, if the expression has an Events leaf
        used = -1
        for o in self.ops:
            if o[0] == _lib.OP_PARAM:
                used = max(used, o[2])
            elif _lib.OP_NORMAL <= o[0] <= _lib.OP_MATERN52_TEXTBOOK:
                used = max(used, o[2], o[3] if o[0] == _lib.OP_PERIODIC else -1)
        self._ntheta = used + 1 if ntheta is None else int(ntheta)
        if self._ntheta < used + 1:
            raise ValueError("ntheta smaller than the largest parameter index used")
        self.is_noise = is_noise

    # gp.Kernel.NTheta (gp/gp.go:16)
    def NTheta(self):
        return self._ntheta

    def WithNTheta(self, n):
        """Declare more parameters than the expression uses (tutorial/anynoise's
        noise allocates one unused parameter, kernel.go:31-35)."""
        return Kernel(self.ops, n, self.is_noise, self.events)

    def Descriptor(self):
        """-> ctypes array of gogp_op"""
        arr = (_lib.Op * len(self.ops))()
        for a, (kind, dim, p0, p1, s0, s1, c) in zip(arr, self.ops):
            a.kind, a.dim = kind, dim
            a.param[0], a.param[1] = p0, p1
            a.scale[0], a.scale[1] = s0, s1
            a.constant = c
        return arr

    @staticmethod
    def _lift(x):
        if isinstance(x, Kernel):
            return x
        return Const(float(x))

    def __add__(self, o):
        o = Kernel._lift(o)
        return Kernel(self.ops + o.ops + [(_lib.OP_ADD, 0, 0, 0, 1.0, 1.0, 0.0)],
                      max(self._ntheta, o._ntheta), self.is_noise and o.is_noise, self.events or o.events)

    __radd__ = __add__

    def __mul__(self, o):
        o = Kernel._lift(o)
        return Kernel(self.ops + o.ops + [(_lib.OP_MUL, 0, 0, 0, 1.0, 1.0, 0.0)],
                      max(self._ntheta, o._ntheta), self.is_noise and o.is_noise, self.events or o.events)

    __rmul__ = __mul__


def Const(c):
    return Kernel([(_lib.OP_CONST, 0, 0, 0, 1.0, 1.0, float(c))], 0, True)


def Param(i, scale=1.0):
    """scale * theta[i]"""
    return Kernel([(_lib.OP_PARAM, 0, int(i), 0, float(scale), 1.0, 0.0)], None, True)


class _Stock(Kernel):
    """A stock 1-D similarity kernel; as a singleton it is the reference's
    kernel.Normal etc. with parameters [l] (or [l, p]) at indices 0 (, 1)."""

    def __init__(self, kind, nparam):
        self.kind, self.nparam = kind, nparam
        p1 = 1 if nparam == 2 else 0
        Kernel.__init__(self, [(kind, 0, 0, p1, 1.0, 1.0, 0.0)], nparam)

    def Of(self, l=0, p=None, dim=0):
        """The same kernel reading its length scale (and period) from other
        parameter slots, optionally constant-scaled, on input coordinate dim."""
        li, ls = _pidx(l)
        pi, ps = (0, 1.0) if p is None else _pidx(p)
        if self.nparam == 2 and p is None:
            raise ValueError("Periodic needs p=")
        return Kernel([(self.kind, int(dim), li, pi, ls, ps, 0.0)])


Normal = _Stock(_lib.OP_NORMAL, 1)                    # kernel/kernel.go:13-26
Periodic = _Stock(_lib.OP_PERIODIC, 2)                # kernel/kernel.go:34-47
Matern32 = _Stock(_lib.OP_MATERN32, 1)                # kernel/kernel.go:60-73
Matern52 = _Stock(_lib.OP_MATERN52, 1)                # kernel/kernel.go:79-92 (5/3 == 1 as shipped)
Matern52Textbook = _Stock(_lib.OP_MATERN52_TEXTBOOK, 1)


def Events(events, dim=0):
    """The discount of tutorial/events/kernel/kernel.go:33-44 as a factor: 1, or the discount of
    the first event (from, to, discount) whose from- or to-boundary separates the two points."""
    ev = [tuple(float(v) for v in e) for e in events]
    if any(len(e) != 3 for e in ev):
        raise ValueError("an event is (from, to, discount)")
    return Kernel([(_lib.OP_EVENTS, int(dim), 0, 0, 1.0, 1.0, 0.0)], 0, False, ev)


def ConstantNoise(std):
    """kernel.ConstantNoise (kernel/noise.go:21-34): variance std^2, no parameters."""
    std = float(std)
    return Const(std * std)


# kernel.UniformNoise (kernel/noise.go:39-53): variance theta[0]^2
UniformNoise = Param(0) * Param(0)

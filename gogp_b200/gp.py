"""Host-side mirror of gp.GP and gp.Model (reference gp/gp.go, gp/model.go) over
the C-ABI.  Same names, argument layout and error behaviour as the reference,
so the parity tests read like gp/gp_test.go:

    g = GP(NDim=1, Simil=kernel.Normal, Noise=kernel.ConstantNoise(0.1))
    ll = g.Observe(x)          # x = [log theta | X flat | Y]  or  [log theta]
    dll = g.Gradient()
    err = g.Absorb(X, Y); mu, sigma, err = g.Produce(Z)

The Go wrapper a maintainer would add is go/gp/gp.go; this
module is what can be executed where no Go toolchain exists.
"""
import ctypes as C

import numpy as np

from . import _lib


class GoGPError(Exception):
    """An ``error`` value of the reference (Absorb / Produce return it)."""

    def __init__(self, status, msg):
        Exception.__init__(self, msg)
        self.status = status


class GoGPPanic(RuntimeError):
    """Where the reference panics (Observe: gp/gp.go:398-405)."""

    def __init__(self, status, msg):
        RuntimeError.__init__(self, msg)
        self.status = status


def _flat(x, ndim):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64)).reshape(-1)
    if ndim and a.size % ndim:
        raise ValueError("input size is not a multiple of NDim")
    return a


class GP:
    """gp.GP (gp/gp.go:20-38)."""

    def __init__(self, NDim=1, Simil=None, Noise=None, ThetaSimil=None, ThetaNoise=None, X=None, Y=None,
                 Parallel=False, Device=0):
        self.NDim = NDim
        self.Simil = Simil
        self.Noise = Noise
        self.ThetaSimil = [] if ThetaSimil is None else list(ThetaSimil)
        self.ThetaNoise = [] if ThetaNoise is None else list(ThetaNoise)
        self.X = [] if X is None else X
        self.Y = [] if Y is None else Y
        self.Parallel = Parallel  # the GPU build is the parallel path; kept for API parity (gp/gp.go:31)
        self.Device = Device
        self._h = None
        self._key = None
        self._with_obs = False
        self._n = 0

    # -- handle management ------------------------------------------------------
    def _handle(self):
        key = (self.NDim, id(self.Simil), id(self.Noise), self.Device)
        if self._h is not None and key == self._key:
            return self._h
        self.close()
        if self.Simil is None:
            raise GoGPPanic(_lib.BAD_ARGUMENT, "GP.Simil is nil")
        L = _lib.lib()
        sd = self.Simil.Descriptor()
        if self.Noise is not None:
            nd = self.Noise.Descriptor()
            nn, ntn = len(nd), self.Noise.NTheta()
        else:
            nd, nn, ntn = None, 0, 0  # defaults(): ConstantNoise(1e-5), gp/gp.go:46-48
        h = C.c_void_p()
        st = L.gogp_create(self.NDim, sd, len(sd), self.Simil.NTheta(), nd, nn, ntn, self.Device, C.byref(h))
        if st != _lib.OK:
            msg = L.gogp_last_error(h).decode() if h else "gogp_create failed"
            if h:
                L.gogp_destroy(h)
            raise GoGPPanic(st, msg)
        self._h, self._key = h, key
        ev = getattr(self.Simil, "events", None)
        if ev:
            flat = np.ascontiguousarray(np.asarray(ev, dtype=np.float64).reshape(-1))
            st = L.gogp_set_events(h, _lib.dptr(flat), len(ev))
            if st != _lib.OK:
                raise GoGPPanic(st, self._err(st))
        return h

    def close(self):
        if self._h is not None:
            _lib.lib().gogp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self, st):
        return _lib.lib().gogp_last_error(self._h).decode() or _lib.lib().gogp_status_string(st).decode()

    def _nts(self):
        return self.Simil.NTheta()

    def _ntn(self):
        return self.Noise.NTheta() if self.Noise is not None else 0

    def _defaults(self):
        # gp/gp.go:45-57
        if len(self.ThetaSimil) == 0:
            self.ThetaSimil = [0.0] * self._nts()
        if len(self.ThetaNoise) == 0:
            self.ThetaNoise = [0.0] * self._ntn()

    # -- gp/gp.go:80-87 ------------------------------------------------------------
    def Absorb(self, x, y):
        """Returns None or a GoGPError (the reference returns ``err``)."""
        self._defaults()
        self.X, self.Y = x, y
        h = self._handle()
        X = _flat(x, self.NDim)
        Y = _flat(y, 0)
        n = len(Y)
        if X.size != n * self.NDim:
            return GoGPError(_lib.BAD_ARGUMENT, "len(x) != len(y)")
        ts = np.array(self.ThetaSimil, dtype=np.float64)
        tn = np.array(self.ThetaNoise, dtype=np.float64)
        st = _lib.lib().gogp_absorb(h, _lib.dptr(ts), _lib.dptr(tn), _lib.dptr(X), _lib.dptr(Y), n)
        self._with_obs, self._n = False, n
        if st != _lib.OK:
            return GoGPError(st, self._err(st))
        return None

    # -- tutorial/tutorial.go:91-179, the window that grows ---------
    def Extend(self, x, y):
        """Append observations to the absorbed ones at UNCHANGED hyper-parameters: the factor is extended
        (O(N^2) per appended point) instead of recomputed.  Returns None or a GoGPError, like Absorb; the result equals
        Absorb on the concatenated data."""
        h = self._handle()
        Xn = _flat(x, self.NDim)
        Yn = _flat(y, 0)
        m = len(Yn)
        if Xn.size != m * self.NDim:
            return GoGPError(_lib.BAD_ARGUMENT, "len(x) != len(y)")
        out = C.c_double(0.0)
        st = _lib.lib().gogp_extend(h, _lib.dptr(Xn), _lib.dptr(Yn), m, C.byref(out))
        if st != _lib.OK and st != _lib.ILL_CONDITIONED:
            return GoGPError(st, self._err(st))
        if m:
            X0 = _flat(self.X, self.NDim).reshape(-1, self.NDim)
            self.X = np.concatenate([X0, Xn.reshape(m, self.NDim)])
            self.Y = np.concatenate([_flat(self.Y, 0), Yn])
        self._n += m
        self._with_obs = False
        return None if st == _lib.OK else GoGPError(st, self._err(st))

    # -- gp/gp.go:244-253 ------------------------------------------------------------
    def LML(self):
        out = C.c_double(0.0)
        st = _lib.lib().gogp_lml(self._handle(), C.byref(out))
        if st != _lib.OK:
            raise GoGPPanic(st, self._err(st))
        return out.value

    # -- gp/gp.go:258-360 ------------------------------------------------------------
    def Produce(self, x):
        """-> (mu, sigma, err)"""
        self._defaults()
        h = self._handle()
        Z = _flat(x, self.NDim)
        m = Z.size // self.NDim
        mu = np.zeros(m)
        sigma = np.zeros(m)
        st = _lib.lib().gogp_produce(h, _lib.dptr(Z), m, _lib.dptr(mu), _lib.dptr(sigma))
        if st != _lib.OK:
            return None, None, GoGPError(st, self._err(st))
        return mu, sigma, None

    # -- gp/gp.go:374-413 ------------------------------------------------------------
    def Observe(self, x):
        """x: float64 numpy array, [log theta] or [log theta | X flat | Y].  As in
        the reference the parameter prefix is exponentiated in place for the
        duration of the call and logged back (gp/gp.go:378-381, 408-410)."""
        self._defaults()
        h = self._handle()
        if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous):
            raise TypeError("Observe takes a contiguous float64 numpy array")
        nts, ntn, D = self._nts(), self._ntn(), self.NDim
        P = nts + ntn
        if len(x) < P:
            raise GoGPPanic(_lib.BAD_ARGUMENT, "len(x)")
        theta = x[:P]
        logt = theta.copy()
        np.exp(theta, out=theta)
        try:
            self.ThetaSimil = list(theta[:nts])
            self.ThetaNoise = list(theta[nts:P])
            rest = x[P:]
            self._with_obs = len(rest) > 0
            if self._with_obs:
                n = len(rest) // (D + 1)
                if n * (D + 1) != len(rest):
                    raise GoGPPanic(_lib.BAD_ARGUMENT, "len(x)")  # gp/gp.go:398-400
                self.X = rest[:n * D].reshape(n, D)  # views into x, as the reference aliases
                self.Y = rest[n * D:]
                X, Y = rest[:n * D], rest[n * D:]
            else:
                X = _flat(self.X, D)
                Y = _flat(self.Y, 0)
                n = len(Y)
                if X.size != n * D:
                    raise GoGPPanic(_lib.BAD_ARGUMENT, "len(gp.X) != len(gp.Y)")
            out = C.c_double(0.0)
            st = _lib.lib().gogp_observe(h, _lib.dptr(logt), 1 if self._with_obs else 0, _lib.dptr(X), _lib.dptr(Y),
                                         n, C.byref(out))
            self._n = n
            if st != _lib.OK:
                raise GoGPPanic(st, self._err(st))  # panic(err), gp/gp.go:403-405
        finally:
            np.log(theta, out=theta)
        return out.value

    # -- gp/gp.go:418-499 ------------------------------------------------------------
    def Gradient(self):
        h = self._handle()
        P = self._nts() + self._ntn()
        n = P + (self._n * (self.NDim + 1) if self._with_obs else 0)
        grad = np.zeros(n)
        st = _lib.lib().gogp_gradient(h, _lib.dptr(grad), n)
        if st != _lib.OK:
            raise GoGPPanic(st, self._err(st))
        return grad

    # -- tutorial/tutorial.go:124-175, run inside the library ----
    def Optimize(self, x, alg="lbfgs", iters=1000, threshold=1e-6, rate=0.01, priors=None, history=0):
        """The tutorial's MLE loop over gp.X, gp.Y without leaving the library: ``alg`` is "lbfgs"
        (optimize.Minimize, tutorial/tutorial.go:131-155) or "adam" (infer.Adam, :156-168); ``iters``,
        ``threshold``, ``rate`` are ITERS, THRESHOLD, RATE.  ``priors`` is gp.Model's Priors (an object
        with Observe(x) and Gradient()) or None.  x (log hyper-parameters, float64) is updated in
        place; returns a dict with iters, evals (Observe calls), grads (Gradient calls), lml0, lml, converged."""
        self._defaults()
        h = self._handle()
        L = _lib.lib()
        P = self._nts() + self._ntn()
        if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous and len(x) == P):
            raise TypeError("Optimize takes a contiguous float64 numpy array of the %d log hyper-parameters" % P)
        X = _flat(self.X, self.NDim)
        Y = _flat(self.Y, 0)
        n = len(Y)
        if X.size != n * self.NDim:
            raise GoGPPanic(_lib.BAD_ARGUMENT, "len(gp.X) != len(gp.Y)")
        st = L.gogp_set_data(h, _lib.dptr(X), _lib.dptr(Y), n)
        if st != _lib.OK:
            raise GoGPPanic(st, self._err(st))
        settings = _lib.OptSettings(method={"adam": 0, "lbfgs": 1}[alg], max_iters=iters, threshold=threshold,
                                    rate=rate, beta1=0.0, beta2=0.0, eps=0.0, history=history)
        result = _lib.OptResult()

        def _prior(ctx, xp, np_, gp_):
            xs = np.ctypeslib.as_array(xp, shape=(np_,)).copy()
            ll = float(priors.Observe(xs))
            g = np.asarray(priors.Gradient(), dtype=np.float64)
            ga = np.ctypeslib.as_array(gp_, shape=(np_,))
            ga[:len(g)] += g
            return ll

        cb = _lib.PRIOR_FN(_prior) if priors is not None else C.cast(None, _lib.PRIOR_FN)
        st = L.gogp_optimize(h, C.byref(settings), _lib.dptr(x), cb, None, C.byref(result))
        if st != _lib.OK:
            raise GoGPPanic(st, self._err(st))
        self._with_obs, self._n = False, n
        self.ThetaSimil = list(np.exp(x[:self._nts()]))
        self.ThetaNoise = list(np.exp(x[self._nts():]))
        return {"iters": result.iters, "evals": result.evals, "grads": result.grads, "lml0": result.lml0,
                "lml": result.lml, "converged": bool(result.converged)}

    # -- gp/gp.go:35-36, 255-257: "Produce works on stored results" ---------------------------
    def State(self):
        """What the reference lets a user keep: the exported L and Alpha (plus the parameters and inputs they
        belong to).  L is the N x N row-major lower Cholesky factor."""
        n = self._n
        return {"ThetaSimil": list(self.ThetaSimil), "ThetaNoise": list(self.ThetaNoise),
                "X": np.array(_flat(self.X, self.NDim)).reshape(n, self.NDim), "Alpha": self.Alpha(), "L": self.L()}

    def Restore(self, state):
        """Restore a stored state before Produce; returns None or a GoGPError."""
        h = self._handle()
        X = _flat(state["X"], self.NDim)
        alpha = _flat(state["Alpha"], 0)
        n = len(alpha)
        Lf = np.ascontiguousarray(np.asarray(state["L"], dtype=np.float64))
        if X.size != n * self.NDim or Lf.shape != (n, n):
            return GoGPError(_lib.BAD_ARGUMENT, "state shapes do not agree")
        self.ThetaSimil, self.ThetaNoise = list(state["ThetaSimil"]), list(state["ThetaNoise"])
        ts = np.array(self.ThetaSimil if self._nts() else [0.0], dtype=np.float64)
        tn = np.array(self.ThetaNoise if self._ntn() else [0.0], dtype=np.float64)
        st = _lib.lib().gogp_set_state(h, _lib.dptr(ts), _lib.dptr(tn), _lib.dptr(X), n, _lib.dptr(alpha),
                                       _lib.dptr(Lf.reshape(-1)))
        if st != _lib.OK:
            return GoGPError(st, self._err(st))
        self.X = X.reshape(n, self.NDim)
        self._with_obs, self._n = False, n
        return None

    # -- extras over the C-ABI ---------------------------------------------------------
    def PhaseTimes(self):
        ms = np.zeros(len(_lib.PHASES))
        _lib.lib().gogp_phase_times(self._handle(), _lib.dptr(ms))
        return dict(zip(_lib.PHASES, ms.tolist()))

    def Launches(self):
        return int(_lib.lib().gogp_launch_count(self._handle()))

    def Alpha(self):
        a = np.zeros(self._n)
        st = _lib.lib().gogp_get_alpha(self._handle(), _lib.dptr(a), self._n)
        if st != _lib.OK:
            raise GoGPPanic(st, self._err(st))
        return a

    def L(self):
        """The N x N row-major lower Cholesky factor (the reference's exported gp.L, gp/gp.go:35)."""
        n = self._n
        out = np.zeros((n, n))
        if n:
            st = _lib.lib().gogp_get_factor(self._handle(), _lib.dptr(out.reshape(-1)), n)
            if st != _lib.OK:
                raise GoGPPanic(st, self._err(st))
        return out


class Model:
    """gp.Model (gp/model.go:9-28): GP plus priors on the hyper-parameters.
    ``Priors`` is any object with Observe(x) -> float and Gradient() -> array."""

    def __init__(self, GP, Priors):
        self.GP = GP
        self.Priors = Priors
        self.gGrad = None
        self.pGrad = None

    def Observe(self, x):
        gll = self.GP.Observe(x)
        self.gGrad = self.GP.Gradient()
        pll = self.Priors.Observe(x)
        self.pGrad = np.asarray(self.Priors.Gradient(), dtype=np.float64)
        return gll + pll

    def Gradient(self):
        self.gGrad[:len(self.pGrad)] += self.pGrad
        return self.gGrad

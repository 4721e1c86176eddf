"""NumPy restatement of gp.GP (reference gp/gp.go) -- TEST INFRASTRUCTURE ONLY.

Follows gp/gp.go step by step: ``absorb`` (:89-239) builds K and, when a
gradient is wanted, the materialised dK list; ``LML`` (:244-253); ``Produce``
(:258-360); ``Observe`` (:374-413) with its argument layout and the in-place
exp/log of the caller's slice; ``Gradient`` (:418-499).  ``gradient()`` has two
modes: ``literal`` repeats the reference's per-parameter GEMM + Cholesky solve +
trace (:473-485); ``fast`` evaluates the same trace as
0.5 * sum_ij (alpha alpha^T - K^-1)_ij dK_ij.  They agree to rounding
(tests/test_oracle.py) and ``fast`` is what larger parity cases use.

gonum's Cholesky is replaced by LAPACK dpotrf/dpotrs through scipy: a different
summation order in the same precision (see SURVEY.md section 7 "hard parts").
"""
import math
import time

import numpy as np
import scipy.linalg as sla

from . import dual as ad
from .kernels import ConstantNoise

NONOISE = 1e-5  # gp/gp.go:43


class NotPositiveDefinite(Exception):
    """absorb's ``Factorize`` failure (gp/gp.go:228-230)."""


def _pairs(kernel, theta, A, B, want):
    """Evaluate kernel.observe on every (a, b) pair of rows of A (n x D) and
    B (m x D).  ``want`` selects which arguments carry derivatives:
    'none', 'theta', or 'all' (theta + both inputs).  Returns (value n x m,
    {arg index: n x m derivative}) like Simil.Observe + model.Gradient."""
    nt = kernel.ntheta
    D = A.shape[1]
    args = []
    for p in range(nt):
        seed = {p: np.float64(1.0)} if want in ("theta", "all") else {}
        args.append(ad.Dual(theta[p], seed))
    for d in range(D):
        seed = {nt + d: np.float64(1.0)} if want == "all" else {}
        args.append(ad.Dual(A[:, d][:, None], seed))
    for d in range(D):
        seed = {nt + D + d: np.float64(1.0)} if want == "all" else {}
        args.append(ad.Dual(B[:, d][None, :], seed))
    out = kernel.observe(args)
    shape = (A.shape[0], B.shape[0])
    val = np.broadcast_to(out.v, shape).astype(np.float64)
    grads = {k: np.broadcast_to(g, shape).astype(np.float64) for k, g in out.d.items()}
    return val, grads


def _noise(kernel, theta, X, want):
    """Noise.Observe([theta_n | x]) for every row of X (gp/gp.go:133-150)."""
    nt = kernel.ntheta
    N, D = X.shape
    args = []
    for p in range(nt):
        seed = {p: np.float64(1.0)} if want in ("theta", "all") else {}
        args.append(ad.Dual(theta[p], seed))
    for d in range(D):
        seed = {nt + d: np.float64(1.0)} if want == "all" else {}
        args.append(ad.Dual(X[:, d], seed))
    out = kernel.observe(args)
    val = np.broadcast_to(out.v, (N,)).astype(np.float64)
    grads = {k: np.broadcast_to(g, (N,)).astype(np.float64) for k, g in out.d.items()}
    return val, grads


def _sym_from_upper(M):
    """What K.SetSym(i, j, v) over j >= i leaves behind (gp/gp.go:155,220-225)."""
    U = np.triu(M)
    return U + np.triu(M, 1).T


class GP:
    """gp.GP (gp/gp.go:20-38)."""

    def __init__(self, ndim, simil, noise=None, theta_simil=None, theta_noise=None):
        self.NDim = ndim
        self.Simil = simil
        self.Noise = noise
        self.ThetaSimil = None if theta_simil is None else np.array(theta_simil, dtype=np.float64)
        self.ThetaNoise = None if theta_noise is None else np.array(theta_noise, dtype=np.float64)
        self.X = np.zeros((0, ndim))
        self.Y = np.zeros(0)
        self.with_obs = False
        self.L = None
        self.Alpha = None
        self.dK = None
        self.K = None
        # wall-clock split used by bench.py's CPU baseline: O(N^2) element work
        # (kernel evaluation, dK assembly, elementwise sums) vs O(N^3) dense algebra
        self.t_elem = 0.0
        self.t_dense = 0.0

    # gp/gp.go:45-57
    def _defaults(self):
        if self.Noise is None:
            self.Noise = ConstantNoise(NONOISE)
        if self.ThetaSimil is None or len(self.ThetaSimil) == 0:
            self.ThetaSimil = np.zeros(self.Simil.ntheta)
        if self.ThetaNoise is None or len(self.ThetaNoise) == 0:
            self.ThetaNoise = np.zeros(self.Noise.ntheta)

    # gp/gp.go:80-87
    def absorb(self, x, y):
        self._defaults()
        self.X = np.asarray(x, dtype=np.float64).reshape(-1, self.NDim)
        self.Y = np.asarray(y, dtype=np.float64)
        self._absorb(False)

    # gp/gp.go:89-239
    def _absorb(self, with_grad):
        nts, ntn, D = self.Simil.ntheta, self.Noise.ntheta, self.NDim
        N = len(self.X)
        self.dK = None
        self._parts = None
        if with_grad:
            self._parts = {}
        if N == 0:
            return
        want = "none" if not with_grad else ("all" if self.with_obs else "theta")
        t0 = time.perf_counter()
        k, kg = _pairs(self.Simil, self.ThetaSimil, self.X, self.X, want)
        n, ng = _noise(self.Noise, self.ThetaNoise, self.X, want)
        K = _sym_from_upper(k)
        K[np.diag_indices(N)] += n
        self.K = K
        if with_grad:
            # chain rule for the log-parameters, gp/gp.go:114-116,138-140
            self._parts = dict(kg=kg, ng=ng)
        t1 = time.perf_counter()
        self.t_elem += t1 - t0
        try:
            self.L = sla.cholesky(K, lower=True, check_finite=False)
        except np.linalg.LinAlgError as e:  # gp/gp.go:228-230
            raise NotPositiveDefinite(str(e))
        self.Alpha = sla.cho_solve((self.L, True), self.Y, check_finite=False)
        self.t_dense += time.perf_counter() - t1

    def _dK_theta(self, p):
        """dK/d log theta_p as the dense symmetric matrix the reference stores
        (gp/gp.go:113-117 for p < nts, :136-141 for the noise parameters)."""
        nts = self.Simil.ntheta
        N = len(self.X)
        if p < nts:
            g = self._parts["kg"].get(p)
            if g is None:
                return np.zeros((N, N))
            return _sym_from_upper(g) * self.ThetaSimil[p]
        q = p - nts
        g = self._parts["ng"].get(q)
        out = np.zeros((N, N))
        if g is not None:
            out[np.diag_indices(N)] = g * self.ThetaNoise[q]
        return out

    def _dK_input(self, i, d):
        """dK/dx_{i,d} (gp/gp.go:118-129 and :143-149): row/column i only."""
        nts, ntn, D = self.Simil.ntheta, self.Noise.ntheta, self.NDim
        N = len(self.X)
        kg, ng = self._parts["kg"], self._parts["ng"]
        ga = kg.get(nts + d)      # d/d xa_d, evaluated at (row, col)
        gb = kg.get(nts + D + d)  # d/d xb_d
        out = np.zeros((N, N))
        if ga is not None:
            # pairs (i, j), j >= i: x_i is the first argument
            out[i, i:] += ga[i, i:]
        if gb is not None:
            # pairs (j, i), j <= i: x_i is the second argument
            out[:i + 1, i] += gb[:i + 1, i]
        gn = ng.get(ntn + d)
        if gn is not None:
            out[i, i] += gn[i]
        # SetSym mirrors every write
        return np.triu(out) + np.triu(out, 1).T

    # gp/gp.go:244-253
    def lml(self):
        N = len(self.X)
        if N == 0:
            return 0.0
        lml = -0.5 * N * math.log(2 * math.pi)
        lml -= 0.5 * (2.0 * np.sum(np.log(np.diag(self.L))))
        lml -= 0.5 * float(self.Y @ self.Alpha)
        return lml

    # gp/gp.go:374-413
    def observe(self, x):
        self._defaults()
        x = np.asarray(x)
        assert x.dtype == np.float64
        nts, ntn, D = self.Simil.ntheta, self.Noise.ntheta, self.NDim
        P = nts + ntn
        x[:P] = np.exp(x[:P])  # in place, like the reference
        self.ThetaSimil[:] = x[:nts]
        self.ThetaNoise[:] = x[nts:P]
        rest = x[P:]
        self.with_obs = len(rest) > 0
        if self.with_obs:
            n = len(rest) // (D + 1)
            if n * (D + 1) != len(rest):
                x[:P] = np.log(x[:P])
                raise ValueError("len(x)")  # panic("len(x)"), gp/gp.go:398-400
            self.X = rest[:n * D].reshape(n, D)
            self.Y = rest[n * D:]
        try:
            self._absorb(True)
        finally:
            x[:P] = np.log(x[:P])
        return self.lml()

    # gp/gp.go:418-499
    def gradient(self, mode="fast"):
        nts, ntn, D = self.Simil.ntheta, self.Noise.ntheta, self.NDim
        P = nts + ntn
        N = len(self.X)
        grad = np.zeros(P + (N * (D + 1) if self.with_obs else 0))
        if N == 0:
            return grad
        a = np.outer(self.Alpha, self.Alpha)
        ndk = P + (N * D if self.with_obs else 0)
        if mode == "literal":
            for p in range(ndk):
                t0 = time.perf_counter()
                dK = self._dK_theta(p) if p < P else self._dK_input((p - P) // D, (p - P) % D)
                t1 = time.perf_counter()
                r0 = a @ dK
                r1 = sla.cho_solve((self.L, True), dK, check_finite=False)
                grad[p] = 0.5 * np.trace(r0 - r1)
                self.t_elem += t1 - t0
                self.t_dense += time.perf_counter() - t1
        else:
            t0 = time.perf_counter()
            Kinv = sla.cho_solve((self.L, True), np.eye(N), check_finite=False)
            self.t_dense += time.perf_counter() - t0
            t0 = time.perf_counter()
            W = a - Kinv
            for p in range(P):
                grad[p] = 0.5 * np.sum(W * self._dK_theta(p))
            if self.with_obs:
                kg, ng = self._parts["kg"], self._parts["ng"]
                up = np.triu(np.ones((N, N), dtype=bool), 1)
                Wu = np.where(up, W, 0.0)
                dg = np.diag(W)
                for d in range(D):
                    ga = kg.get(nts + d)
                    gb = kg.get(nts + D + d)
                    g = np.zeros(N)
                    if ga is not None:
                        g += np.sum(Wu * ga, axis=1) + 0.5 * dg * np.diag(ga)
                    if gb is not None:
                        g += np.sum(Wu * gb, axis=0) + 0.5 * dg * np.diag(gb)
                    gn = ng.get(ntn + d)
                    if gn is not None:
                        g += 0.5 * dg * gn
                    grad[P + d:P + N * D:D] = g
            self.t_elem += time.perf_counter() - t0
        if self.with_obs:
            grad[ndk:] = -self.Alpha
        self.dK = None
        return grad

    # gp/gp.go:258-360
    def produce(self, z, clamp=False):
        self._defaults()
        Z = np.asarray(z, dtype=np.float64).reshape(-1, self.NDim)
        M = len(Z)
        variance = np.empty(M)
        for i in range(M):  # prior variance, gp/gp.go:270-278
            v, _ = _pairs(self.Simil, self.ThetaSimil, Z[i:i + 1], Z[i:i + 1], "none")
            variance[i] = v[0, 0]
        if len(self.X) > 0:
            Kstar, _ = _pairs(self.Simil, self.ThetaSimil, self.X, Z, "none")
            mean = Kstar.T @ self.Alpha
            v = sla.cho_solve((self.L, True), Kstar, check_finite=False)
            cov = np.einsum("ij,ij->j", Kstar, v)  # diagonal of Kstar^T v
        else:
            mean = np.zeros(M)
            cov = np.zeros(M)
        rad = variance - cov
        if clamp:
            rad = np.maximum(rad, 0.0)
        with np.errstate(invalid="ignore"):
            sigma = np.sqrt(rad)
        return mean, sigma


class Model:
    """gp.Model (gp/model.go:9-28): GP + priors with additive log-density."""

    def __init__(self, gp, priors):
        self.gp = gp
        self.priors = priors  # object with observe(x) -> float and gradient() -> array
        self.g_grad = None
        self.p_grad = None

    def observe(self, x, mode="fast"):
        gll = self.gp.observe(x)
        self.g_grad = self.gp.gradient(mode)
        pll = self.priors.observe(x)
        self.p_grad = self.priors.gradient()
        return gll + pll

    def gradient(self):
        self.g_grad[:len(self.p_grad)] += self.p_grad
        return self.g_grad


def mean_std(y):
    """gonum stat.MeanStdDev(y, nil): mean and the unbiased (n-1) standard
    deviation, as used by tutorial/tutorial.go:78-86."""
    y = np.asarray(y, dtype=np.float64)
    return float(np.mean(y)), float(np.std(y, ddof=1))

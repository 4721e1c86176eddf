"""Host-side mirror of tutorial.Evaluate (reference tutorial/tutorial.go:56-230) over the device GP: the caller
and the data formats either side of the hot path.

    load(rdr)       CSV rows "x_1,...,x_D,y" -> (X, Y)                                   tutorial.go:234-272
    Evaluate(...)   the expanding window: for end = 0 .. len(X)-1 fit the first `end` observations, forecast the
                    next one, write "z...,y,mu,sigma,lml0,lml,theta..." (%f) per step             tutorial.go:91-197

The reference re-jitters the hyper-parameters for every window (tutorial.go:119-121) and then optimises them, so
consecutive windows never share a factorisation.  With ``jitter=0`` and ``optimise=False`` they do, and the window
grows by GP.Extend -- the factor is extended in O(N^2) per step instead of recomputed in O(N^3) (gogp_extend).
"""
import csv
import io
import math

import numpy as np

# the reference's package-level settings (tutorial.go:21-33)
ALG, ITERS, MINOPT, THRESHOLD, RATE = "lbfgs", 1000, 0, 1e-6, 0.01


class HyperPriors:
    """The Priors of tutorial/hyperpriors (tutorial/hyperpriors/model/model.go:10-40) restated for gp.Model: Normal log-densities on the log
    hyper-parameters (c1, c2, l1, l2, p, s); c2's mean depends on c1."""

    @staticmethod
    def _logp(mu, sigma, x):
        z = (x - mu) / sigma
        return -0.5 * z * z - math.log(sigma) - 0.5 * math.log(2 * math.pi)

    def Observe(self, x):
        self.x = np.array(x, dtype=np.float64)
        c1, c2, l1, l2, p, s = self.x
        return (self._logp(-1, 1, c1) + self._logp(c1 - math.log(2), 1, c2) + self._logp(0, 2, l1) +
                self._logp(0, 2, l2) + self._logp(0, 1, p) + self._logp(0, 1, s))

    def Gradient(self):
        c1, c2, l1, l2, p, s = self.x
        d2 = c2 - (c1 - math.log(2))
        return np.array([-(c1 + 1) + d2, -d2, -l1 / 4, -l2 / 4, -p, -s])

    def sample(self, rng):
        c1 = -1 + rng.standard_normal()
        return np.array([c1, c1 - math.log(2) + rng.standard_normal(), 2 * rng.standard_normal(),
                         2 * rng.standard_normal(), rng.standard_normal(), rng.standard_normal()])



def _normal_logp(mu, sigma, x):
    z = (x - mu) / sigma
    return -0.5 * z * z - math.log(sigma) - 0.5 * math.log(2 * math.pi)


class AnynoisePriors:
    """The Priors of tutorial/anynoise (tutorial/anynoise/model/model.go:8-45) for x = [log c, log l, log s | X | Y]:
    Normal priors on the three log hyper-parameters and LAPLACIAN observation noise -- Expon.Logp(1 / e^s, |y_i - x_y_i|)
    between the memoised initial outputs y and the inferred ones.  Gradient in closed form (the reference gets it from
    infergo's tape)."""

    def __init__(self):
        self.Y = None

    def Observe(self, x):
        self.x = np.array(x, dtype=np.float64)
        n = (len(self.x) - 3) // 2
        if self.Y is None or len(self.Y) != n:
            self.Y = self.x[3 + n:].copy()            # first call: memoise the initial outputs (model.go:21-25)
        c, l, s = self.x[:3]
        lam = math.exp(-s)
        r = np.abs(self.Y - self.x[3 + n:])
        return (_normal_logp(-1, 1, c) + _normal_logp(0, 2, l) + _normal_logp(-1, 2, s) +
                float(np.sum(math.log(lam) - lam * r)))

    def Gradient(self):
        n = (len(self.x) - 3) // 2
        c, l, s = self.x[:3]
        lam = math.exp(-s)
        d = self.Y - self.x[3 + n:]
        g = np.zeros(len(self.x))
        g[0] = -(c + 1)
        g[1] = -l / 4
        g[2] = -(s + 1) / 4 + float(np.sum(-1.0 + lam * np.abs(d)))
        g[3 + n:] = lam * np.sign(d)
        return g


class WarpedtimePriors:
    """The Priors of tutorial/warpedtime (tutorial/warpedtime/model/model.go:8-62): Normal priors on the log
    hyper-parameters, and every step between consecutive inputs, relative to its memoised initial length, is
    Normal(1, e^LogSigma) -- the inputs may move slightly."""

    def __init__(self, LogSigma=math.log(0.5)):
        self.LogSigma = LogSigma
        self.step = None

    def Observe(self, x):
        self.x = np.array(x, dtype=np.float64)
        n = (len(self.x) - 3) // 2
        xi = self.x[3:3 + n]
        if self.step is None or len(self.step) != max(n - 1, 0):
            self.step = np.diff(xi).copy() if n > 1 else np.zeros(0)   # memoised, not differentiated (model.go:22-40)
        c, l, s = self.x[:3]
        ll = _normal_logp(-1, 1, c) + _normal_logp(0, 2, l) + _normal_logp(0.5, 1, s)
        sig = math.exp(self.LogSigma)
        for i in range(len(self.step)):
            ll += _normal_logp(1, sig, (xi[i + 1] - xi[i]) / self.step[i])
        return ll

    def Gradient(self):
        n = (len(self.x) - 3) // 2
        xi = self.x[3:3 + n]
        c, l, s = self.x[:3]
        g = np.zeros(len(self.x))
        g[0], g[1], g[2] = -(c + 1), -l / 4, -(s - 0.5)
        sig2 = math.exp(2 * self.LogSigma)
        for i in range(len(self.step)):
            dz = -((xi[i + 1] - xi[i]) / self.step[i] - 1) / sig2 / self.step[i]
            g[3 + i + 1] += dz
            g[3 + i] -= dz
        return g


class _MaskedModel:
    """gp.Model whose gradient is wiped on some arguments (the tutorials' AnyNoise / WarpedTime wrappers)."""

    def __init__(self, GP, Priors):
        from .gp import Model
        self.Model = Model(GP, Priors)
        self.GP, self.Priors = GP, Priors

    def Observe(self, x):
        return self.Model.Observe(x)


class AnyNoise(_MaskedModel):
    """tutorial/anynoise/main.go:28-41: the inputs stay where they are, the outputs are inferred."""

    def Gradient(self):
        grad = self.Model.Gradient()
        first = self.GP.Simil.NTheta() + self.GP.Noise.NTheta()
        grad[first:first + len(self.GP.X)] = 0.0
        return grad


class WarpedTime(_MaskedModel):
    """tutorial/warpedtime/main.go:44-57: the first and the last input and all outputs stay where they are."""

    def Gradient(self):
        grad = self.Model.Gradient()
        first = self.GP.Simil.NTheta() + self.GP.Noise.NTheta()
        last = first + len(self.GP.X) - 1
        grad[first] = 0.0
        grad[last:] = 0.0
        return grad


def lbfgs_ascent(value_grad, x0, iters, threshold, history=15):
    """Host-driven L-BFGS (two-loop recursion, backtracking Armijo line search), MAXIMISING: what
    optimize.Minimize(FuncGrad(m)) does for the reference when the inputs are optimised too (OPTINP,
    tutorial.go:131-155) and the argument is [log theta | X | Y], which gogp_optimize (hyper-parameters over resident
    data) does not cover.  value_grad(x) -> (value, gradient).  Returns (x, value, iterations)."""
    x = np.array(x0, dtype=np.float64)
    f, g = value_grad(x.copy())
    S, Y = [], []
    it = 0
    for it in range(1, iters + 1):
        if np.all(np.abs(g) < threshold):
            it -= 1
            break
        q = g.copy()
        al = []
        for s_, y_ in zip(reversed(S), reversed(Y)):
            a = float(s_ @ q) / float(y_ @ s_)
            al.append(a)
            q -= a * y_
        if S:
            q *= float(S[-1] @ Y[-1]) / float(Y[-1] @ Y[-1])
        for (s_, y_), a in zip(zip(S, Y), reversed(al)):
            q += (a - float(y_ @ q) / float(y_ @ s_)) * s_
        d = q if S else g / max(1.0, float(np.linalg.norm(g)))
        slope = float(g @ d)
        if not slope > 0.0:                      # not an ascent direction: restart from the gradient
            S, Y = [], []
            d = g / max(1.0, float(np.linalg.norm(g)))
            slope = float(g @ d)
        t, ok = 1.0, False
        for _ in range(30):
            try:
                fn, gn = value_grad(x + t * d)
            except Exception:                    # covariance not positive definite at the trial point: a rejected step
                fn, gn = -np.inf, None
            if np.isfinite(fn) and fn >= f + 1e-4 * t * slope:
                ok = True
                break
            t *= 0.5
        if not ok:
            break
        s_, y_ = t * d, g - gn                   # ascent: curvature pair of -f
        if float(s_ @ y_) > 1e-12 * float(np.linalg.norm(s_) * np.linalg.norm(y_)):
            S.append(s_)
            Y.append(y_)
            if len(S) > history:
                S.pop(0)
                Y.pop(0)
        x, f, g = x + t * d, fn, gn
    return x, f, it


def parse_events(spec):
    """tutorial/events/main.go:17-63: the -events flag, "from:to:discount,..." (e.g. "1.:2.5:0.3,3:6:0.5") -> rows for
    kernel.Events / gogp_set_events.  A malformed number raises, as the reference panics."""
    if not spec:
        return []
    return [tuple(float(v) for v in event.split(":")) for event in spec.split(",")]


def load(rdr):
    """tutorial.go:234-272: every record is D inputs followed by one output."""
    if isinstance(rdr, (str, bytes)):
        rdr = io.StringIO(rdr.decode() if isinstance(rdr, bytes) else rdr)
    X, Y = [], []
    for record in csv.reader(rdr):
        if not record:
            continue
        vals = [float(v) for v in record]          # a data error raises, as the reference returns it
        X.append(vals[:-1])
        Y.append(vals[-1])
    return np.array(X, dtype=np.float64).reshape(len(Y), -1), np.array(Y, dtype=np.float64)


def mean_std(y):
    """gonum stat.MeanStdDev: the unbiased (n - 1) standard deviation (tutorial.go:83)."""
    y = np.asarray(y, dtype=np.float64)
    return float(y.mean()), float(y.std(ddof=1))


def Evaluate(gp, m, theta, rdr, wtr, alg=None, iters=None, threshold=None, rate=None, minopt=None, normalize=True,
             jitter=0.1, optimise=True, out_of_sample=False, rng=None, optinp=False):
    """gp: gogp_b200.GP; m: the optimisation model (gp itself or gp.Model with priors -- anything with
    Observe(x) / Gradient()); theta: initial log hyper-parameters; rdr / wtr: CSV in, forecasts out.
    optinp: the reference's OPTINP (tutorial.go:97-109; anynoise and warpedtime set it): the window's inputs and outputs
    are appended to the argument of Observe and optimised with the hyper-parameters -- the loop is then driven from
    the host over m.Observe / m.Gradient (lbfgs_ascent or the Adam loop).
    Returns the list of rows written (floats), for tests."""
    alg = ALG if alg is None else alg
    iters = ITERS if iters is None else iters
    threshold = THRESHOLD if threshold is None else threshold
    rate = RATE if rate is None else rate
    minopt = MINOPT if minopt is None else minopt
    rng = np.random.default_rng() if rng is None else rng
    X, Y = load(rdr)
    if normalize:
        meany, stdy = mean_std(Y)
        Y = (Y - meany) / stdy
    else:
        meany, stdy = 0.0, 1.0
    theta = np.asarray(theta, dtype=np.float64)
    priors = getattr(m, "Priors", None)
    rows = []
    shared = jitter == 0.0 and not optimise      # every window at the same point: the factor can grow in place
    for end in range(len(X)):
        x = theta.copy()
        x += jitter * rng.standard_normal(len(x)) if jitter else 0.0
        if optinp:
            P = len(theta)
            x = np.concatenate([x, X[:end].reshape(-1), Y[:end]])     # tutorial.go:97-109
            lml0 = m.Observe(x.copy())
            if optimise and end > minopt:
                def value_grad(p):
                    v = m.Observe(p.copy())
                    return v, np.asarray(m.Gradient(), dtype=np.float64)
                if alg == "lbfgs":
                    x, _, _ = lbfgs_ascent(value_grad, x, iters, threshold)
                else:                                                  # infer.Adam loop, tutorial.go:156-168
                    mom, vel = np.zeros_like(x), np.zeros_like(x)
                    for t in range(1, iters + 1):
                        _, g = value_grad(x)
                        if np.all(np.abs(g) < threshold):
                            break
                        mom = 0.9 * mom + 0.1 * g
                        vel = 0.999 * vel + 0.001 * g * g
                        x = x + rate * (mom / (1 - 0.9 ** t)) / (np.sqrt(vel / (1 - 0.999 ** t)) + 1e-8)
            lml = m.Observe(x.copy())                                 # leaves gp.X, gp.Y at the inferred values
            x = x[:P]
        elif shared and end > 0:
            err = gp.Extend(X[end - 1:end], Y[end - 1:end])     # O(N^2): gogp_extend
            if err is not None:
                raise RuntimeError(str(err))
            lml0 = lml = gp.LML()
        else:
            gp.X, gp.Y = X[:end], Y[:end]
            if shared:                                           # the first window: Absorb at exp(theta)
                gp.ThetaSimil = list(np.exp(x[:gp.Simil.NTheta()]))
                gp.ThetaNoise = list(np.exp(x[gp.Simil.NTheta():]))
                err = gp.Absorb(X[:end], Y[:end])
                if err is not None:
                    raise RuntimeError(str(err))
                lml0 = lml = gp.LML()
            else:
                lml0 = m.Observe(x)                              # initial log likelihood (tutorial.go:124)
                if optimise and end > minopt:
                    # the MLE loop runs inside the library over the resident window (gogp_optimize)
                    gp.Optimize(x, alg=alg, iters=iters, threshold=threshold, rate=rate, priors=priors)
                lml = m.Observe(x)                               # final log likelihood (tutorial.go:173)
        mu, sigma, err = gp.Produce(X[end:end + 1])              # one step out of sample (tutorial.go:178-182)
        if err is not None:
            raise RuntimeError(str(err))
        row = list(X[end]) + [Y[end] * stdy + meany, mu[0] * stdy + meany, sigma[0] * stdy, lml0, lml] + \
            [math.exp(v) for v in x]
        rows.append(row)
        wtr.write(",".join("%f" % v for v in row) + "\n")       # tutorial.go:185-197
    if out_of_sample and len(X) > 1:                             # tutorial.go:200-224
        Z = (X + X[-1])[1:]
        mu, sigma, err = gp.Produce(Z)
        if err is not None:
            raise RuntimeError(str(err))
        for i in range(len(Z)):
            wtr.write(",".join("%f" % v for v in Z[i]) + ",nan,%f,%f\n" % (mu[i] * stdy + meany, sigma[i] * stdy))
    return rows

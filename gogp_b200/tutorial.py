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
             jitter=0.1, optimise=True, out_of_sample=False, rng=None):
    """gp: gogp_b200.GP; m: the optimisation model (gp itself or gp.Model with priors -- anything with
    Observe(x) / Gradient()); theta: initial log hyper-parameters; rdr / wtr: CSV in, forecasts out.
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
        if shared and end > 0:
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

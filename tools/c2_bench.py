"""BASELINE configs[1]: synthetic 1-D RBF + noise, N = 4096, LML + gradient and Produce at 1024
test points on one GPU.  Prints one JSON line: device phase times (CUDA events inside the library)
and host wall time per call, averaged over fresh parameter points."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from gogp_b200 import GP, kernel as k

N, M, REPS = 4096, 1024, 20
rng = np.random.default_rng(0)
X = rng.uniform(0.0, N / 50.0, size=(N, 1))
y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(N)
y = (y - y.mean()) / y.std(ddof=1)
Z = rng.uniform(0.0, N / 50.0, size=(M, 1))
truth = np.array([0.0, 0.0, np.log(0.1)])
g = GP(NDim=1, Simil=k.Param(0) * k.Normal.Of(l=1), Noise=k.UniformNoise)
g.X, g.Y = X, y
for r in range(3):
    g.Observe(truth + 0.05 * rng.standard_normal(3)); g.Gradient(); g.Produce(Z)
ph = {}
t_eval = t_prod = 0.0
for r in range(REPS):
    th = truth + 0.1 * rng.standard_normal(3)
    t0 = time.perf_counter()
    lml = g.Observe(th)
    gr = g.Gradient()
    t1 = time.perf_counter()
    mu, sigma, err = g.Produce(Z)
    t2 = time.perf_counter()
    t_eval += t1 - t0
    t_prod += t2 - t1
    for a, b in g.PhaseTimes().items():
        ph[a] = ph.get(a, 0.0) + b
ph = {a: round(b / REPS, 4) for a, b in ph.items()}
dev_eval = sum(ph[a] for a in ("upload", "build", "potrf", "solve", "potri", "trace"))
print(json.dumps({"workload": "configs[1]: 1-D RBF + noise, N=4096, M=1024", "phases_ms": ph,
                  "eval_device_ms": dev_eval, "eval_wall_ms": 1e3 * t_eval / REPS,
                  "evals_per_s_wall": REPS / t_eval, "produce_device_ms": ph["predict"],
                  "produce_wall_ms": 1e3 * t_prod / REPS, "lml": lml}))

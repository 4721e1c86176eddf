"""One LML+gradient evaluation at N with the configs[1] kernel (1-D scaled Normal + noise): the
cheapest element function, so the build and trace kernels show their HBM-bound side (for ncu)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from gogp_b200 import GP, kernel as k

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
rng = np.random.default_rng(0)
X = rng.uniform(0.0, N / 50.0, size=(N, 1))
y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(N)
y = (y - y.mean()) / y.std(ddof=1)
g = GP(NDim=1, Simil=k.Param(0) * k.Normal.Of(l=1), Noise=k.UniformNoise)
g.X, g.Y = X, y
lml = g.Observe(np.array([0.0, 0.0, np.log(0.1)]))
gr = g.Gradient()
print(lml, gr, g.PhaseTimes(), flush=True)

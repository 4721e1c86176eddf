"""Robustness probe at N = 65536 (LML + gradient, 2 x 34 GB buffers) on one GPU."""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
from gogp_b200 import GP, kernel as k
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
e = k.Param(0)
for d in range(8):
    e = e * k.Normal.Of(l=1 + d, dim=d)
e = e * k.Periodic.Of(l=9, p=10, dim=0)
X, y, truth = bench.synth(N, 0)
g = GP(NDim=8, Simil=e, Noise=k.UniformNoise)
g.X, g.Y = X, y
t0 = bench.theta_for(truth, 0, 0)
lml = g.Observe(t0.copy()); gr = g.Gradient()
ph = g.PhaseTimes()
v = np.random.default_rng(1).standard_normal(len(t0)); v /= np.linalg.norm(v)
fd = (g.Observe(t0 + 1e-4 * v) - g.Observe(t0 - 1e-4 * v)) / 2e-4
print("N", N, "lml", lml, "phases", {a: round(b, 1) for a, b in ph.items()})
print("potrf TF", N**3 / 3 / ph["potrf"] / 1e9, "potri TF", 2 * N**3 / 3 / ph["potri"] / 1e9, "evals/s", 1e3 / sum(ph.values()))
print("directional derivative", gr @ v, "central difference", fd, "rel", abs(fd - gr @ v) / max(1, abs(fd)))

"""One warm-up + `reps` LML+gradient evaluations of the C3 workload at N (for ncu launch lists)."""
import sys


sys.path.insert(0, ".")
import bench
from gogp_b200 import GP, kernel as k

N = int(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
e = k.Param(0)
for d in range(8):
    e = e * k.Normal.Of(l=1 + d, dim=d)
e = e * k.Periodic.Of(l=9, p=10, dim=0)
X, y, truth = bench.synth(N, 0)
g = GP(NDim=8, Simil=e, Noise=k.UniformNoise)
g.X, g.Y = X, y
for r in range(reps + 1):
    lml = g.Observe(bench.theta_for(truth, 0, r))
    gr = g.Gradient()
    print(r, lml, g.PhaseTimes(), flush=True)

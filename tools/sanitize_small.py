"""A small pass over every kernel family for compute-sanitizer (memcheck / racecheck): Observe + Gradient + Produce
through the C-ABI for a few kernel expressions and sizes (tile kernel, cp.async and TMA GEMMs, the single-launch
triangular solves, build / trace / input-gradient kernels), the stored-state and window-extension paths and the
block-cyclic grid on one rank.  GOGP_TMA_MIN_K=128 routes the K >= 128 products through the TMA kernel."""
import sys

import numpy as np

sys.path.insert(0, ".")
from tests import cases  # noqa: E402

for name, N in (("hyperpriors", 300), ("c5_matern4", 700), ("c3_ard8", 520)):
    X, y, logt = cases.synth(name, N, seed=2)
    g = cases.make_device_gp(name)
    g.X, g.Y = X, y
    lml = g.Observe(logt.copy())
    grad = g.Gradient()
    mu, sigma, err = g.Produce(X[:40] + 0.01)
    assert err is None and np.isfinite(lml) and np.all(np.isfinite(grad))
    st = g.State()
    g2 = cases.make_device_gp(name)
    assert g2.Restore(st) is None
    mu2, _, _ = g2.Produce(X[:40] + 0.01)
    assert np.max(np.abs(mu2 - mu)) < 1e-9
    g2.close()
    th = np.exp(logt)
    nts = g.Simil.NTheta()
    g.ThetaSimil, g.ThetaNoise = list(th[:nts]), list(th[nts:])
    assert g.Absorb(X[:N - 140], y[:N - 140]) is None
    assert g.Extend(X[N - 140:], y[N - 140:]) is None
    g.close()
    print(name, N, "ok", lml, flush=True)

# with_obs: the input gradient kernel
name, N = "warpedtime", 150
X, y, logt = cases.synth(name, N, seed=3)
g = cases.make_device_gp(name)
g.Observe(np.concatenate([logt, X.reshape(-1), y]))
gr = g.Gradient()
assert np.all(np.isfinite(gr))
g.close()
print("with_obs ok", flush=True)

# the block-cyclic path on one rank (masked GEMMs, fused sweep, local trace)
from gogp_b200 import GridGP  # noqa: E402
name, N = "c5_matern4", 700
ndim, ds, dn, _, _ = cases.CASES[name]
X, y, logt = cases.synth(name, N, seed=4)
gg = GridGP(NDim=ndim, Simil=ds, Noise=dn, Devices=[0], Block=256)
gg.X, gg.Y = X, y
l = gg.Observe(logt.copy())
gr = gg.Gradient()
gg.close()
assert np.isfinite(l) and np.all(np.isfinite(gr))
print("grid ok", l, flush=True)

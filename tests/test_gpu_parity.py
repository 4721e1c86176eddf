"""Parity tests proper: the CUDA path, called through the C-ABI (via the ctypes
host mirror gogp_b200.gp), against the reference's golden vectors and against the
CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): relative 1e-9 on the LML (relative to
max(|LML|, N)), 1e-7 on gradients and predictive moments
(max-norm error over max(norm, 1))."""
import numpy as np
import pytest

from tests import cases
from tests.golden_ref import DX, ELEMENTAL, EPS, PRODUCE, arr

pytestmark = pytest.mark.gpu

LML_TOL, GRAD_TOL, PRED_TOL = 1e-9, 1e-7, 1e-7


def _dev_noise(n):
    from gogp_b200 import kernel as k
    return k.UniformNoise if n == "uniform" else k.ConstantNoise(n)


def _grad_err(g, ref):
    return float(np.max(np.abs(g - ref)) / max(1.0, np.max(np.abs(ref)))) if len(ref) else 0.0


# ---- the reference's own tables (gp/gp_test.go) through the C-ABI ---------------------
@pytest.mark.parametrize("case", PRODUCE, ids=[c[0] for c in PRODUCE])
def test_produce_goldens(case):
    from gogp_b200 import GP, kernel as k
    name, nstd, theta, x, y, z, mu, sigma = case
    for parallel in (False, True):  # gp/gp_test.go:123-132
        g = GP(NDim=1, Simil=k.Normal, Noise=k.ConstantNoise(nstd), ThetaSimil=theta)
        g.Parallel = parallel
        err = g.Absorb(x, y)
        assert err is None, err
        m, s, err = g.Produce(z)
        assert err is None, err
        assert len(m) == len(mu) and len(s) == len(sigma)
        assert np.all(np.abs(m - arr(mu)) <= 1e-6), (m, mu)
        assert np.all(np.abs(s - arr(sigma)) <= 1e-6), (s, sigma)


@pytest.mark.parametrize("case", ELEMENTAL, ids=[c[0] for c in ELEMENTAL])
def test_elemental_goldens(case):
    from gogp_b200 import GP, kernel as k
    name, n, x, ll = case
    x = arr(x)
    g = GP(NDim=1, Simil=k.Normal, Noise=_dev_noise(n))
    v = g.Observe(x)
    dll = g.Gradient()
    assert abs(v - ll) < 1e-6
    assert len(dll) == len(x)
    for j in range(len(x)):  # forward difference, gp/gp_test.go:242-252
        x0 = x[j]
        x[j] += DX
        vj = g.Observe(x)
        x[j] = x0
        assert abs(dll[j] - (vj - v) / DX) <= EPS, (j, dll[j], (vj - v) / DX)
    P = g.Simil.NTheta() + g.Noise.NTheta()
    v2 = g.Observe(x[:P].copy())  # hyper-parameters only, gp/gp_test.go:254-267
    assert abs(v2 - ll) < 1e-6
    assert len(g.Gradient()) == P


# ---- covariance build / descriptor interpreter, element by element ---------------------
@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_cov_build_matches_oracle(name):
    import ctypes as C
    from gogp_b200 import _lib
    N = 200
    X, y, logt = cases.synth(name, N, seed=5)
    dg = cases.make_device_gp(name)
    og = cases.make_oracle_gp(name)
    og.X, og.Y = X, y
    og.observe(logt.copy())
    nts = dg.Simil.NTheta()
    th = np.exp(logt)
    out = np.zeros((N, N))
    Xf = np.ascontiguousarray(X).ravel()
    ts, tn = np.ascontiguousarray(th[:nts]), np.ascontiguousarray(th[nts:])
    st = _lib.lib().gogp_debug_build(dg._handle(), _lib.dptr(ts), _lib.dptr(tn), _lib.dptr(Xf), N, _lib.dptr(out))
    assert st == _lib.OK
    assert np.max(np.abs(out - og.K) / np.maximum(np.abs(og.K), 1e-300)) < 1e-12  # exp argument to a few ulp
    assert np.max(np.abs(out - og.K)) < 1e-14 * max(1.0, np.abs(og.K).max())


# ---- LML + gradient + Produce against the oracle -----------------------------------------
@pytest.mark.parametrize("name", sorted(cases.CASES))
@pytest.mark.parametrize("N", [1, 2, 37, 128, 129, 300])
def test_hyper_only_parity(name, N):
    X, y, logt = cases.synth(name, N, seed=N)
    dg = cases.make_device_gp(name)
    og = cases.make_oracle_gp(name)
    dg.X, dg.Y = X, y
    og.X, og.Y = X, y
    x = logt.copy()
    lml = dg.Observe(x)
    assert np.allclose(x, logt, rtol=0, atol=1e-15)  # exp/log round trip of the caller's slice
    ref = og.observe(logt.copy())
    assert abs(lml - ref) <= LML_TOL * max(abs(ref), N), (lml, ref)
    g = dg.Gradient()
    gref = og.gradient()
    assert len(g) == len(gref) == len(logt)
    assert _grad_err(g, gref) <= GRAD_TOL, (g, gref)
    # Produce at in-sample, interpolated and far-away points
    rng = np.random.default_rng(1000 + N)
    Z = np.concatenate([X[:min(N, 5)], rng.uniform(X.min() - 1, X.max() + 1, size=(70, X.shape[1]))])
    mu, sigma, err = dg.Produce(Z)
    assert err is None
    mref, sref = og.produce(Z, clamp=True)
    assert np.max(np.abs(mu - mref)) <= PRED_TOL * max(1.0, np.abs(mref).max())
    # sigma = sqrt(small difference): compare variances, where the rounding lives
    assert np.max(np.abs(sigma ** 2 - sref ** 2)) <= PRED_TOL * max(1.0, np.abs(sref).max() ** 2)


@pytest.mark.parametrize("name", ["normal_uniform", "anynoise", "warpedtime", "hyperpriors", "events", "c3_ard3", "c5_matern4"])
@pytest.mark.parametrize("N", [1, 3, 43, 150])
def test_with_obs_parity(name, N):
    """Observe's [theta | X | Y] layout: gradient w.r.t. inputs and outputs too
    (tutorial anynoise optimises Y, warpedtime optimises X)."""
    X, y, logt = cases.synth(name, N, seed=7 * N)
    x = np.concatenate([logt, X.ravel(), y])
    dg = cases.make_device_gp(name)
    og = cases.make_oracle_gp(name)
    lml = dg.Observe(x.copy())
    ref = og.observe(x.copy())
    assert abs(lml - ref) <= LML_TOL * max(abs(ref), N)
    g = dg.Gradient()
    gref = og.gradient()
    assert len(g) == len(x)
    assert _grad_err(g, gref) <= GRAD_TOL, np.max(np.abs(g - gref))


def test_barebones_kat():
    """Config 1 (tutorial/barebones on tutorial/data), SURVEY.md section 8(c)."""
    from oracle.gp import mean_std
    d = np.loadtxt(cases.GOLDEN + "/barebones.csv", delimiter=",")
    m, s = mean_std(d[:, 1])
    g = cases.make_device_gp("barebones")
    g.X, g.Y = d[:, :1].copy(), (d[:, 1] - m) / s
    assert abs(g.Observe(np.zeros(3)) - (-8.482987052156)) < 1e-9
    assert np.allclose(g.Gradient(), [-3.2320901320, 6.9054771170, -0.4752473692], atol=1e-8)
    assert abs(g.Observe(arr([-0.5, 0.3, 1.0])) - (-9.103987864150)) < 1e-9
    assert np.allclose(g.Gradient(), [-0.1009289296, 2.3769085404, -7.5447626235], atol=1e-8)


@pytest.mark.parametrize("name", ["barebones", "hyperpriors", "anynoise", "warpedtime", "events"])
def test_tutorial_expanding_window(name):
    """tutorial.Evaluate's loop shape (tutorial/tutorial.go:91-179): N = 0..len-1 on the
    shipped data, one Observe+Gradient and a one-step forecast per prefix."""
    from oracle.gp import mean_std
    d = np.loadtxt(cases.GOLDEN + "/%s.csv" % name, delimiter=",")
    m, s = mean_std(d[:, 1])
    X, Y = d[:, :1].copy(), (d[:, 1] - m) / s
    dg = cases.make_device_gp(name)
    og = cases.make_oracle_gp(name)
    P = dg.Simil.NTheta() + dg.Noise.NTheta()
    logt = np.zeros(P)
    for end in range(0, len(X), 3):
        dg.X, dg.Y = X[:end], Y[:end]
        og.X, og.Y = X[:end], Y[:end]
        lml = dg.Observe(logt.copy())
        ref = og.observe(logt.copy())
        assert abs(lml - ref) <= LML_TOL * max(abs(ref), end, 1)
        assert _grad_err(dg.Gradient(), og.gradient()) <= GRAD_TOL
        mu, sigma, err = dg.Produce(X[end:end + 1])
        mref, sref = og.produce(X[end:end + 1], clamp=True)
        assert err is None
        assert abs(mu[0] - mref[0]) <= PRED_TOL * max(1.0, abs(mref[0]))
        assert abs(sigma[0] ** 2 - sref[0] ** 2) <= PRED_TOL


# ---- error behaviour -------------------------------------------------------------------------
def test_not_positive_definite_panics_in_observe_and_errors_in_absorb():
    from gogp_b200 import GP, GoGPPanic, _lib, kernel as k
    g = GP(NDim=1, Simil=k.Normal, Noise=k.ConstantNoise(0.0), ThetaSimil=[1.0])
    err = g.Absorb([[0.0], [0.0]], [1.0, 2.0])  # duplicate input, zero noise: singular K
    assert err is not None and err.status == _lib.NOT_POSITIVE_DEFINITE
    with pytest.raises(GoGPPanic):
        g.Observe(arr([0.0, 0.0, 0.0, 1.0, 2.0]))


def test_bad_length_panics():
    from gogp_b200 import GoGPPanic
    g = cases.make_device_gp("c5_matern4")
    with pytest.raises(GoGPPanic):
        g.Observe(np.zeros(5 + 1 + 7))  # 7 is not a multiple of NDim+1 = 5


def test_default_noise_is_1e_5():
    """Noise == nil -> ConstantNoise(1e-5), gp/gp.go:43-48."""
    from gogp_b200 import GP, kernel as k
    from oracle import kernels as ok
    from oracle.gp import GP as OGP
    X, y, _ = cases.synth("normal_const", 12, seed=2, spread=30.0)
    g = GP(NDim=1, Simil=k.Normal)
    o = OGP(1, ok.Normal)
    g.X, g.Y = X, y
    o.X, o.Y = X, y
    ref = o.observe(np.zeros(1))
    assert abs(g.Observe(np.zeros(1)) - ref) <= 1e-7 * max(abs(ref), 12)  # cond(K) is large here


# ---- mid-size differential + size-independent properties --------------------------------------
@pytest.mark.parametrize("name,N", [("c2_rbf", 1000), ("c3_ard8", 1500), ("c5_matern4", 2048), ("hyperpriors", 1100)])
def test_midsize_parity(name, N):
    X, y, logt = cases.synth(name, N, seed=N)
    dg = cases.make_device_gp(name)
    og = cases.make_oracle_gp(name)
    dg.X, dg.Y = X, y
    og.X, og.Y = X, y
    lml = dg.Observe(logt.copy())
    ref = og.observe(logt.copy())
    assert abs(lml - ref) <= LML_TOL * max(abs(ref), N), (lml, ref)
    g, gref = dg.Gradient(), og.gradient()
    assert _grad_err(g, gref) <= GRAD_TOL, (g, gref)
    Z = np.random.default_rng(1).uniform(X.min(), X.max(), size=(300, X.shape[1]))
    mu, sigma, err = dg.Produce(Z)
    mref, sref = og.produce(Z, clamp=True)
    assert err is None
    assert np.max(np.abs(mu - mref)) <= PRED_TOL * max(1.0, np.abs(mref).max())
    assert np.max(np.abs(sigma ** 2 - sref ** 2)) <= PRED_TOL * max(1.0, np.abs(sref).max() ** 2)


def test_factor_and_inverse_residuals_4096():
    """C2 size (N = 4096): residuals ||L L^T - K|| / ||K||, ||K K^-1 - I||, ||K alpha - y|| / ||y||
    computed on the host from the fetched device state."""
    import ctypes as C
    from gogp_b200 import _lib
    N = 4096
    X, y, logt = cases.synth("c2_rbf", N, seed=0)
    dg = cases.make_device_gp("c2_rbf")
    og = cases.make_oracle_gp("c2_rbf")
    dg.X, dg.Y = X, y
    og.X, og.Y = X, y
    lml = dg.Observe(logt.copy())
    ref = og.observe(logt.copy())
    assert abs(lml - ref) <= LML_TOL * max(abs(ref), N)
    Lm = np.zeros((N, N))
    assert _lib.lib().gogp_get_factor(dg._handle(), _lib.dptr(Lm), N) == _lib.OK
    K = og.K
    assert np.linalg.norm(Lm @ Lm.T - K) / np.linalg.norm(K) < 1e-13
    alpha = dg.Alpha()
    assert np.linalg.norm(K @ alpha - y) / np.linalg.norm(y) < 1e-9
    g, gref = dg.Gradient(), og.gradient()
    assert _grad_err(g, gref) <= GRAD_TOL
    Kinv = np.zeros((N, N))
    assert _lib.lib().gogp_debug_fetch(dg._handle(), 2, _lib.dptr(Kinv), N) == _lib.OK
    assert np.linalg.norm(K @ Kinv - np.eye(N)) / np.sqrt(N) < 1e-9
    Z = np.random.default_rng(3).uniform(X.min(), X.max(), size=(1024, 1))
    mu, sigma, err = dg.Produce(Z)
    mref, sref = og.produce(Z, clamp=True)
    assert np.max(np.abs(mu - mref)) <= PRED_TOL * max(1.0, np.abs(mref).max())
    assert np.max(np.abs(sigma ** 2 - sref ** 2)) <= PRED_TOL


@pytest.mark.parametrize("nb", ["256", "512,0,128", "0,0,128", "640,256,128", "384,128"])
def test_lookahead_factorisation_matches_recursion_and_oracle(nb, monkeypatch):
    """Blocked::potrf_la (nested block columns of widths GOGP_LA_NB, next column on the priority
    stream, bulk updates on lower-priority streams) against the one-stream recursion and the oracle;
    repeated runs must be bit-identical (a missing cross-stream dependency shows up as a changing
    result)."""
    import ctypes as C
    from gogp_b200 import _lib
    N = 2000  # Npad = 2048 >= 3 nb
    X, y, logt = cases.synth("c2_rbf", N, seed=4)
    og = cases.make_oracle_gp("c2_rbf")
    og.X, og.Y = X, y
    ref = og.observe(logt.copy())
    gref = og.gradient()
    out = {}
    for mode in ("0", nb, nb, nb, nb):
        monkeypatch.setenv("GOGP_LA_NB", mode)
        dg = cases.make_device_gp("c2_rbf")
        dg.X, dg.Y = X, y
        lml = dg.Observe(logt.copy())
        g = dg.Gradient()
        Lm = np.zeros((N, N))
        assert _lib.lib().gogp_get_factor(dg._handle(), _lib.dptr(Lm), N) == _lib.OK
        assert abs(lml - ref) <= LML_TOL * max(abs(ref), N)
        assert _grad_err(g, gref) <= GRAD_TOL
        assert np.linalg.norm(Lm @ Lm.T - og.K) / np.linalg.norm(og.K) < 1e-13
        if mode in out:
            assert lml == out[mode][0] and np.array_equal(Lm, out[mode][1])
        out[mode] = (lml, Lm)
    assert np.abs(out["0"][1] - out[nb][1]).max() < 1e-11


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_tile_kernel_against_numpy(variant):
    """The 128 x 128 tile kernel alone (0: blocked shipped version, 1: the first version, 2: blocked with
    the shared-memory column broadcast): factor, inverse, and the index of the first bad pivot."""
    import ctypes as C
    from gogp_b200 import _lib
    dg = cases.make_device_gp("c2_rbf")
    h = dg._handle()
    L = _lib.lib()
    rng = np.random.default_rng(11)
    for trial in range(4):
        G = rng.standard_normal((128, 128))
        K = G @ G.T / 128 + np.eye(128) * (10.0 ** -trial)
        A = np.tril(K) + np.triu(rng.standard_normal((128, 128)), 1)  # the upper triangle must be ignored
        Lo, Wo, info = np.zeros((128, 128)), np.zeros((128, 128)), C.c_int(-1)
        assert L.gogp_debug_leaf_run(h, variant, _lib.dptr(A), _lib.dptr(Lo), _lib.dptr(Wo), C.byref(info)) == _lib.OK
        ref = np.linalg.cholesky(K)
        assert info.value == 0
        assert np.abs(Lo - ref).max() <= 1e-13 * np.abs(ref).max() * np.linalg.cond(ref)
        assert np.abs(np.triu(Lo, 1)).max() == 0.0 and np.abs(np.triu(Wo, 1)).max() == 0.0
        assert np.abs(Wo @ ref - np.eye(128)).max() <= 1e-13 * np.linalg.cond(ref) ** 2
    for bad in (0, 31, 32, 77, 127):
        K = np.eye(128) * 2.0
        K[bad, bad] = -1.0
        Lo, Wo, info = np.zeros((128, 128)), np.zeros((128, 128)), C.c_int(-1)
        assert L.gogp_debug_leaf_run(h, variant, _lib.dptr(K), _lib.dptr(Lo), _lib.dptr(Wo), C.byref(info)) == _lib.OK
        assert info.value == bad + 1


def test_repeatable_and_handle_reuse():
    """Same inputs -> bit-identical results (fixed reduction orders); a handle survives
    growing and shrinking N."""
    dg = cases.make_device_gp("c3_ard3")
    res = []
    for N in (300, 50, 700, 300):
        X, y, logt = cases.synth("c3_ard3", N, seed=1)
        dg.X, dg.Y = X, y
        res.append((dg.Observe(logt.copy()), dg.Gradient()))
    assert res[0][0] == res[3][0]
    assert np.array_equal(res[0][1], res[3][1])


def test_evaluation_memo_for_funcgrad_pattern():
    """infer.FuncGrad evaluates Observe(x) for the value and again for the gradient
    (SURVEY.md section 3.4): the repeat at the same point costs no kernel launch."""
    name, N = "hyperpriors", 400
    X, y, logt = cases.synth(name, N, seed=8)
    dg = cases.make_device_gp(name)
    og = cases.make_oracle_gp(name)
    dg.X, dg.Y = X, y
    og.X, og.Y = X, y
    v1 = dg.Observe(logt.copy())
    n1 = dg.Launches()
    v2 = dg.Observe(logt.copy())          # Func(x) then Grad(x): same x
    assert v2 == v1 and dg.Launches() == n1
    g = dg.Gradient()
    ref = og.observe(logt.copy())
    assert abs(v1 - ref) <= LML_TOL * max(abs(ref), N)
    assert _grad_err(g, og.gradient()) <= GRAD_TOL
    y2 = y.copy()
    y2[0] += 1e-3                          # same theta, different data: must recompute
    dg.Y = y2
    v3 = dg.Observe(logt.copy())
    assert dg.Launches() > n1 and v3 != v1
    og.Y = y2
    ref3 = og.observe(logt.copy())
    assert abs(v3 - ref3) <= LML_TOL * max(abs(ref3), N)


def test_full_size_properties_n32768():
    """BASELINE size (N = 32768, the bench workload): the oracle cannot run here in test time, so
    size-independent properties instead -- (i) the gradient agrees with a central difference of the
    LML along a random direction, (ii) two different factorisation paths of the library (recursive
    single-GPU, block-cyclic with 2048-blocks) give the same LML, (iii) repeat evaluation is
    bit-identical, (iv) predictions at training inputs reproduce y within the noise level."""
    import bench
    from gogp_b200 import GP, GridGP, kernel as k
    N, D = 32768, 8
    e = k.Param(0)
    for d in range(D):
        e = e * k.Normal.Of(l=1 + d, dim=d)
    e = e * k.Periodic.Of(l=1 + D, p=2 + D, dim=0)
    X, y, truth = bench.synth(N, 0)
    g = GP(NDim=D, Simil=e, Noise=k.UniformNoise)
    g.X, g.Y = X, y
    t0 = bench.theta_for(truth, 0, 0)
    lml = g.Observe(t0.copy())
    grad = g.Gradient()
    assert np.all(np.isfinite(grad)) and np.isfinite(lml)
    # (iv) before theta changes: posterior mean at 64 training points
    mu, sigma, err = g.Produce(X[:64])
    assert err is None
    assert np.max(np.abs(mu - y[:64])) < 0.5 and np.all(sigma < 0.2)
    # (i) directional derivative, central difference with h = 1e-4 (truncation ~1e-8 relative)
    rng = np.random.default_rng(5)
    v = rng.standard_normal(len(t0))
    v /= np.linalg.norm(v)
    h = 1e-4
    fd = (g.Observe(t0 + h * v) - g.Observe(t0 - h * v)) / (2 * h)
    assert abs(fd - grad @ v) <= 1e-5 * max(1.0, abs(grad @ v)), (fd, grad @ v)
    # (iii) bit-repeatability
    lml2 = g.Observe(t0.copy() + 0.0)
    g2 = g.Gradient()
    assert lml2 == lml and np.array_equal(g2, grad)
    g.close()
    # (ii) the block-cyclic path (gogp_grid_*, one rank, 2048-blocks: right-looking factorisation and the fused
    # V / K^-1 sweep instead of the recursion) on the same inputs
    gg = GridGP(NDim=D, Simil=e, Noise=k.UniformNoise, Devices=[0], Block=2048)
    gg.X, gg.Y = X, y
    lml_bc = gg.Observe(t0.copy())
    grad_bc = gg.Gradient()
    gg.close()
    assert abs(lml_bc - lml) <= 1e-11 * max(abs(lml), N), (lml_bc, lml)
    assert np.max(np.abs(grad_bc - grad)) <= 1e-9 * max(1.0, np.max(np.abs(grad))), (grad_bc, grad)


def test_produce_chunks_many_test_points():
    """Produce walks the test points in chunks of 8192: M = 9000 crosses a chunk boundary."""
    name, N, M = "c2_rbf", 500, 9000
    X, y, logt = cases.synth(name, N, seed=12)
    dg = cases.make_device_gp(name)
    og = cases.make_oracle_gp(name)
    dg.X, dg.Y = X, y
    og.X, og.Y = X, y
    dg.Observe(logt.copy())
    og.observe(logt.copy())
    Z = np.random.default_rng(4).uniform(X.min() - 1, X.max() + 1, size=(M, 1))
    mu, sigma, err = dg.Produce(Z)
    assert err is None and len(mu) == M
    mref, sref = og.produce(Z, clamp=True)
    assert np.max(np.abs(mu - mref)) <= PRED_TOL * max(1.0, np.abs(mref).max())
    assert np.max(np.abs(sigma ** 2 - sref ** 2)) <= PRED_TOL * max(1.0, np.abs(sref).max() ** 2)


def test_produce_on_stored_results():
    """gp/gp.go:255-257: "Produce works on stored results" -- L and Alpha are exported so that a user can keep them
    and restore them before Produce.  A second GP restored from State() predicts what the first one does."""
    from gogp_b200 import GoGPPanic, _lib
    name, N = "hyperpriors", 300
    X, y, logt = cases.synth(name, N, seed=21)
    dg = cases.make_device_gp(name)
    dg.X, dg.Y = X, y
    dg.Observe(logt.copy())
    Z = np.linspace(X.min() - 1.0, X.max() + 1.0, 200).reshape(-1, 1)
    mu, sigma, err = dg.Produce(Z)
    assert err is None
    state = dg.State()
    assert state["L"].shape == (N, N) and np.allclose(np.triu(state["L"], 1), 0.0)
    dg.close()
    g2 = cases.make_device_gp(name)
    assert g2.Restore(state) is None
    mu2, sigma2, err = g2.Produce(Z)
    assert err is None
    assert np.max(np.abs(mu2 - mu)) <= 1e-10 * max(1.0, np.abs(mu).max())
    assert np.max(np.abs(sigma2 ** 2 - sigma ** 2)) <= 1e-10
    with pytest.raises(GoGPPanic) as e:      # no Y in the stored state
        g2.LML()
    assert e.value.status == _lib.NOT_READY
    with pytest.raises(GoGPPanic):
        g2.Gradient()
    bad = dict(state)
    bad["Alpha"] = state["Alpha"][:-1]
    assert g2.Restore(bad) is not None
    # and the restored handle goes on working as a GP
    g2.X, g2.Y = X, y
    og = cases.make_oracle_gp(name)
    og.X, og.Y = X, y
    assert abs(g2.Observe(logt.copy()) - og.observe(logt.copy())) <= LML_TOL * N
    g2.close()


def test_ill_conditioned_is_an_error_with_a_complete_state():
    """gonum's SolveVecTo returns a Condition error when cond(K) > 1e16 and the reference passes it on
    (gp/gp.go:233-236: Absorb returns it, Observe panics); the factor, alpha and LML exist all the same.
    K = 1 1^T + s I (N identical inputs): cond_2 = (N + s) / s; s = 1e-13 -> 4e16, s = 1e-9 -> 4e12."""
    from gogp_b200 import GP, GoGPPanic, _lib, kernel as k
    N = 4096
    X = np.zeros((N, 1))
    y = np.sin(np.arange(N) * 0.01)
    g = GP(NDim=1, Simil=k.Normal, Noise=k.ConstantNoise(np.sqrt(1e-13)), ThetaSimil=[1.0])
    err = g.Absorb(X, y)
    assert err is not None and err.status == _lib.ILL_CONDITIONED, err
    assert "condition number" in str(err)
    assert np.isfinite(g.LML())
    with pytest.raises(GoGPPanic) as e:
        g.Observe(np.array([0.0]))
    assert e.value.status == _lib.ILL_CONDITIONED
    g.close()
    g = GP(NDim=1, Simil=k.Normal, Noise=k.ConstantNoise(np.sqrt(1e-9)), ThetaSimil=[1.0])
    assert g.Absorb(X, y) is None          # suspicious diagonal (stage 1), cleared by the estimate (stage 2)
    g.close()


# ---- the window that grows (tutorial/tutorial.go:91-179) -------------------------
@pytest.mark.parametrize("name,N0,m", [("c5_matern4", 0, 5), ("c5_matern4", 1, 1), ("c5_matern4", 100, 29),
                                       ("c5_matern4", 127, 1), ("c5_matern4", 128, 1), ("hyperpriors", 300, 200),
                                       ("c3_ard3", 1000, 300)])
def test_extend_equals_absorb(name, N0, m):
    """gogp_extend: the factor of N0 observations extended by m at unchanged hyper-parameters is the factor of
    N0 + m -- LML, alpha and the predictions agree with a fresh Absorb of everything, and with the oracle."""
    X, y, logt = cases.synth(name, N0 + m, seed=17)
    theta = np.exp(logt)
    nts = cases.CASES[name][1].NTheta()
    g = cases.make_device_gp(name)
    g.ThetaSimil, g.ThetaNoise = list(theta[:nts]), list(theta[nts:])
    assert g.Absorb(X[:N0], y[:N0]) is None
    assert g.Extend(X[N0:], y[N0:]) is None
    ref = cases.make_device_gp(name)
    ref.ThetaSimil, ref.ThetaNoise = list(theta[:nts]), list(theta[nts:])
    assert ref.Absorb(X, y) is None
    N = N0 + m
    assert abs(g.LML() - ref.LML()) <= 1e-11 * max(abs(ref.LML()), N)
    a, b = g.Alpha(), ref.Alpha()
    assert len(a) == N and np.max(np.abs(a - b)) <= 1e-9 * max(1.0, np.max(np.abs(b)))
    Z = X[:7] + 0.01
    mu, sg, err = g.Produce(Z)
    mu2, sg2, err2 = ref.Produce(Z)
    assert err is None and err2 is None
    assert np.max(np.abs(mu - mu2)) <= 1e-9 * max(1.0, np.max(np.abs(mu2))) and np.max(np.abs(sg ** 2 - sg2 ** 2)) <= 1e-9
    og = cases.make_oracle_gp(name)
    og.ThetaSimil, og.ThetaNoise = theta[:nts].copy(), theta[nts:].copy()
    og.absorb(X, y)
    assert abs(g.LML() - og.lml()) <= LML_TOL * max(abs(og.lml()), N)
    # the extended handle goes on as a GP: Observe at a new point refactors everything
    og2 = cases.make_oracle_gp(name)
    og2.X, og2.Y = X, y
    assert abs(g.Observe(logt.copy() + 0.05) - og2.observe(logt.copy() + 0.05)) <= LML_TOL * N
    g.close()
    ref.close()


def test_extend_point_by_point_across_tile_boundaries():
    name = "barebones"
    X, y, logt = cases.synth(name, 270, seed=19)
    theta = np.exp(logt)
    nts = cases.CASES[name][1].NTheta()
    g = cases.make_device_gp(name)
    g.ThetaSimil, g.ThetaNoise = list(theta[:nts]), list(theta[nts:])
    assert g.Absorb(X[:120], y[:120]) is None
    og = cases.make_oracle_gp(name)
    og.ThetaSimil, og.ThetaNoise = theta[:nts].copy(), theta[nts:].copy()
    for end in range(120, 270):
        assert g.Extend(X[end:end + 1], y[end:end + 1]) is None
        if end in (126, 127, 128, 200, 255, 256, 269):
            og.absorb(X[:end + 1], y[:end + 1])
            assert abs(g.LML() - og.lml()) <= LML_TOL * max(abs(og.lml()), end + 1), end
    g.close()


def test_tutorial_evaluate_driver():
    """gogp_b200.tutorial.Evaluate (tutorial/tutorial.go:56-230) on the shipped barebones data: with the
    hyper-parameters held (jitter 0, no optimisation) the window grows by GP.Extend and must print what a loop of
    fresh fits prints; with the reference's defaults (jitter, L-BFGS) the rows have the reference's columns."""
    import io
    from gogp_b200 import tutorial
    name = "barebones"
    text = open(cases.GOLDEN + "/%s.csv" % name).read()
    X, Y = tutorial.load(text)
    mean, std = tutorial.mean_std(Y)
    Yn = (Y - mean) / std
    g = cases.make_device_gp(name)
    P = g.Simil.NTheta() + g.Noise.NTheta()
    out = io.StringIO()
    rows = tutorial.Evaluate(g, g, np.zeros(P), text, out, jitter=0.0, optimise=False)
    assert len(rows) == len(X) and len(out.getvalue().splitlines()) == len(X)
    ref = cases.make_device_gp(name)
    ref.ThetaSimil, ref.ThetaNoise = [1.0] * ref.Simil.NTheta(), [1.0] * ref.Noise.NTheta()
    for end in range(len(X)):
        assert ref.Absorb(X[:end], Yn[:end]) is None
        mu, sigma, err = ref.Produce(X[end:end + 1])
        want = list(X[end]) + [Y[end], mu[0] * std + mean, sigma[0] * std, ref.LML(), ref.LML()] + [1.0] * P
        assert np.max(np.abs(np.array(rows[end]) - np.array(want))) <= 1e-9 * max(1.0, np.max(np.abs(want))), end
    # the reference's defaults: jittered start, L-BFGS inside the library; 1 + 5 + P columns, "%f" formatted
    out2 = io.StringIO()
    rows2 = tutorial.Evaluate(g, g, np.zeros(P), text, out2, iters=20, rng=np.random.default_rng(3))
    lines = out2.getvalue().splitlines()
    assert len(lines) == len(X) and all(len(l.split(",")) == 1 + 5 + P for l in lines)
    assert all(np.isfinite(r).all() for r in rows2)
    assert all(r[5] >= r[4] - 1e-9 for r in rows2[2:])   # optimisation does not lower the LML (lml >= lml0)
    g.close()
    ref.close()


# ---- the largest sizes the oracle can still do on the host (SURVEY.md section 8c iii) --------------------
@pytest.mark.parametrize("name", ["c5_matern4", "hyperpriors", "c3_ard8"])
def test_n8192_vs_oracle(name):
    """LML, alpha and gradient at N = 8192 (64 tiles: every level of the look-ahead factorisation, the TMA GEMM,
    the single-launch solves, the specialised and the multi-term trace) against the oracle's `fast` mode."""
    N = 8192
    X, y, logt = cases.synth(name, N, seed=41)
    dg = cases.make_device_gp(name)
    og = cases.make_oracle_gp(name)
    dg.X, dg.Y = X, y
    og.X, og.Y = X, y
    lml = dg.Observe(logt.copy())
    grad = dg.Gradient()
    ref = og.observe(logt.copy())
    gref = og.gradient("fast")
    assert abs(lml - ref) <= LML_TOL * max(abs(ref), N), (lml, ref)
    assert _grad_err(grad, gref) <= GRAD_TOL, (grad, gref)
    a = dg.Alpha()
    assert np.max(np.abs(a - og.Alpha)) <= GRAD_TOL * max(1.0, np.max(np.abs(og.Alpha)))
    dg.close()


@pytest.mark.parametrize("name,N", [("warpedtime", 2048), ("c3_ard3", 1024)])
def test_with_obs_large_vs_oracle(name, N):
    """The tutorial anynoise / warpedtime layout, Observe([log theta | X | Y]) + Gradient with the inputs' and
    outputs' gradient, beyond toy sizes (the reference needs N D dense N x N matrices for the same numbers)."""
    ndim = cases.CASES[name][0]
    X, y, logt = cases.synth(name, N, seed=43)
    x = np.concatenate([logt, X.reshape(-1), y])
    dg = cases.make_device_gp(name)
    og = cases.make_oracle_gp(name)
    lml = dg.Observe(x.copy())
    grad = dg.Gradient()
    ref = og.observe(x.copy())
    gref = og.gradient("fast")
    assert len(grad) == len(logt) + N * (ndim + 1)
    assert abs(lml - ref) <= LML_TOL * max(abs(ref), N)
    P = len(logt)
    assert _grad_err(grad[:P], gref[:P]) <= GRAD_TOL
    assert _grad_err(grad[P:P + N * ndim], gref[P:P + N * ndim]) <= GRAD_TOL
    assert _grad_err(grad[P + N * ndim:], gref[P + N * ndim:]) <= GRAD_TOL
    dg.close()

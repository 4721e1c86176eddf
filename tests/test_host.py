"""CPU-side checks of the product's host logic: the C-ABI library loads and
exports every declared symbol, descriptor lowering accepts/rejects what it
should (it runs before any CUDA call), and the blocked recursion's index
arithmetic is right (driven through a naive host backend that lives in tests/)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
import scipy.linalg as sla

from tests.conftest import ROOT


def test_library_exports_every_declared_symbol(built_lib):
    from gogp_b200 import _lib
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "gogp_b200.h")).read()
    declared = set(re.findall(r"\b(gogp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s


def test_no_cpu_fallback_without_device(built_lib):
    """On a machine without a GPU the product path must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import gogp_b200 as g
    from gogp_b200 import kernel as k
    gp = g.GP(1, k.Normal, k.ConstantNoise(0.1))
    with pytest.raises(g.GoGPPanic) as e:
        gp.Observe(np.zeros(1))
    assert e.value.status == g._lib.CUDA_ERROR


def _create(ndim, simil, noise):
    from gogp_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    sd = simil.Descriptor()
    nd = noise.Descriptor() if noise is not None else None
    st = L.gogp_create(ndim, sd, len(sd), simil.NTheta(), nd, len(nd) if nd is not None else 0,
                       noise.NTheta() if noise is not None else 0, 0, C.byref(h))
    msg = L.gogp_last_error(h).decode() if h else ""
    if h:
        L.gogp_destroy(h)
    return st, msg


def test_descriptor_validation(built_lib):
    from gogp_b200 import _lib, kernel as k
    # a similarity leaf inside a noise program is rejected before any CUDA call
    st, msg = _create(1, k.Normal, k.Normal)
    assert st == _lib.UNSUPPORTED and "noise" in msg
    # input dimension out of range
    st, msg = _create(1, k.Normal.Of(l=0, dim=3), None)
    assert st == _lib.UNSUPPORTED and "dimension" in msg
    # too many product terms after distribution: (a+b)^4 = 16 terms > 8
    s = k.Param(0) * k.Normal.Of(l=1) + k.Param(2)
    st, msg = _create(1, s * s * s * s, None)
    assert st == _lib.UNSUPPORTED and "terms" in msg
    # a fine descriptor gets as far as the device check
    st, msg = _create(1, k.Param(0) * k.Matern52.Of(l=2) + k.Param(1) * k.Periodic.Of(l=3, p=(4, 10.0)),
                      0.01 * k.UniformNoise)
    assert st in (_lib.OK, _lib.CUDA_ERROR)


def test_kernel_expression_ntheta():
    from gogp_b200 import kernel as k
    assert k.Normal.NTheta() == 1 and k.Periodic.NTheta() == 2
    assert k.UniformNoise.NTheta() == 1 and k.ConstantNoise(0.1).NTheta() == 0
    assert (k.Param(0) * k.Matern32.Of(l=1)).NTheta() == 2
    assert k.Const(1e-5).WithNTheta(1).NTheta() == 1
    with pytest.raises(ValueError):
        k.Periodic.Of(l=0)


@pytest.fixture(scope="module")
def cpu_blocked():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libcpu_blocked.so")
    src = os.path.join(ROOT, "tests", "cpu_blocked_backend.cc")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, src])
    L = C.CDLL(so)
    L.cpu_blocked_potrf.restype = C.c_int
    return L


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("rl_max", [0, 256, 1024])
@pytest.mark.parametrize("n", [128, 384, 640])
def test_blocked_recursion_on_host_backend(cpu_blocked, n, rl_max):
    cpu_blocked.cpu_blocked_set_rl_max(C.c_int64(rl_max))
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n, n))
    K = G @ G.T / n + np.eye(n)
    A = K.copy()
    winv = np.zeros((n // 128, 128, 128))
    assert cpu_blocked.cpu_blocked_potrf(_p(A), C.c_int64(n), _p(winv)) == 0
    Lref = np.linalg.cholesky(K)
    assert np.abs(np.tril(A) - Lref).max() < 1e-13
    Bm = np.full((n, n), np.nan)
    dg = np.zeros((n // 128, 128, 128))
    cpu_blocked.cpu_blocked_kinv(_p(A), C.c_int64(n), _p(winv), _p(Bm), _p(dg))
    out = np.tril(Bm, -1)
    for t in range(n // 128):
        out[t * 128:(t + 1) * 128, t * 128:(t + 1) * 128] = dg[t]
    assert np.abs(np.tril(out) - np.tril(np.linalg.inv(K))).max() < 1e-12
    m = 256
    B = rng.standard_normal((m, n))
    B0 = B.copy()
    cpu_blocked.cpu_blocked_trsm(_p(A), C.c_int64(n), _p(winv), _p(B), C.c_int64(m))
    assert np.abs(B - sla.solve_triangular(Lref, B0.T, lower=True).T).max() < 1e-12


@pytest.mark.parametrize("n,nb,nb1,nb2", [(384, 128, 0, 0), (640, 128, 0, 0), (640, 256, 0, 0), (896, 256, 0, 0),
                                          (256, 128, 0, 0), (1280, 384, 128, 0), (1536, 512, 0, 128),
                                          (1024, 0, 128, 0), (640, 0, 0, 0), (2304, 768, 256, 128)])
def test_blocked_lookahead_factorisation_on_host_backend(cpu_blocked, n, nb, nb1, nb2):
    """Blocked::potrf_la (nested block columns, next column first) gives the same factor."""
    cpu_blocked.cpu_blocked_set_rl_max(C.c_int64(256))
    cpu_blocked.cpu_blocked_potrf_la.restype = C.c_int
    rng = np.random.default_rng(n + nb)
    G = rng.standard_normal((n, n))
    K = G @ G.T / n + np.eye(n)
    A = K.copy()
    winv = np.zeros((n // 128, 128, 128))
    cpu_blocked.cpu_blocked_potrf_la.argtypes = [C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double)] + [C.c_int64] * 3
    assert cpu_blocked.cpu_blocked_potrf_la(_p(A), C.c_int64(n), _p(winv), C.c_int64(nb), C.c_int64(nb1),
                                            C.c_int64(nb2)) == 0
    Lref = np.linalg.cholesky(K)
    assert np.abs(np.tril(A) - Lref).max() < 1e-13
    for t in range(n // 128):
        d = Lref[t * 128:(t + 1) * 128, t * 128:(t + 1) * 128]
        assert np.abs(winv[t] @ d - np.eye(128)).max() < 1e-12


def test_blocked_potrf_flags_non_pd(cpu_blocked):
    cpu_blocked.cpu_blocked_set_rl_max(C.c_int64(0))
    n = 256
    K = np.eye(n)
    K[200, 200] = -1.0
    winv = np.zeros((2, 128, 128))
    assert cpu_blocked.cpu_blocked_potrf(_p(K), C.c_int64(n), _p(winv)) == 201


def test_grid_has_no_cpu_fallback_either(built_lib):
    """gogp_create_grid on a machine without a GPU fails loudly (and an impossible grid shape is rejected first)."""
    import torch
    import gogp_b200 as g
    from gogp_b200 import kernel as k
    with pytest.raises(g.GoGPPanic) as e:
        g.GridGP(NDim=1, Simil=k.Normal, Noise=None, Devices=[0], Grid=(2, 1))
    assert e.value.status == g._lib.BAD_ARGUMENT
    with pytest.raises(g.GoGPPanic) as e:
        g.GridGP(NDim=1, Simil=k.Normal, Noise=None, Devices=[0], Block=100)
    assert e.value.status == g._lib.BAD_ARGUMENT
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(g.GoGPPanic) as e:
        g.GridGP(NDim=1, Simil=k.Normal, Noise=k.ConstantNoise(0.1), Devices=[0])
    assert e.value.status == g._lib.CUDA_ERROR


def test_tutorial_csv_format():
    """tutorial.load (tutorial/tutorial.go:234-272): D inputs then one output per record; a bad field is an error."""
    from gogp_b200 import tutorial
    X, Y = tutorial.load("0.5,1.5,2\n1,2,3.25\n")
    assert X.shape == (2, 2) and np.array_equal(Y, [2.0, 3.25]) and X[1, 1] == 2.0
    with pytest.raises(ValueError):
        tutorial.load("1,x\n")
    m, s = tutorial.mean_std([1.0, 2.0, 3.0, 4.0])
    assert m == 2.5 and abs(s - np.std([1, 2, 3, 4], ddof=1)) < 1e-15

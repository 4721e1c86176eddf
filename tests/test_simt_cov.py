"""The covariance-side kernels (gogp_b200/csrc/cov_kernels.cuh: build -- interpreted and specialised --,
fused gradient trace incl. the block mode of the distributed path, input gradient) compiled UNMODIFIED for the
host under the SIMT emulator of tests/simt/ and checked against the oracle.  No GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import cases
from tests.conftest import ROOT

SIMT = os.path.join(ROOT, "tests", "simt")
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def simt():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libsimt_cov.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-Wno-unknown-pragmas", "-pthread", "-shared", "-fPIC", "-o", so,
                           os.path.join(SIMT, "cov_host.cc")])
    from gogp_b200._lib import Op
    L = C.CDLL(so)
    head = [C.POINTER(Op), C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, dp, C.c_int64]   # descriptor, theta, events, X, N
    L.simt_cov_build.argtypes = head + [C.c_double, C.c_int, dp]
    L.simt_grad_trace.argtypes = head + [dp, dp, dp, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                         C.POINTER(C.c_int), C.c_int, dp]
    L.simt_grad_inputs.argtypes = head + [dp, dp, dp, dp]
    for f in (L.simt_cov_build, L.simt_grad_trace, L.simt_grad_inputs):
        f.restype = C.c_int
    return L


def _p(a):
    return a.ctypes.data_as(dp)


def _problem(name, N, seed):
    ndim, ds, _, _, _ = cases.CASES[name]
    X, y, logt = cases.synth(name, N, seed=seed)
    og = cases.make_oracle_gp(name)
    og.X, og.Y = X, y
    og.observe(logt.copy())
    nts = ds.NTheta()
    theta = np.exp(logt[:nts])
    sd = ds.Descriptor()
    ev = np.asarray(getattr(ds, "events", None) or np.zeros((0, 3)), dtype=np.float64).reshape(-1)
    noise_var = float(og.K[0, 0] - _k00(og, name, theta, X))
    args = (sd, len(sd), nts, ndim, _p(theta), _p(ev) if len(ev) else None, len(ev) // 3, _p(np.ascontiguousarray(X)), N)
    keep = (theta, ev, X)   # keep the arrays alive while ctypes holds their pointers
    return og, args, keep, nts, ndim, noise_var


def _k00(og, name, theta, X):
    from oracle.gp import _pairs
    osim = cases.CASES[name][3]
    v, _ = _pairs(osim, theta, X[:1], X[:1], "none")
    return float(v[0, 0])


@pytest.mark.parametrize("name,fast_expected", [("c2_rbf", True), ("c3_ard3", True), ("barebones", True),
                                                ("hyperpriors", False), ("events", True), ("c5_matern4", False)])
def test_build_kernels_on_the_emulator(simt, name, fast_expected):
    N = 200                                            # Npad = 256: diagonal, interior and padded tiles
    og, args, keep, nts, ndim, noise = _problem(name, N, seed=3)
    Npad = 256
    out = np.full((Npad, Npad), np.nan)
    assert simt.simt_cov_build(*args, noise, 0, _p(out)) == 0
    K = og.K
    tril = np.tril_indices(N)
    assert np.max(np.abs(out[:N, :N][tril] - K[tril]) / np.maximum(np.abs(K[tril]), 1e-300)) <= 1e-12
    pad = out[N:, :]
    assert np.array_equal(pad[:, N:][np.tril_indices(Npad - N)], np.eye(Npad - N)[np.tril_indices(Npad - N)])
    assert np.all(pad[:, :N] == 0.0)
    fast = np.full((Npad, Npad), np.nan)
    rc = simt.simt_cov_build(*args, noise, 1, _p(fast))
    assert (rc == 0) == fast_expected
    if fast_expected:                                  # same operations in the same order: bit-identical
        assert np.array_equal(np.nan_to_num(fast, nan=-1.0), np.nan_to_num(out, nan=-1.0))


@pytest.mark.parametrize("name", ["c2_rbf", "c3_ard3", "hyperpriors", "events", "c5_matern4", "c3_ard8", "sum_times"])
@pytest.mark.parametrize("fast", [0, 1])
def test_trace_kernel_on_the_emulator(simt, name, fast):
    """fast = 0: the descriptor interpreter; fast = 1: the specialised kernel (Normal factors unrolled, value and
    log-derivatives of the other leaves evaluated together, everything in registers)."""
    N = 200
    og, args, keep, nts, ndim, _ = _problem(name, N, seed=4)
    Npad, T = 256, 2
    gref = og.gradient()[:nts]
    Kinv = np.linalg.inv(og.K)
    alpha = np.zeros(Npad)
    alpha[:N] = og.Alpha
    dense = np.eye(Npad)
    dense[:N, :N] = Kinv
    kinv = np.full((Npad, Npad), np.nan)               # strictly-lower tiles only; diagonal tiles live in kdiag
    kinv[128:, :128] = dense[128:, :128]
    kdiag = np.stack([dense[t * 128:(t + 1) * 128, t * 128:(t + 1) * 128] for t in range(T)]).copy()
    out = np.zeros(nts + 1)
    assert simt.simt_grad_trace(*args, _p(alpha), _p(kinv), _p(kdiag), 0, 0, 0, 0, 0, None, fast, _p(out)) == 0
    scale = max(1.0, np.max(np.abs(gref)))
    assert np.max(np.abs(out[:nts] - gref)) <= 1e-10 * scale, (out[:nts], gref)
    trw = float(og.Alpha @ og.Alpha - np.trace(Kinv))
    assert abs(out[nts] - trw) <= 1e-10 * max(1.0, abs(trw))
    # the block mode of the distributed path: the three lower 128-blocks, accumulated, give the same sums
    acc = np.zeros(nts + 1)
    for (r0, c0) in ((0, 0), (128, 0), (128, 128)):
        assert simt.simt_grad_trace(*args, _p(alpha), _p(dense), None, 1, r0, 128, c0, 128, None, fast, _p(acc)) == 0
    assert np.max(np.abs(acc - out)) <= 1e-12 * max(1.0, np.max(np.abs(out)))


@pytest.mark.parametrize("fast", [0, 1])
@pytest.mark.parametrize("Pr,Pc,tb", [(2, 1, 1), (1, 2, 1), (2, 2, 1), (1, 1, 2), (2, 1, 2)])
def test_trace_over_block_cyclic_local_matrices(simt, fast, Pr, Pc, tb):
    """Mode 2 (grid.hpp's single trace launch per rank): every rank's local matrix of a Pr x Pc block-cyclic
    distribution (blocks above the global diagonal poisoned) -- the ranks' sums add up to the whole trace."""
    name, N = "c3_ard3", 500
    og, args, keep, nts, ndim, _ = _problem(name, N, seed=6)
    Npad, T = 512, 4
    gref = og.gradient()[:nts]
    Kinv = np.linalg.inv(og.K)
    alpha = np.zeros(Npad)
    alpha[:N] = og.Alpha
    dense = np.eye(Npad)
    dense[:N, :N] = Kinv
    nb, NB = T // tb, 128 * tb
    acc = np.zeros(nts + 1)
    for r in range(Pr):
        for c in range(Pc):
            rows = [I for I in range(nb) if I % Pr == r]
            cols = [J for J in range(nb) if J % Pc == c]
            if not rows or not cols:
                continue
            local = np.full((len(rows) * NB, len(cols) * NB), np.nan)
            for i, I in enumerate(rows):
                for j, J in enumerate(cols):
                    if J <= I:
                        local[i * NB:(i + 1) * NB, j * NB:(j + 1) * NB] = dense[I * NB:(I + 1) * NB, J * NB:(J + 1) * NB]
            if tb > 1:   # within a diagonal block the tiles above the diagonal must not be read either
                for i, I in enumerate(rows):
                    for j, J in enumerate(cols):
                        if I == J:
                            for a in range(tb):
                                for b in range(a + 1, tb):
                                    local[i * NB + a * 128:i * NB + (a + 1) * 128, j * NB + b * 128:j * NB + (b + 1) * 128] = np.nan
            bc = (C.c_int * 3)(tb, Pr, Pc)
            local = np.ascontiguousarray(local)
            assert simt.simt_grad_trace(*args, _p(alpha), _p(local), None, 2, r, local.shape[0], c, local.shape[1], bc,
                                        fast, _p(acc)) == 0
    assert np.max(np.abs(acc[:nts] - gref)) <= 1e-10 * max(1.0, np.max(np.abs(gref))), (acc[:nts], gref)


def test_input_gradient_kernel_on_the_emulator(simt):
    name, N = "c3_ard3", 43
    ndim, ds, _, _, _ = cases.CASES[name]
    X, y, logt = cases.synth(name, N, seed=5)
    og = cases.make_oracle_gp(name)
    x = np.concatenate([logt, X.reshape(-1), y])
    og.observe(x)                                      # with_obs layout: [log theta | X flat | Y]
    g = og.gradient()
    P = len(logt)
    gx_ref = g[P:P + N * ndim]
    nts = ds.NTheta()
    theta = np.exp(logt[:nts])
    sd = ds.Descriptor()
    Kinv = np.linalg.inv(og.K)
    Npad = 128
    kinv = np.full((Npad, Npad), np.nan)
    kdiag = np.eye(Npad)
    kdiag[:N, :N] = Kinv
    alpha = np.zeros(Npad)
    alpha[:N] = og.Alpha
    Xc = np.ascontiguousarray(X)
    gx = np.zeros(N * ndim)
    assert simt.simt_grad_inputs(sd, len(sd), nts, ndim, _p(theta), None, 0, _p(Xc), N, _p(alpha), _p(kinv),
                                 _p(kdiag.reshape(1, Npad, Npad).copy()), _p(gx)) == 0
    assert np.max(np.abs(gx - gx_ref)) <= 1e-9 * max(1.0, np.max(np.abs(gx_ref)))

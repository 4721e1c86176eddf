"""Descriptor lowering + closed-form partials on the CPU: the product's own lowering
(csrc/program.cc) and element functions (csrc/kexpr.cuh, compiled as host C++ by
tests/cpu_kexpr_backend.cc) against the oracle's forward-mode AD of the restated reference kernels,
for every kernel configuration of the parity suite.  The device kernels call exactly these functions."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import cases
from tests.conftest import ROOT


@pytest.fixture(scope="module")
def kexpr():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libcpu_kexpr.so")
    src = os.path.join(ROOT, "tests", "cpu_kexpr_backend.cc")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-shared", "-fPIC", "-o", so, src])
    L = C.CDLL(so)
    L.cpu_kexpr_eval.restype = C.c_int
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_host_evaluation_of_the_descriptor_matches_the_oracle(kexpr, name):
    from oracle.gp import _pairs
    ndim, ds, _, osim, _ = cases.CASES[name]
    nt = ds.NTheta()
    assert nt == osim.ntheta
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
    theta = np.exp(0.4 * rng.standard_normal(nt))
    if name == "hyperpriors":
        theta[4] = 0.3
    A = rng.uniform(0.0, 7.0, size=(6, ndim))
    B = rng.uniform(0.0, 7.0, size=(5, ndim))
    B[0] = A[0]                                   # r = 0: |.| and sign at the kink
    val, grads = _pairs(osim, theta, A, B, "all")
    sd = ds.Descriptor()
    ev = np.asarray(getattr(ds, "events", None) or np.zeros((0, 3)), dtype=np.float64).reshape(-1)
    for i in range(len(A)):
        for j in range(len(B)):
            v = C.c_double()
            dlog, dxa = np.zeros(max(nt, 1)), np.zeros(ndim)
            xa, xb = np.ascontiguousarray(A[i]), np.ascontiguousarray(B[j])
            rc = kexpr.cpu_kexpr_eval(sd, len(sd), nt, ndim, _dp(theta), _dp(ev) if len(ev) else None, len(ev) // 3,
                                      _dp(xa), _dp(xb), C.byref(v), _dp(dlog), _dp(dxa))
            assert rc == 0
            scale = max(abs(val[i, j]), 1e-300)
            assert abs(v.value - val[i, j]) <= 1e-12 * scale, (name, i, j, v.value, val[i, j])
            for q in range(nt):
                ref = theta[q] * grads[q][i, j] if q in grads else 0.0
                assert abs(dlog[q] - ref) <= 1e-11 * max(abs(ref), scale), (name, "theta", q, dlog[q], ref)
            for d in range(ndim):
                ra = grads[nt + d][i, j] if nt + d in grads else 0.0
                rb = grads[nt + ndim + d][i, j] if nt + ndim + d in grads else 0.0
                assert abs(dxa[d] - ra) <= 1e-11 * max(abs(ra), scale), (name, "xa", d, dxa[d], ra)
                assert abs(ra + rb) <= 1e-11 * max(abs(ra), scale)      # stationary leaves: d/dxb = -d/dxa

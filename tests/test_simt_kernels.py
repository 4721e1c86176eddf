"""The product's two most intricate kernels -- the blocked 128 x 128 tile Cholesky + inverse and the
single-launch (ticket / release-acquire) triangular solves, gogp_b200/csrc/leaf_kernels.cuh -- compiled
UNMODIFIED for the host under the lock-step SIMT emulator of tests/simt/ and checked against NumPy, plus a
ThreadSanitizer run of the same code as a shared-memory race check.  No GPU needed; the device build of
the same header is what tests/test_gpu_parity.py exercises on the B200."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg as sla

from tests.conftest import ROOT

SIMT = os.path.join(ROOT, "tests", "simt")
FLAGS = ["-std=c++20", "-O1", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-variable", "-pthread"]
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def simt():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libsimt_leaf.so")
    subprocess.check_call(["g++"] + FLAGS + ["-shared", "-fPIC", "-o", so, os.path.join(SIMT, "leaf_host.cc")])
    return C.CDLL(so)


def _leaf(simt, A, ld=128, base=0, variant=0):
    buf = np.full((128, ld), np.nan)
    buf[:, :128] = A
    W = np.zeros((128, 128))
    info = C.c_int(0)
    simt.simt_potrf_leaf(buf.ctypes.data_as(dp), C.c_int64(ld), W.ctypes.data_as(dp), C.byref(info), base, variant)
    return buf[:, :128].copy(), W, info.value


@pytest.mark.parametrize("variant", [0, 2])
@pytest.mark.parametrize("ld,shift", [(128, 1.0), (384, 1e-3)])
def test_tile_kernel_source_on_the_emulator(simt, ld, shift, variant):
    rng = np.random.default_rng(int(ld))
    G = rng.standard_normal((128, 128))
    K = G @ G.T / 128 + shift * np.eye(128)
    A = np.tril(K) + np.triu(rng.standard_normal((128, 128)), 1)   # the upper triangle must be ignored
    Lo, W, info = _leaf(simt, A, ld, variant=variant)
    if variant == 2:   # same operations on the same values as the shipped variant: bit-identical
        L0, W0, _ = _leaf(simt, A, ld, variant=0)
        assert np.array_equal(Lo, L0) and np.array_equal(W, W0)
    ref = np.linalg.cholesky(K)
    cond = np.linalg.cond(ref)
    assert info == 0
    assert np.abs(Lo - ref).max() <= 1e-14 * cond * np.abs(ref).max()
    assert np.abs(np.triu(Lo, 1)).max() == 0.0 and np.abs(np.triu(W, 1)).max() == 0.0
    assert np.abs(W @ ref - np.eye(128)).max() <= 1e-14 * cond ** 2


def test_tile_kernel_flags_the_first_bad_pivot(simt):
    for bad in (0, 40, 127):
        K = 2.0 * np.eye(128)
        K[bad, bad] = -1.0
        for variant in (0, 2):
            _, _, info = _leaf(simt, K, base=256, variant=variant)
            assert info == 256 + bad + 1


def test_chained_solves_source_on_the_emulator(simt):
    rng = np.random.default_rng(5)
    T = 3
    n = T * 128
    G = rng.standard_normal((n, n))
    Lm = np.linalg.cholesky(G @ G.T / n + np.eye(n))
    Lp = np.tril(Lm) + np.triu(np.full((n, n), np.nan), 1)          # poison: never read above the diagonal
    winv = np.ascontiguousarray(np.stack([np.tril(np.linalg.inv(Lm[t * 128:(t + 1) * 128, t * 128:(t + 1) * 128]))
                                          for t in range(T)]))
    rhs = rng.standard_normal(n)
    sync = np.zeros(T + 1, dtype=np.uint32)
    for transposed in (0, 1):
        out = np.full(n, np.nan)
        simt.simt_trsv(Lp.ctypes.data_as(dp), C.c_int64(n), winv.ctypes.data_as(dp), rhs.ctypes.data_as(dp),
                       out.ctypes.data_as(dp), T, transposed, sync.ctypes.data_as(C.POINTER(C.c_uint)))
        ref = sla.solve_triangular(Lm, rhs, lower=True, trans="T" if transposed else "N")
        assert np.abs(out - ref).max() <= 1e-13 * np.abs(ref).max()
        assert sync[0] == T and np.all(sync[1:] == 1)                 # every CTA took a ticket and released its flag


def test_no_data_race_under_thread_sanitizer():
    """Shared-memory protocol of the tile kernel (twelve CTA barriers, warp-synchronous phases) and of the
    solves: ThreadSanitizer sees every CUDA thread as an OS thread; a missing barrier is a reported race."""
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "simt_tsan")
    r = subprocess.run(["g++"] + FLAGS + ["-g", "-fsanitize=thread", "-o", exe, os.path.join(SIMT, "tsan_main.cc"),
                                          os.path.join(SIMT, "leaf_host.cc")], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr[-200:])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "data race" not in r.stdout + r.stderr, (r.stdout + r.stderr)[-2000:]
    assert "leaf info 0" in r.stdout

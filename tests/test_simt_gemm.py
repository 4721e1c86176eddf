"""Both FP64 DMMA GEMM kernels -- the cp.async one (gogp_b200/csrc/dgemm_kernels.cuh, both shipped CTA shapes) and
the TMA + mbarrier warp-specialised one (gogp_b200/csrc/dgemm_tma_kernel.cuh) -- compiled UNMODIFIED for the host under
the SIMT emulator and checked against NumPy in every tile-map mode the blocked algebra uses.  The m8n8k4 MMA is the
emulator's warp collective with the PTX fragment layout; cp.async is an immediate copy; the emulator models mbarrier
phases / transaction counts, the 2-D tensor copy with its 128-byte swizzle and the named barrier, so the TMA
kernel's ring protocol and swizzled fragment addressing run -- and are race-checked under ThreadSanitizer -- on a CPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.conftest import ROOT

dp, i64 = C.POINTER(C.c_double), C.c_int64
FULL, LOWER, KTRI, DIAG_OUT, INPLACE = 0, 1, 2, 4, 8


@pytest.fixture(scope="module")
def gemm():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libsimt_gemm.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-Wno-unknown-pragmas", "-pthread", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "simt", "gemm_host.cc")])
    L = C.CDLL(so)
    L.simt_dgemm.argtypes = [dp, i64, dp, i64, dp, i64, i64, i64, i64, C.c_double, C.c_double, C.c_int, dp, C.c_int]
    L.simt_dgemm.restype = None
    return L


@pytest.fixture(scope="module")
def tma():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libsimt_gemm_tma.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-Wno-unknown-pragmas", "-pthread", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "simt", "gemm_tma_host.cc")])
    L = C.CDLL(so)
    L.simt_dgemm_tma.argtypes = [dp, i64, dp, i64, dp, i64, i64, i64, i64, C.c_double, C.c_double, C.c_int, dp]
    L.simt_dgemm_tma.restype = None
    return L


def _p(a):
    return a.ctypes.data_as(dp)


def _tiles(M, t=128):
    return M.reshape(M.shape[0] // t, t, M.shape[1] // t, t).swapaxes(1, 2)


@pytest.mark.parametrize("shape", [0, 1, 2, 3])
def test_full_rectangular_update(gemm, shape):
    rng = np.random.default_rng(shape)
    m, n, k = 256, 128, 256
    A, B = rng.standard_normal((m, k + 16))[:, :k], rng.standard_normal((n, k))      # lda != k
    Cm = rng.standard_normal((m, n))
    ref = 0.5 * Cm - 1.0 * (A @ B.T)
    gemm.simt_dgemm(_p(Cm), n, _p(A), A.strides[0] // 8, _p(B), k, m, n, k, -1.0, 0.5, FULL, None, shape)
    assert np.abs(Cm - ref).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("shape", [0, 1, 2])
def test_lower_tiles_only_syrk(gemm, shape):
    rng = np.random.default_rng(10 + shape)
    n, k = 256, 128
    A = rng.standard_normal((n, k))
    Cm = rng.standard_normal((n, n))
    C0 = Cm.copy()
    gemm.simt_dgemm(_p(Cm), n, _p(A), k, _p(A), k, n, n, k, -1.0, 1.0, LOWER, None, shape)
    ref = C0 - A @ A.T
    t = 64 if shape == 2 else 128                            # the latency shape walks 64 x 64 tiles
    T, Tc, Tr = _tiles(Cm, t), _tiles(C0, t), _tiles(ref, t)
    for i in range(n // t):
        for j in range(n // t):
            want = Tr[i, j] if j <= i else Tc[i, j]          # tiles above the diagonal are not touched
            assert np.abs(T[i, j] - want).max() <= 1e-13 * np.abs(ref).max()


def test_latency_shape_triangular_k_range(gemm):
    """U12 = -U11 L21^T (Blocked::trtri_t) with the 64 x 64 shape: A upper triangular w.r.t. its own origin, the k
    range of a 64-row tile starts at its first row; what lies below the diagonal inside A is never read."""
    rng = np.random.default_rng(21)
    m, n = 256, 128
    U = np.triu(rng.standard_normal((m, m)))
    Up = U.copy()
    for t in range(m // 64):                                 # poison everything left of each 64-row tile's k range
        Up[t * 64:(t + 1) * 64, :t * 64] = np.nan
    L21 = rng.standard_normal((n, m))
    Cm = np.full((m, n), np.nan)
    gemm.simt_dgemm(_p(Cm), n, _p(Up), m, _p(L21), m, m, n, m, -1.0, 0.0, KTRI, None, 2)
    ref = -U @ L21.T
    assert np.abs(Cm - ref).max() <= 1e-13 * np.abs(ref).max()


def test_triangular_k_range_and_diagonal_tiles_to_side_buffer(gemm):
    """K^-1 = U U^T as one launch (Blocked::lauum): A upper triangular w.r.t. its own origin (k starts at the
    row tile), lower tiles only, diagonal tiles into cdiag; beta = 0 must not read C."""
    rng = np.random.default_rng(20)
    n = 256
    U = np.triu(rng.standard_normal((n, n)))
    Up = U + np.tril(np.full((n, n), np.nan), -129)           # poison far below the diagonal tiles: never read
    Cm = np.full((n, n), np.nan)
    Cm[:128, 128:] = 7.0                                      # the upper tile must stay as it is
    cdiag = np.full((2, 128, 128), np.nan)
    gemm.simt_dgemm(_p(Cm), n, _p(Up), n, _p(Up), n, n, n, n, 1.0, 0.0, LOWER | KTRI | DIAG_OUT, _p(cdiag), 1)
    ref = U @ U.T
    assert np.abs(Cm[128:, :128] - ref[128:, :128]).max() <= 1e-13 * np.abs(ref).max()
    assert np.all(Cm[:128, 128:] == 7.0)
    for t in range(2):
        assert np.abs(cdiag[t] - ref[t * 128:(t + 1) * 128, t * 128:(t + 1) * 128]).max() <= 1e-13 * np.abs(ref).max()
    assert np.all(np.isnan(Cm[:128, :128])) and np.all(np.isnan(Cm[128:, 128:]))   # diagonal tiles went to cdiag


def test_in_place_solve_with_a_diagonal_block(gemm):
    """X = B Winv^T with C aliasing A (the base case of Blocked::trsm): one CTA owns whole rows."""
    rng = np.random.default_rng(30)
    m = 256
    Bm = rng.standard_normal((m, 384))                        # the 128 columns live inside a wider panel
    W = np.tril(rng.standard_normal((128, 128)))
    ref = Bm[:, 128:256] @ W.T
    keep = Bm.copy()
    view = Bm[:, 128:256]
    ptr = C.cast(Bm.ctypes.data + 128 * 8, dp)
    gemm.simt_dgemm(ptr, 384, ptr, 384, _p(W), 128, m, 128, 128, 1.0, 0.0, INPLACE, None, 0)
    assert np.abs(view - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.array_equal(Bm[:, :128], keep[:, :128]) and np.array_equal(Bm[:, 256:], keep[:, 256:])
    # the 32 x 128 latency shape: eight CTAs, each owning 32 whole rows
    Bm[:] = keep
    gemm.simt_dgemm(ptr, 384, ptr, 384, _p(W), 128, m, 128, 128, 1.0, 0.0, INPLACE, None, 3)
    assert np.abs(view - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.array_equal(Bm[:, :128], keep[:, :128]) and np.array_equal(Bm[:, 256:], keep[:, 256:])


def test_tma_kernel_full_update_with_a_wrapping_ring(tma):
    rng = np.random.default_rng(40)
    m, n, k = 128, 256, 512                                   # 32 k-tiles through a 6-stage ring
    A, B = rng.standard_normal((m, k + 32))[:, 16:16 + k], rng.standard_normal((n, k))   # offset, strided view of A
    Cm = rng.standard_normal((m, n))
    ref = 0.25 * Cm - 1.0 * (A @ B.T)
    ptr = C.cast(A.ctypes.data, dp)
    tma.simt_dgemm_tma(_p(Cm), n, ptr, A.strides[0] // 8, _p(B), k, m, n, k, -1.0, 0.25, FULL, None)
    assert np.abs(Cm - ref).max() <= 1e-13 * np.abs(ref).max()


def test_tma_kernel_lauum_modes(tma):
    rng = np.random.default_rng(41)
    n = 256
    U = np.triu(rng.standard_normal((n, n)))
    Up = U + np.tril(np.full((n, n), np.nan), -129)
    Cm = np.full((n, n), np.nan)
    Cm[:128, 128:] = 7.0
    cdiag = np.full((2, 128, 128), np.nan)
    tma.simt_dgemm_tma(_p(Cm), n, _p(Up), n, _p(Up), n, n, n, n, 1.0, 0.0, LOWER | KTRI | DIAG_OUT, _p(cdiag))
    ref = U @ U.T
    assert np.abs(Cm[128:, :128] - ref[128:, :128]).max() <= 1e-13 * np.abs(ref).max()
    assert np.all(Cm[:128, 128:] == 7.0)
    for t in range(2):
        assert np.abs(cdiag[t] - ref[t * 128:(t + 1) * 128, t * 128:(t + 1) * 128]).max() <= 1e-13 * np.abs(ref).max()


def test_tma_kernel_in_place_solve(tma):
    rng = np.random.default_rng(42)
    m = 256
    Bm = rng.standard_normal((m, 384))
    W = np.tril(rng.standard_normal((128, 128)))
    ref = Bm[:, 128:256] @ W.T
    keep = Bm.copy()
    ptr = C.cast(Bm.ctypes.data + 128 * 8, dp)
    tma.simt_dgemm_tma(ptr, 384, ptr, 384, _p(W), 128, m, 128, 128, 1.0, 0.0, INPLACE, None)
    assert np.abs(Bm[:, 128:256] - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.array_equal(Bm[:, :128], keep[:, :128]) and np.array_equal(Bm[:, 256:], keep[:, 256:])


def test_tma_ring_protocol_under_thread_sanitizer():
    """Producer warp / eight consumer warps, full and empty mbarriers, stage reuse: any read of a stage before its
    data landed, or a refill before every consumer released it, is a reported race."""
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "simt_tsan_tma")
    simt = os.path.join(ROOT, "tests", "simt")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=thread", "-Wno-unknown-pragmas", "-pthread", "-o", exe,
                        os.path.join(simt, "tsan_tma_main.cc"), os.path.join(simt, "gemm_tma_host.cc")],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr[-200:])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "data race" not in r.stdout + r.stderr, (r.stdout + r.stderr)[-2000:]

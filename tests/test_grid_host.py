"""The C++ block-cyclic orchestration of the library (gogp_b200/csrc/grid.hpp: factor, solve, the fused
V = L^-T / K^-1 = V V^T sweep, alpha, trace) on a machine without a GPU: tests/cpu_grid_backend.cc runs it
unmodified with the ranks of a Pr x Pc grid as host threads, against the oracle's LML, alpha and gradient."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import cases
from tests.conftest import ROOT


@pytest.fixture(scope="module")
def grid():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libcpu_grid.so")
    src = os.path.join(ROOT, "tests", "cpu_grid_backend.cc")
    deps = [src, os.path.join(ROOT, "tests", "host_tiles.h")] + [
        os.path.join(ROOT, "gogp_b200", "csrc", f) for f in ("grid.hpp", "blocked.hpp", "kexpr.cuh", "program.cc")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-pthread", "-shared", "-fPIC", "-o", so, src])
    L = C.CDLL(so)
    L.cpu_grid_eval.restype = C.c_int
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _run(grid, name, N, NB, Pr, Pc, seed=3):
    ndim, ds, dn, _, _ = cases.CASES[name]
    X, y, logt = cases.synth(name, N, seed=seed)
    nts, ntn = ds.NTheta(), dn.NTheta()
    theta = np.exp(logt)
    ts, tn = np.ascontiguousarray(theta[:nts]), np.ascontiguousarray(theta[nts:] if ntn else np.zeros(1))
    sd, nd = ds.Descriptor(), dn.Descriptor()
    lml = C.c_double()
    grad, alpha, stats = np.zeros(nts + ntn), np.zeros(N), np.zeros(4)
    Xc = np.ascontiguousarray(X.reshape(-1))
    rc = grid.cpu_grid_eval(sd, len(sd), nts, nd, len(nd), ntn, ndim, _dp(ts), _dp(tn), _dp(Xc), _dp(y),
                            C.c_int64(N), C.c_int64(NB), Pr, Pc, C.byref(lml), _dp(grad), _dp(alpha), _dp(stats))
    return rc, lml.value, grad, alpha, stats, (X, y, logt)


def _oracle(name, X, y, logt):
    og = cases.make_oracle_gp(name)
    og.X, og.Y = X, y
    ref = og.observe(logt.copy())
    return ref, og.gradient(), np.asarray(og.Alpha)


GRIDS = [(1, 1), (2, 1), (1, 2), (2, 2), (4, 2), (2, 4), (3, 2)]


@pytest.mark.parametrize("Pr,Pc", GRIDS)
def test_block_cyclic_lml_alpha_gradient_match_the_oracle(grid, Pr, Pc):
    name, N, NB = "c5_matern4", 700, 128
    rc, lml, grad, alpha, stats, (X, y, logt) = _run(grid, name, N, NB, Pr, Pc)
    assert rc == 0
    ref, gref, aref = _oracle(name, X, y, logt)
    assert abs(lml - ref) <= 1e-9 * max(abs(ref), N)
    assert np.max(np.abs(alpha - aref)) <= 1e-7 * max(1.0, np.max(np.abs(aref)))
    assert np.max(np.abs(grad - gref)) <= 1e-7 * max(1.0, np.max(np.abs(gref)))


def test_result_does_not_depend_on_the_grid_beyond_rounding(grid):
    name, N, NB = "hyperpriors", 600, 128
    base = _run(grid, name, N, NB, 1, 1)
    assert base[0] == 0
    for Pr, Pc in [(2, 2), (4, 2)]:
        rc, lml, grad, alpha, _, _ = _run(grid, name, N, NB, Pr, Pc)
        assert rc == 0
        assert abs(lml - base[1]) <= 1e-11 * max(abs(base[1]), N)
        assert np.max(np.abs(grad - base[2])) <= 1e-9 * max(1.0, np.max(np.abs(base[2])))


@pytest.mark.parametrize("name,N,NB,Pr,Pc", [
    ("c3_ard3", 520, 256, 2, 1),      # 2 x 2 tiles per distribution block: the tile-level mask on diagonal blocks
    ("c2_rbf", 1000, 256, 2, 2),
    ("sum_times", 300, 128, 2, 2),    # fewer block rows than some ranks need: empty owners
    ("barebones", 100, 128, 2, 2),    # one block, three idle ranks
    ("c3_ard8", 384, 128, 4, 2),      # N a multiple of the block, more process rows than blocks
])
def test_shapes_and_kernels(grid, name, N, NB, Pr, Pc):
    rc, lml, grad, alpha, stats, (X, y, logt) = _run(grid, name, N, NB, Pr, Pc)
    assert rc == 0
    ref, gref, aref = _oracle(name, X, y, logt)
    assert abs(lml - ref) <= 1e-9 * max(abs(ref), N)
    assert np.max(np.abs(alpha - aref)) <= 1e-7 * max(1.0, np.max(np.abs(aref)))
    assert np.max(np.abs(grad - gref)) <= 1e-7 * max(1.0, np.max(np.abs(gref)))


@pytest.mark.parametrize("Pr,Pc,N,NB", [(8, 1, 1100, 128), (1, 4, 1100, 128), (4, 1, 900, 128), (3, 3, 1300, 128),
                                         (2, 2, 1100, 384), (4, 2, 2100, 256)])
def test_more_grid_shapes(grid, Pr, Pc, N, NB):
    """Grids taller, wider and odder than the shipped 2 x 1 / 2 x 2 / 4 x 2, block counts that do not divide evenly,
    blocks of three tiles: LML and gradient against the oracle."""
    name = "barebones"
    rc, lml, grad, alpha, stats, (X, y, logt) = _run(grid, name, N, NB, Pr, Pc, seed=9)
    assert rc == 0
    ref, gref, aref = _oracle(name, X, y, logt)
    assert abs(lml - ref) <= 1e-9 * max(abs(ref), N)
    assert np.max(np.abs(alpha - aref)) <= 1e-7 * max(1.0, np.max(np.abs(aref)))
    assert np.max(np.abs(grad - gref)) <= 1e-7 * max(1.0, np.max(np.abs(gref)))


def test_the_mask_skips_the_upper_half_and_flops_stay_at_n_cubed(grid):
    """GEMM tiles actually computed over the whole grid ~ (N/128)^3 tile steps per LML + gradient evaluation:
    the masked launches must not compute (or write) blocks above the global diagonal."""
    name, N, NB = "c5_matern4", 1536, 128
    rc, _, _, _, stats, _ = _run(grid, name, N, NB, 2, 2)
    assert rc == 0
    T = N // 128
    # tile-products of K = NB: factor T^3/6, V T^3/6 + solves, K^-1 T^3/6 (+ lower-order terms) -> count tiles
    computed, skipped = stats[0], stats[1]
    assert skipped > 0
    # every computed tile is one 128x128xNB product; the three phases together need ~ T^3/2 of them
    assert computed <= 0.75 * T ** 3


def test_not_positive_definite_is_reported_by_every_rank(grid):
    name, N, NB = "normal_const", 300, 128
    ndim, ds, dn, _, _ = cases.CASES[name]
    X, y, logt = cases.synth(name, N, seed=1)
    X[150] = X[10]            # a duplicated input with (almost) no noise: K is singular to rounding
    nts = ds.NTheta()
    sd = ds.Descriptor()
    from gogp_b200 import kernel as k
    nd = k.ConstantNoise(0.0).Descriptor()
    ts = np.ascontiguousarray(np.exp(logt[:nts]) * 50.0)   # long length scale: numerically rank deficient
    tn = np.zeros(1)
    lml = C.c_double()
    grad, alpha, stats = np.zeros(nts), np.zeros(N), np.zeros(4)
    Xc = np.ascontiguousarray(X.reshape(-1))
    rc = grid.cpu_grid_eval(sd, len(sd), nts, nd, len(nd), 0, ndim, _dp(ts), _dp(tn), _dp(Xc), _dp(y),
                            C.c_int64(N), C.c_int64(NB), 2, 2, C.byref(lml), _dp(grad), _dp(alpha), _dp(stats))
    assert rc == 2
    assert 1 <= int(lml.value) <= N

"""The 2D block-cyclic Cholesky orchestration (gogp_b200/dist_chol.py) on CPU:
gloo ranks + a NumPy compute backend defined here (test infrastructure), against the
oracle's log marginal likelihood.  On the GPU box the same orchestration runs with the
CUDA backend (tests/test_gpu_dist.py, tools/dist_bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.linalg as sla
import torch
import torch.multiprocessing as mp

from tests.conftest import ROOT

NAME, N, NB = "c5_matern4", 700, 128


class NumpyBlocks:
    """Same interface as dist_chol.CudaBlocks, NumPy arithmetic on torch CPU tensors."""

    def __init__(self, oracle_simil, oracle_noise):
        self.simil, self.noise, self.bad = oracle_simil, oracle_noise, 0

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=torch.float64)

    def empty(self, *shape):
        return torch.full(tuple(shape), float("nan"), dtype=torch.float64)  # poison: nothing may depend on it

    def set_inputs(self, X, block=128):
        self.X = np.asarray(X, dtype=np.float64)

    def cov_block(self, ts, tn, row0, rows, col0, cols, diagonal, out):
        from oracle.gp import _noise, _pairs
        n = len(self.X)
        blk = np.zeros((rows, cols))
        r1, c1 = min(n, row0 + rows), min(n, col0 + cols)
        if r1 > row0 and c1 > col0:
            k, _ = _pairs(self.simil, np.asarray(ts), self.X[col0:c1], self.X[row0:r1], "none")
            blk[:r1 - row0, :c1 - col0] = k.T
        if diagonal:
            nv, _ = _noise(self.noise, np.asarray(tn), self.X[row0:r1], "none")
            for i in range(rows):
                blk[i, i] = blk[i, i] + nv[i] if row0 + i < n else 1.0
            blk = np.tril(blk) + np.triu(np.full((rows, cols), np.nan), 1)  # the device builds lower tiles only
        out.copy_(torch.from_numpy(blk))

    def potrf(self, A, winv, base):
        a = np.tril(A.numpy())
        a = a + np.tril(a, -1).T
        try:
            L = np.linalg.cholesky(a)
        except np.linalg.LinAlgError:
            self.bad = base + 1
            L = np.eye(len(a))
        A.copy_(torch.from_numpy(L))
        for t in range(len(a) // 128):
            blk = L[t * 128:(t + 1) * 128, t * 128:(t + 1) * 128]
            winv[t].copy_(torch.from_numpy(np.linalg.inv(blk)))

    def trsm(self, B, L, winv):
        x = sla.solve_triangular(np.tril(L.numpy()), B.numpy().T, lower=True).T
        B.copy_(torch.from_numpy(np.ascontiguousarray(x)))

    def gemm(self, Cm, A, B, alpha, beta, lower=False):
        Cm.copy_(alpha * (A @ B.T) if beta == 0.0 else beta * Cm + alpha * (A @ B.T))  # beta = 0: C is not read

    def sumlogdiag(self, Lb, nvalid, out2):
        out2[0] = float(np.sum(np.log(np.diag(Lb.numpy())[:max(nvalid, 0)])))

    def gemv_sub(self, B, v, acc, scratch):
        acc -= B @ v

    def trsv(self, Lb, winv, rhs, z):
        z.copy_(torch.from_numpy(sla.solve_triangular(np.tril(Lb.numpy()), rhs.numpy(), lower=True)))

    def trtri_t(self, Lb, winv, out):
        out.copy_(torch.from_numpy(np.ascontiguousarray(np.linalg.inv(np.tril(Lb.numpy())).T)))

    def trace_block(self, theta_s, alpha, blk, row0, col0, acc, scratch):
        """sum over the block's elements with global row >= global column of w W dK/dlog theta_q
        (w = 1 below the diagonal, 1/2 on it), W = alpha alpha^T - K^-1; last slot: tr(W)."""
        from oracle.gp import _pairs
        n = len(self.X)
        rows, cols = blk.shape
        r1, c1 = min(n, row0 + rows), min(n, col0 + cols)
        if r1 <= row0 or c1 <= col0:
            return
        a = alpha.numpy()
        W = np.outer(a[row0:r1], a[col0:c1]) - blk.numpy()[:r1 - row0, :c1 - col0]
        gi = np.arange(row0, r1)[:, None]
        gj = np.arange(col0, c1)[None, :]
        w = np.where(gi > gj, 1.0, np.where(gi == gj, 0.5, 0.0))
        _, grads = _pairs(self.simil, np.asarray(theta_s), self.X[col0:c1], self.X[row0:r1], "theta")
        for q in range(self.simil.ntheta):
            if q in grads:
                acc[q] += float(np.sum(w * W * grads[q].T) * theta_s[q])
        acc[self.simil.ntheta] += float(np.sum(np.where(gi == gj, W, 0.0)))

    def noise_eval(self, theta_n):
        from oracle.gp import _noise
        nv, g = _noise(self.noise, np.asarray(theta_n), self.X[:1], "theta")
        return float(nv[0]), np.array([float(np.ravel(g.get(q, np.zeros(1)))[0]) * theta_n[q]
                                       for q in range(self.noise.ntheta)])

    def nsimil(self):
        return self.simil.ntheta

    def side(self, fn):
        fn()
        return None

    def wait_event(self, ev):
        pass

    def wait_side(self):
        pass

    def bad_pivot(self):
        return self.bad


def _problem():
    from tests import cases
    X, y, logt = cases.synth(NAME, N, seed=21)
    return X, y, logt


def _run(rank, world, grid, port, out, with_grad=False):
    sys.path.insert(0, ROOT)
    from gogp_b200.dist_chol import BlockCyclicCholesky
    from tests import cases
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    X, y, logt = _problem()
    _, _, _, osim, onoise = cases.CASES[NAME]
    be = NumpyBlocks(osim, onoise)
    be.set_inputs(X)
    ch = BlockCyclicCholesky(be, N, NB, rank, world, grid, dist)
    th = np.exp(logt)
    ch.build(th[:osim.ntheta], th[osim.ntheta:])
    ch.factor()
    lml = ch.solve_lml(y)
    res = np.array([ch.logdet(), lml, float(be.bad_pivot())])
    if with_grad:
        ch.invert()
        alpha = ch.solve_alpha().numpy()[:N].copy()
        ch.kinv()
        grad = ch.gradient(th[:osim.ntheta], th[osim.ntheta:])
        res = np.concatenate([res, grad, alpha])
    if out is not None:
        np.save(os.path.join(out, "r%d.npy" % rank), res)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return res


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle():
    from tests import cases
    X, y, logt = _problem()
    g = cases.make_oracle_gp(NAME)
    g.X, g.Y = X, y
    lml = g.observe(logt.copy())
    return 2.0 * float(np.sum(np.log(np.diag(g.L)))), lml


def test_default_grids():
    from gogp_b200.dist_chol import default_grid
    assert [default_grid(w) for w in (1, 2, 4, 8)] == [(1, 1), (2, 1), (2, 2), (4, 2)]


def test_single_rank_matches_oracle():
    logdet, lml = _oracle()
    res = _run(0, 1, (1, 1), 0, None)
    assert res[2] == 0
    assert abs(res[0] - logdet) < 1e-9 * abs(logdet)
    assert abs(res[1] - lml) < 1e-9 * max(abs(lml), N)


@pytest.mark.parametrize("world,grid", [(2, (2, 1)), (2, (1, 2)), (4, (2, 2))])
def test_gloo_ranks_match_oracle(tmp_path, world, grid):
    logdet, lml = _oracle()
    mp.spawn(_run, args=(world, grid, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = np.load(tmp_path / ("r%d.npy" % r))
        assert res[2] == 0
        assert abs(res[0] - logdet) < 1e-9 * abs(logdet), (r, res, logdet)
        assert abs(res[1] - lml) < 1e-9 * max(abs(lml), N), (r, res, lml)


def _oracle_grad():
    from tests import cases
    X, y, logt = _problem()
    g = cases.make_oracle_gp(NAME)
    g.X, g.Y = X, y
    g.observe(logt.copy())
    return g.gradient(), np.asarray(g.Alpha).ravel()


def _check_grad(res, grad, alpha):
    P = len(grad)
    got_g, got_a = res[3:3 + P], res[3 + P:]
    assert np.max(np.abs(got_g - grad)) <= 1e-7 * max(1.0, np.max(np.abs(grad))), (got_g, grad)
    assert np.max(np.abs(got_a - alpha)) <= 1e-7 * max(1.0, np.max(np.abs(alpha)))


def test_single_rank_gradient_matches_oracle():
    """distributed K^-1 + gradient, one rank: V = L^-T by block columns,
    alpha = V z, K^-1 = V V^T, fused trace per block."""
    grad, alpha = _oracle_grad()
    res = _run(0, 1, (1, 1), 0, None, True)
    _check_grad(res, grad, alpha)


@pytest.mark.parametrize("world,grid", [(2, (2, 1)), (2, (1, 2)), (4, (2, 2))])
def test_gloo_ranks_gradient_matches_oracle(tmp_path, world, grid):
    grad, alpha = _oracle_grad()
    mp.spawn(_run, args=(world, grid, _free_port(), str(tmp_path), True), nprocs=world, join=True)
    for r in range(world):
        _check_grad(np.load(tmp_path / ("r%d.npy" % r)), grad, alpha)

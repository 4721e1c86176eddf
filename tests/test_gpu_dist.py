"""The block-cyclic orchestration with the CUDA backend on one GPU (world = 1, several
block sizes) against the oracle; multi-rank runs on real GPUs are tools/dist_bench.py
--check (the same code under torchrun/NCCL), their CPU twin is tests/test_dist_chol.py."""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,N,NB", [("c5_matern4", 700, 128), ("c5_matern4", 1500, 512), ("hyperpriors", 1100, 256)])
def test_block_cyclic_single_rank_matches_oracle(name, N, NB):
    from gogp_b200.dist_chol import BlockCyclicCholesky, CudaBlocks
    ndim, ds, dn, _, _ = cases.CASES[name]
    X, y, logt = cases.synth(name, N, seed=31)
    og = cases.make_oracle_gp(name)
    og.X, og.Y = X, y
    ref = og.observe(logt.copy())
    be = CudaBlocks(ds, dn, ndim, 0)
    be.set_inputs(X, NB)
    ch = BlockCyclicCholesky(be, N, NB)
    th = np.exp(logt)
    nts = ds.NTheta()
    ch.build(th[:nts], th[nts:])
    ch.factor()
    lml = ch.solve_lml(y)
    assert be.bad_pivot() == 0
    assert abs(ch.logdet() - 2.0 * float(np.sum(np.log(np.diag(og.L))))) <= 1e-9 * N
    assert abs(lml - ref) <= 1e-9 * max(abs(ref), N), (lml, ref)
    be.close()


@pytest.mark.parametrize("name,N,NB", [("c5_matern4", 700, 128), ("c5_matern4", 1500, 512), ("hyperpriors", 1100, 256),
                                       ("c3_ard3", 900, 384)])
def test_block_cyclic_gradient_single_rank_matches_oracle(name, N, NB):
    """Distributed K^-1 + gradient path (invert / solve_alpha / kinv / per-block fused trace) with the
    CUDA backend on one rank, against the oracle's gradient and alpha."""
    from gogp_b200.dist_chol import BlockCyclicCholesky, CudaBlocks
    ndim, ds, dn, _, _ = cases.CASES[name]
    X, y, logt = cases.synth(name, N, seed=33)
    og = cases.make_oracle_gp(name)
    og.X, og.Y = X, y
    ref = og.observe(logt.copy())
    gref = og.gradient()
    be = CudaBlocks(ds, dn, ndim, 0)
    be.set_inputs(X, NB)
    ch = BlockCyclicCholesky(be, N, NB)
    th = np.exp(logt)
    nts = ds.NTheta()
    lml, grad = ch.lml_and_gradient(th[:nts], th[nts:], y)
    assert be.bad_pivot() == 0
    assert abs(lml - ref) <= 1e-9 * max(abs(ref), N)
    alpha = ch.alpha.cpu().numpy()[:N]
    assert np.max(np.abs(alpha - og.Alpha)) <= 1e-7 * max(1.0, np.max(np.abs(og.Alpha)))
    assert np.max(np.abs(grad - gref)) <= 1e-7 * max(1.0, np.max(np.abs(gref))), (grad, gref)
    be.close()

"""gp.GP across GPUs (gogp_grid_*: block-cyclic K, NCCL inside the library) on the B200 box, through the C-ABI:
one rank on one GPU against the oracle and against the single-GPU path; two ranks (host threads of this process,
one per GPU, NCCL between them) when the box shows at least two GPUs.  The orchestration's CPU twin is
tests/test_grid_host.py."""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu

LML_TOL, GRAD_TOL = 1e-9, 1e-7


def _ngpus():
    import torch
    return torch.cuda.device_count()


def _grid_eval(name, X, y, logt, devices, grid=(0, 0), block=128):
    import torch  # noqa: F401  (its libnccl.so.2 is the copy the library binds)
    from gogp_b200 import GridGP
    ndim, ds, dn, _, _ = cases.CASES[name]
    g = GridGP(NDim=ndim, Simil=ds, Noise=dn, Devices=devices, Grid=grid, Block=block)
    g.X, g.Y = X, y
    lml = g.Observe(logt.copy())
    grad = g.Gradient()
    alpha = g.Alpha()
    stats = g.Stats()
    g.close()
    return lml, grad, alpha, stats


@pytest.mark.parametrize("name,N,NB", [("c5_matern4", 700, 128), ("c5_matern4", 1500, 512), ("hyperpriors", 1100, 256),
                                       ("c3_ard3", 900, 384), ("c2_rbf", 2500, 128)])
def test_one_rank_matches_the_oracle(name, N, NB):
    X, y, logt = cases.synth(name, N, seed=33)
    og = cases.make_oracle_gp(name)
    og.X, og.Y = X, y
    ref = og.observe(logt.copy())
    gref = og.gradient()
    lml, grad, alpha, _ = _grid_eval(name, X, y, logt, [0], block=NB)
    assert abs(lml - ref) <= LML_TOL * max(abs(ref), N), (lml, ref)
    assert np.max(np.abs(alpha - og.Alpha)) <= GRAD_TOL * max(1.0, np.max(np.abs(og.Alpha)))
    assert np.max(np.abs(grad - gref)) <= GRAD_TOL * max(1.0, np.max(np.abs(gref))), (grad, gref)


def test_one_rank_default_block_matches_the_single_gpu_path():
    """block = 2048 (the shipped default: TMA GEMM with the tile mask, K = 2048) at N = 9000, against gp.GP on
    the same GPU; evaluated twice: a missing cross-stream dependency shows as a changing result."""
    from gogp_b200 import GP, GridGP
    name, N = "c5_matern4", 9000
    ndim, ds, dn, _, _ = cases.CASES[name]
    X, y, logt = cases.synth(name, N, seed=5)
    g1 = GP(NDim=ndim, Simil=ds, Noise=dn)
    g1.X, g1.Y = X, y
    ref = g1.Observe(logt.copy())
    gref = g1.Gradient()
    g1.close()
    g = GridGP(NDim=ndim, Simil=ds, Noise=dn, Devices=[0], Block=0)
    g.X, g.Y = X, y
    lml = g.Observe(logt.copy())
    grad = g.Gradient()
    lml2 = g.Observe(logt.copy())
    grad2 = g.Gradient()
    ms, _ = g.PhaseTimes()
    g.close()
    assert lml == lml2 and np.array_equal(grad, grad2)
    assert abs(lml - ref) <= LML_TOL * max(abs(ref), N), (lml, ref)
    assert np.max(np.abs(grad - gref)) <= GRAD_TOL * max(1.0, np.max(np.abs(gref))), (grad, gref)
    assert ms["factor"] > 0 and ms["sweep"] > 0


def test_not_positive_definite_and_bad_calls():
    from gogp_b200 import GridGP, GoGPPanic, _lib, kernel as k
    g = GridGP(NDim=1, Simil=k.Normal, Noise=k.ConstantNoise(0.0), Devices=[0], Block=128)
    X = np.linspace(0.0, 1.0, 300).reshape(-1, 1)
    X[150] = X[10]
    g.X, g.Y = X, np.sin(X[:, 0])
    with pytest.raises(GoGPPanic) as e:
        g.Observe(np.array([np.log(50.0)]))
    assert e.value.status == _lib.NOT_POSITIVE_DEFINITE
    with pytest.raises(GoGPPanic):
        g.Observe(np.zeros(3))       # len(x), gp/gp.go:398-400
    with pytest.raises(GoGPPanic):
        g.Gradient()                 # nothing observed
    # the handle is still usable after a failed evaluation
    g.Noise = k.ConstantNoise(0.0)
    g.close()
    with pytest.raises(GoGPPanic):
        GridGP(NDim=1, Simil=k.Normal, Noise=None, Devices=[0], Grid=(2, 1))   # 2 x 1 grid on one device


@pytest.mark.parametrize("grid", [(2, 1), (1, 2)])
def test_two_ranks_nccl_match_one_rank(grid):
    """Two GPUs, two host threads, NCCL broadcasts between them: same LML / alpha / gradient as one rank."""
    if _ngpus() < 2:
        pytest.skip("needs two GPUs")
    name, N, NB = "c5_matern4", 3000, 256
    X, y, logt = cases.synth(name, N, seed=7)
    lml1, grad1, alpha1, _ = _grid_eval(name, X, y, logt, [0], block=NB)
    lml2, grad2, alpha2, st = _grid_eval(name, X, y, logt, [0, 1], grid=grid, block=NB)
    assert st["world"] == 2 and st["collective_bytes_received"] > 0 and (st["pr"], st["pc"]) == grid
    assert abs(lml2 - lml1) <= 1e-11 * max(abs(lml1), N)
    assert np.max(np.abs(alpha2 - alpha1)) <= 1e-9 * max(1.0, np.max(np.abs(alpha1)))
    assert np.max(np.abs(grad2 - grad1)) <= 1e-9 * max(1.0, np.max(np.abs(grad1)))
    og = cases.make_oracle_gp(name)
    og.X, og.Y = X, y
    ref = og.observe(logt.copy())
    gref = og.gradient()
    assert abs(lml2 - ref) <= LML_TOL * max(abs(ref), N)
    assert np.max(np.abs(grad2 - gref)) <= GRAD_TOL * max(1.0, np.max(np.abs(gref)))

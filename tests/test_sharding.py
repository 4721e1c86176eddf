"""N > 1 path on CPU: two gloo ranks shard multi-start restarts (no data-path
collective) and agree with a single-rank run.  The evaluator here is the CPU
oracle -- this test checks the sharding/gather logic, not the kernels."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from tests.conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _evaluator():
    from tests import cases
    X, y, _ = cases.synth("hyperpriors", 40, seed=4)
    g = cases.make_oracle_gp("hyperpriors")
    g.X, g.Y = X, y

    def evaluate(theta):
        lml = g.observe(theta)
        return lml, g.gradient()
    return evaluate


def _starts():
    rng = np.random.default_rng(9)
    s = 0.3 * rng.standard_normal((7, 6))  # 7 restarts: not a multiple of the world size
    s[:, 4] += np.log(0.3)
    s[:, 5] += np.log(3.0)
    return s


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gogp_b200 import restarts
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lmls, thetas, best = restarts.multi_start(_evaluator(), _starts(), rank, world, iters=3, dist=dist)
    np.save(os.path.join(out, "r%d.npy" % rank), np.concatenate([lmls, thetas.ravel(), [best]]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_partitions_units():
    from gogp_b200 import restarts
    for n in (0, 1, 7, 64):
        for world in (1, 2, 8):
            owned = [restarts.shard(n, r, world) for r in range(world)]
            assert sorted(i for o in owned for i in o) == list(range(n))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def test_two_rank_gloo_matches_single_rank(tmp_path):
    from gogp_b200 import restarts
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "r0.npy")
    r1 = np.load(tmp_path / "r1.npy")
    assert np.array_equal(r0, r1)  # every rank holds the gathered result
    lmls, thetas, best = restarts.multi_start(_evaluator(), _starts(), 0, 1, iters=3)
    ref = np.concatenate([lmls, thetas.ravel(), [best]])
    assert np.allclose(r0, ref, rtol=0, atol=1e-12)
    assert np.all(np.isfinite(lmls))


def test_several_handles_per_gpu_deal_the_share_round_robin():
    """run_share_concurrent: this rank's restarts are dealt to one host thread per handle; every restart is
    optimised exactly once and lands in its own row."""
    from gogp_b200 import restarts
    starts = np.arange(22, dtype=np.float64).reshape(11, 2)
    seen = [[], [], []]

    def make(hh):
        def optimise(x0):
            seen[hh].append(int(x0[0]) // 2)
            return float(-x0.sum()), x0 + 1.0
        return optimise

    lmls, thetas = restarts.run_share_concurrent([make(0), make(1), make(2)], starts, rank=1, world=2)
    mine = restarts.shard(11, 1, 2)
    assert sorted(seen[0] + seen[1] + seen[2]) == mine and seen[0] == mine[0::3] and seen[1] == mine[1::3]
    for r in range(11):
        if r in mine:
            assert lmls[r] == -starts[r].sum() and np.array_equal(thetas[r], starts[r] + 1.0)
        else:
            assert lmls[r] == -np.inf and np.array_equal(thetas[r], starts[r])

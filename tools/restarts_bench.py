"""BASELINE configs[3]: the hyperpriors model (tutorial/hyperpriors), N = 16384, 64 multi-start
restarts sharded across the GPUs of one box -- independent units, no data-path collective, one
host-side gather of (objective, theta) per restart at the end.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
      --master-port 29512 tools/restarts_bench.py [--size 16384] [--restarts 64] [--alg adam] [--iters 5]

Every restart is the tutorial's MLE loop run INSIDE the library (gogp_optimize) on the rank's own
handle, with the tutorial's priors through the C callback.  Prints one JSON line from rank 0.
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gogp_b200 import GP, kernel as k, restarts  # noqa: E402
from gogp_b200.tutorial import HyperPriors  # noqa: E402


def synth(N, seed=0):
    """x = 0.39269908 i (the spacing of tutorial/data/hyperpriors.csv),
    y = trend + season (period 8) + 0.1 noise, normalised."""
    rng = np.random.default_rng(seed)
    x = 0.39269908 * np.arange(N)
    y = 0.002 * x + np.sin(2 * np.pi * x / 8.0) + 0.1 * rng.standard_normal(N)
    y = (y - y.mean()) / y.std(ddof=1)
    return x[:, None], y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=16384)
    ap.add_argument("--restarts", type=int, default=64)
    ap.add_argument("--alg", default="adam", choices=["adam", "lbfgs"])
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--handles", type=int, default=1,
                    help="device handles per GPU, one host thread each: one restart's tile chains overlap another's GEMMs")
    ap.add_argument("--threshold", type=float, default=1e-4,
                    help="gradient threshold (tutorial/tutorial.go:28 THRESHOLD = 1e-6 with ITERS = 1000)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")  # the only exchange is a host-side gather of R x (P+1) doubles
    X, y = synth(a.n)
    simil = k.Param(0) * k.Matern52.Of(l=2) + k.Param(1) * k.Periodic.Of(l=3, p=(4, 10.0))
    gps = [GP(NDim=1, Simil=simil, Noise=0.01 * k.UniformNoise, Device=local) for _ in range(max(1, a.handles))]
    for gp_ in gps:
        gp_.X, gp_.Y = X, y
    g = gps[0]
    priors = HyperPriors()
    rng = np.random.default_rng(7)
    starts = np.stack([priors.sample(rng) for _ in range(a.restarts)])
    starts[:, 2:4] = np.clip(starts[:, 2:4], -1.5, 1.5)   # keep the sampled length scales where K stays well conditioned
    evals = [0]
    stats = []

    def make_optimise(gp_):
        pri = HyperPriors()  # the priors object keeps the last point: one per handle

        def optimise(x0):
            x = np.ascontiguousarray(x0, dtype=np.float64)
            try:
                res = gp_.Optimize(x, alg=a.alg, iters=a.iters, threshold=a.threshold, rate=0.05, priors=pri)
            except Exception:  # a start where K is not positive definite: the restart is dropped
                return -np.inf, x0
            evals[0] += res["evals"]
            stats.append((res["iters"], res["evals"], res["grads"], bool(res["converged"])))
            return res["lml"], x
        return optimise

    optimisers = [make_optimise(gp_) for gp_ in gps]
    optimise = optimisers[0]

    for o in optimisers:
        o(starts[0].copy())  # warm-up (allocations); not counted
    evals[0] = 0
    del stats[:]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    import ctypes as C
    from gogp_b200 import _lib
    L, h, ms = _lib.lib(), g._handle(), C.c_double()
    L.gogp_timer_start(h)       # CUDA events on the handle's stream bracket this rank's share
    t0 = time.perf_counter()
    if len(optimisers) > 1:
        obj, thetas = restarts.run_share_concurrent(optimisers, starts, rank, world)
    else:
        obj, thetas = restarts.run_share(None, starts, rank, world, optimise=optimise)
    L.gogp_timer_stop(h, C.byref(ms))
    wall = time.perf_counter() - t0
    dt = ms.value * 1e-3 if len(optimisers) == 1 else wall  # several handles: their streams overlap, wall clock counts
    obj, thetas, best = restarts.gather_results(obj, thetas, rank, world, dist if world > 1 else None)
    tt = torch.tensor([dt, float(evals[0]), wall], dtype=torch.float64)
    if world > 1:
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        dt, total_evals, wall = float(tmax[0]), float(tt[1]), float(tmax[2])
    else:
        total_evals = float(tt[1])
    if rank == 0:
        print(json.dumps({
            "workload": "configs[3]: hyperpriors model, multi-start restarts sharded across GPUs (no collective)",
            "N": a.n, "restarts": a.restarts, "n_gpus": world, "handles_per_gpu": len(optimisers), "alg": a.alg, "iters": a.iters, "threshold": a.threshold,
            "rank0_restarts": [{"iters": s_[0], "evals": s_[1], "grads": s_[2], "converged": s_[3]} for s_ in stats],
            "seconds": dt, "timing": "device time of each rank's share (CUDA events on its handle's stream), max over ranks",
            "wall_seconds": wall, "restarts_per_s": a.restarts / dt, "evaluations": total_evals,
            "evals_per_s_total": total_evals / dt, "best_restart": best, "best_objective": float(obj[best]),
            "best_theta": [float(v) for v in np.exp(thetas[best])],
            "finite_restarts": int(np.sum(np.isfinite(obj))),
        }), flush=True)
    for gp_ in gps:
        gp_.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Block-cyclic Cholesky across the GPUs of one box (BASELINE configs[4]).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
      --master-port 29511 tools/dist_bench.py --size 131072 --block 2048 [--check]

Prints one JSON line from rank 0: build / factor / solve device times (CUDA events on
the compute stream, max over ranks), Cholesky TFLOP/s (N^3/3 over all GPUs) and, with
--check (N <= 32768), the difference to the single-GPU path of the same library.
"""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gogp_b200 import GP, kernel as k  # noqa: E402
from gogp_b200.dist_chol import BlockCyclicCholesky, CudaBlocks, default_grid  # noqa: E402

NDIM = 4


def kernel_c5():
    e = k.Param(0)
    for d in range(NDIM):
        e = e * k.Matern32.Of(l=1 + d, dim=d)
    return e, k.UniformNoise


def synth(N, seed=0):
    """x ~ U(0,8)^4, y = sum sin(x_d) + 0.1 N(0,1) normalised, sigma = 0.1."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(0.0, 8.0, size=(N, NDIM))
    y = np.sin(X).sum(axis=1) + 0.1 * rng.standard_normal(N)
    y = (y - y.mean()) / y.std(ddof=1)
    logt = np.zeros(NDIM + 2)
    logt[NDIM + 1] = math.log(0.1)
    return X, y, logt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=131072)
    ap.add_argument("--block", dest="nb", type=int, default=2048)
    ap.add_argument("--pr", type=int, default=0)
    ap.add_argument("--pc", type=int, default=0)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--no-lookahead", action="store_true")
    ap.add_argument("--grad", action="store_true", help="also K^-1 and the gradient (SURVEY.md section 8 f-2)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    grid = (a.pr, a.pc) if a.pr and a.pc else default_grid(world)
    simil, noise = kernel_c5()
    X, y, logt = synth(a.n)
    th = np.exp(logt)
    be = CudaBlocks(simil, noise, NDIM, local)
    be.set_inputs(X, a.nb)
    ch = BlockCyclicCholesky(be, a.n, a.nb, rank, world, grid, dist if world > 1 else None)

    def timed(fn):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    res = []
    # the main stream carries the latency-critical chain (diagonal factor, panel solve, broadcasts):
    # give it priority over the side stream that runs the bulk of the trailing update
    torch.cuda.set_stream(torch.cuda.Stream(priority=-1))
    for rep in range(a.reps + 1):  # first pass warms up (allocations, NCCL channels)
        l0 = be.launches()
        t_build, _ = timed(lambda: ch.build(th[:NDIM + 1], th[NDIM + 1:]))
        t_fac, _ = timed(lambda: ch.factor(lookahead=not a.no_lookahead))
        t_sol, lml = timed(lambda: ch.solve_lml(y))
        gr = None
        if a.grad:
            t_inv, _ = timed(ch.invert)
            t_alpha, _ = timed(ch.solve_alpha)
            t_kinv, _ = timed(ch.kinv)
            t_trace, grad = timed(lambda: ch.gradient(th[:NDIM + 1], th[NDIM + 1:]))
            gr = (t_inv, t_alpha, t_kinv, t_trace, grad)
        res.append((t_build, t_fac, t_sol, lml, be.launches() - l0, gr))
    t_build, t_fac, t_sol, lml, launches, gr = res[-1]
    bad = be.bad_pivot()
    out = {
        "workload": "configs[4]: synthetic 4-D Matern32 + noise, block-cyclic Cholesky", "N": a.n, "NB": a.nb,
        "n_gpus": world, "grid": list(grid), "lookahead": not a.no_lookahead, "build_ms": t_build, "factor_ms": t_fac, "solve_ms": t_sol,
        "cholesky_tflops_total": a.n ** 3 / 3 / (t_fac * 1e-3) / 1e12,
        "cholesky_tflops_per_gpu": a.n ** 3 / 3 / (t_fac * 1e-3) / 1e12 / world,
        "lml": lml, "bad_pivot": bad, "launches_rank0": launches,
        "mem_gb_rank0": torch.cuda.max_memory_allocated() / 1e9,
    }
    if gr is not None:
        t_inv, t_alpha, t_kinv, t_trace, grad = gr
        total = t_build + t_fac + t_sol + t_inv + t_alpha + t_kinv + t_trace
        out.update({"invert_ms": t_inv, "alpha_ms": t_alpha, "kinv_ms": t_kinv, "trace_ms": t_trace,
                    "potri_tflops_total": 2 * a.n ** 3 / 3 / ((t_inv + t_kinv) * 1e-3) / 1e12,
                    "eval_ms": total, "evals_per_s": 1e3 / total,
                    "eval_tflops_total": a.n ** 3 / (total * 1e-3) / 1e12, "grad": [float(v) for v in grad]})
    if a.check and rank == 0:
        del ch
        torch.cuda.empty_cache()
        g = GP(NDim=NDIM, Simil=simil, Noise=noise, Device=local)
        g.X, g.Y = X, y
        ref = g.Observe(logt.copy())
        out["single_gpu_lml"] = ref
        out["rel_diff"] = abs(lml - ref) / max(abs(ref), a.n)
        if gr is not None:
            gref = g.Gradient()
            out["single_gpu_grad"] = [float(v) for v in gref]
            out["grad_rel_diff"] = float(np.max(np.abs(np.asarray(gr[4]) - gref)) / max(1.0, np.max(np.abs(gref))))
        g.close()
    if rank == 0:
        print(json.dumps(out), flush=True)
    be.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

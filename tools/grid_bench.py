"""LML + gradient across the GPUs of one box through gogp_grid_* (BASELINE configs[4]: synthetic 4-D Matern32 +
noise, 2D block-cyclic K, NCCL inside the library).

  one process, one host thread per GPU (what a Go host does):
      python tools/grid_bench.py --size 131072 --gpus 8 [--pr 4 --pc 2] [--block 2048] [--check]
  SPMD, one process per GPU:
      python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
          tools/grid_bench.py --size 131072

Prints one JSON line (rank 0): per-phase device times (max over ranks), TFLOP/s of the factorisation (N^3/3), of
the fused V / K^-1 sweep (2N^3/3) and of the whole evaluation (N^3), NCCL bytes received, and with --check
(N <= 40000) the difference to the single-GPU path.
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

NDIM = 4


def kernel_c5():
    from gogp_b200 import kernel as k
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


def run(n, block, pr, pc, gpus, reps, check, lml_only=False):
    import torch
    from gogp_b200 import GP, GridGP
    from gogp_b200 import grid as G
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    simil, noise = kernel_c5()
    X, y, logt = synth(n)
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(G.unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        g = GridGP(NDim=NDIM, Simil=simil, Noise=noise, Grid=(pr, pc), Block=block, Rank=rank, World=world,
                   Device=local, UniqueId=bytes(idt.cpu().tolist()))
        ngpu = world
    else:
        g = GridGP(NDim=NDIM, Simil=simil, Noise=noise, Devices=list(range(gpus)), Grid=(pr, pc), Block=block)
        ngpu = gpus
    g.X, g.Y = X, y
    res = []
    for rep in range(reps + 1):  # the first pass warms up (allocations, NCCL channels)
        th = logt + (0.01 * rep)
        t0 = time.perf_counter()
        lml = g.Observe(th.copy())
        grad = None if lml_only else g.Gradient()
        wall = time.perf_counter() - t0
        ms, cm = g.PhaseTimes()
        res.append((lml, grad, ms, cm, wall, g.Stats()["eval_ms"]))
    lml, grad, ms, cm, wall, total = res[-1]
    if world > 1:
        t = torch.tensor([ms[p] for p in ms] + [wall, total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = t.cpu().tolist()
        ms = dict(zip(ms.keys(), vals[:-2]))
        wall, total = vals[-2], vals[-1]
    st = g.Stats()
    out = {
        "workload": "configs[4]: synthetic 4-D Matern32 + noise, LML + gradient, 2D block-cyclic over %d GPU(s)" % ngpu,
        "N": n, "block": st["block"], "n_gpus": ngpu, "grid": [st["pr"], st["pc"]],
        "mode": "spmd (one process per GPU)" if world > 1 else "one process, one host thread per GPU",
        "phases_ms": {k: round(v, 3) for k, v in ms.items()}, "comm_ms_on_priority_stream": {k: round(v, 3) for k, v in cm.items()},
        "eval_ms": total, "eval_ms_note": "slowest rank's own sum of phases (device time); phases_ms are per-phase maxima over ranks",
        "wall_ms": wall * 1e3, "evals_per_s": 1e3 / total,
        "cholesky_tflops_total": n ** 3 / 3 / (ms["factor"] * 1e-3) / 1e12,
        "cholesky_tflops_per_gpu": n ** 3 / 3 / (ms["factor"] * 1e-3) / 1e12 / ngpu,
        "lml": lml, "collective_bytes_received_rank0": st["collective_bytes_received"], "of_which_peer_copy_engine": st["peer_copy_bytes_received"], "device_gb_rank0": st["device_bytes"] / 1e9,
        "nccl_version": st["nccl_version"], "launches_rank0": st["launches"],
    }
    if grad is not None:
        out.update({
            "sweep_tflops_total": 2 * n ** 3 / 3 / (ms["sweep"] * 1e-3) / 1e12,
            "sweep_tflops_per_gpu": 2 * n ** 3 / 3 / (ms["sweep"] * 1e-3) / 1e12 / ngpu,
            "eval_tflops_total": n ** 3 / (total * 1e-3) / 1e12, "eval_tflops_per_gpu": n ** 3 / (total * 1e-3) / 1e12 / ngpu,
            "grad": [float(v) for v in grad]})
    g.close()
    if check and rank == 0:
        g1 = GP(NDim=NDIM, Simil=simil, Noise=noise, Device=local)
        g1.X, g1.Y = X, y
        th = logt + 0.01 * reps
        ref = g1.Observe(th.copy())
        out["single_gpu_lml"] = ref
        out["lml_rel_diff"] = abs(lml - ref) / max(abs(ref), n)
        if grad is not None:
            gref = g1.Gradient()
            out["grad_rel_diff"] = float(np.max(np.abs(np.asarray(grad) - gref)) / max(1.0, np.max(np.abs(gref))))
        out["single_gpu_phases_ms"] = {k: round(float(v), 3) for k, v in g1.PhaseTimes().items()}
        g1.close()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=131072)
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--pr", type=int, default=0)
    ap.add_argument("--pc", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--lml-only", action="store_true")
    a = ap.parse_args()
    run(a.n, a.block, a.pr, a.pc, a.gpus, a.reps, a.check, a.lml_only)

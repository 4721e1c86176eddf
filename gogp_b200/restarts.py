"""Multi-start restart sharding (BASELINE configs[3]).

Independent hyper-parameter restarts are separate optimisations over the same
(X, Y): they shard across GPUs with NO data-path collective -- one process per
GPU, one device handle per process; the only exchange is a host-side gather of
(LML, theta) per restart at the end.  `evaluate` is any callable
theta -> (lml, grad); on the product path it is gp.GP.Observe + Gradient.
"""
import numpy as np


def shard(n_restarts, rank, world):
    """Restart indices owned by `rank`: round-robin, so every rank gets
    floor or ceil of n/world units."""
    return list(range(rank, n_restarts, world))


def adam_ascent(evaluate, theta0, iters, rate=0.01, beta1=0.9, beta2=0.999, eps=1e-8, threshold=1e-6):
    """The tutorial's Adam loop (tutorial/tutorial.go:156-168: RATE 0.01, stop when
    every |g| < THRESHOLD), maximising the log marginal likelihood."""
    theta = np.array(theta0, dtype=np.float64)
    m = np.zeros_like(theta)
    v = np.zeros_like(theta)
    lml = None
    for t in range(1, iters + 1):
        lml, g = evaluate(theta.copy())
        if np.all(np.abs(g) < threshold):
            break
        m = beta1 * m + (1 - beta1) * g
        v = beta2 * v + (1 - beta2) * g * g
        theta = theta + rate * (m / (1 - beta1 ** t)) / (np.sqrt(v / (1 - beta2 ** t)) + eps)
    lml, _ = evaluate(theta.copy())
    return lml, theta


def run_share(evaluate, starts, rank=0, world=1, iters=0, optimise=None):
    """This rank's share of `starts` (R x P array, identical on every rank): returns
    (lmls[R], thetas[R, P]) with -inf / the start itself in the rows other ranks own.
    `optimise`, if given, is start -> (objective, theta) and replaces the host-driven Adam
    loop -- on the product path gp.GP.Optimize, i.e. the loop inside the library."""
    starts = np.asarray(starts, dtype=np.float64)
    R, P = starts.shape
    lmls = np.full(R, -np.inf)
    thetas = np.array(starts, copy=True)
    for r in shard(R, rank, world):
        if optimise is not None:
            lmls[r], thetas[r] = optimise(starts[r].copy())
        elif iters > 0:
            lmls[r], thetas[r] = adam_ascent(evaluate, starts[r], iters)
        else:
            lmls[r], _ = evaluate(starts[r].copy())
    return lmls, thetas


def run_share_concurrent(optimisers, starts, rank=0, world=1):
    """run_share with several device handles on ONE GPU ("optionally 2 handles per GPU to
    overlap one restart's panel latency with another's GEMM"): `optimisers` holds one start -> (objective, theta)
    callable per handle; this rank's restarts are dealt round-robin to one host thread per handle.  A handle is not
    thread-safe, distinct handles are (include/gogp_b200.h); the C calls release the interpreter lock."""
    import threading
    starts = np.asarray(starts, dtype=np.float64)
    R, P = starts.shape
    lmls = np.full(R, -np.inf)
    thetas = np.array(starts, copy=True)
    mine = shard(R, rank, world)
    H = len(optimisers)

    def work(hh):
        for r in mine[hh::H]:
            lmls[r], thetas[r] = optimisers[hh](starts[r].copy())

    threads = [threading.Thread(target=work, args=(hh,)) for hh in range(H)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return lmls, thetas


def gather_results(lmls, thetas, rank=0, world=1, dist=None):
    """The only exchange of the sharded restarts: every rank receives every restart's
    (objective, theta), R*(P+1) doubles per rank.  Returns (lmls, thetas, best index)."""
    R = len(lmls)
    if world > 1:
        import torch
        mine = torch.from_numpy(np.concatenate([lmls[:, None], thetas], axis=1))
        if dist.get_backend() == "nccl":  # NCCL moves device memory only: stage the (tiny) payload on this rank's GPU
            mine = mine.cuda()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        for w, part in enumerate(parts):
            idx = shard(R, w, world)
            a = part.cpu().numpy()
            lmls[idx] = a[idx, 0]
            thetas[idx] = a[idx, 1:]
    return lmls, thetas, int(np.argmax(lmls))


def multi_start(evaluate, starts, rank=0, world=1, iters=0, dist=None, optimise=None):
    """Run this rank's share of `starts` and gather all results.  Returns (lmls[R],
    thetas[R, P], best index); every rank gets the same answer.  `dist` is torch.distributed
    (gloo: host tensors; nccl: the payload is staged on the current CUDA device) or None."""
    lmls, thetas = run_share(evaluate, starts, rank, world, iters, optimise)
    return gather_results(lmls, thetas, rank, world, dist)

"""GEMM throughput at the shapes the block-cyclic trailing update uses (M x N x K, C -= A B^T)."""
import sys

import torch

sys.path.insert(0, ".")
from gogp_b200 import kernel as k
from gogp_b200.dist_chol import CudaBlocks

be = CudaBlocks(k.Normal, None, 1, 0)
shapes = [(2048, 32768, 2048), (2048, 65536, 2048), (2048, 8192, 2048), (4096, 32768, 4096), (16384, 16384, 2048),
          (32768, 2048, 2048), (8192, 8192, 8192), (16384, 128, 128)]
for (m, n, kk) in shapes:
    ldc = 131072 if n <= 65536 else n
    Cm = torch.zeros(m, ldc, dtype=torch.float64, device="cuda")[:, :n]
    A = torch.rand(m, kk, dtype=torch.float64, device="cuda") - 0.5
    B = torch.rand(n, kk, dtype=torch.float64, device="cuda") - 0.5
    be.gemm(Cm, A, B, -1e-6, 1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        be.gemm(Cm, A, B, -1e-6, 1.0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("M=%d N=%d K=%d: %.3f ms  %.2f TFLOP/s" % (m, n, kk, ms, 2.0 * m * n * kk / ms / 1e9), flush=True)
    del Cm, A, B
be.close()

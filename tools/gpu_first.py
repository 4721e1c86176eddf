"""First contact with the GPU: FP64 peaks, GEMM throughput, phase timings."""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from gogp_b200 import _lib
from tests import cases

L = _lib.lib()
res = {}
g = cases.make_device_gp("c2_rbf")
h = g._handle()
out = C.c_double()
for which, nm in ((0, "dmma"), (1, "dfma")):
    st = L.gogp_debug_fp64_peak(h, which, C.byref(out))
    res["fp64_%s_tflops" % nm] = out.value
    print(nm, st, out.value, flush=True)
for n, k, mode in ((8192, 8192, 0), (8192, 8192, 1), (16384, 16384, 1), (16384, 256, 0), (16384, 128, 0), (2048, 2048, 1)):
    st = L.gogp_debug_gemm(h, n, k, mode, 3, C.byref(out))
    res["gemm_n%d_k%d_m%d" % (n, k, mode)] = out.value
    print("gemm", n, k, mode, st, out.value, flush=True)
for name, N in (("c2_rbf", 4096), ("c3_ard8", 8192), ("c3_ard8", 16384), ("c3_ard8", 32768)):
    if len(sys.argv) > 1 and N > int(sys.argv[1]):
        continue
    X, y, logt = cases.synth(name, N, seed=0)
    dg = cases.make_device_gp(name)
    dg.X, dg.Y = X, y
    for rep in range(2):
        t0 = time.time()
        lml = dg.Observe(logt.copy())
        t1 = time.time()
        gr = dg.Gradient()
        t2 = time.time()
        ph = dg.PhaseTimes()
        print(name, N, rep, "lml", lml, "observe %.1f ms gradient %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3),
              {k_: round(v, 2) for k_, v in ph.items()}, "launches", dg.Launches(), flush=True)
    res["%s_%d" % (name, N)] = dict(lml=lml, grad=gr.tolist(), phases=ph, observe_ms=(t1 - t0) * 1e3,
                                   gradient_ms=(t2 - t1) * 1e3)
    dg.close()
json.dump(res, open("gpurun_out/first.json", "w"), indent=1)

"""GEMM throughput probe: python tools/gemm_bench.py n k mode iters"""
import ctypes as C
import sys

sys.path.insert(0, ".")
from gogp_b200 import _lib
from gogp_b200 import kernel as k

L = _lib.lib()
h = C.c_void_p()
sd = k.Normal.Descriptor()
assert L.gogp_create(1, sd, len(sd), 1, None, 0, 0, 0, C.byref(h)) == 0
n, kk, mode, iters = (int(v) for v in sys.argv[1:5])
out = C.c_double()
st = L.gogp_debug_gemm(h, n, kk, mode, iters, C.byref(out))
print("gemm n=%d k=%d mode=%d iters=%d status=%d tflops=%.3f" % (n, kk, mode, iters, st, out.value))
L.gogp_destroy(h)

import ctypes as C, sys
sys.path.insert(0, ".")
from gogp_b200 import _lib, kernel as k
L = _lib.lib(); h = C.c_void_p(); sd = k.Normal.Descriptor()
assert L.gogp_create(1, sd, len(sd), 1, None, 0, 0, 0, C.byref(h)) == 0
out = C.c_double()
for v in (0, 1, 2):
    st = L.gogp_debug_leaf(h, v, 20, C.byref(out))
    print("leaf variant", v, "status", st, "%.1f us" % out.value, flush=True)
L.gogp_destroy(h)

"""ctypes binding of the C-ABI (include/gogp_b200.h).

This is the Python counterpart of the cgo stub in INTEGRATION.md: the same
entry points, the same plain pointers and sizes.  There is no fallback: if the
shared library is missing the import fails, and if no CUDA device is present
``gogp_create`` reports GOGP_CUDA_ERROR.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgogp_b200.so")

# gogp_status
OK, BAD_ARGUMENT, NOT_POSITIVE_DEFINITE, ILL_CONDITIONED, CUDA_ERROR, NCCL_ERROR, OUT_OF_MEMORY, NOT_READY, \
    UNSUPPORTED = range(9)

# gogp_op_kind
OP_CONST, OP_PARAM, OP_ADD, OP_MUL, OP_NORMAL, OP_PERIODIC, OP_MATERN32, OP_MATERN52, OP_MATERN52_TEXTBOOK, \
    OP_EVENTS = range(10)

PHASES = ("upload", "build", "potrf", "solve", "potri", "trace", "predict")

# every symbol include/gogp_b200.h declares
SYMBOLS = (
    "gogp_create", "gogp_destroy", "gogp_set_events", "gogp_set_data", "gogp_observe", "gogp_gradient", "gogp_absorb", "gogp_extend", "gogp_lml",
    "gogp_produce", "gogp_optimize", "gogp_get_alpha", "gogp_get_factor", "gogp_set_state", "gogp_last_error", "gogp_status_string",
    "gogp_phase_times", "gogp_launch_count", "gogp_debug_fetch", "gogp_debug_build", "gogp_debug_fp64_peak",
    "gogp_debug_gemm", "gogp_debug_leaf", "gogp_debug_leaf_run", "gogp_dev_set_inputs", "gogp_dev_cov_block", "gogp_dev_potrf", "gogp_dev_trsm",
    "gogp_dev_gemm", "gogp_dev_gemm_bc", "gogp_dev_reserve", "gogp_dev_sumlogdiag", "gogp_dev_gemv_sub", "gogp_dev_trsv", "gogp_dev_trtri_t", "gogp_dev_trace_block", "gogp_dev_trace_local", "gogp_noise_eval", "gogp_timer_start", "gogp_timer_stop", "gogp_profile_enable", "gogp_profile_read",
    "gogp_create_grid", "gogp_grid_unique_id", "gogp_grid_create_rank", "gogp_grid_destroy", "gogp_grid_set_data",
    "gogp_grid_observe", "gogp_grid_gradient", "gogp_grid_absorb", "gogp_grid_lml", "gogp_grid_get_alpha",
    "gogp_grid_phase_times", "gogp_grid_stats", "gogp_grid_last_error",
)

GRID_PHASES = ("build", "factor", "solve", "sweep", "alpha", "trace")
GRID_ID_BYTES = 128


class OptSettings(C.Structure):
    """gogp_opt_settings"""
    _fields_ = [("method", C.c_int), ("max_iters", C.c_int), ("threshold", C.c_double), ("rate", C.c_double),
                ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double), ("history", C.c_int)]


class OptResult(C.Structure):
    """gogp_opt_result"""
    _fields_ = [("iters", C.c_int), ("evals", C.c_int), ("lml0", C.c_double), ("lml", C.c_double),
                ("converged", C.c_int), ("grads", C.c_int)]


PRIOR_FN = C.CFUNCTYPE(C.c_double, C.c_void_p, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double))


class Op(C.Structure):
    """gogp_op"""
    _fields_ = [
        ("kind", C.c_uint8),
        ("dim", C.c_uint8),
        ("param", C.c_int16 * 2),
        ("scale", C.c_double * 2),
        ("constant", C.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "gogp_b200: %s is missing -- build the CUDA extension first "
            "(python -m gogp_b200.build); there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    H = C.c_void_p
    L.gogp_create.argtypes = [C.c_int, C.POINTER(Op), C.c_int, C.c_int, C.POINTER(Op), C.c_int, C.c_int, C.c_int,
                              C.POINTER(H)]
    L.gogp_create.restype = C.c_int
    L.gogp_destroy.argtypes = [H]
    L.gogp_destroy.restype = None
    L.gogp_set_events.argtypes = [H, dp, C.c_int]
    L.gogp_set_events.restype = C.c_int
    L.gogp_set_data.argtypes = [H, dp, dp, C.c_int64]
    L.gogp_set_data.restype = C.c_int
    L.gogp_observe.argtypes = [H, dp, C.c_int, dp, dp, C.c_int64, dp]
    L.gogp_observe.restype = C.c_int
    L.gogp_gradient.argtypes = [H, dp, C.c_int64]
    L.gogp_gradient.restype = C.c_int
    L.gogp_absorb.argtypes = [H, dp, dp, dp, dp, C.c_int64]
    L.gogp_absorb.restype = C.c_int
    L.gogp_extend.argtypes = [H, dp, dp, C.c_int64, dp]
    L.gogp_extend.restype = C.c_int
    L.gogp_lml.argtypes = [H, dp]
    L.gogp_lml.restype = C.c_int
    L.gogp_produce.argtypes = [H, dp, C.c_int64, dp, dp]
    L.gogp_produce.restype = C.c_int
    L.gogp_optimize.argtypes = [H, C.POINTER(OptSettings), dp, PRIOR_FN, C.c_void_p, C.POINTER(OptResult)]
    L.gogp_optimize.restype = C.c_int
    L.gogp_get_alpha.argtypes = [H, dp, C.c_int64]
    L.gogp_get_alpha.restype = C.c_int
    L.gogp_get_factor.argtypes = [H, dp, C.c_int64]
    L.gogp_get_factor.restype = C.c_int
    L.gogp_set_state.argtypes = [H, dp, dp, dp, C.c_int64, dp, dp]
    L.gogp_set_state.restype = C.c_int
    L.gogp_last_error.argtypes = [H]
    L.gogp_last_error.restype = C.c_char_p
    L.gogp_status_string.argtypes = [C.c_int]
    L.gogp_status_string.restype = C.c_char_p
    L.gogp_phase_times.argtypes = [H, dp]
    L.gogp_phase_times.restype = C.c_int
    L.gogp_launch_count.argtypes = [H]
    L.gogp_launch_count.restype = C.c_int64
    L.gogp_debug_fetch.argtypes = [H, C.c_int, dp, C.c_int64]
    L.gogp_debug_fetch.restype = C.c_int
    L.gogp_debug_build.argtypes = [H, dp, dp, dp, C.c_int64, dp]
    L.gogp_debug_build.restype = C.c_int
    L.gogp_debug_fp64_peak.argtypes = [H, C.c_int, dp]
    L.gogp_debug_fp64_peak.restype = C.c_int
    L.gogp_debug_gemm.argtypes = [H, C.c_int64, C.c_int64, C.c_int, C.c_int, dp]
    L.gogp_debug_gemm.restype = C.c_int
    vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double  # device pointers travel as integers
    L.gogp_dev_set_inputs.argtypes = [H, dp, i64, i64]
    L.gogp_dev_cov_block.argtypes = [H, dp, dp, i64, i64, i64, i64, C.c_int, vp, i64, vp]
    L.gogp_dev_potrf.argtypes = [H, vp, i64, i64, vp, vp, C.c_int, vp]
    L.gogp_dev_trsm.argtypes = [H, vp, i64, i64, vp, i64, i64, vp, vp]
    L.gogp_dev_gemm.argtypes = [H, vp, i64, vp, i64, vp, i64, i64, i64, i64, dbl, dbl, C.c_int, vp]
    L.gogp_dev_sumlogdiag.argtypes = [H, vp, i64, i64, vp, vp]
    L.gogp_dev_gemv_sub.argtypes = [H, vp, i64, i64, i64, vp, vp, vp, vp]
    L.gogp_dev_trsv.argtypes = [H, vp, i64, vp, vp, vp, i64, vp]
    L.gogp_dev_trtri_t.argtypes = [H, vp, i64, i64, vp, vp, vp]
    L.gogp_dev_trace_block.argtypes = [H, dp, vp, vp, i64, i64, i64, i64, i64, vp, vp, vp]
    L.gogp_dev_trace_local.argtypes = [H, dp, vp, vp, i64, i64, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp,
                                       vp]
    L.gogp_dev_trace_local.restype = C.c_int
    L.gogp_noise_eval.argtypes = [H, dp, dp, dp]
    for f in (L.gogp_dev_trtri_t, L.gogp_dev_trace_block, L.gogp_noise_eval):
        f.restype = C.c_int
    for f in (L.gogp_dev_set_inputs, L.gogp_dev_cov_block, L.gogp_dev_potrf, L.gogp_dev_trsm, L.gogp_dev_gemm,
              L.gogp_dev_sumlogdiag, L.gogp_dev_gemv_sub, L.gogp_dev_trsv):
        f.restype = C.c_int
    L.gogp_dev_gemm_bc.argtypes = [H, vp, i64, vp, i64, vp, i64, i64, i64, i64, dbl, dbl, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, vp]
    L.gogp_dev_gemm_bc.restype = C.c_int
    L.gogp_dev_reserve.argtypes = [H, i64]
    L.gogp_dev_reserve.restype = C.c_int
    # gp.GP across the GPUs of one box (2D block-cyclic, NCCL inside the library)
    G = C.c_void_p
    idp = C.POINTER(C.c_ubyte)
    L.gogp_create_grid.argtypes = [C.c_int, C.POINTER(Op), C.c_int, C.c_int, C.POINTER(Op), C.c_int, C.c_int,
                                   C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, i64, C.POINTER(G)]
    L.gogp_grid_unique_id.argtypes = [idp]
    L.gogp_grid_create_rank.argtypes = [C.c_int, C.POINTER(Op), C.c_int, C.c_int, C.POINTER(Op), C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i64, idp, C.POINTER(G)]
    L.gogp_grid_destroy.argtypes = [G]
    L.gogp_grid_destroy.restype = None
    L.gogp_grid_set_data.argtypes = [G, dp, dp, i64]
    L.gogp_grid_observe.argtypes = [G, dp, dp]
    L.gogp_grid_gradient.argtypes = [G, dp, i64]
    L.gogp_grid_absorb.argtypes = [G, dp, dp]
    L.gogp_grid_lml.argtypes = [G, dp]
    L.gogp_grid_get_alpha.argtypes = [G, dp, i64]
    L.gogp_grid_phase_times.argtypes = [G, dp, dp]
    L.gogp_grid_stats.argtypes = [G, dp]
    L.gogp_grid_last_error.argtypes = [G]
    L.gogp_grid_last_error.restype = C.c_char_p
    for f in (L.gogp_create_grid, L.gogp_grid_unique_id, L.gogp_grid_create_rank, L.gogp_grid_set_data,
              L.gogp_grid_observe, L.gogp_grid_gradient, L.gogp_grid_absorb, L.gogp_grid_lml, L.gogp_grid_get_alpha,
              L.gogp_grid_phase_times, L.gogp_grid_stats):
        f.restype = C.c_int
    L.gogp_debug_leaf.argtypes = [H, C.c_int, C.c_int, dp]
    L.gogp_debug_leaf.restype = C.c_int
    L.gogp_debug_leaf_run.argtypes = [H, C.c_int, dp, dp, dp, C.POINTER(C.c_int)]
    L.gogp_debug_leaf_run.restype = C.c_int
    L.gogp_timer_start.argtypes = [H]
    L.gogp_timer_start.restype = C.c_int
    L.gogp_timer_stop.argtypes = [H, dp]
    L.gogp_timer_stop.restype = C.c_int
    L.gogp_profile_enable.argtypes = [H, C.c_int]
    L.gogp_profile_enable.restype = C.c_int
    L.gogp_profile_read.argtypes = [H, dp, dp, C.POINTER(C.c_int64)]
    L.gogp_profile_read.restype = C.c_int
    _lib = L
    return L


def dptr(a):
    """float64 C-contiguous numpy array (or None) -> double*"""
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(C.c_double))

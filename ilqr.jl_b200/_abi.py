"""ctypes declarations of include/ilqr_b200.h (the C ABI a Julia host reaches with ccall)."""
import ctypes
import os

from . import _build

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)

ABI_VERSION = 5
MAX_N, MAX_M = 16, 8
MODEL_TWO_LINK, MODEL_SERIAL_CHAIN, MODEL_FLOATING_CHAIN, MODEL_CUSTOM = 1, 2, 3, 4
MAX_JOINTS, CHAIN_STRIDE = 8, 20
VARIANT_AUTO, VARIANT_LANE_PER_TRAJ, VARIANT_WARP_PER_TRAJ = 0, 1, 2

STATUS_NAN_GAINS, STATUS_NAN_ROLLOUT, STATUS_LS_EXHAUSTED = 1, 2, 4
STATUS_NOT_DECREASED, STATUS_CONVERGED, STATUS_MAX_ITER = 8, 16, 32

(X, U, XBAR, UBAR, DUFF, K, NEW_COST, PREV_COST, ALPHA, DU2, COST_TRACE, ALPHA_TRACE, DU2_TRACE, STATUS, ITERS,
 ACTIVE) = range(16)


class Problem(ctypes.Structure):
    """ilqr_problem (include/ilqr_b200.h)."""
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("model_id", ctypes.c_int32),
        ("n", ctypes.c_int32), ("m", ctypes.c_int32), ("H", ctypes.c_int32), ("B", ctypes.c_int32),
        ("n_alpha", ctypes.c_int32), ("trace_iters", ctypes.c_int32),
        ("device", ctypes.c_int32), ("variant", ctypes.c_int32),
        ("dt", ctypes.c_double), ("reg", ctypes.c_double),
        ("model_params", ctypes.c_double * 32),
        ("x_target", ctypes.c_double * MAX_N), ("w_x", ctypes.c_double * MAX_N),
        ("w_u", ctypes.c_double * MAX_M), ("w_xf", ctypes.c_double * MAX_N),
        ("nq", ctypes.c_int32), ("custom_cost", ctypes.c_int32), ("gravity", ctypes.c_double * 3),
        ("chain", ctypes.c_double * ((MAX_JOINTS + 1) * CHAIN_STRIDE)),
        ("custom_src", ctypes.c_char_p),
    ]


# every symbol include/ilqr_b200.h declares: name -> (restype, argtypes)
_H = ctypes.c_void_p
SYMBOLS = {
    "ilqr_abi_version": (ctypes.c_int32, []),
    "ilqr_problem_two_link": (ctypes.c_int32, [ctypes.POINTER(Problem), ctypes.c_int32, ctypes.c_int32]),
    "ilqr_problem_serial_chain": (ctypes.c_int32, [ctypes.POINTER(Problem), ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_int32, ctypes.c_int32]),
    "ilqr_problem_floating_chain": (ctypes.c_int32, [ctypes.POINTER(Problem), ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                                    ctypes.c_int32, ctypes.c_int32]),
    "ilqr_problem_custom": (ctypes.c_int32, [ctypes.POINTER(Problem), ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                            ctypes.c_double, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int32]),
    "ilqr_custom_compile_check": (ctypes.c_int32, [ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32]),
    "ilqr_create": (ctypes.c_int32, [ctypes.POINTER(Problem), ctypes.POINTER(_H)]),
    "ilqr_destroy": (ctypes.c_int32, [_H]),
    "ilqr_last_error": (ctypes.c_char_p, [_H]),
    "ilqr_upload": (ctypes.c_int32, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "ilqr_upload_device": (ctypes.c_int32, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "ilqr_upload_x0": (ctypes.c_int32, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "ilqr_upload_gains": (ctypes.c_int32, [_H, ctypes.c_void_p, ctypes.c_void_p]),
    "ilqr_backward_pass": (ctypes.c_int32, [_H]),
    "ilqr_forward_pass": (ctypes.c_int32, [_H, ctypes.c_void_p]),
    "ilqr_commit": (ctypes.c_int32, [_H, ctypes.c_double, c_int32_p]),
    "ilqr_set_reg": (ctypes.c_int32, [_H, ctypes.c_double]),
    "ilqr_set_active": (ctypes.c_int32, [_H, ctypes.c_void_p]),
    "ilqr_iterate": (ctypes.c_int32, [_H, ctypes.c_double, c_int32_p]),
    "ilqr_fit": (ctypes.c_int32, [_H, ctypes.c_int32, ctypes.c_double, c_int32_p]),
    "ilqr_download": (ctypes.c_int32, [_H, ctypes.c_int32, ctypes.c_void_p]),
    "ilqr_download_device": (ctypes.c_int32, [_H, ctypes.c_int32, ctypes.c_void_p]),
    "ilqr_solve": (ctypes.c_int32, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                    ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p]),
    "ilqr_stream_solve_device": (ctypes.c_int32, [_H, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                                 ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]),
    "ilqr_mpc_start": (ctypes.c_int32, [_H, ctypes.c_void_p, ctypes.c_void_p]),
    "ilqr_mpc_step": (ctypes.c_int32, [_H, ctypes.c_int32, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]),
    "ilqr_pool_create": (ctypes.c_int32, [ctypes.POINTER(Problem), ctypes.c_int32, ctypes.POINTER(_H)]),
    "ilqr_pool_destroy": (ctypes.c_int32, [_H]),
    "ilqr_pool_last_error": (ctypes.c_char_p, [_H]),
    "ilqr_pool_submit": (ctypes.c_int64, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                          ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p]),
    "ilqr_pool_submit_device": (ctypes.c_int64, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                                 ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.c_void_p]),
    "ilqr_pool_wait": (ctypes.c_int32, [_H, ctypes.c_int64]),
    "ilqr_pool_wait_all": (ctypes.c_int32, [_H]),
    "ilqr_pool_launch_count": (ctypes.c_int64, [_H]),
    "ilqr_streamer_create": (ctypes.c_int32, [ctypes.POINTER(Problem), ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                              ctypes.c_double, ctypes.POINTER(_H)]),
    "ilqr_streamer_destroy": (ctypes.c_int32, [_H]),
    "ilqr_streamer_last_error": (ctypes.c_char_p, [_H]),
    "ilqr_streamer_submit": (ctypes.c_int64, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "ilqr_streamer_submit_device": (ctypes.c_int64, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "ilqr_streamer_submit_traj": (ctypes.c_int64, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "ilqr_streamer_submit_traj_device": (ctypes.c_int64, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "ilqr_streamer_submit_x0": (ctypes.c_int64, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "ilqr_streamer_submit_x0_device": (ctypes.c_int64, [_H, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "ilqr_streamer_wait": (ctypes.c_int32, [_H, ctypes.c_int64]),
    "ilqr_streamer_wait_all": (ctypes.c_int32, [_H]),
    "ilqr_streamer_launch_count": (ctypes.c_int64, [_H]),
    "ilqr_streamer_rounds": (ctypes.c_int64, [_H]),
    "ilqr_streamer_profile": (ctypes.c_int32, [_H, c_double_p]),
    "ilqr_host_alloc": (ctypes.c_int32, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint64]),
    "ilqr_host_free": (ctypes.c_int32, [ctypes.c_void_p]),
    "ilqr_launch_count": (ctypes.c_int64, [_H]),
    "ilqr_last_kernel_ms": (ctypes.c_int32, [_H, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
    "ilqr_profile": (ctypes.c_int32, [_H, c_double_p]),
    "ilqr_stream_profile": (ctypes.c_int32, [_H, c_double_p]),
    "ilqr_set_variant": (ctypes.c_int32, [_H, ctypes.c_int32]),
    "ilqr_set_tuning": (ctypes.c_int32, [_H, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]),
    "ilqr_sync": (ctypes.c_int32, [_H]),
    "ilqr_stream": (ctypes.c_void_p, [_H]),
}

_lib = None


def load_library(rebuild=False):
    """dlopen libilqr_b200.so (building it first if stale).  Raises if it cannot be built/loaded:
    there is deliberately no fallback implementation."""
    global _lib
    if _lib is None or rebuild:
        path = os.environ.get("ILQR_LIB") or _build.build(force=rebuild)   # ILQR_LIB: a pre-built variant (experiments)
        lib = ctypes.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)   # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.ilqr_abi_version() != ABI_VERSION:
            raise RuntimeError("libilqr_b200.so ABI version mismatch")
        _lib = lib
    return _lib

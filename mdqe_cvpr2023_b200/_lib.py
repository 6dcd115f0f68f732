"""ctypes binding of the C ABI declared in include/msda_b200.h.

There is deliberately no fallback: if libmsda_b200.so is missing or a call fails, a RuntimeError with
the library's own message is raised (the reference only printf'd kernel launch errors,
/root/reference/mdqe/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:948-952).
"""
import ctypes
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# MSDA_B200_LIB: experiment builds only (tools/whatif_bench.py); the product library is always the in-tree one
LIB_PATH = os.environ.get("MSDA_B200_LIB") or os.path.join(PKG_DIR, "libmsda_b200.so")

MSDA_F32, MSDA_BF16, MSDA_F64, MSDA_BF16_LOC32, MSDA_F16 = 0, 1, 2, 3, 4
ABI_VERSION = 6
BWD_ACC_ZEROED = 1

_c_int, _c_vp, _c_i64, _c_sz = ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t
_SEVEN = [_c_int] * 7

# name -> (restype, argtypes); mirrors include/msda_b200.h one to one (tests/test_abi.py checks it)
PROTOTYPES = {
    "msda_abi_version": (_c_int, []),
    "msda_last_error": (ctypes.c_char_p, []),
    "msda_set_option": (_c_int, [ctypes.c_char_p, _c_int]),
    "msda_get_option": (_c_int, [ctypes.c_char_p, ctypes.POINTER(_c_int)]),
    "msda_forward": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp] + _SEVEN + [_c_vp]),
    "msda_backward_workspace_bytes": (_c_sz, [_c_int] * 5),
    "msda_backward": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp] + _SEVEN
                      + [_c_vp, _c_vp, _c_vp, _c_vp, _c_sz]),
    "msda_forward_grouped": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp] + [_c_int] * 8 + [ctypes.c_float, _c_vp]),
    "msda_backward_grouped": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp] + [_c_int] * 8
                              + [ctypes.c_float, _c_vp, _c_vp, _c_vp, _c_vp, _c_sz]),
    "msda_zero_fill": (_c_int, [_c_vp, _c_vp, _c_sz]),
    "msda_backward_grouped_flags": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp] + [_c_int] * 8
                                    + [ctypes.c_float, _c_vp, _c_vp, _c_vp, _c_vp, _c_sz, _c_int]),
    "msda_fused_backward_flags": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_int, ctypes.c_float,
                                           _c_vp] + [_c_int] * 8 + [ctypes.c_float, _c_vp, _c_vp, _c_vp, _c_int]),
    "msda_fused_forward": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_int, ctypes.c_float]
                           + [_c_int] * 8 + [ctypes.c_float, _c_vp]),
    "msda_fused_backward": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_int, ctypes.c_float,
                                     _c_vp] + [_c_int] * 8 + [ctypes.c_float, _c_vp, _c_vp, _c_vp]),
    "msda_fused_forward_joint": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_int, _c_vp, _c_int, ctypes.c_float]
                                 + [_c_int] * 8 + [ctypes.c_float, _c_vp]),
    "msda_fused_backward_joint": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_int, _c_vp, _c_int, ctypes.c_float,
                                           _c_vp] + [_c_int] * 8 + [ctypes.c_float, _c_vp, _c_vp, _c_int]),
    "tc_linear_forward_packed": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_int, _c_vp]),
    "msda_fused_forward_packed_joint": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_int, _c_vp, _c_int, ctypes.c_float]
                                        + [_c_int] * 8 + [_c_vp]),
    "mask_logits_forward": (_c_int, [_c_vp, _c_int, _c_int, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_i64, _c_vp]),
    "mask_logits_backward": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_i64, _c_vp, _c_vp]),
    "tc_linear_forward": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_int, _c_int, _c_vp]),
    "tc_linear_backward": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_int, _c_int, _c_vp, _c_vp]),
    "tc_linear_bias_grad": (_c_int, [_c_vp, _c_vp, _c_i64, _c_int, _c_vp]),
    "tc_linear_backward_bias": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_int, _c_int, _c_vp, _c_vp, _c_vp]),
    "msda_forward_host": (_c_int, [_c_int, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp] + _SEVEN + [_c_vp]),
    "msda_backward_host": (_c_int, [_c_int, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp] + _SEVEN
                           + [_c_vp, _c_vp, _c_vp]),
    "mask_logits_forward_host": (_c_int, [_c_int, _c_int, _c_int, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_i64, _c_vp]),
    "msda_host_sync": (_c_int, []),
    "msda_host_fence": (_c_int, [ctypes.POINTER(_c_i64)]),
    "msda_host_wait": (_c_int, [_c_i64]),
    "msda_forward_host_saved": (_c_int, [_c_int, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp] + [_c_int] * 8
                                + [ctypes.c_float, _c_vp, ctypes.POINTER(_c_i64)]),
    "msda_backward_host_saved": (_c_int, [_c_i64, _c_vp, _c_vp, _c_vp, _c_vp]),
    "mask_logits_forward_host_saved": (_c_int, [_c_int, _c_int, _c_int, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_i64, _c_vp,
                                                ctypes.POINTER(_c_i64)]),
    "mask_logits_backward_host_saved": (_c_int, [_c_i64, _c_vp, _c_vp, _c_vp]),
    "msda_host_saved_release": (_c_int, [_c_i64]),
    "msda_host_arena_release": (_c_int, []),
    "mask_match_cost_workspace_bytes": (ctypes.c_size_t, []),
    "mask_match_cost": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_i64, _c_vp, _c_vp, _c_vp]),
    "mask_losses_workspace_bytes": (ctypes.c_size_t, []),
    "mask_losses_forward": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_i64, ctypes.c_float, _c_vp, _c_vp, _c_vp]),
    "mask_losses_backward": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_i64, ctypes.c_float,
                                      _c_vp, _c_vp]),
    "mask_nms_siou_workspace_bytes": (ctypes.c_size_t, []),
    "mask_nms_siou": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "mask_track_siou": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    "aligned_bilinear_sigmoid": (_c_int, [_c_vp, _c_vp, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_vp]),
    "query_init_sample_forward": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp]),
    "query_init_sample_backward": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int,
                                            _c_vp, _c_vp]),
    "msda_packed_value_bytes": (_c_sz, [_c_int] * 4),
    "msda_pack_value": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_vp] + [_c_int] * 5 + [_c_vp]),
    "msda_forward_packed": (_c_int, [_c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp] + _SEVEN + [_c_vp]),
    "msda_allreduce_max_ranks": (_c_int, []),
    "msda_allreduce_flag_bytes": (_c_sz, [_c_int]),
    "msda_allreduce_f32": (_c_int, [_c_vp, _c_int, _c_int, _c_int, ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint64,
                                    ctypes.POINTER(ctypes.c_uint64), _c_vp, _c_i64, _c_i64, ctypes.c_float, _c_int]),
    "msda_profile_read": (_c_int, [_c_int, _c_i64, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_c_i64)]),
    "msda_debug_read": (_c_int, [ctypes.POINTER(ctypes.c_longlong)]),
    "msda_launch_count": (_c_i64, []),
    "msda_launch_count_reset": (None, []),
    "msda_gemm_flag_timeouts": (_c_i64, []),
}

_lib = None


def load():
    """Load (once) and return the ctypes handle; raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m mdqe_cvpr2023_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback for this operator.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    got = lib.msda_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"libmsda_b200.so ABI version {got} != expected {ABI_VERSION}; rebuild it")
    _lib = lib
    return lib


def last_error():
    msg = load().msda_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed (status {rc}): {last_error()}")


def set_option(key, value):
    check(load().msda_set_option(key.encode(), int(value)), f"msda_set_option({key})")


def get_option(key):
    v = ctypes.c_int(0)
    check(load().msda_get_option(key.encode(), ctypes.byref(v)), f"msda_get_option({key})")
    return v.value


def launch_count():
    return int(load().msda_launch_count())


def launch_count_reset():
    load().msda_launch_count_reset()


PROF_MSDA_FWD, PROF_MSDA_BWD, PROF_MASK_FWD, PROF_MASK_BWD = 0, 1, 2, 3


def profile_read(kind, min_units=0):
    """-> (total_ms, count) of the profiled launches of `kind` since the last read."""
    tot, n = ctypes.c_double(0.0), _c_i64(0)
    check(load().msda_profile_read(kind, min_units, ctypes.byref(tot), ctypes.byref(n)), "msda_profile_read")
    return tot.value, n.value

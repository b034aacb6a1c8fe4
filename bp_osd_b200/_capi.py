"""ctypes binding of libbposd_b200.so (the C ABI in include/bposd_b200.h).

The library is built in-tree by :mod:`bp_osd_b200.build`; loading fails loudly if it is
missing -- there is no CPU fallback behind this module.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BPOSD_LIB") or os.path.join(_HERE, "libbposd_b200.so")  # BPOSD_LIB: A/B builds

BP_PRODUCT_SUM, BP_MINIMUM_SUM = 0, 1
OSD_0, OSD_E, OSD_CS, OSD_OFF = 0, 1, 2, 3
OK, EINVAL, ECUDA, ENOMEM, EUNSUP = 0, -1, -2, -3, -4

P = C.c_void_p


class Out(C.Structure):
    _fields_ = [("d_osdw", P), ("d_osd0", P), ("d_bp", P), ("d_llr", P), ("d_converge", P), ("d_iter", P)]


class CssSector(C.Structure):
    _fields_ = [("d_fail_x", P), ("d_fail_z", P), ("d_weight_x", P), ("d_weight_z", P)]


class Info(C.Structure):
    _fields_ = [(k, C.c_int32) for k in (
        "m", "n", "nnz", "rank", "k", "max_iter", "bp_method", "osd_method", "osd_order", "precision",
        "device", "bp_kernel", "bp_threads", "bp_ctas_per_sm", "bp_smem_bytes", "osd_threads",
        "osd_smem_bytes", "sm_count", "osd_variant", "bp_layout_excess", "bp_cluster_size", "reserved")] + [("ms_scaling_factor", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("shots", C.c_int64), ("bp_converged", C.c_int64), ("osd_invocations", C.c_int64),
                ("bp_iterations", C.c_int64), ("ms_bp", C.c_float), ("ms_osd", C.c_float),
                ("launches", C.c_int32), ("chunks", C.c_int32)]


# name -> (restype, argtypes); also the list tests check against include/bposd_b200.h
SIGNATURES = {
    "bposd_create": (C.c_int, [P, P, C.c_int32, C.c_int32, P, C.c_int32, C.c_int32, C.c_double, C.c_int32,
                               C.c_int32, C.c_int32, C.c_int32, C.POINTER(P)]),
    "bposd_update_channel_probs": (C.c_int, [P, P]),
    "bposd_decode_batch": (C.c_int, [P, P, C.c_int64, C.POINTER(Out), P, P, P]),
    "bposd_decode_host": (C.c_int, [P, P, C.c_int64, P, P, P, P, P, P]),
    "bposd_decode_batch_packed": (C.c_int, [P, P, C.c_int64, C.POINTER(Out), P, P, P]),
    "bposd_decode_host_packed": (C.c_int, [P, P, C.c_int64, P, P, P, P, P, P]),
    "bposd_sample_syndromes_packed": (C.c_int, [P, C.c_uint64, C.c_uint64, C.c_int64, C.c_int32, P, P, P]),
    "bposd_set_channel_thresholds": (C.c_int, [P, P, P, P]),
    "bposd_sample_syndromes": (C.c_int, [P, C.c_uint64, C.c_uint64, C.c_int64, C.c_int32, P, P, P]),
    "bposd_set_logicals": (C.c_int, [P, P, P, C.c_int32]),
    "bposd_logical_check": (C.c_int, [P, P, P, C.c_int64, P, P, P, P, P]),
    "bposd_channel_update": (C.c_int, [P, P, C.c_int64, P, P, P, P, P]),
    "bposd_css_counters": (C.c_int, [P, C.c_int64, P, P, P, P, P, P, P]),
    "bposd_sample_and_decode": (C.c_int, [P, C.c_uint64, C.c_uint64, C.c_int64, C.c_int32, P, P]),
    "bposd_get_info": (C.c_int, [P, C.POINTER(Info)]),
    "bposd_get_stats": (C.c_int, [P, C.POINTER(Stats)]),
    "bposd_set_tuning": (C.c_int, [P, C.c_int32, C.c_int32, C.c_int64]),
    "bposd_int32_peak": (C.c_int, [P, C.POINTER(C.c_double)]),
    "bposd_smem_peak": (C.c_int, [P, C.POINTER(C.c_double)]),
    "bposd_fp64_peak": (C.c_int, [P, C.POINTER(C.c_double)]),
    "bposd_math_probe": (C.c_int, [P, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "bposd_set_schedule": (C.c_int, [P, C.c_int32, C.c_void_p]),
    "bposd_set_cluster_size": (C.c_int, [P, C.c_int32]),
    "bposd_set_osd_variant": (C.c_int, [P, C.c_int32, C.c_int64]),
    "bposd_last_error": (C.c_char_p, [P]),
    "bposd_version": (C.c_char_p, []),
    "bposd_destroy": (None, [P]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m bp_osd_b200.build` "
                "(nvcc, sm_100a).  bp_osd_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class BposdError(RuntimeError):
    pass


def check(handle, rc):
    if rc == OK:
        return
    msg = load().bposd_last_error(handle)
    msg = msg.decode() if msg else "unknown error"
    if rc == EINVAL:
        raise ValueError(msg)
    if rc == ENOMEM:
        raise MemoryError(msg)
    if rc == EUNSUP:
        raise NotImplementedError(msg)
    raise BposdError(msg)

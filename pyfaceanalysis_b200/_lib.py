"""ctypes binding of libhgsfa.so (the C ABI declared in ``include/hgsfa.h``).

There is no CPU fallback: if the library is missing or a call fails, the caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# HGSFA_LIB: a differently built libhgsfa.so (kernel experiments: tools/build_variant.py); default = the in-tree build
LIB_PATH = os.environ.get("HGSFA_LIB") or os.path.join(HERE, "libhgsfa.so")

U8, F32, F64 = 0, 1, 2
ROWMAJOR, TILED = 0, 1
TILE = 128
NEAREST, BILINEAR, BICUBIC = 0, 2, 3

_DTYPES = {np.dtype(np.uint8): U8, np.dtype(np.float32): F32, np.dtype(np.float64): F64}

# every symbol include/hgsfa.h declares: (name, restype, argtypes)
_i64, _int, _vp, _sz = C.c_int64, C.c_int, C.c_void_p, C.c_size_t
_pi64, _pd, _pint = C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int)
SYMBOLS = [
    ("hgsfa_last_error", C.c_char_p, []),
    ("hgsfa_version", _int, []),
    ("hgsfa_device_count", _int, [_pint]),
    ("hgsfa_plan_create", _int, [_vp, _sz, _int, C.POINTER(_vp)]),
    ("hgsfa_plan_destroy", _int, [_vp]),
    ("hgsfa_plan_info", _int, [_vp, _pi64, _pi64, _pi64]),
    ("hgsfa_plan_flops", _int, [_vp, _i64, _int, _pd, _pd, _pd]),
    ("hgsfa_plan_execute", _int, [_vp, _vp, _int, _i64, _i64, _vp, _int, _i64, _vp]),
    ("hgsfa_plan_execute_device", _int, [_vp, _vp, _int, _int, _i64, _i64, _vp, _int, _i64, _vp]),
    ("hgsfa_plan_stats", _int, [_vp, _pi64, _pd]),
    ("hgsfa_plan_profile", _int, [_vp, _int]),
    ("hgsfa_plan_op_stats", _int, [_vp, _i64, _pd, _vp, _pd, _pd]),
    ("hgsfa_plan_set_chunks", _int, [_vp, _i64, _i64]),
    ("hgsfa_crop_extent", _int, [_vp, _int, _int, _vp, _vp, _i64, _int, _int, _int, _vp, _int, _int, _vp]),
    ("hgsfa_crop_extent_device", _int, [_vp, _int, _int, _vp, _vp, _i64, _int, _int, _int, _vp, _int, _int, _vp]),
    ("hgsfa_crop_extent_batch_device", _int, [_vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _int, _vp, _int, _int, _vp]),
    ("hgsfa_contrast_avg_std_device", _int, [_vp, _i64, _i64, C.c_double, C.c_double, _vp]),
    ("hgsfa_age_crop_device", _int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _int, _int, _vp, _vp]),
    ("hgsfa_tile_windows_device", _int, [_vp, _int, _i64, _i64, _i64, _vp, _int, _vp]),
    ("hgsfa_gauss_create", _int, [_vp, _vp, _vp, _vp, _int, _int, _int, C.POINTER(_vp)]),
    ("hgsfa_gauss_destroy", _int, [_vp]),
    ("hgsfa_gauss_regress", _int, [_vp, _vp, _int, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("hgsfa_gauss_regress_device", _int, [_vp, _vp, _int, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("hgsfa_cascade_update_device", _int, [_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    ("hgsfa_compact_index_device", _int, [_vp, _i64, _vp, _vp, _vp, _i64, _vp]),
    ("hgsfa_gather_rows_device", _int, [_vp, _vp, _vp, _i64, _i64, _vp]),
]

_lib = None


class HgsfaError(RuntimeError):
    pass


def load():
    """Load libhgsfa.so, binding every declared symbol.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HgsfaError(
            "libhgsfa.so is not built (%s). Run `python -m pyfaceanalysis_b200.build` "
            "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().hgsfa_last_error()
        raise HgsfaError(msg.decode("utf-8", "replace") if msg else "libhgsfa call failed (rc=%d)" % rc)


def dtype_code(dt):
    try:
        return _DTYPES[np.dtype(dt)]
    except KeyError:
        raise TypeError("unsupported dtype %r (uint8, float32, float64 only)" % (dt,))


def ptr(a):
    """void* of a numpy array (or None)."""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def device_count():
    n = C.c_int(0)
    check(load().hgsfa_device_count(C.byref(n)))
    return n.value

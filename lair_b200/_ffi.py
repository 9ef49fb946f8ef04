"""ctypes binding of the C ABI in include/lair_b200.h (the same symbols the Rust shim binds).

There is no CPU fallback: if liblair_b200.so is missing or no sm_100 device is usable, every
compute call raises.  Nothing in this package imports `oracle/`.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liblair_b200.so")

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_ALLOC, ERR_UNSUPPORTED, ERR_NCCL = range(7)


class LairB200Error(RuntimeError):
    """A non-zero lair_b200_status from the native library."""

    def __init__(self, status: int, message: str):
        super().__init__(f"lair_b200 status {status}: {message}")
        self.status = status


_lib = None

i64, i32p, vp, cint = ctypes.c_int64, ctypes.POINTER(ctypes.c_int32), ctypes.c_void_p, ctypes.c_int

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/lair_b200.h
_SIGNATURES = {
    "lair_b200_version": [],
    "lair_b200_last_error": [],
    "lair_b200_device_count": [ctypes.POINTER(cint)],
    "lair_b200_init": [cint],
    "lair_b200_shutdown": [],
    "lair_b200_set_option": [ctypes.c_char_p, i64],
    "lair_b200_get_option": [ctypes.c_char_p, ctypes.POINTER(i64)],
    "lair_b200_launch_count": [],
    "lair_b200_check_fault": [vp],
    "lair_b200_debug_panel_timing": [ctypes.POINTER(ctypes.c_longlong), cint],
    "lair_b200_mg_unique_id": [vp],
    "lair_b200_mg_init": [cint, cint, vp],
    "lair_b200_mg_finalize": [],
    "lair_b200_mg_timeline": [cint],
    "lair_b200_mg_timeline_read": [vp, i64, ctypes.POINTER(i64)],
    "lair_b200_dgetrf_mg_dev": [i64, i64, vp, i64, vp, vp, vp],
    "lair_b200_sgetrf_mg_dev": [i64, i64, vp, i64, vp, vp, vp],
    "lair_b200_profile_begin": [],
    "lair_b200_profile_end": [],
    "lair_b200_profile_get": [ctypes.c_char_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64),
                              ctypes.POINTER(ctypes.c_double)],
}
_SIGNATURES.update({
    # device-resident factors (lu::Factorized behind a handle)
    "lair_b200_lu_solve": [vp, i64, vp, i64, i64, vp, i64, i64],
    "lair_b200_lu_pivots": [vp, vp],
    "lair_b200_lu_factors": [vp, vp, i64, i64],
    "lair_b200_lu_view": [vp, cint, vp, i64, i64],
    "lair_b200_lu_shape": [vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(cint), ctypes.POINTER(i64)],
    "lair_b200_lu_destroy": [vp],
})
for _p in "sdcz":
    _SIGNATURES[f"lair_b200_{_p}lu_factor"] = [i64, i64, vp, i64, i64, ctypes.POINTER(vp), ctypes.POINTER(i64)]
    _SIGNATURES[f"lair_b200_{_p}getrf"] = [i64, i64, vp, i64, i64, vp, vp]
    _SIGNATURES[f"lair_b200_{_p}getrs"] = [i64, i64, vp, i64, i64, vp, vp, i64, i64, vp, i64, i64]
for _p in "sdcz":
    _SIGNATURES[f"lair_b200_{_p}geqrf"] = [i64, i64, vp, i64, i64, vp]
    _SIGNATURES[f"lair_b200_{_p}qr_q"] = [i64, i64, vp, i64, i64, vp, vp, i64, i64]
    _SIGNATURES[f"lair_b200_{_p}geqrf_dev"] = [i64, i64, vp, i64, vp, vp]
for _p in "cz":
    _SIGNATURES[f"lair_b200_{_p}getrf_dev"] = [i64, i64, vp, i64, vp, vp, vp]
for _p in "sd":
    _SIGNATURES[f"lair_b200_{_p}gesv"] = [i64, i64, vp, i64, i64, vp, i64, i64, vp, i64, i64, vp]
    _SIGNATURES[f"lair_b200_{_p}getrf_batched"] = [i64, i64, vp, vp, vp]
    _SIGNATURES[f"lair_b200_{_p}getrf_batched_mg"] = [i64, i64, vp, vp, vp, cint]
    _SIGNATURES[f"lair_b200_{_p}getrf_dev"] = [i64, i64, vp, i64, vp, vp, vp]
    _SIGNATURES[f"lair_b200_{_p}getrs_dev"] = [i64, i64, vp, i64, vp, vp, i64, vp]
    _SIGNATURES[f"lair_b200_{_p}getrf_batched_dev"] = [i64, i64, vp, vp, vp, vp]
    _SIGNATURES[f"lair_b200_{_p}laswp_dev"] = [i64, vp, i64, i64, i64, vp, vp]
    _SIGNATURES[f"lair_b200_{_p}trsm_dev"] = [i64, i64, vp, i64, vp, i64, vp]
    _SIGNATURES[f"lair_b200_{_p}gemm_minus_dev"] = [i64, i64, i64, vp, i64, vp, i64, vp, i64, vp]
_RESTYPES = {"lair_b200_last_error": ctypes.c_char_p, "lair_b200_launch_count": i64}

EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))


def lib() -> ctypes.CDLL:
    """Load liblair_b200.so (built in-tree by lair_b200/build.py).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LairB200Error(ERR_NO_DEVICE, f"{LIB_PATH} not found: build it with `python -m lair_b200.build` "
                                               "(there is no CPU fallback)")
        l = ctypes.CDLL(LIB_PATH)
        for name, args in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, cint)
        _lib = l
    return _lib


def check(status: int) -> None:
    if status != OK:
        msg = lib().lair_b200_last_error()
        raise LairB200Error(status, msg.decode() if msg else "")


def set_option(name: str, value: int) -> None:
    check(lib().lair_b200_set_option(name.encode(), int(value)))


def get_option(name: str) -> int:
    v = i64(0)
    check(lib().lair_b200_get_option(name.encode(), ctypes.byref(v)))
    return int(v.value)


def launch_count() -> int:
    return int(lib().lair_b200_launch_count())


def check_fault(stream=None) -> None:
    """Wait for `stream` and raise if a device-side cross-CTA wait timed out since the last check
    (the results of the asynchronous _dev calls issued before are then invalid)."""
    check(lib().lair_b200_check_fault(ctypes.c_void_p(stream or 0)))


def profile_begin() -> None:
    check(lib().lair_b200_profile_begin())


def profile_end() -> dict:
    """Stop live kernel timing; returns {family: {"ms", "launches", "work"}}."""
    check(lib().lair_b200_profile_end())
    out = {}
    for fam in ("gemm", "panel", "laswp", "trsm", "batched", "small"):
        ms, n, w = ctypes.c_double(0), i64(0), ctypes.c_double(0)
        check(lib().lair_b200_profile_get(fam.encode(), ctypes.byref(ms), ctypes.byref(n), ctypes.byref(w)))
        out[fam] = {"ms": ms.value, "launches": int(n.value), "work": w.value}
    return out


def device_count() -> int:
    c = cint(0)
    lib().lair_b200_device_count(ctypes.byref(c))
    return int(c.value)

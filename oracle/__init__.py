"""TEST INFRASTRUCTURE ONLY -- ctypes face of the CPU oracle (oracle/lair_oracle.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product package (lair_b200/) never does.

Parity pin: the reference (pure Rust) cannot be compiled in this image; the oracle is
pinned against the reference's own golden vectors in tests/golden/lair_golden.json.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblair_oracle.so")
_lib = None

_PREFIX = {
    np.dtype(np.float32): "s",
    np.dtype(np.float64): "d",
    np.dtype(np.complex64): "c",
    np.dtype(np.complex128): "z",
}


def build(force: bool = False) -> str:
    """Compile liblair_oracle.so with the committed Makefile (g++ -O2 -ffp-contract=off)."""
    src_newer = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("lair_oracle.hpp", "lair_oracle_c.cpp", "Makefile")
    )
    if force or src_newer:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        i64, vp, dbl = ctypes.c_int64, ctypes.c_void_p, ctypes.c_double
        for p in "sdcz":
            getattr(_lib, f"oracle_{p}getrf").restype = i64
            getattr(_lib, f"oracle_{p}getrf").argtypes = [i64, i64, vp, i64, i64, vp]
            getattr(_lib, f"oracle_{p}getrf_variant").restype = i64
            getattr(_lib, f"oracle_{p}getrf_variant").argtypes = [ctypes.c_int, i64, i64, vp, i64, i64, vp]
            getattr(_lib, f"oracle_{p}getrf_recursive").restype = i64
            getattr(_lib, f"oracle_{p}getrf_recursive").argtypes = [i64, i64, vp, i64, i64, vp]
            getattr(_lib, f"oracle_{p}getrs").restype = None
            getattr(_lib, f"oracle_{p}getrs").argtypes = [i64, vp, i64, i64, vp, vp, i64, vp]
            getattr(_lib, f"oracle_{p}laswp").restype = None
            getattr(_lib, f"oracle_{p}laswp").argtypes = [i64, vp, i64, i64, i64, vp, i64]
            getattr(_lib, f"oracle_{p}trsm").restype = None
            getattr(_lib, f"oracle_{p}trsm").argtypes = [vp, i64, i64, vp, i64, i64, i64, i64]
            getattr(_lib, f"oracle_{p}gemm_minus").restype = None
            getattr(_lib, f"oracle_{p}gemm_minus").argtypes = [vp, i64, i64, i64, i64, vp, i64, i64, i64, vp, i64, i64]
            getattr(_lib, f"oracle_{p}into_pl").restype = None
            getattr(_lib, f"oracle_{p}into_pl").argtypes = [i64, i64, vp, i64, i64, vp, i64]
            getattr(_lib, f"oracle_{p}geqrf").restype = None
            getattr(_lib, f"oracle_{p}geqrf").argtypes = [i64, i64, vp, i64, i64, vp]
            getattr(_lib, f"oracle_{p}qr_q").restype = None
            getattr(_lib, f"oracle_{p}qr_q").argtypes = [i64, i64, vp, i64, i64, vp, vp]
            getattr(_lib, f"oracle_{p}larfg").restype = None
            getattr(_lib, f"oracle_{p}larfg").argtypes = [vp, i64, vp, i64, vp]
            getattr(_lib, f"oracle_{p}getrf_batched").restype = None
            getattr(_lib, f"oracle_{p}getrf_batched").argtypes = [i64, i64, vp, vp, vp]
        _lib.oracle_diamax.restype = i64
        _lib.oracle_diamax.argtypes = [i64, vp, i64, vp]
        _lib.oracle_siamax.restype = i64
        _lib.oracle_siamax.argtypes = [i64, vp, i64, vp]
        _lib.oracle_ziamax.restype = i64
        _lib.oracle_ziamax.argtypes = [i64, vp, i64, vp]
        _lib.oracle_time_dgetrf_dgetrs.restype = dbl
        _lib.oracle_time_dgetrf_dgetrs.argtypes = [i64, vp, vp, i64, vp, vp, vp, vp]
    return _lib


def _pfx(a: np.ndarray) -> str:
    try:
        return _PREFIX[a.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {a.dtype}") from None


def _strides2(a: np.ndarray):
    assert a.ndim == 2
    isz = a.itemsize
    assert a.strides[0] % isz == 0 and a.strides[1] % isz == 0
    return a.strides[0] // isz, a.strides[1] // isz


def getrf(a: np.ndarray, variant: str | None = None):
    """In-place `lapack::getrf` on a 2-D numpy view of ANY strides.

    Returns (pivots: list[int], singular: int | None) exactly like
    src/lapack/getrf.rs:12-27.  `variant` in {None, "row", "col"} forces one body.
    """
    m, n = a.shape
    rs, cs = _strides2(a)
    piv = np.zeros(max(min(m, n), 1), dtype=np.uint64)
    p = _pfx(a)
    if variant is None:
        info = getattr(lib(), f"oracle_{p}getrf")(m, n, a.ctypes.data, rs, cs, piv.ctypes.data)
    else:
        which = {"row": 0, "col": 1}[variant]
        info = getattr(lib(), f"oracle_{p}getrf_variant")(which, m, n, a.ctypes.data, rs, cs, piv.ctypes.data)
    return [int(v) for v in piv[: min(m, n)]], (None if info < 0 else int(info))


def getrf_recursive(a: np.ndarray):
    """`getrf_recursive` (src/lapack/getrf.rs:30-40): returns (pivots[len m], err_row|None)."""
    m, n = a.shape
    rs, cs = _strides2(a)
    piv = np.zeros(max(m, 1), dtype=np.uint64)
    err = getattr(lib(), f"oracle_{_pfx(a)}getrf_recursive")(m, n, a.ctypes.data, rs, cs, piv.ctypes.data)
    return [int(v) for v in piv[:m]], (None if err < 0 else int(err))


def getrs(a: np.ndarray, p, b: np.ndarray) -> np.ndarray:
    """`lapack::getrs` (src/lapack/getrs.rs:12-38), single right-hand side."""
    assert a.shape[0] == len(p) and len(p) == b.shape[0] and a.shape[1] >= len(p)
    n = len(p)
    rs, cs = _strides2(a)
    piv = np.asarray(p, dtype=np.uint64)
    x = np.empty(n, dtype=a.dtype)
    bb = np.asarray(b, dtype=a.dtype)
    incb = bb.strides[0] // bb.itemsize if n else 1
    getattr(lib(), f"oracle_{_pfx(a)}getrs")(
        n, a.ctypes.data, rs, cs, piv.ctypes.data, bb.ctypes.data, incb, x.ctypes.data)
    return x


def laswp(a: np.ndarray, piv, begin: int = 0) -> None:
    """`lapack::laswp` on all columns of a 2-D view (src/lapack/laswp.rs:11-40)."""
    rs, cs = _strides2(a)
    pv = np.asarray(piv, dtype=np.uint64)
    if a.shape[1] == 0:
        return
    getattr(lib(), f"oracle_{_pfx(a)}laswp")(a.shape[1], a.ctypes.data, rs, cs, begin, pv.ctypes.data, len(pv))


def iamax(x: np.ndarray):
    """`blas::iamax` (src/blas/iamax.rs:6-21) -> (index, value)."""
    inc = x.strides[0] // x.itemsize if x.size else 1
    if x.dtype == np.float64:
        v = ctypes.c_double()
        i = lib().oracle_diamax(x.size, x.ctypes.data, inc, ctypes.addressof(v))
    elif x.dtype == np.float32:
        v = ctypes.c_float()
        i = lib().oracle_siamax(x.size, x.ctypes.data, inc, ctypes.addressof(v))
    elif x.dtype == np.complex128:
        v = ctypes.c_double()
        i = lib().oracle_ziamax(x.size, x.ctypes.data, inc, ctypes.addressof(v))
    else:
        raise TypeError(x.dtype)
    return int(i), float(v.value)


def into_pl(lu: np.ndarray, piv) -> None:
    """`lu::Factorized::into_pl` in place (src/decomposition/lu.rs:107-153)."""
    rs, cs = _strides2(lu)
    pv = np.asarray(piv, dtype=np.uint64)
    getattr(lib(), f"oracle_{_pfx(lu)}into_pl")(lu.shape[0], lu.shape[1], lu.ctypes.data, rs, cs, pv.ctypes.data, len(pv))


def getrf_batched(a: np.ndarray):
    """Loop of `getrf` over a contiguous [batch, n, n] row-major array (timing helper)."""
    assert a.ndim == 3 and a.shape[1] == a.shape[2] and a.flags.c_contiguous
    batch, n, _ = a.shape
    piv = np.zeros((batch, n), dtype=np.uint64)
    info = np.zeros(batch, dtype=np.int64)
    getattr(lib(), f"oracle_{_pfx(a)}getrf_batched")(batch, n, a.ctypes.data, piv.ctypes.data, info.ctypes.data)
    return piv.astype(np.int64), info


def time_dgetrf_dgetrs(a: np.ndarray, b: np.ndarray | None):
    """Factor (+ solve each RHS column with the reference's single-RHS getrs) on one core.

    Returns (secs_getrf, secs_getrs, piv, x).  `a` (row-major f64) is overwritten.
    """
    assert a.dtype == np.float64 and a.flags.c_contiguous and a.shape[0] == a.shape[1]
    n = a.shape[0]
    piv = np.zeros(n, dtype=np.uint64)
    nrhs = 0 if b is None else b.shape[1]
    x = np.zeros((n, max(nrhs, 1)), dtype=np.float64)
    tg, ts = ctypes.c_double(), ctypes.c_double()
    bb = np.ascontiguousarray(b) if b is not None else x
    lib().oracle_time_dgetrf_dgetrs(n, a.ctypes.data, piv.ctypes.data, nrhs, bb.ctypes.data, x.ctypes.data,
                                    ctypes.addressof(tg), ctypes.addressof(ts))
    return tg.value, ts.value, piv.astype(np.int64), x[:, :nrhs]


# ---- QR path ---------------------------------------------------------------
def larfg(alpha, x: np.ndarray):
    """`lapack::larfg` (src/lapack/larfg.rs:9-42): x (1-D) is overwritten; returns (beta, tau)."""
    ab = np.array([alpha], dtype=x.dtype)
    tau = np.zeros(1, dtype=x.dtype)
    inc = x.strides[0] // x.itemsize if x.size else 1
    getattr(lib(), f"oracle_{_pfx(x)}larfg")(ab.ctypes.data, x.shape[0], x.ctypes.data, inc, tau.ctypes.data)
    return float(ab[0].real), tau[0]


def geqrf(a: np.ndarray) -> np.ndarray:
    """In-place `lapack::geqrf` (src/lapack/geqrf.rs:9-30) on a 2-D view of any strides; returns tau."""
    m, n = a.shape
    rs, cs = _strides2(a)
    tau = np.zeros(max(min(m, n), 1), dtype=a.dtype)
    getattr(lib(), f"oracle_{_pfx(a)}geqrf")(m, n, a.ctypes.data, rs, cs, tau.ctypes.data)
    return tau[: min(m, n)]


def qr_q(qr: np.ndarray, tau: np.ndarray) -> np.ndarray:
    """`qr::Factorized::q` (src/decomposition/qr.rs:27-59): the m x m unitary factor."""
    m, n = qr.shape
    rs, cs = _strides2(qr)
    q = np.zeros((m, m), dtype=qr.dtype)
    t = np.ascontiguousarray(tau, dtype=qr.dtype)
    if t.size == 0:
        t = np.zeros(1, dtype=qr.dtype)
    getattr(lib(), f"oracle_{_pfx(qr)}qr_q")(m, n, qr.ctypes.data, rs, cs, t.ctypes.data, q.ctypes.data)
    return q


def qr_r(qr: np.ndarray) -> np.ndarray:
    """`qr::Factorized::r` (src/decomposition/qr.rs:62-70): qr with the strict lower triangle zeroed."""
    return np.triu(np.array(qr, copy=True))

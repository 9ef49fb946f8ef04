"""`lair::lapack::{getrf, getrs, laswp}` over numpy arrays, running on the B200.

Host-side mirror of the reference's crate-private LAPACK layer for the LU path; names,
argument meaning and error behaviour follow the Rust signatures:

* getrf(a) -> (pivots, singular)        src/lapack/getrf.rs:12-27  (in place, any strides)
* getrs(a, p, b) -> x                   src/lapack/getrs.rs:12-38  (panics -> AssertionError)
* laswp(ncols, a, ..., begin, piv)      src/lapack/laswp.rs:11-40  (host-side index work)
* geqrf(a) -> tau                       src/lapack/geqrf.rs:9-30   (in place, any strides)

All arithmetic happens in liblair_b200.so (CUDA, sm_100a); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _ffi

_PREFIX = {np.dtype(np.float32): "s", np.dtype(np.float64): "d", np.dtype(np.complex64): "c",
           np.dtype(np.complex128): "z"}


def _prefix(a: np.ndarray) -> str:
    try:
        return _PREFIX[a.dtype]
    except KeyError:
        # the reference is generic over `A: Scalar`; only f32/f64/Complex<f32>/Complex<f64> have kernels
        raise TypeError(f"unsupported scalar {a.dtype}: lair_b200 implements f32, f64, c64, c128") from None


def _elem_strides(a: np.ndarray):
    isz = a.itemsize
    if any(s % isz for s in a.strides):
        raise ValueError("strides must be a multiple of the element size")
    return [s // isz for s in a.strides]


def getrf(a: np.ndarray):
    """LU-factorize the 2-D array (view) `a` in place: P*A = L*U.

    Returns `(pivots, singular)`: the 0-based sequential interchange vector of length
    min(m, n) and the LAST step whose pivot was exactly zero (or None) -- getrf.rs:12-27.
    """
    if a.ndim != 2:
        raise ValueError("getrf expects a 2-D array")
    if not a.flags.writeable:
        raise ValueError("getrf factors in place: array must be writeable")
    m, n = a.shape
    k = min(m, n)
    rs, cs = _elem_strides(a)
    piv = np.zeros(max(k, 1), dtype=np.int64)
    info = ctypes.c_int64(-1)
    fn = getattr(_ffi.lib(), f"lair_b200_{_prefix(a)}getrf")
    _ffi.check(fn(m, n, a.ctypes.data, rs, cs, piv.ctypes.data, ctypes.addressof(info)))
    return [int(p) for p in piv[:k]], (None if info.value < 0 else int(info.value))


def getrs(a: np.ndarray, p, b: np.ndarray) -> np.ndarray:
    """Solve `a * x = b` from the factors and pivots produced by `getrf` (getrs.rs:12-38).

    `b` may be 1-D (the reference signature) or 2-D `n x nrhs` (the multi-RHS extension:
    column r of the result equals the reference's getrs on column r).  Shape violations
    raise AssertionError, mirroring the reference's `assert!` panics (getrs.rs:18-20).
    """
    assert a.ndim == 2 and a.shape[0] == len(p), "assertion failed: a.nrows() == p.len()"
    assert len(p) == b.shape[0], "assertion failed: p.len() == b.len()"
    assert a.shape[1] >= len(p), "assertion failed: a.ncols() >= p.len()"
    n = len(p)
    pfx = _prefix(a)
    b = np.asarray(b)
    if b.dtype != a.dtype:
        raise TypeError("a and b must have the same scalar type")
    one_d = b.ndim == 1
    b2 = b.reshape(n, 1) if one_d else b
    nrhs = b2.shape[1]
    x = np.empty((n, nrhs), dtype=a.dtype)
    piv = np.ascontiguousarray(np.asarray(p, dtype=np.int64))
    lrs, lcs = _elem_strides(a)
    brs, bcs = _elem_strides(b2)
    if one_d:
        brs = _elem_strides(b)[0]
        bcs = 1
    fn = getattr(_ffi.lib(), f"lair_b200_{pfx}getrs")
    _ffi.check(fn(n, nrhs, a.ctypes.data, lrs, lcs, piv.ctypes.data, b2.ctypes.data, brs, bcs, x.ctypes.data, nrhs, 1))
    return x[:, 0].copy() if one_d else x


def geqrf(a: np.ndarray) -> np.ndarray:
    """QR-factorize the 2-D array (view) `a` in place and return tau (geqrf.rs:9-30): R on and above the
    diagonal, the Householder vectors below it, `tau` of length min(m, n)."""
    if a.ndim != 2:
        raise ValueError("geqrf expects a 2-D array")
    if not a.flags.writeable:
        raise ValueError("geqrf factors in place: array must be writeable")
    m, n = a.shape
    k = min(m, n)
    rs, cs = _elem_strides(a)
    tau = np.zeros(max(k, 1), dtype=a.dtype)
    fn = getattr(_ffi.lib(), f"lair_b200_{_prefix(a)}geqrf")
    _ffi.check(fn(m, n, a.ctypes.data, rs, cs, tau.ctypes.data))
    return tau[:k]


def qr_q(qr: np.ndarray, tau: np.ndarray) -> np.ndarray:
    """The m x m unitary factor from `geqrf`'s output (qr::Factorized::q, qr.rs:27-59)."""
    if qr.ndim != 2:
        raise ValueError("qr_q expects a 2-D array")
    m, n = qr.shape
    assert len(tau) == min(m, n), "assertion failed: tau.len() == min(nrows, ncols)"
    q = np.empty((m, m), dtype=qr.dtype)
    t = np.ascontiguousarray(tau, dtype=qr.dtype)
    if t.size == 0:
        t = np.zeros(1, dtype=qr.dtype)
    rs, cs = _elem_strides(qr)
    fn = getattr(_ffi.lib(), f"lair_b200_{_prefix(qr)}qr_q")
    _ffi.check(fn(m, n, qr.ctypes.data, rs, cs, t.ctypes.data, q.ctypes.data, m, 1))
    return q


def laswp(a: np.ndarray, piv, begin: int = 0) -> None:
    """Apply the interchanges `piv[begin:]` to the rows of `a` in place (laswp.rs:11-40).

    Works on any element type (the reference requires only `T: Copy`; it is also used on
    the `usize` permutation vector in lu.rs:31), so this stays host-side index work.
    """
    view = a if a.ndim == 2 else a.reshape(-1, 1)
    for i in range(begin, len(piv)):
        p = int(piv[i])
        if i == p:
            continue
        tmp = view[i].copy()
        view[i] = view[p]
        view[p] = tmp


def getrf_batched(a: np.ndarray, ngpu: int = 1):
    """`batch` independent getrf calls on a C-contiguous [batch, n, n] array (n <= 32), in place.

    Returns (ipiv [batch, n] int32, info [batch] int32 with -1 = None).  `ngpu` > 1 spreads contiguous slices of the
    batch over the first `ngpu` devices of the node from this one process (no collective; same bits).
    """
    if a.ndim != 3 or a.shape[1] != a.shape[2] or not a.flags.c_contiguous:
        raise ValueError("getrf_batched expects a C-contiguous [batch, n, n] array")
    batch, n, _ = a.shape
    ipiv = np.zeros((batch, n), dtype=np.int32)
    info = np.full(batch, -1, dtype=np.int32)
    pfx = _prefix(a)
    if pfx not in "sd":
        raise TypeError("getrf_batched supports f32 and f64")
    if ngpu != 1:
        fn = getattr(_ffi.lib(), f"lair_b200_{pfx}getrf_batched_mg")
        _ffi.check(fn(batch, n, a.ctypes.data, ipiv.ctypes.data, info.ctypes.data, int(ngpu)))
        return ipiv, info
    fn = getattr(_ffi.lib(), f"lair_b200_{pfx}getrf_batched")
    _ffi.check(fn(batch, n, a.ctypes.data, ipiv.ctypes.data, info.ctypes.data))
    return ipiv, info


def gesv(a: np.ndarray, b: np.ndarray):
    """Factor a copy of `a` and solve for `b` with the factors kept device-resident.

    Returns (x, singular); x is None when singular (equation.rs:55-57).
    """
    b = np.asarray(b)
    if a.ndim != 2 or a.shape[0] != a.shape[1]:
        raise ValueError("gesv expects a square 2-D matrix")
    n = a.shape[0]
    pfx = _prefix(a)
    if pfx not in "sd":
        raise TypeError("gesv supports f32 and f64")
    if b.ndim not in (1, 2) or b.shape[0] != n:
        raise ValueError(f"gesv: b has {b.shape[0] if b.ndim else 0} rows, a has {n}")
    if b.dtype != a.dtype:  # the C side reinterprets b's bytes as a's scalar type
        raise TypeError("a and b must have the same scalar type")
    one_d = b.ndim == 1
    b2 = b.reshape(n, 1) if one_d else b
    nrhs = b2.shape[1]
    x = np.empty((n, nrhs), dtype=a.dtype)
    ars, acs = _elem_strides(a)
    brs, bcs = _elem_strides(b2)
    if one_d:
        brs, bcs = _elem_strides(b)[0], 1
    info = ctypes.c_int64(-1)
    fn = getattr(_ffi.lib(), f"lair_b200_{pfx}gesv")
    _ffi.check(fn(n, nrhs, a.ctypes.data, ars, acs, b2.ctypes.data, brs, bcs, x.ctypes.data, nrhs, 1,
                  ctypes.addressof(info)))
    if info.value >= 0:
        return None, int(info.value)
    return (x[:, 0].copy() if one_d else x), None

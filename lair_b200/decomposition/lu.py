"""`lair::decomposition::lu` -- LU decomposition factors (src/decomposition/lu.rs).

`Factorized` owns the factors the way the reference's does (lu.rs:12-20), except that they live in
HBM behind a `lair_b200_lu_t` handle: `from_` uploads and factors, `solve` sends only the
right-hand side(s), `p / l / u / into_pl` are built by device kernels, and L\\U itself comes to the
host only if `.lu` is read.
"""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _ffi
from ..errors import InvalidInput
from ..lapack import _elem_strides, _prefix

VIEW_L, VIEW_U, VIEW_P, VIEW_PL = 0, 1, 2, 3


class Factorized:
    """LU decomposition factors: `lu`, `pivots`, `singular` (lu.rs:12-20), device-resident.

    `Factorized.from_(a)` is `From<ArrayBase<S, Ix2>>` (lu.rs:156-171).  The reference consumes the
    array by value, so the caller can no longer observe it; here `a` is only read (the H2D copy is
    the copy) and left untouched.
    """

    def __init__(self, handle, shape, dtype, singular):
        self._h = handle
        self._shape = tuple(shape)
        self._dtype = np.dtype(dtype)
        self._singular = singular
        self._lu = None
        self._pivots = None

    @classmethod
    def from_(cls, a: np.ndarray) -> "Factorized":
        if a.ndim != 2:
            raise ValueError("Factorized.from_ expects a 2-D array")
        m, n = a.shape
        rs, cs = _elem_strides(a)
        h = ctypes.c_void_p(None)
        info = ctypes.c_int64(-1)
        fn = getattr(_ffi.lib(), f"lair_b200_{_prefix(a)}lu_factor")
        _ffi.check(fn(m, n, a.ctypes.data, rs, cs, ctypes.byref(h), ctypes.byref(info)))
        return cls(h, (m, n), a.dtype, None if info.value < 0 else int(info.value))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and h.value:
            try:
                _ffi.lib().lair_b200_lu_destroy(h)
            except Exception:  # interpreter shutdown
                pass

    def _view(self, which: int, rows: int, cols: int) -> np.ndarray:
        out = np.empty((rows, cols), dtype=self._dtype)
        if rows and cols:
            _ffi.check(_ffi.lib().lair_b200_lu_view(self._h, which, out.ctypes.data, cols, 1))
        return out

    # -- lu.rs:28-39 --
    def p(self) -> np.ndarray:
        """Permutation matrix P with P[perm[i], i] = 1."""
        m = self._shape[0]
        return self._view(VIEW_P, m, m)

    # -- lu.rs:42-57 --
    def l(self) -> np.ndarray:
        m, n = self._shape
        return self._view(VIEW_L, m, min(m, n))

    # -- lu.rs:60-72 --
    def u(self) -> np.ndarray:
        m, n = self._shape
        return self._view(VIEW_U, min(m, n), n)

    # -- lu.rs:75-77 --
    def is_singular(self) -> bool:
        return self._singular is not None

    # -- lu.rs:87-98 --
    def solve(self, b: np.ndarray) -> np.ndarray:
        """Solve P*L*U*x = b.  Raises InvalidInput.Shape when b has the wrong length.

        `b` is a vector as in the reference, or (an addition) an n x nrhs matrix of right-hand sides.
        """
        m, n = self._shape
        if b.shape[0] != m:
            raise InvalidInput.Shape(f"b must have {m} elements")
        if b.dtype != self._dtype:
            raise TypeError(f"b must be {self._dtype}")
        one_d = b.ndim == 1
        nrhs = 1 if one_d else b.shape[1]
        if one_d:
            brs, bcs = _elem_strides(b)[0], 1
        else:
            brs, bcs = _elem_strides(b)
        x = np.empty((m, nrhs), dtype=self._dtype)
        _ffi.check(_ffi.lib().lair_b200_lu_solve(self._h, nrhs, b.ctypes.data, brs, bcs, x.ctypes.data, nrhs, 1))
        return x[:, 0].copy() if one_d else x

    # -- lu.rs:107-153 --
    def into_pl(self) -> np.ndarray:
        """P*L in the first min(m, n) columns of the factor array (consumes self)."""
        m, n = self._shape
        k = min(m, n)
        out = self._view(VIEW_PL, m, k) if k == n else self.lu.copy()
        if k != n:
            out[:, :k] = self._view(VIEW_PL, m, k)
        return out

    @property
    def pivots(self):
        if self._pivots is None:
            k = min(self._shape)
            piv = np.zeros(max(k, 1), dtype=np.int64)
            _ffi.check(_ffi.lib().lair_b200_lu_pivots(self._h, piv.ctypes.data))
            self._pivots = [int(v) for v in piv[:k]]
        return list(self._pivots)

    @property
    def singular(self):
        return self._singular

    @property
    def lu(self) -> np.ndarray:
        """Packed L\\U, downloaded on first use."""
        if self._lu is None:
            m, n = self._shape
            out = np.empty((m, n), dtype=self._dtype)
            if m and n:
                _ffi.check(_ffi.lib().lair_b200_lu_factors(self._h, out.ctypes.data, n, 1))
            self._lu = out
        return self._lu

"""`lair::decomposition::lu` -- LU decomposition factors (src/decomposition/lu.rs)."""
from __future__ import annotations

import numpy as np

from .. import lapack
from ..errors import InvalidInput


class Factorized:
    """LU decomposition factors: owns `lu`, `pivots`, `singular` (lu.rs:12-20).

    `Factorized.from_(a)` is `From<ArrayBase<S, Ix2>>` (lu.rs:156-171): it takes the array
    and factors it IN PLACE (the reference consumes the array by value).
    """

    def __init__(self, lu: np.ndarray, pivots, singular):
        self._lu = lu
        self._pivots = list(pivots)
        self._singular = singular

    @classmethod
    def from_(cls, a: np.ndarray) -> "Factorized":
        pivots, singular = lapack.getrf(a)
        return cls(a, pivots, singular)

    # -- lu.rs:28-39 --
    def p(self) -> np.ndarray:
        """Permutation matrix P with P[perm[i], i] = 1."""
        n = self._lu.shape[0]
        perm = np.arange(n)
        lapack.laswp(perm, self._pivots)
        out = np.zeros((n, n), dtype=self._lu.dtype)
        out[perm, np.arange(n)] = 1
        return out

    # -- lu.rs:42-57 --
    def l(self) -> np.ndarray:
        m, n = self._lu.shape
        rank = min(m, n)
        out = np.tril(self._lu[:, :rank], -1).astype(self._lu.dtype, copy=True)
        idx = np.arange(rank)
        out[idx, idx] = 1
        return out

    # -- lu.rs:60-72 --
    def u(self) -> np.ndarray:
        m, n = self._lu.shape
        rank = min(m, n)
        return np.triu(self._lu[:rank, :]).astype(self._lu.dtype, copy=True)

    # -- lu.rs:75-77 --
    def is_singular(self) -> bool:
        return self._singular is not None

    # -- lu.rs:87-98 --
    def solve(self, b: np.ndarray) -> np.ndarray:
        """Solve P*L*U*x = b.  Raises InvalidInput.Shape when b has the wrong length."""
        if b.shape[0] != self._lu.shape[0]:
            raise InvalidInput.Shape(f"b must have {self._lu.shape[0]} elements")
        return lapack.getrs(self._lu, self._pivots, b)

    # -- lu.rs:107-153 --
    def into_pl(self) -> np.ndarray:
        """P*L in the first min(m, n) columns of the factor array (consumes self)."""
        lu = self._lu
        m, n = lu.shape
        k = min(m, n)
        perm = np.arange(m)
        lapack.laswp(perm, self._pivots)  # perm[i] = original row now at position i
        l_full = np.zeros((m, k), dtype=lu.dtype)
        l_full[:, :] = np.tril(lu[:, :k], -1)
        idx = np.arange(k)
        l_full[idx, idx] = 1
        pl = np.zeros_like(l_full)
        pl[perm] = l_full
        lu[:, :k] = pl
        return lu

    @property
    def pivots(self):
        return list(self._pivots)

    @property
    def lu(self) -> np.ndarray:
        return self._lu

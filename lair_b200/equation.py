"""`lair::equation` -- matrix equation solvers (src/equation.rs)."""
from __future__ import annotations

import numpy as np

from . import lapack
from .errors import InvalidInput


def solve(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Solve a system of linear scalar equations `a * x = b` (equation.rs:32-60).

    Raises InvalidInput.Shape if `a` is not square or `b`'s length differs from `a`'s rows,
    and InvalidInput.Value if `a` is singular.  `a` is not modified (the reference factors
    `a.to_owned()`); on the device the copy is the H2D transfer itself and the factors never
    come back to the host between factorization and solve.
    """
    if a.ndim != 2 or a.shape[0] != a.shape[1]:
        raise InvalidInput.Shape("input matrix is not square")
    if b.shape[0] != a.shape[0]:
        raise InvalidInput.Shape(
            f"The number of elements in `b`, {b.shape[0]}, must be the same as the number of rows in `a`, {a.shape[0]}")
    if a.dtype in (np.float32, np.float64):
        x, singular = lapack.gesv(a, b)
        if singular is not None:
            raise InvalidInput.Value("`a` is a singular matrix")
        return x
    # complex: factor a copy, then solve (two C-ABI calls)
    from .decomposition import lu
    f = lu.Factorized.from_(np.array(a, copy=True))
    if f.is_singular():
        raise InvalidInput.Value("`a` is a singular matrix")
    return f.solve(b)

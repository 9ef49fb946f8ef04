"""`lair::decomposition::qr` -- QR decomposition factors (src/decomposition/qr.rs).

`Factorized.from_(a)` is `From<ArrayBase<S, Ix2>>` (qr.rs:73-84: `geqrf` on the array it consumes); `q()` and
`r()` follow qr.rs:27-59 and :62-70.  geqrf and the accumulation of Q run on the B200 through the C ABI
(`lair_b200_*geqrf`, `lair_b200_*qr_q`); `r()` is the upper triangle of the stored factors (no arithmetic).
"""
from __future__ import annotations

import numpy as np

from .. import lapack


class Factorized:
    """QR decomposition factors: `qr` (R and the reflectors) and `tau` (qr.rs:12-19)."""

    def __init__(self, qr: np.ndarray, tau: np.ndarray):
        self.qr = qr
        self.tau = tau

    @classmethod
    def from_(cls, a: np.ndarray) -> "Factorized":
        if a.ndim != 2:
            raise ValueError("Factorized.from_ expects a 2-D array")
        qr = np.array(a, copy=True, order="C")  # the reference consumes `a`; here it is left untouched
        tau = lapack.geqrf(qr)
        return cls(qr, tau)

    def q(self) -> np.ndarray:
        """*Q* of the decomposition, nrows x nrows (qr.rs:27-59)."""
        return lapack.qr_q(self.qr, self.tau)

    def r(self) -> np.ndarray:
        """*R* of the decomposition, nrows x ncols: the factors with the strict lower triangle zeroed (qr.rs:62-70)."""
        return np.triu(self.qr)

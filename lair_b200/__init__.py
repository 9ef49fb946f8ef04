"""lair_b200 -- B200-native LU factorization and solve behind lair's API.

Host-side mirror of the reference's public path (`decomposition::lu`, `equation::solve`,
`decomposition::qr`) and its crate-private `lapack::{getrf, getrs, laswp, geqrf}`, over the C ABI in
include/lair_b200.h.  All arithmetic runs in hand-written CUDA for sm_100a.
"""
from . import _ffi, decomposition, equation, lapack  # noqa: F401
from .errors import InvalidInput  # noqa: F401

__all__ = ["decomposition", "equation", "lapack", "InvalidInput"]

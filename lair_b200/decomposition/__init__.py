"""`lair::decomposition` -- the LU module and the QR module."""
from . import lu, qr  # noqa: F401

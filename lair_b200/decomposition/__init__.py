"""`lair::decomposition` -- only the LU module is in scope."""
from . import lu  # noqa: F401

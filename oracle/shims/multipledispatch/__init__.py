"""Minimal stand-in for the `multipledispatch` package (see ../README.md).  Test infrastructure only."""
from .core import dispatch
from .dispatcher import Dispatcher, MDNotImplementedError, halt_ordering, restart_ordering

__version__ = "0.6.0+rbnics_b200_shim"
__all__ = ["dispatch", "Dispatcher", "MDNotImplementedError", "halt_ordering", "restart_ordering"]

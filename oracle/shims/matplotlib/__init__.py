"""No-op stand-in for `matplotlib` (RBniCS imports it in rbnics/utils/test/matplotlib.py:7-9 only to switch plotting
back-ends in its test helpers).  Test infrastructure only; see ../README.md."""
__version__ = "0.0+rbnics_b200_shim"
_backend = "agg"


def get_backend():
    return _backend


def use(backend, **kwargs):
    global _backend
    _backend = backend

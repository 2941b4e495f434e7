"""No-op `matplotlib.pyplot` (see __init__.py)."""
import matplotlib


def switch_backend(name):
    matplotlib.use(name)


def close(*args, **kwargs):
    pass


def _unavailable(*args, **kwargs):
    raise RuntimeError("matplotlib is not installed: this is the no-op stand-in of oracle/shims")


figure = plot = show = savefig = subplots = _unavailable

"""Import-only stand-in for `cvxopt` (used by RBniCS solely inside the SCM linear-program solver,
rbnics/backends/common/linear_program_solver.py:7 -- not on the greedy-sweep path).  Test infrastructure only."""
__version__ = "0.0+rbnics_b200_shim"


def matrix(*args, **kwargs):
    raise RuntimeError("cvxopt is not installed: this is the import-only stand-in of oracle/shims")


class solvers(object):
    options = {}

    @staticmethod
    def lp(*args, **kwargs):
        raise RuntimeError("cvxopt is not installed: this is the import-only stand-in of oracle/shims")

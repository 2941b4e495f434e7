"""`rbnics.backends.online.b200` -- the online backend that sits beside `rbnics/backends/online/numpy` and is selected
with

    [backends]
    online backend = b200

in `.rbnicsrc` (rbnics/backends/online/__init__.py:19-53 imports this package and copies the names of __all__ as
OnlineXxx / online_xxx).  The four containers (AffineExpansionStorage, Function, Matrix, Vector) are the reference's
own classes re-instantiated as distinct types; the arithmetic on the greedy-sweep path (product / sum, LinearSolver,
transpose) goes through the C ABI of librbnics_b200.so (ctypes, `rbnics_b200.online`) into CUDA on the B200 -- there is
no CPU fallback for those calls.  The remaining 21 names of the backend interface (rbnics/backends/online/numpy/
__init__.py:38-68) are not on that path and are the numpy backend's, which accept the B200 containers by inheritance.

The batched many-query sweep -- the point of this backend -- has no counterpart in the reference's per-parameter API;
it enters through the `_greedy` overrides of `rbnics_b200_plugin`."""
from rbnics.backends.online.numpy import (
    abs, assign, BasisFunctionsMatrix, copy, EigenSolver, evaluate, export, FunctionsList, GramSchmidt,
    HighOrderProperOrthogonalDecomposition, import_, max, NonAffineExpansionStorage, NonlinearSolver,
    ProperOrthogonalDecomposition, SnapshotsMatrix, TensorBasisList, TensorSnapshotsList, TensorsList, TimeQuadrature,
    TimeStepping)
from rbnics.backends.online.b200.affine_expansion_storage import AffineExpansionStorage
from rbnics.backends.online.b200.function import Function
from rbnics.backends.online.b200.linear_solver import LinearSolver
from rbnics.backends.online.b200.matrix import Matrix
from rbnics.backends.online.b200.product import product
from rbnics.backends.online.b200.sum import sum
from rbnics.backends.online.b200.transpose import transpose
from rbnics.backends.online.b200.vector import Vector

__all__ = [
    "abs",
    "AffineExpansionStorage",
    "assign",
    "BasisFunctionsMatrix",
    "copy",
    "EigenSolver",
    "evaluate",
    "export",
    "Function",
    "FunctionsList",
    "GramSchmidt",
    "HighOrderProperOrthogonalDecomposition",
    "import_",
    "LinearSolver",
    "Matrix",
    "max",
    "NonAffineExpansionStorage",
    "NonlinearSolver",
    "product",
    "ProperOrthogonalDecomposition",
    "SnapshotsMatrix",
    "sum",
    "TensorBasisList",
    "TensorSnapshotsList",
    "TensorsList",
    "TimeQuadrature",
    "TimeStepping",
    "transpose",
    "Vector"
]

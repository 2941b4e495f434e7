"""B200 online backend: Vector.  Same container as the numpy backend (rbnics/backends/online/numpy/vector.py:14-47 on
top of rbnics/backends/online/basic/vector.py), as a DISTINCT subclass so that the dispatcher
(rbnics/utils/decorators/dispatch.py:76-186) can route the hot-path calls (product, LinearSolver, transpose) to CUDA
while every other numpy-backend function keeps accepting it by inheritance."""
from rbnics.backends.online.numpy.vector import Vector as NumpyVector
from rbnics.utils.decorators import backend_for, OnlineSizeType


class _Vector_Type(NumpyVector.Type()):
    pass


@backend_for("b200", inputs=(OnlineSizeType, ), replaces=NumpyVector)
def Vector(N):
    return _Vector_Type(N)


def Type():
    return _Vector_Type


Vector.Type = Type

"""B200 online backend: Function = a Vector with the reference's Function interface
(rbnics/backends/online/basic/function.py:10-30, numpy/function.py:11-24)."""
from rbnics.backends.online.numpy.function import Function as NumpyFunction
from rbnics.backends.online.b200.vector import Vector
from rbnics.utils.decorators import backend_for, OnlineSizeType


class _Function_Type(NumpyFunction.Type()):
    def __init__(self, arg):
        if isinstance(arg, (int, dict)):
            arg = Vector(arg)   # a B200 vector, not a numpy-backend one
        NumpyFunction.Type().__init__(self, arg)


def Function(arg):
    return _Function_Type(arg)


# the size signatures replace the numpy registration in rbnics.backends.Function; the vector one is new
backend_for("b200", inputs=(OnlineSizeType, ), replaces=NumpyFunction)(Function)
backend_for("b200", inputs=(Vector.Type(), ))(Function)


def Type():
    return _Function_Type


Function.Type = Type

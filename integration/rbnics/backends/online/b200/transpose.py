"""B200 online backend: transpose().  v^T w and v^T M w of rbnics/backends/basic/transpose.py:126-237 as ONE device
call each (rb_bilinear) for B200 Functions / Vectors; every other argument type (basis functions matrices, functions
lists, tensors lists: not on the greedy-sweep path) is the numpy backend's."""
from rbnics.backends.online.numpy.transpose import transpose as numpy_transpose
from rbnics.backends.online.b200.function import Function
from rbnics.backends.online.b200.matrix import Matrix
from rbnics.backends.online.b200.vector import Vector
from rbnics.utils.decorators import backend_for


def _content(arg):
    return arg.vector().content if isinstance(arg, Function.Type()) else arg.content


class _Vector_Transpose(object):
    def __init__(self, arg):
        self._arg = arg

    def __mul__(self, other):
        from rbnics_b200 import online as _dev
        if isinstance(other, Matrix.Type()):
            return _Vector_Transpose__times__Matrix(self._arg, other)
        if isinstance(other, (Function.Type(), Vector.Type())):
            return _dev.bilinear(_dev.get_engine(), _content(self._arg), None, _content(other))
        return numpy_transpose(self._arg) * other


class _Vector_Transpose__times__Matrix(object):
    def __init__(self, arg, matrix):
        self._arg = arg
        self._matrix = matrix

    def __mul__(self, other):
        from rbnics_b200 import online as _dev
        if isinstance(other, (Function.Type(), Vector.Type())):
            return _dev.bilinear(_dev.get_engine(), _content(self._arg), self._matrix.content, _content(other))
        return (numpy_transpose(self._arg) * self._matrix) * other


@backend_for("b200", inputs=((Function.Type(), Vector.Type()), ))
def transpose(arg):
    return _Vector_Transpose(arg)

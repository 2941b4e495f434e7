"""B200 online backend: Matrix (see vector.py; numpy counterpart rbnics/backends/online/numpy/matrix.py:15-66)."""
from rbnics.backends.online.numpy.matrix import Matrix as NumpyMatrix
from rbnics.backends.online.b200.vector import Vector
from rbnics.utils.decorators import backend_for, OnlineSizeType


class _Matrix_Type(NumpyMatrix.Type()):
    def __mul__(self, other):
        if isinstance(other, Vector.Type()):
            # matrix * vector on the device: M v = sum_j v_j M[:, j] through rb_affine_combine (one launch)
            from rbnics_b200 import online as _dev
            self._arithmetic_operations_assert_attributes(other, other_order=1)
            import numpy
            content = _dev.affine_combine(_dev.get_engine(), other.content,
                                          numpy.ascontiguousarray(self.content.T))
            output = Vector.Type()(self.M, content)
            self._arithmetic_operations_preserve_attributes(output, other_order=1)
            return output
        return NumpyMatrix.Type().__mul__(self, other)


@backend_for("b200", inputs=(OnlineSizeType, OnlineSizeType), replaces=NumpyMatrix)
def Matrix(M, N):
    return _Matrix_Type(M, N)


def Type():
    return _Matrix_Type


Matrix.Type = Type

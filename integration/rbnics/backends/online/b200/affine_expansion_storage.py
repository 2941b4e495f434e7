"""B200 online backend: AffineExpansionStorage.  The reference's storage (object array of Matrix / Vector / scalar
items, slicing with component dictionaries, save / load: rbnics/backends/online/basic/affine_expansion_storage.py:22-349)
instantiated for the B200 containers, plus a contiguous FP64 copy of all items -- what the C ABI takes -- cached per
(sliced) storage, so that the per-parameter product() does not re-flatten the operators."""
import numpy
from numbers import Number
from rbnics.backends.online.basic import AffineExpansionStorage as BasicAffineExpansionStorage
from rbnics.backends.online.numpy.affine_expansion_storage import AffineExpansionStorage as NumpyAffineExpansionStorage
from rbnics.backends.online.numpy.copy import function_copy, tensor_copy
from rbnics.backends.online.numpy.function import Function as NumpyFunction
from rbnics.backends.online.numpy.matrix import Matrix as NumpyMatrix
from rbnics.backends.online.numpy.vector import Vector as NumpyVector
from rbnics.backends.online.numpy.wrapping import function_load, function_save, tensor_load, tensor_save
from rbnics.backends.online.b200.function import Function
from rbnics.backends.online.b200.matrix import Matrix
from rbnics.backends.online.b200.vector import Vector
from rbnics.utils.decorators import BackendFor, ModuleWrapper, tuple_of

backend = ModuleWrapper(Function, Matrix, Vector)
wrapping = ModuleWrapper(function_load, function_save, tensor_load, tensor_save, function_copy=function_copy,
                         tensor_copy=tensor_copy)
AffineExpansionStorage_Base = BasicAffineExpansionStorage(backend, wrapping)


def _as_b200(item):
    """Tensors produced by numpy-backend code that is not on the hot path (e.g. transpose(Z) * A * Z of the offline
    stage builds its result with the numpy backend's own Matrix) enter the B200 backend here: same content, same sizes
    and component dictionaries, B200 type."""
    if isinstance(item, (Matrix.Type(), Vector.Type(), Function.Type())):
        return item
    if isinstance(item, NumpyMatrix.Type()):
        out = Matrix.Type()(item.M, item.N, item.content)
    elif isinstance(item, NumpyVector.Type()):
        out = Vector.Type()(item.N, item.content)
    elif isinstance(item, NumpyFunction.Type()):
        return Function.Type()(_as_b200(item.vector()))
    else:
        return item
    out._component_name_to_basis_component_index = item._component_name_to_basis_component_index
    out._component_name_to_basis_component_length = item._component_name_to_basis_component_length
    return out


class AffineExpansionStorage(AffineExpansionStorage_Base):
    def __init__(self, arg1, arg2=None):
        if isinstance(arg1, tuple):
            arg1 = tuple(_as_b200(a) for a in arg1)
        AffineExpansionStorage_Base.__init__(self, arg1, arg2)
        self._b200_stacked = None

    def __setitem__(self, key, item):
        self._b200_stacked = None
        return AffineExpansionStorage_Base.__setitem__(self, key, _as_b200(item))

    def b200_stacked(self):
        """(Q, ...) / (Q1, Q2, ...) contiguous FP64 tensor of the items (None if an item is not a tensor / scalar)."""
        if getattr(self, "_b200_stacked", None) is None:
            first = self._content.flat[0]
            if isinstance(first, Number):
                shape = ()
            elif isinstance(first, (Matrix.Type(), Vector.Type())):
                shape = first.content.shape
            else:
                return None
            out = numpy.empty(self._content.shape + shape)
            for idx in numpy.ndindex(self._content.shape):
                item = self._content[idx]
                out[idx] = item if isinstance(item, Number) else item.content
            self._b200_stacked = out
        return self._b200_stacked


# creation from sizes replaces the numpy registration in rbnics.backends.AffineExpansionStorage; creation from tuples of
# B200 tensors is a new, more specific signature
BackendFor("b200", inputs=(int, (int, None)), replaces=NumpyAffineExpansionStorage)(AffineExpansionStorage)
BackendFor("b200", inputs=((tuple_of(Matrix.Type()), tuple_of(Vector.Type())), None))(AffineExpansionStorage)

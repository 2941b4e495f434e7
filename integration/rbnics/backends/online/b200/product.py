"""B200 online backend: product().  Order-1 and order-2 affine sums of rbnics/backends/online/basic/product.py:22-46 on
the device (rb_affine_combine / rb_affine_combine2: same term order, zero-theta skipping and rounding as the
reference's loops -- bit-identical, tests/test_golden.py).  Storages whose items are not tensors or scalars
(NonAffineExpansionStorage, storages of Functions) are outside the greedy-sweep path and go to the reference's generic
loop unchanged."""
from numbers import Number
from rbnics.backends.online.basic import product as basic_product
from rbnics.backends.online.numpy.non_affine_expansion_storage import NonAffineExpansionStorage
from rbnics.backends.online.numpy.transpose import DelayedTransposeWithArithmetic
from rbnics.backends.online.b200.affine_expansion_storage import AffineExpansionStorage
from rbnics.backends.online.b200.function import Function
from rbnics.backends.online.b200.matrix import Matrix
from rbnics.backends.online.b200.vector import Vector
from rbnics.utils.decorators import backend_for, ModuleWrapper, ThetaType

backend = ModuleWrapper(AffineExpansionStorage, Function, Matrix, NonAffineExpansionStorage, Vector)
wrapping = ModuleWrapper(DelayedTransposeWithArithmetic=DelayedTransposeWithArithmetic)
(product_base, ProductOutput) = basic_product(backend, wrapping)


def _like(first, content):
    """A tensor of the type, sizes and component dictionaries of `first` around `content`."""
    if isinstance(first, Matrix.Type()):
        return type(first)(first.M, first.N, content)
    return type(first)(first.N, content)


@backend_for("b200", inputs=(ThetaType, (AffineExpansionStorage, ), ThetaType + (None,)))
def product(thetas, operators, thetas2=None):
    from rbnics_b200 import online as _dev
    stacked = operators.b200_stacked()
    if stacked is None:
        return product_base(thetas, operators, thetas2)
    order = operators.order()
    assert order in (1, 2)
    engine = _dev.get_engine()
    if order == 1:
        assert thetas2 is None
        assert len(thetas) == len(operators)
        first = operators[0]
        content = _dev.affine_combine(engine, thetas, stacked if stacked.ndim > 1 else stacked[:, None])
    else:
        assert thetas2 is not None
        first = operators[0, 0]
        content = _dev.affine_combine2(engine, thetas, thetas2, stacked if stacked.ndim > 2 else stacked[..., None])
    if isinstance(first, Number):
        return ProductOutput(float(content[0]))
    return ProductOutput(_like(first, content))

"""Per-parameter online backend on the B200: host-side mirror of ``rbnics/backends/online/numpy`` for the calls on the
greedy-sweep path, same names and argument meaning:

    AffineExpansionStorage   rbnics/backends/online/basic/affine_expansion_storage.py:22-54,329-349
    product / sum            rbnics/backends/online/basic/product.py:22-46, sum.py:8-9, numpy/product.py:23-25
    LinearSolver             rbnics/backends/online/basic/linear_solver.py:19-71, numpy/linear_solver.py:25-33
    transpose                rbnics/backends/basic/transpose.py:126-237 (vector^T * matrix * vector, vector^T * vector)
    Matrix / Vector / Function   rbnics/backends/online/numpy/{matrix,vector,function}.py

Every arithmetic call goes through the C ABI (``rb_affine_combine[2]``, ``rb_solve_dense``, ``rb_bilinear``) into
CUDA; there is no numpy fallback.  The affine sums keep the reference's order AND rounding (multiply then add, zero
thetas skipped except the first term), so they are bit-identical to the numpy backend.  The greedy loop itself does
not use these one-at-a-time calls: it uses the batched sweep (``rbnics_b200.sweep``).
"""
from __future__ import annotations

import builtins
import ctypes as C
from numbers import Number

import numpy as np

from . import _lib as L

builtins_sum = builtins.sum  # this module defines its own sum(), as the reference's backend does
_engine = None


def set_engine(engine):
    """Engine (= GPU) used by the class-based API below."""
    global _engine
    _engine = engine


def get_engine():
    global _engine
    if _engine is None:
        from .engine import SweepEngine
        _engine = SweepEngine(0)  # raises without a B200: no CPU fallback
    return _engine


# ----------------------------------------------------------------------------------------------------------------------
# functional layer (explicit engine)
# ----------------------------------------------------------------------------------------------------------------------
def affine_combine(engine, thetas, ops):
    """sum_q thetas[q] * ops[q] in the reference's order and rounding.  ops: (Q, ...) -> (...)."""
    ops = np.ascontiguousarray(ops, dtype=np.float64)
    th = np.ascontiguousarray(thetas, dtype=np.float64)
    Q = ops.shape[0]
    if th.shape != (Q,):
        raise ValueError("thetas and operators have different lengths")
    out = np.empty(ops.shape[1:])
    engine._check(engine._lib.rb_affine_combine(engine._h, Q, out.size, L.dptr(th), L.dptr(ops), L.dptr(out)))
    return out


def affine_combine2(engine, thetas, thetas2, ops):
    """sum_ij thetas[i] * ops[i, j] * thetas2[j], row-major (i, j) order.  ops: (Q1, Q2, ...) -> (...)."""
    ops = np.ascontiguousarray(ops, dtype=np.float64)
    t1 = np.ascontiguousarray(thetas, dtype=np.float64)
    t2 = np.ascontiguousarray(thetas2, dtype=np.float64)
    Q1, Q2 = ops.shape[0], ops.shape[1]
    if t1.shape != (Q1,) or t2.shape != (Q2,):
        raise ValueError("thetas and operators have different lengths")
    out = np.empty(ops.shape[2:])
    engine._check(engine._lib.rb_affine_combine2(engine._h, Q1, Q2, out.size, L.dptr(t1), L.dptr(t2), L.dptr(ops),
                                                 L.dptr(out)))
    return out


def solve_dense(engine, lhs, rhs, bc_rows=None, bc_vals=None):
    """x = lhs^-1 rhs after the lifting rows (LinearSolver.solve)."""
    lhs = np.ascontiguousarray(lhs, dtype=np.float64)
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    N = rhs.shape[0]
    if lhs.shape != (N, N):
        raise ValueError("lhs must be N x N and rhs N")
    rows = np.ascontiguousarray(bc_rows if bc_rows is not None else [], dtype=np.int32)
    vals = np.ascontiguousarray(bc_vals if bc_vals is not None else [], dtype=np.float64)
    if rows.shape != vals.shape:
        raise ValueError("bc_rows and bc_vals have different lengths")
    x = np.empty(N)
    engine._check(engine._lib.rb_solve_dense(engine._h, N, L.dptr(lhs), L.dptr(rhs), int(rows.size),
                                             rows.ctypes.data_as(L.c_int_p), L.dptr(vals), L.dptr(x)))
    return x


def bilinear(engine, u, M, v):
    """u^T M v (M is None: u^T v)."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    if M is not None:
        M = np.ascontiguousarray(M, dtype=np.float64)
        if M.shape != (u.shape[0], v.shape[0]):
            raise ValueError("transpose(u)*M*v: inconsistent shapes")
    elif u.shape != v.shape:
        raise ValueError("transpose(u)*v: inconsistent shapes")
    out = C.c_double()
    engine._check(engine._lib.rb_bilinear(engine._h, u.shape[0], v.shape[0], L.dptr(u), L.dptr(M), L.dptr(v),
                                          C.byref(out)))
    return out.value


# ----------------------------------------------------------------------------------------------------------------------
# Matrix / Vector / Function (thin ndarray subclasses; the reference's wrappers forward arithmetic to ndarray too,
# rbnics/backends/online/basic/matrix.py:216-235)
# ----------------------------------------------------------------------------------------------------------------------
class _MatrixType(np.ndarray):
    @property
    def M(self):
        return self.shape[0]

    @property
    def N(self):
        return self.shape[1]


class _VectorType(np.ndarray):
    @property
    def N(self):
        return self.shape[0]


class _FunctionType:
    def __init__(self, vec):
        self._v = vec

    def vector(self):
        return self._v

    @property
    def N(self):
        return self._v.shape[0]


def Matrix(M, N=None):
    """Matrix(M, N) -> zero matrix; Matrix(ndarray) wraps (copy)."""
    if isinstance(M, np.ndarray):
        return np.array(M, dtype=np.float64).view(_MatrixType)
    return np.zeros((M, N)).view(_MatrixType)


def Vector(N):
    if isinstance(N, np.ndarray):
        return np.array(N, dtype=np.float64).view(_VectorType)
    return np.zeros(N).view(_VectorType)


def Function(N):
    if isinstance(N, _FunctionType):
        return _FunctionType(N._v.copy())
    return _FunctionType(Vector(N))


Matrix.Type = lambda: _MatrixType
Vector.Type = lambda: _VectorType
Function.Type = lambda: _FunctionType


# ----------------------------------------------------------------------------------------------------------------------
# Component dictionaries of block problems (Stokes: components u, s, p)
# rbnics/utils/io/online_size_dict.py, component_name_to_basis_component_index_dict.py: ordered dicts whose repr() is
# what the reference writes into component_name_to_basis_component_{index,length}.txt and content_item_shape.txt
# ----------------------------------------------------------------------------------------------------------------------
class OnlineSizeDict(dict):
    """component name -> number of basis functions (insertion-ordered, like the reference's OrderedDict)."""

    def __repr__(self):
        return "OnlineSizeDict(" + dict.__repr__(self) + ")"


class ComponentNameToBasisComponentIndexDict(dict):
    """component name -> position of the component's block."""

    def __repr__(self):
        return "ComponentNameToBasisComponentIndexDict(" + dict.__repr__(self) + ")"


_DICT_CTORS = {"OnlineSizeDict": OnlineSizeDict,
               "ComponentNameToBasisComponentIndexDict": ComponentNameToBasisComponentIndexDict}


def _parse_reference_text(text):
    """Safe replacement of the reference's eval() of its text files (rbnics/utils/io/text_io.py:26-44): literals plus
    the two dict constructors above; anything else raises ValueError."""
    import ast

    def convert(node):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in _DICT_CTORS \
                and len(node.args) <= 1 and not node.keywords:
            return _DICT_CTORS[node.func.id](convert(node.args[0]) if node.args else {})
        if isinstance(node, ast.Tuple):
            return tuple(convert(e) for e in node.elts)
        if isinstance(node, ast.List):
            return [convert(e) for e in node.elts]
        if isinstance(node, ast.Dict):
            return {convert(k): convert(v) for k, v in zip(node.keys, node.values)}
        return ast.literal_eval(node)

    return convert(ast.parse(text.strip(), mode="eval").body)


def _block_indices(stop, start, length_dict, index_dict):
    """Index set of storage[start:stop] along one axis when start / stop are component dicts
    (rbnics/backends/online/basic/wrapping/slice_to_array.py:27-45): the blocks are laid out in the order of
    index_dict with the lengths of length_dict; component c contributes offset_c + [start_c, stop_c)."""
    names = sorted(index_dict, key=lambda c: index_dict[c])
    lengths = [int(length_dict[c]) for c in names]
    offsets = np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(int)
    idx = []
    for c, off, ln in zip(names, offsets, lengths):
        lo = int(start[c]) if start is not None else 0
        hi = int(stop[c])
        if not (0 <= lo <= hi <= ln):
            raise ValueError("slice %r:%r outside component %r of length %d" % (lo, hi, c, ln))
        idx.extend(range(int(off) + lo, int(off) + hi))
    return np.array(idx, dtype=np.intp)


# ----------------------------------------------------------------------------------------------------------------------
# AffineExpansionStorage
# ----------------------------------------------------------------------------------------------------------------------
class AffineExpansionStorage:
    """Order-1 (Q,) or order-2 (Q1, Q2) storage of reduced operators: matrices, vectors or scalars.

    ``AffineExpansionStorage(Q)`` / ``(Q1, Q2)`` creates empty storage filled by ``storage[q] = item``;
    ``AffineExpansionStorage((op_0, op_1, ...))`` wraps a tuple.  ``storage[:N, :N]`` / ``storage[:N]`` slices
    every item (affine_expansion_storage.py:329-349); slices are cached like the reference's precomputed slices."""

    def __init__(self, arg1, arg2=None):
        if isinstance(arg1, (tuple, list)):
            self._content = np.empty((len(arg1),), dtype=object)
            for i, a in enumerate(arg1):
                self[i] = a
        elif arg2 is None:
            self._content = np.empty((int(arg1),), dtype=object)
        else:
            self._content = np.empty((int(arg1), int(arg2)), dtype=object)
        self._slices = {}
        self._stacked = None
        # per axis of the items: component name -> block position / block length (None: single block)
        self._component_name_to_basis_component_index = None
        self._component_name_to_basis_component_length = None

    def set_components(self, length_dicts, index_dicts=None):
        """Declare the block structure of the items: one OnlineSizeDict per item axis (matrix: rows, columns), as the
        reference's Matrix(M, N) / Vector(N) do when M, N are dicts (online/basic/matrix.py:32-65)."""
        if isinstance(length_dicts, dict):
            length_dicts = (length_dicts,)
        length_dicts = tuple(OnlineSizeDict(d) for d in length_dicts)
        if index_dicts is None:
            index_dicts = tuple(ComponentNameToBasisComponentIndexDict((c, i) for i, c in enumerate(d))
                                for d in length_dicts)
        elif isinstance(index_dicts, dict):
            index_dicts = (index_dicts,)
        self._component_name_to_basis_component_length = length_dicts
        self._component_name_to_basis_component_index = tuple(ComponentNameToBasisComponentIndexDict(d)
                                                              for d in index_dicts)
        self._slices = {}
        return self

    def order(self):
        return self._content.ndim

    def __len__(self):
        assert self.order() == 1
        return self._content.shape[0]

    @property
    def shape(self):
        return self._content.shape

    def __setitem__(self, key, item):
        if isinstance(item, Number):
            item = float(item)
        else:
            item = np.ascontiguousarray(item, dtype=np.float64)
        self._content[key] = item
        self._slices = {}
        self._stacked = None

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)) or (isinstance(key, tuple) and all(isinstance(k, (int, np.integer)) for k in key)):
            return self._content[key]
        ck = repr(key)
        if ck not in self._slices:
            keys = key if isinstance(key, tuple) else (key,)
            assert all(isinstance(k, slice) and k.step is None for k in keys)
            with_dicts = any(isinstance(k.stop, dict) for k in keys)
            out = AffineExpansionStorage.__new__(AffineExpansionStorage)
            out._content = np.empty(self._content.shape, dtype=object)
            out._component_name_to_basis_component_index = self._component_name_to_basis_component_index
            out._component_name_to_basis_component_length = self._component_name_to_basis_component_length
            if with_dicts:
                # block slicing (slice_to_array.py:27-45): storage[:N] / storage[:N, :N] with N = {"u": .., "s": .., "p": ..}
                lens, inds = self._component_name_to_basis_component_length, self._component_name_to_basis_component_index
                if lens is None or len(lens) != len(keys):
                    raise ValueError("component-dict slicing needs set_components() with one dict per item axis")
                index_sets = [_block_indices(k.stop, k.start if isinstance(k.start, dict) else None, lens[a], inds[a])
                              for a, k in enumerate(keys)]
                take = np.ix_(*index_sets)
                out._component_name_to_basis_component_length = tuple(
                    OnlineSizeDict((c, int(k.stop[c]) - (int(k.start[c]) if isinstance(k.start, dict) else 0))
                                   for c in sorted(inds[a], key=lambda c: inds[a][c])) for a, k in enumerate(keys))
            else:
                take = key
            for idx in np.ndindex(self._content.shape):
                it = self._content[idx]
                out._content[idx] = it if isinstance(it, float) else np.ascontiguousarray(it[take])
            out._slices = {}
            out._stacked = None
            self._slices[ck] = out
        return self._slices[ck]

    # -- on-disk format of the reference (rbnics/backends/online/basic/affine_expansion_storage.py:56-153,155-290;
    #    rbnics/utils/io/text_io.py:11-44 -> repr()/eval() text files; numpy_io.py:12-40 -> numpy.save):
    #    <directory>/<filename>/content_item_type.txt, content_item_shape.txt, content_item_<C-order index>.npy
    #    (.txt for scalars), component_name_to_basis_component_{index,length}.txt
    def save(self, directory, filename):
        import os
        full = os.path.join(str(directory), filename)
        os.makedirs(full, exist_ok=True)
        if self._content.size == 0:
            return
        first = self._content.flat[0]

        def text(name, value):
            with open(os.path.join(full, name + ".txt"), "w") as f:
                f.write(repr(value))

        if first is None:
            text("content_item_type", "empty")
            text("content_item_shape", None)
        elif isinstance(first, float):
            text("content_item_type", "scalar")
            text("content_item_shape", None)
            for k, idx in enumerate(np.ndindex(self._content.shape)):
                text("content_item_" + str(k), self._content[idx])
        else:
            lens = self._component_name_to_basis_component_length
            if first.ndim == 2:
                text("content_item_type", "matrix")
                text("content_item_shape", tuple(lens) if lens is not None
                     else (int(first.shape[0]), int(first.shape[1])))
            else:
                text("content_item_type", "vector")
                text("content_item_shape", lens[0] if lens is not None else int(first.shape[0]))
            for k, idx in enumerate(np.ndindex(self._content.shape)):
                np.save(os.path.join(full, "content_item_" + str(k) + ".npy"), np.asarray(self._content[idx]),
                        allow_pickle=False)

        def dicts(value, is_vector):
            # the reference stores one dict per item axis for matrices and a bare dict for vectors
            if value is None:
                return (None, None) if (not is_vector and first is not None and not isinstance(first, float)) else None
            return value[0] if is_vector else tuple(value)

        is_vector = isinstance(first, np.ndarray) and first.ndim == 1
        text("component_name_to_basis_component_index", dicts(self._component_name_to_basis_component_index, is_vector))
        text("component_name_to_basis_component_length", dicts(self._component_name_to_basis_component_length, is_vector))

    def load(self, directory, filename, globals=None):
        """Fill an (empty) storage of the right order/shape from the reference's directory layout.  Returns False if
        the storage already holds content (the reference avoids loading twice), True otherwise."""
        import os
        if any(it is not None for it in self._content.flat):
            return False
        full = os.path.join(str(directory), filename)
        if self._content.size == 0:
            return True

        def text(name):
            # the reference eval()s these files; only literals and its two dict classes can legitimately appear
            with open(os.path.join(full, name + ".txt")) as f:
                return _parse_reference_text(f.read())

        kind = text("content_item_type")
        assert kind in ("matrix", "vector", "function", "scalar", "empty"), kind
        shape = text("content_item_shape")
        for k, idx in enumerate(np.ndindex(self._content.shape)):
            if kind == "scalar":
                self._content[idx] = float(text("content_item_" + str(k)))
            elif kind == "empty":
                self._content[idx] = None
            else:
                arr = np.load(os.path.join(full, "content_item_" + str(k) + ".npy"), allow_pickle=False)
                expect = tuple(shape) if kind == "matrix" else (shape,)
                expect = tuple(builtins_sum(x.values()) if isinstance(x, dict) else int(x) for x in expect)
                assert arr.shape == expect, (arr.shape, expect)
                self._content[idx] = np.ascontiguousarray(arr, dtype=np.float64)

        def as_tuple(value):
            if value is None or (isinstance(value, tuple) and all(v is None for v in value)):
                return None
            return (value,) if isinstance(value, dict) else tuple(value)

        self._component_name_to_basis_component_index = as_tuple(text("component_name_to_basis_component_index"))
        self._component_name_to_basis_component_length = as_tuple(text("component_name_to_basis_component_length"))
        self._slices = {}
        self._stacked = None
        return True

    def stacked(self):
        """Contiguous (Q, ...) / (Q1, Q2, ...) FP64 tensor of all items: what the C ABI takes."""
        if self._stacked is None:
            first = self._content.flat[0]
            shp = () if isinstance(first, float) else first.shape
            out = np.empty(self._content.shape + shp)
            for idx in np.ndindex(self._content.shape):
                out[idx] = self._content[idx]
            self._stacked = out
        return self._stacked


# ----------------------------------------------------------------------------------------------------------------------
# product / sum
# ----------------------------------------------------------------------------------------------------------------------
class ProductOutput:
    def __init__(self, sum_product_return_value):
        self.sum_product_return_value = sum_product_return_value


def product(thetas, operators, thetas2=None):
    """``sum(product(thetas, operators[, thetas2]))`` assembles the affine expansion (the product carries out both,
    as in the reference)."""
    eng = get_engine()
    ops = operators.stacked()
    if operators.order() == 1:
        assert thetas2 is None
        assert len(thetas) == len(operators)
        out = affine_combine(eng, thetas, ops)
    else:
        assert thetas2 is not None
        assert (len(thetas), len(thetas2)) == operators.shape
        out = affine_combine2(eng, thetas, thetas2, ops if ops.ndim > 2 else ops[..., None])
        if ops.ndim == 2:
            out = float(out[0])
    if isinstance(out, np.ndarray):
        if out.ndim == 2:
            out = out.view(_MatrixType)
        elif out.ndim == 1:
            out = out.view(_VectorType)
        else:
            out = float(out)
    return ProductOutput(out)


def sum(product_output):  # noqa: A001  (the reference shadows the builtin in the same way)
    return product_output.sum_product_return_value


# ----------------------------------------------------------------------------------------------------------------------
# transpose
# ----------------------------------------------------------------------------------------------------------------------
class _Transposed:
    def __init__(self, u, M=None):
        self._u = u.vector() if isinstance(u, _FunctionType) else u
        self._M = M

    def __mul__(self, other):
        if isinstance(other, _FunctionType):
            other = other.vector()
        other = np.asarray(other)
        if self._M is None and other.ndim == 2:
            assert other.shape[0] == self._u.shape[0]
            return _Transposed(self._u, other)
        return bilinear(get_engine(), self._u, self._M, other)


def transpose(arg):
    return _Transposed(arg)


# ----------------------------------------------------------------------------------------------------------------------
# LinearSolver
# ----------------------------------------------------------------------------------------------------------------------
class LinearSolver:
    """``LinearSolver(lhs, solution, rhs, bcs=None)`` or ``LinearSolver(problem_wrapper, solution)`` with a wrapper
    exposing ``matrix_eval() / vector_eval() / bc_eval() / monitor(solution)``
    (rbnics/backends/abstract/linear_solver.py:29-44).  ``bcs`` is a tuple of theta_bc values: rows i < len(bcs) become
    identity rows with rhs ``bcs[i]`` (rbnics/backends/online/basic/wrapping/DirichletBC.py:37-56)."""

    def __init__(self, lhs, solution, rhs=None, bcs=None):
        if rhs is None and hasattr(lhs, "matrix_eval"):
            wrapper = lhs
            lhs, rhs, bcs = wrapper.matrix_eval(), wrapper.vector_eval(), wrapper.bc_eval()
            self.monitor = getattr(wrapper, "monitor", None)
        else:
            self.monitor = None
        self.lhs, self.rhs, self.solution = lhs, rhs, solution
        if bcs is not None and not isinstance(bcs, tuple):
            raise TypeError("bcs must be a tuple of theta values or None")
        self.bcs = bcs

    def set_parameters(self, parameters):
        assert len(parameters) == 0, "the B200 linear solver does not accept parameters"

    def solve(self):
        rows = list(range(len(self.bcs))) if self.bcs else None
        x = solve_dense(get_engine(), self.lhs, self.rhs, rows, self.bcs if self.bcs else None)
        self.solution.vector()[:] = x
        if self.monitor is not None:
            self.monitor(self.solution)


__all__ = ["AffineExpansionStorage", "ComponentNameToBasisComponentIndexDict", "Function", "LinearSolver", "Matrix",
           "OnlineSizeDict", "product", "sum", "transpose", "Vector"]

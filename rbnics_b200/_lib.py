"""ctypes binding of ``librbnics_b200.so`` (C ABI declared in ``include/rbnics_b200.h``).

The library is the product: if it is missing this module raises, it never falls back to numpy."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librbnics_b200.so")

RB_OK = 0
RB_ERR_CUDA = -1
RB_ERR_ARG = -2
RB_ERR_STATE = -3
RB_ERR_SINGULAR = -4
RB_ERR_UNSUPPORTED = -5

RB_EST_SQRT_EPS2_OVER_BETA = 0
RB_EST_SQRT_OF_EPS2_OVER_BETA = 1

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_float_p = C.POINTER(C.c_float)

# name -> (restype, argtypes); mirrors include/rbnics_b200.h one to one
SIGNATURES = {
    "rb_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "rb_ctx_destroy": (None, [C.c_void_p]),
    "rb_last_error": (C.c_char_p, [C.c_void_p]),
    "rb_device_info": (C.c_int, [C.c_void_p, c_int_p, c_int_p, c_int_p]),
    "rb_set_system": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, C.c_int, c_int_p]),
    "rb_set_estimator": (C.c_int, [C.c_void_p, C.c_int, c_int_p, c_int_p, c_int_p, C.c_int, C.c_int, c_double_p]),
    "rb_set_estimator_ex": (C.c_int, [C.c_void_p, C.c_int, c_int_p, c_int_p, c_int_p, C.c_int, C.c_int, c_double_p,
                                      C.c_int]),
    "rb_sweep_parabolic": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, c_double_p, c_double_p, c_int64_p,
                                     c_double_p, c_double_p, c_double_p]),
    "rb_sweep_parabolic_ic": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, c_double_p, c_double_p, c_double_p,
                                        c_double_p, c_int64_p, c_double_p, c_double_p, c_double_p]),
    "rb_compute_outputs": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p]),
    "rb_supremizers": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, C.c_int64, c_double_p, c_double_p,
                                 c_double_p]),
    "rb_set_training_set": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, c_double_p, C.c_int, c_double_p, c_double_p,
                                      C.c_int64, C.c_int64]),
    "rb_sweep_resident": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_int64_p, c_double_p, c_double_p, c_double_p]),
    "rb_sweep": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, c_double_p, C.c_int, c_double_p, c_double_p, C.c_int64,
                           C.c_int64, C.c_int, c_double_p, c_int64_p, c_double_p, c_double_p, c_double_p]),
    "rb_result_device_ptrs": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(c_double_p)]),
    "rb_last_timings": (C.c_int, [C.c_void_p, c_float_p, c_int_p]),
    "rb_estimator_work": (C.c_int, [C.c_void_p, c_int64_p, c_int_p]),
    "rb_affine_combine": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_double_p, c_double_p, c_double_p]),
    "rb_affine_combine2": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, c_double_p, c_double_p, c_double_p,
                                     c_double_p]),
    "rb_solve_dense": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, C.c_int, c_int_p, c_double_p,
                                 c_double_p]),
    "rb_bilinear": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "rb_host_alloc": (C.c_void_p, [C.c_int64]),
    "rb_host_free": (None, [C.c_void_p]),
    "rb_gram": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, c_double_p, c_double_p, c_int64_p, c_int32_p,
                          c_double_p, C.c_int64, C.c_int64, c_double_p]),
    "rb_lincomb": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]),
    "rb_gram_schmidt": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, c_double_p, c_int64_p, c_int32_p, c_double_p,
                                  c_double_p, c_double_p]),
    "rb_project": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, c_double_p, c_int64_p, c_int32_p, c_double_p,
                             c_int64_p, c_double_p]),
    "rb_off_csr": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_int64_p, c_int32_p, c_double_p]),
    "rb_off_block": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int]),
    "rb_off_append": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p]),
    "rb_off_append_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_int64, C.c_int64]),
    "rb_off_size": (C.c_int, [C.c_void_p, C.c_int, c_int64_p, c_int_p]),
    "rb_off_read": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p]),
    "rb_off_spmm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "rb_off_gram": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                              C.c_int64, c_double_p]),
    "rb_off_gram_schmidt": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_int, c_double_p]),
    "rb_set_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rb_off_lincomb": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, C.c_int, C.c_int, C.c_int]),
    "rb_off_lincomb_add": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, C.c_int, C.c_int, C.c_int,
                                     C.c_int]),
    "rb_off_column_norms2": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, c_double_p]),
    "rb_off_scale_columns": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p]),
    "rb_off_gram_schmidt_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_int, C.c_int64, C.c_int64,
                                           c_double_p]),
    "rb_off_gram_schmidt_from": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_int, C.c_int, C.c_int64,
                                           C.c_int64, c_double_p]),
}

# int (*rb_allreduce_fn)(void* user, double* device_buffer, int64_t count)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)

_lib = None


class B200LibraryError(RuntimeError):
    """The CUDA library is missing or a call failed (mirrors the reference raising Python exceptions)."""


class SingularReducedMatrixError(B200LibraryError):
    """A reduced matrix was singular: numpy.linalg.solve would have raised LinAlgError
    (rbnics/backends/online/numpy/linear_solver.py:29-31)."""


def load():
    """Load the shared library (once) and attach the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200LibraryError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C rbnics_b200/csrc`.  rbnics_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dptr(a):
    """numpy float64 array -> double*, or NULL for None."""
    if a is None:
        return None
    return a.ctypes.data_as(c_double_p)

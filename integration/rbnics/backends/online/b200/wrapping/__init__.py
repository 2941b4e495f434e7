"""`rbnics.backends.online.b200.wrapping`: the 15 names of rbnics/backends/online/numpy/wrapping/__init__.py:23-39.
The two products used by transpose() run on the device; I/O, slicing and the offline helpers are the numpy ones."""
from rbnics.backends.online.numpy.wrapping import (
    basis_functions_matrix_mul_online_matrix, basis_functions_matrix_mul_online_vector, function_load, function_save,
    function_to_vector, functions_list_mul_online_matrix, functions_list_mul_online_vector, get_mpi_comm,
    gram_schmidt_projection_step, Slicer, tensor_load, tensor_save, vectorized_matrix_inner_vectorized_matrix)


def matrix_mul_vector(matrix, vector):
    return matrix * vector  # _Matrix_Type.__mul__ -> rb_affine_combine


def vector_mul_vector(vector1, vector2):
    from rbnics_b200 import online as _dev
    return _dev.bilinear(_dev.get_engine(), vector1.content, None, vector2.content)


__all__ = [
    "basis_functions_matrix_mul_online_matrix",
    "basis_functions_matrix_mul_online_vector",
    "function_load",
    "function_save",
    "function_to_vector",
    "functions_list_mul_online_matrix",
    "functions_list_mul_online_vector",
    "get_mpi_comm",
    "gram_schmidt_projection_step",
    "matrix_mul_vector",
    "Slicer",
    "tensor_load",
    "tensor_save",
    "vector_mul_vector",
    "vectorized_matrix_inner_vectorized_matrix"
]

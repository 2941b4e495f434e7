"""A FEniCS-free TRUTH problem for the real reference: dense numpy operators served by RBniCS' own numpy online backend.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  With this module the UNMODIFIED RBniCS package runs its complete
offline stage here -- `ReducedBasis(problem).offline()`: greedy loop, truth solves, Gram-Schmidt, Galerkin projection,
Riesz representers, error-estimation operators -- so the greedy-selected parameters and the error estimators of the
real reference are available as golden data (tests/golden/make_reference_golden.py) and the `_greedy` override of
`integration/rbnics_b200_plugin` can be exercised inside the reference's own loop.

How: RBniCS dispatches on argument types, and its numpy online backend is a complete backend for "truth" problems
whose space is the span of a basis of a higher-level space (two-level reduction).  `DenseSpace(Nh)` is an `int` (so
`Function(V)`, `Matrix`, `Vector` build Nh-sized numpy containers) that is ALSO an `AbstractFunctionsList` (so
`BasisFunctionsMatrix(V)`, `GramSchmidt(V, X)` ... accept it as the root space R^Nh).  The helpers the numpy backend
leaves unimplemented for this use ("TODO" / missing) are filled in by `install_numpy_truth_backend_helpers()`, restated
from the dolfin wrappers of the same name; they are outside the hot path of this repository.
"""
from __future__ import annotations

import importlib
import pkgutil

import numpy as np


def make_dense_space_class():
    from mpi4py import MPI
    from rbnics.backends.abstract import FunctionsList as AbstractFunctionsList

    class DenseSpace(int, AbstractFunctionsList):
        """R^Nh as root space: an online size that the online backend also accepts as the basis of a higher level."""
        mpi_comm = MPI.COMM_WORLD

        def __new__(cls, n):
            return int.__new__(cls, n)

        def __init__(self, n):
            pass

        def __len__(self):
            return int(self)

    def _unavailable(self, *args, **kwargs):
        raise NotImplementedError("DenseSpace is a size, not a list of functions")

    for name in list(getattr(AbstractFunctionsList, "__abstractmethods__", ())):
        if name != "__len__":
            setattr(DenseSpace, name, _unavailable)
    DenseSpace.__abstractmethods__ = frozenset()
    return DenseSpace


def install_numpy_truth_backend_helpers():
    """(1) function_extend_or_restrict, same-space case of rbnics/backends/dolfin/wrapping/
    function_extend_or_restrict.py:38-56 (copy, times weight); (2) Z * u_N, "TODO" in
    rbnics/backends/online/numpy/wrapping/basis_functions_matrix_mul.py:11-13, restated from the dolfin wrapper
    (rbnics/backends/dolfin/wrapping/basis_functions_matrix_mul.py:37-53); used by the diagnostic prints of
    RBReduction._greedy (rb_reduction.py:167-170) and by error_analysis."""
    import rbnics.backends.online.numpy as numpy_backend
    import rbnics.backends.online.numpy.basis_functions_matrix as bfm
    from rbnics.backends.online import OnlineFunction
    from rbnics.backends.online.numpy.copy import function_copy

    def function_extend_or_restrict(function, function_components, V, V_components, weight, copy):
        assert function_components is None and V_components is None
        if not copy:
            assert weight is None
            return function
        out = function_copy(function)
        if weight is not None:
            out.vector().content[:] *= weight
        return out

    for m in pkgutil.iter_modules(numpy_backend.__path__):
        if m.ispkg:
            continue
        mod = importlib.import_module(numpy_backend.__name__ + "." + m.name)
        if hasattr(mod, "wrapping") and not hasattr(mod.wrapping, "function_extend_or_restrict"):
            mod.wrapping.function_extend_or_restrict = function_extend_or_restrict

    def basis_functions_matrix_mul_online_vector(basis_functions_matrix, online_vector):
        output = OnlineFunction(basis_functions_matrix.space)
        i = 0
        for component_name in basis_functions_matrix._components_name:
            for fun_i in basis_functions_matrix._components[component_name]:
                output.vector().content[:] += fun_i.vector().content * online_vector[i]
                i += 1
        return output

    bfm.wrapping.basis_functions_matrix_mul_online_vector = basis_functions_matrix_mul_online_vector

    # (3) S * eigenvector of the POD (proper_orthogonal_decomposition_base.py:80), "TODO" in
    # rbnics/backends/online/numpy/wrapping/functions_list_mul.py:12-13, restated from the dolfin wrapper
    # (rbnics/backends/dolfin/wrapping/functions_list_mul.py:26-38): sum_i s_i * v_i in list order
    import rbnics.backends.online.numpy.functions_list as fl

    def functions_list_mul_online_vector(functions_list, online_vector):
        output = OnlineFunction(functions_list.space)
        for (i, fun_i) in enumerate(functions_list):
            output.vector().content[:] += fun_i.vector().content * online_vector[i]
        return output

    fl.wrapping.functions_list_mul_online_vector = functions_list_mul_online_vector


def ram_cache_only():
    """`[problems] cache = RAM`: the default disk cache of truth solutions would write into the working directory."""
    from rbnics.utils.config import config
    config.set("problems", "cache", {"RAM"})


def make_thermal_block_problem(tb, compliant, theta_f_scale=1.0, decorator=None):
    """The reference's EllipticCoercive[Compliant]Problem over `tb` (tests/truth_problems.ThermalBlock): theta_a =
    (mu_0, ..., mu_{nb-2}, 1), theta_f = (mu_{nb-1}, ...), beta = min(theta_a) -- tutorials/01_thermal_block/
    tutorial_thermal_block.ipynb:158-169 -- with dense numpy operators."""
    from rbnics import EllipticCoerciveCompliantProblem, EllipticCoerciveProblem
    from rbnics.backends.online import OnlineMatrix, OnlineVector
    DenseSpace = make_dense_space_class()
    Nh, nb, nl = tb.Nh, tb.nblocks, tb.nloads

    def M(a):
        m = OnlineMatrix(Nh, Nh)
        m.content[:] = a.toarray() if hasattr(a, "toarray") else a
        return m

    def V(a):
        v = OnlineVector(Nh)
        v.content[:] = a
        return v

    Base = EllipticCoerciveCompliantProblem if compliant else EllipticCoerciveProblem

    class DenseThermalBlock(Base):
        def __init__(self, V, **kwargs):
            Base.__init__(self, V, **kwargs)

        def name(self):
            return "DenseThermalBlock"

        def compute_theta(self, term):
            mu = self.mu
            if term == "a":
                return tuple(mu[:nb - 1]) + (1.,)
            if term == "f":
                return tuple(mu[nb - 1:nb - 1 + nl])
            raise ValueError("Invalid term for compute_theta().")

        def get_stability_factor_lower_bound(self):
            return min(self.compute_theta("a"))

        def assemble_operator(self, term):
            if term == "a":
                return tuple(M(a) for a in tb.operators_a)
            if term == "f":
                return tuple(V(f) for f in tb.operators_f)
            if term == "inner_product":
                return (M(tb.inner_product),)
            raise ValueError("Invalid term for assemble_operator().")

    if decorator is not None:
        DenseThermalBlock = decorator(DenseThermalBlock)
    return DenseThermalBlock(DenseSpace(Nh))


def make_unsteady_thermal_block_problem(tb, decorator=None):
    """The reference's ParabolicCoerciveProblem over `tb` (tests/truth_problems.ThermalBlockUnsteady): terms m, a, f of
    tutorials/06_thermal_block_unsteady/tutorial_thermal_block_unsteady.ipynb with theta_m = (1,), homogeneous initial
    condition; time stepping and POD-Greedy are the reference's own (numpy online backend at truth level)."""
    from rbnics import ParabolicCoerciveProblem
    from rbnics.backends.online import OnlineMatrix, OnlineVector
    DenseSpace = make_dense_space_class()
    Nh, nb, nl = tb.Nh, tb.nblocks, tb.nloads

    def M(a):
        m = OnlineMatrix(Nh, Nh)
        m.content[:] = a.toarray() if hasattr(a, "toarray") else a
        return m

    def V(a):
        v = OnlineVector(Nh)
        v.content[:] = a
        return v

    class DenseThermalBlockUnsteady(ParabolicCoerciveProblem):
        def __init__(self, V, **kwargs):
            ParabolicCoerciveProblem.__init__(self, V, **kwargs)

        def name(self):
            return "DenseThermalBlockUnsteady"

        def compute_theta(self, term):
            mu = self.mu
            if term == "m":
                return (1.,)
            if term == "a":
                return tuple(mu[:nb - 1]) + (1.,)
            if term == "f":
                return tuple(mu[nb - 1:nb - 1 + nl])
            raise ValueError("Invalid term for compute_theta().")

        def get_stability_factor_lower_bound(self):
            return min(self.compute_theta("a"))

        def assemble_operator(self, term):
            if term == "m":
                return tuple(M(a) for a in tb.operators_m)
            if term == "a":
                return tuple(M(a) for a in tb.operators_a)
            if term == "f":
                return tuple(V(f) for f in tb.operators_f)
            if term in ("inner_product", "projection_inner_product"):
                return (M(tb.inner_product),)
            raise ValueError("Invalid term for assemble_operator().")

    if decorator is not None:
        DenseThermalBlockUnsteady = decorator(DenseThermalBlockUnsteady)
    problem = DenseThermalBlockUnsteady(DenseSpace(Nh))
    problem.set_mu_range(tb.mu_range)
    problem.set_time_step_size(tb.dt)
    problem.set_final_time(tb.dt * tb.K)
    return problem


def run_pod_greedy_offline(problem, ntrain, Nmax, N1, tol, seed):
    """ReducedBasis(parabolic problem).offline() with set_Nmax(Nmax, POD_Greedy=N1): the reference's POD-Greedy."""
    from rbnics import ReducedBasis
    rm = ReducedBasis(problem)
    rm.set_Nmax(Nmax, POD_Greedy=N1)
    rm.set_tolerance(tol, POD_Greedy=0.0)
    np.random.seed(seed)
    rm.initialize_training_set(ntrain)
    rp = rm.offline()
    return rm, rp


def run_offline(problem, mu_range, ntrain, Nmax, tol, seed):
    """ReducedBasis(problem).offline() of the reference; returns (reduction_method, reduced_problem)."""
    from rbnics import ReducedBasis
    problem.set_mu_range(mu_range)
    rm = ReducedBasis(problem)
    rm.set_Nmax(Nmax)
    rm.set_tolerance(tol)
    np.random.seed(seed)  # sampling uses the global numpy RNG (rbnics/sampling/distributions/uniform_distribution.py:11-19)
    rm.initialize_training_set(ntrain)
    rp = rm.offline()
    return rm, rp

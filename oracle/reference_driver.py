"""Drive the UNMODIFIED reference (RBniCS) through its own public classes, without FEniCS.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/, tests/golden/make_reference_golden.py and
`bench.py --impl reference`; nothing under rbnics_b200/ imports it.

RBniCS is pure Python.  It imports `multipledispatch`, `mpi4py`, `matplotlib`, `cvxopt` at module level, none of which is
in this image; `oracle/shims` holds stand-ins for the small surface it uses (oracle/shims/README.md).  The FEniCS
backend is only loaded when `dolfin` is already imported (rbnics/utils/config/config.py:45-51), so with the shims on the
path `import rbnics` succeeds and the ONLINE stage -- exactly the hot path of this repository -- runs for real:

    EllipticCoerciveRBReducedProblem / EllipticCoerciveCompliantRBReducedProblem
        .set_mu / .solve / .estimate_error            rbnics/problems/elliptic/*.py, problems/base/*.py
    OnlineAffineExpansionStorage, online_product, online_sum, OnlineLinearSolver, transpose
                                                      rbnics/backends/online/{basic,numpy}/*.py
    ParameterSpaceSubset.max + the closure of RBReduction._greedy
                                                      rbnics/sampling/parameter_space_subset.py:52-77,
                                                      rbnics/reduction_methods/base/rb_reduction.py:160-185

The truth problem (FEniCS forms, truth solves) is replaced by `SyntheticTruthProblem`, which answers what the reduced
problem asks of it online: name, mu, terms, Q, compute_theta, get_stability_factor_lower_bound.  The reduced operators
are handed over as arrays and wrapped in the reference's own containers.

The reference is looked for in `baseline/_ref` (pip --target install of /root/reference, travels to the GPU box) and
then in /root/reference.
"""
from __future__ import annotations

import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(_ROOT, "oracle", "shims")
_CANDIDATES = [os.path.join(_ROOT, "baseline", "_ref"), "/root/reference"]


def reference_root():
    for c in _CANDIDATES:
        if os.path.isdir(os.path.join(c, "rbnics")):
            return c
    return None


def setup_reference_import(backend="numpy", rbnicsrc_dir=None):
    """Put the shims, the reference and (for backend == "b200") the plug-in on sys.path.  Must run before the first
    `import rbnics`.  Returns the reference root or raises RuntimeError."""
    root = reference_root()
    if root is None:
        raise RuntimeError("the reference is neither under baseline/_ref nor under /root/reference")
    if "rbnics" in sys.modules:
        raise RuntimeError("rbnics is already imported")
    for p in (root, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    if _ROOT not in sys.path:
        sys.path.append(_ROOT)
    if backend == "b200":
        integ = os.path.join(_ROOT, "integration")
        if integ not in sys.path:
            sys.path.append(integ)
        import rbnics_b200_plugin
        rbnics_b200_plugin.register(write_rbnicsrc_to=rbnicsrc_dir)
    return root


def make_truth_problem_class():
    from rbnics.problems.base import ParametrizedProblem

    class SyntheticTruthProblem(ParametrizedProblem):
        """What a reduced problem needs from its truth problem ONLINE.  mu = (theta_a..., theta_f..., [beta]): the
        affine coefficients themselves are the parameters, so compute_theta is a slice of mu (the reference's
        compute_theta is user code, rbnics/problems/base/parametrized_reduced_differential_problem.py:669-676)."""

        def __init__(self, terms, terms_order, Q, components=("u",), theta_of_mu=None, beta_of_mu=None, V=1):
            ParametrizedProblem.__init__(self, "SyntheticTruthProblem")
            self.terms = list(terms)
            self.terms_order = dict(terms_order)
            self.components = list(components)
            self.Q = dict(Q)
            self.V = V
            self._theta_of_mu = theta_of_mu
            self._beta_of_mu = beta_of_mu
            self._combined_inner_product = None
            self._combined_and_homogenized_dirichlet_bc = None
            off, self._slices = 0, {}
            for t in self.terms:
                self._slices[t] = slice(off, off + self.Q[t])
                off += self.Q[t]

        def name(self):
            return "SyntheticTruthProblem"

        def compute_theta(self, term):
            if self._theta_of_mu is not None:
                return self._theta_of_mu(term, self.mu)
            return tuple(self.mu[self._slices[term]])

        def get_stability_factor_lower_bound(self):
            if self._beta_of_mu is not None:
                return self._beta_of_mu(self.mu)
            return min(self.compute_theta("a"))

    return SyntheticTruthProblem


def _tensor(arr, component="u"):
    from rbnics.backends.online import OnlineMatrix, OnlineVector
    arr = np.asarray(arr, dtype=np.float64)
    if arr.ndim == 2:
        t = OnlineMatrix({component: arr.shape[0]}, {component: arr.shape[1]})
    elif arr.ndim == 1:
        t = OnlineVector({component: arr.shape[0]})
    else:
        return float(arr)
    t.content[:] = arr
    return t


def storage(arrs):
    """Order-1 / order-2 OnlineAffineExpansionStorage of the reference around (Q, ...) / (Q1, Q2, ...) arrays."""
    from rbnics.backends.online import OnlineAffineExpansionStorage
    arrs = np.asarray(arrs, dtype=np.float64)
    return _storage_of(OnlineAffineExpansionStorage, arrs, order=1)


def _storage_of(cls, arrs, order):
    if order == 1:
        st = cls(arrs.shape[0])
        for q in range(arrs.shape[0]):
            st[q] = _tensor(arrs[q])
    else:
        st = cls(arrs.shape[0], arrs.shape[1])
        for i in range(arrs.shape[0]):
            for j in range(arrs.shape[1]):
                st[i, j] = _tensor(arrs[i, j])
    return st


def elliptic_reduced_problem(A, F, Gff, Gaf, Gaa, compliant=False):
    """The reference's EllipticCoercive[Compliant]RBReducedProblem filled with the given reduced operators
    (A: Qa x N x N, F: Qf x N, Gff: Qf x Qf, Gaf: Qa x Qf x N, Gaa: Qa x Qa x N x N)."""
    from rbnics.backends.online import OnlineAffineExpansionStorage
    from rbnics.problems.elliptic import EllipticCoerciveCompliantRBReducedProblem, EllipticCoerciveRBReducedProblem
    A, F = np.asarray(A), np.asarray(F)
    Qa, Qf, N = A.shape[0], F.shape[0], A.shape[1]
    Truth = make_truth_problem_class()
    tp = Truth(["a", "f"], {"a": 2, "f": 1}, {"a": Qa, "f": Qf})
    tp.set_mu_range([(0.0, 1.0)] * (Qa + Qf))
    cls = EllipticCoerciveCompliantRBReducedProblem if compliant else EllipticCoerciveRBReducedProblem
    rp = cls(tp)
    rp.N, rp.N_bc = N, 0
    rp.dirichlet_bc, rp.dirichlet_bc_are_homogeneous = False, False
    rp.Q = {"a": Qa, "f": Qf}
    rp.operator["a"] = _storage_of(OnlineAffineExpansionStorage, A, 1)
    rp.operator["f"] = _storage_of(OnlineAffineExpansionStorage, F, 1)
    if compliant:
        rp.Q["s"] = Qf  # compliant output: s = f (elliptic_coercive_compliant_reduced_problem.py)
    rp.error_estimation_operator["f", "f"] = _storage_of(OnlineAffineExpansionStorage, np.asarray(Gff), 2)
    rp.error_estimation_operator["a", "f"] = _storage_of(OnlineAffineExpansionStorage, np.asarray(Gaf), 2)
    rp.error_estimation_operator["a", "a"] = _storage_of(OnlineAffineExpansionStorage, np.asarray(Gaa), 2)
    return rp


def _block_tensor(arr, sizes):
    """Matrix / vector of the reference with COMPONENT-DICTIONARY sizes (Stokes: u, s, p) around a block-embedded array."""
    from rbnics.backends.online import OnlineMatrix, OnlineVector
    from rbnics.utils.io import OnlineSizeDict
    arr = np.asarray(arr, dtype=np.float64)
    if arr.ndim == 0:
        return float(arr)

    def N():
        d = OnlineSizeDict()
        for c, n in sizes:
            d[c] = n
        return d

    t = OnlineMatrix(N(), N()) if arr.ndim == 2 else OnlineVector(N())
    t.content[:] = arr
    return t


def stokes_reduced_problem(op, E, n):
    """The reference's StokesRBReducedProblem (components u, s, p with n basis functions each) filled with
    block-embedded reduced operators op[term] (Q x 3n x 3n / Q x 3n) and error-estimation storages E[(t1, t2)].
    mu = (theta_a..., theta_b..., theta_bt..., theta_f..., theta_g..., beta)."""
    from rbnics.backends.online import OnlineAffineExpansionStorage
    from rbnics.problems.stokes import StokesRBReducedProblem
    from rbnics.utils.io import OnlineSizeDict
    terms = ["a", "b", "bt", "f", "g"]
    Q = {t: int(np.asarray(op[t]).shape[0]) for t in terms}
    Truth = make_truth_problem_class()
    off, sl = 0, {}
    for t in terms:
        sl[t] = slice(off, off + Q[t])
        off += Q[t]
    tp = Truth(terms + ["bt_restricted"], {"a": 2, "b": 2, "bt": 2, "f": 1, "g": 1, "bt_restricted": 2},
               dict(Q, bt_restricted=Q["bt"]), components=("u", "s", "p"),
               theta_of_mu=lambda term, mu: tuple(mu[sl[term]]), beta_of_mu=lambda mu: mu[-1])
    tp.set_mu_range([(0.0, 1.0)] * (off + 1))
    rp = StokesRBReducedProblem(tp)
    sizes = (("u", n), ("s", n), ("p", n))

    def sd(v):
        d = OnlineSizeDict()
        for c, _ in sizes:
            d[c] = v
        return d

    rp.N, rp.N_bc = sd(n), sd(0)
    rp.dirichlet_bc = {c: False for c, _ in sizes}
    rp.dirichlet_bc_are_homogeneous = {c: False for c, _ in sizes}
    rp.Q = dict(Q)
    for t in terms:
        arr = np.asarray(op[t], dtype=np.float64)
        st = OnlineAffineExpansionStorage(arr.shape[0])
        for q in range(arr.shape[0]):
            st[q] = _block_tensor(arr[q], sizes)
        rp.operator[t] = st
    for (t1, t2), arr in E.items():
        arr = np.asarray(arr, dtype=np.float64)
        st = OnlineAffineExpansionStorage(arr.shape[0], arr.shape[1])
        for i in range(arr.shape[0]):
            for j in range(arr.shape[1]):
                st[i, j] = _block_tensor(arr[i, j], sizes)
        rp.error_estimation_operator[t1, t2] = st
    return rp


def parabolic_reduced_problem(M, A, F, E, dt, K):
    """The reference's ParabolicCoerciveRBReducedProblem (implicit Euler through its numpy TimeStepping, per-step error
    estimator with the m-terms) with homogeneous initial condition, final time K dt.
    mu = (theta_m..., theta_a..., theta_f..., beta)."""
    from rbnics.backends.online import OnlineAffineExpansionStorage
    from rbnics.problems.parabolic import ParabolicCoerciveRBReducedProblem
    M, A, F = np.asarray(M), np.asarray(A), np.asarray(F)
    Qm, Qa, Qf, N = M.shape[0], A.shape[0], F.shape[0], A.shape[1]
    sl = {"m": slice(0, Qm), "a": slice(Qm, Qm + Qa), "f": slice(Qm + Qa, Qm + Qa + Qf)}
    Truth = make_truth_problem_class()
    tp = Truth(["m", "a", "f"], {"m": 2, "a": 2, "f": 1}, {"m": Qm, "a": Qa, "f": Qf},
               theta_of_mu=lambda term, mu: tuple(mu[sl[term]]), beta_of_mu=lambda mu: mu[-1])
    tp.t, tp.t0, tp.dt, tp.T = 0., 0., float(dt), K * float(dt)
    tp._time_stepping_parameters = {}
    tp.set_time = lambda t: setattr(tp, "t", t)
    tp.set_initial_time = lambda t0: setattr(tp, "t0", t0)
    tp.set_time_step_size = lambda d: setattr(tp, "dt", d)
    tp.set_final_time = lambda T: setattr(tp, "T", T)
    tp.set_mu_range([(0.0, 1.0)] * (Qm + Qa + Qf + 1))
    rp = ParabolicCoerciveRBReducedProblem(tp)
    rp.N, rp.N_bc = N, 0
    rp.dirichlet_bc, rp.dirichlet_bc_are_homogeneous = False, False
    rp.initial_condition, rp.initial_condition_is_homogeneous = False, True
    rp.Q = {"m": Qm, "a": Qa, "f": Qf}
    for t, arr in (("m", M), ("a", A), ("f", F)):
        rp.operator[t] = _storage_of(OnlineAffineExpansionStorage, arr, 1)
    for (t1, t2), arr in E.items():
        rp.error_estimation_operator[t1, t2] = _storage_of(OnlineAffineExpansionStorage, np.asarray(arr), 2)
    rp._init_time_series("online")
    return rp


def pod_greedy_sweep(reduced_problem, ts, collect=None):
    """`training_set.max(solve_and_estimate_error)` with the closure of the POD-Greedy `_greedy`
    (rbnics/reduction_methods/base/time_dependent_rb_reduction.py:280-288): time quadrature of the squared error bound."""
    from math import sqrt
    from rbnics.backends import TimeQuadrature

    def solve_and_estimate_error(mu):
        reduced_problem.set_mu(mu)
        reduced_problem.solve()
        error_estimator_over_time = reduced_problem.estimate_error()
        error_estimator_squared_over_time = [v**2 for v in error_estimator_over_time]
        time_quadrature = TimeQuadrature((0., reduced_problem.truth_problem.T), error_estimator_squared_over_time)
        error_estimator = sqrt(time_quadrature.integrate())
        if collect is not None:
            collect.setdefault("delta", []).append(float(error_estimator))
        return error_estimator

    return ts.max(solve_and_estimate_error)


def training_set(Theta):
    """The reference's ParameterSpaceSubset holding the rows of Theta as tuples of Python floats."""
    from rbnics.sampling import ParameterSpaceSubset
    ts = ParameterSpaceSubset()
    ts._list = [tuple(float(x) for x in row) for row in np.asarray(Theta)]
    return ts


def greedy_sweep(reduced_problem, ts, collect=None):
    """`training_set.max(solve_and_estimate_error)` with the closure of RBReduction._greedy
    (rbnics/reduction_methods/base/rb_reduction.py:173-185): returns (max, argmax) exactly as the reference does.
    `collect`: optional dict receiving per-parameter 'u' and 'delta' lists."""
    def solve_and_estimate_error(mu):
        reduced_problem.set_mu(mu)
        u = reduced_problem.solve()
        error_estimator = reduced_problem.estimate_error()
        if collect is not None:
            collect.setdefault("u", []).append(np.array(u.vector().content, dtype=np.float64))
            collect.setdefault("delta", []).append(float(error_estimator))
        return error_estimator

    return ts.max(solve_and_estimate_error)


def disable_reduced_solution_cache():
    """`[reduced problems] cache =` (empty): the RAM-unlimited default memoises every reduced solution
    (rbnics/utils/config/config.py:33-36) and would grow without bound in a 10^6-parameter sweep."""
    from rbnics.utils.config import config
    config.set("reduced problems", "cache", set())

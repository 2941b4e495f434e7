"""The batched greedy sweep as a RBniCS decorator: `_greedy` of the reduction method is overridden, nothing else.

    from rbnics_b200_plugin.sweep import B200Sweep

    @B200Sweep()
    class ThermalBlock(EllipticCoerciveCompliantProblem):
        ...
    reduction_method = ReducedBasis(problem)       # the factory composes the decorated reduction method
    reduction_method.offline()                     # RBniCS' own offline loop; every `training_set.max(closure)`
                                                   # becomes ONE sweep on the B200

Hooks used (all the reference's own): `ProblemDecoratorFor` / `ReductionMethodDecoratorFor`
(rbnics/utils/decorators/problem_decorator_for.py, reduction_method_decorator_for.py; composition in
rbnics/utils/factories/reduction_method_factory.py:34-61; precedent overriding `_greedy`:
tutorials/10_weighted_uq/reduction_methods/weighted_uncertainty_quantification_decorated_reduction_method.py:47-82).

What the override replaces, per problem family:
  elliptic (coercive, compliant)  rbnics/reduction_methods/base/rb_reduction.py:160-185,
                                  rbnics/problems/elliptic/elliptic_rb_reduced_problem.py:35-64,
                                  elliptic_coercive_compliant_rb_reduced_problem.py:22-31
  Stokes                          rbnics/reduction_methods/stokes/stokes_rb_reduction.py:16-65 (same closure),
                                  rbnics/problems/stokes/stokes_rb_reduced_problem.py:35-83 (9 term pairs)
  parabolic (POD-Greedy)          rbnics/reduction_methods/base/time_dependent_rb_reduction.py:262-295,
                                  rbnics/problems/parabolic/abstract_parabolic_rb_reduced_problem.py:29-92
`greedy()` (rb_reduction.py:144-158) keeps its bookkeeping on the returned (max, argmax) pair.

theta(mu) / beta(mu) are evaluated for the WHOLE training set once per training set (they do not change during the
greedy loop).  `compute_theta` is user code written for one mu; it is first tried on a tuple of COLUMN ARRAYS (every
arithmetic expression of the coordinates vectorises as is), checked against per-mu calls, and only falls back to the
per-mu loop when that fails (e.g. a `min(...)` over thetas).
"""
import os

import numpy as np
from rbnics.utils.decorators import (PreserveClassName, ProblemDecoratorFor, ReducedProblemDecoratorFor,
                                     ReductionMethodDecoratorFor)

from rbnics_b200 import _lib as L
from rbnics_b200.engine import SweepEngine
from rbnics_b200.sweep import DistributedMaxLoc, TrainingSetSweep


# ----------------------------------------------------------------------------------------------------------------------
# theta / beta for a whole training set
# ----------------------------------------------------------------------------------------------------------------------
def _per_mu(problem, training_set, fn):
    old = problem.mu
    try:
        rows = []
        for mu in training_set:
            problem.mu = mu
            rows.append(np.atleast_1d(np.asarray(fn(), dtype=np.float64)))
        return np.stack(rows, axis=0)
    finally:
        problem.mu = old


def batched(problem, training_set, fn, n_check=4):
    """(B x Q) array of `fn()` (a tuple or a scalar depending on problem.mu) over the training set."""
    mus = np.asarray([tuple(mu) for mu in training_set], dtype=np.float64)
    B = mus.shape[0]
    if B == 0 or mus.ndim != 2:
        return _per_mu(problem, training_set, fn)
    old = problem.mu
    try:
        problem.mu = tuple(mus[:, k] for k in range(mus.shape[1]))
        out = fn()
        out = out if isinstance(out, (tuple, list)) else (out,)
        table = np.stack([np.broadcast_to(np.asarray(c, dtype=np.float64), (B,)) for c in out], axis=1)
        for i in np.unique(np.linspace(0, B - 1, n_check).astype(int)):  # the vectorised evaluation must be the scalar one
            problem.mu = training_set[int(i)]
            if not np.array_equal(np.atleast_1d(np.asarray(fn(), dtype=np.float64)), table[i]):
                raise ValueError("vectorised evaluation differs")
        return np.ascontiguousarray(table)
    except Exception:
        return _per_mu(problem, training_set, fn)
    finally:
        problem.mu = old


# ----------------------------------------------------------------------------------------------------------------------
# reduced operators: the reference's storages -> contiguous FP64 tensors
# ----------------------------------------------------------------------------------------------------------------------
def _stack(storage):
    if hasattr(storage, "b200_stacked"):  # B200 online backend: cached
        st = storage.b200_stacked()
        if st is not None:
            return st
    content = storage._content
    first = content.flat[0]
    shape = () if isinstance(first, (int, float)) else first.content.shape
    out = np.empty(content.shape + shape)
    for idx in np.ndindex(content.shape):
        item = content[idx]
        out[idx] = item if isinstance(item, (int, float)) else item.content
    return out


def _online_size(rp):
    """N (+ N_bc) as the reference's solve() computes it (parametrized_reduced_differential_problem.py:328-329)."""
    N, _ = rp._online_size_from_kwargs(None)
    N += rp.N_bc
    return N


def _total(N):
    return int(sum(N.values())) if isinstance(N, dict) else int(N)


# ----------------------------------------------------------------------------------------------------------------------
# one object per reduction method: engine, training-set tables, the sweep
# ----------------------------------------------------------------------------------------------------------------------
class GreedySweep(object):
    def __init__(self, reduction_method, device=None):
        self.rm = reduction_method
        ts = reduction_method.training_set
        comm = getattr(ts, "mpi_comm", None)
        size = getattr(comm, "size", 1)
        rank = getattr(comm, "rank", 0)
        self.dist = None
        if size > 1 and getattr(ts, "distributed_max", True):
            os.environ.setdefault("RANK", str(rank))
            os.environ.setdefault("WORLD_SIZE", str(size))
            self.dist = DistributedMaxLoc(rank % _device_count())
        self.engine = SweepEngine(device if device is not None else rank % _device_count())
        self._tables = None
        self._sweep = None
        self._key = None

    def _family(self, rp):
        if "m" in rp.terms:
            return "parabolic"
        if "bt" in rp.terms:
            return "stokes"
        return "elliptic"

    def _prepare_tables(self, rp):
        tp, ts = self.rm.truth_problem, self.rm.training_set
        key = (id(ts), len(ts))
        if self._key == key:
            return
        fam = self._family(rp)
        terms = {"elliptic": ("a", "f"), "stokes": ("a", "b", "bt", "f", "g"), "parabolic": ("m", "a", "f")}[fam]
        theta = {t: batched(tp, ts, lambda t=t: tp.compute_theta(t)) for t in terms}
        beta = batched(tp, ts, tp.get_stability_factor_lower_bound)[:, 0]
        self._tables = (fam, terms, theta, beta)
        self._sweep = None
        self._key = key

    def max(self, rp):
        """(max estimator, arg-max index) over the training set == training_set.max(solve_and_estimate_error)."""
        self._prepare_tables(rp)
        fam, terms, theta, beta = self._tables
        N = _online_size(rp)
        if _total(N) == 0 and fam == "stokes":
            raise NotImplementedError("B200Sweep: empty reduced space for " + fam)
        eng = self.engine
        if fam == "parabolic":
            if _total(N) == 0:
                return self._max_parabolic_empty(rp, theta, beta)
            return self._max_parabolic(rp, N, theta, beta)
        key2 = tuple(slice(None, N) for _ in range(2))
        op = {t: _stack(rp.operator[t][key2 if rp.terms_order[t] == 2 else key2[0]]) for t in terms}
        e = rp.error_estimation_operator
        if fam == "elliptic":
            eng.set_system(op["a"], op["f"], bc_rows=self._bc_rows(rp))
            eng.set_elliptic_estimator(_stack(e["f", "f"]), _stack(e["a", "f"][key2[0]]), _stack(e["a", "a"][key2]))
            compliant = "Compliant" in type(rp).__name__
            kind = L.RB_EST_SQRT_OF_EPS2_OVER_BETA if compliant else L.RB_EST_SQRT_EPS2_OVER_BETA
        else:  # Stokes: every lhs / rhs term is block-embedded in the (u, s, p) system (stokes_reduced_problem.py:39-53)
            A = np.concatenate([op["a"], op["b"], op["bt"]])
            F = np.concatenate([op["f"], op["g"]])
            eng.set_system(A, F, bc_rows=self._bc_rows(rp))
            E = {}
            for (t1, t2) in rp.error_estimation_terms:
                o1, o2 = rp.terms_order[t1], rp.terms_order[t2]
                k = key2 if (o1 == 2 and o2 == 2) else (key2[0] if (o1 == 2 or o2 == 2) else None)
                E[t1, t2] = _stack(e[t1, t2][k] if k is not None else e[t1, t2])
            eng.set_stokes_estimator(E, {t: op[t].shape[0] for t in terms})
            kind = L.RB_EST_SQRT_OF_EPS2_OVER_BETA
        if self._sweep is None:
            Theta = np.concatenate([theta[t] for t in terms], axis=1)
            self._sweep = TrainingSetSweep(eng, Theta, beta, ThetaBC=self._theta_bc(rp), dist=self.dist)
        return self._sweep.max(kind=kind)

    def _bc_rows(self, rp):
        """Lifting rows: the first N_bc rows of every component block (online/basic/wrapping/DirichletBC.py:65-98)."""
        if not rp.N_bc or _total(rp.N_bc) == 0:
            return None
        N = _online_size(rp)
        if not isinstance(N, dict):
            return list(range(int(rp.N_bc)))
        rows, off = [], 0
        for c in rp.components:
            nbc = rp.N_bc[c] if isinstance(rp.N_bc, dict) else 0
            rows.extend(range(off, off + int(nbc)))
            off += int(N[c])
        return rows or None

    def _theta_bc(self, rp):
        if self._bc_rows(rp) is None:
            return None
        tp, ts = self.rm.truth_problem, self.rm.training_set
        if len(rp.components) > 1:
            cols = [batched(tp, ts, lambda c=c: tp.compute_theta("dirichlet_bc_" + c)) for c in rp.components
                    if rp.dirichlet_bc[c] and not rp.dirichlet_bc_are_homogeneous[c]]
            return np.concatenate(cols, axis=1)
        return batched(tp, ts, lambda: tp.compute_theta("dirichlet_bc"))

    def _max_parabolic_empty(self, rp, theta, beta):
        """First POD-Greedy iteration (N = 0): no reduced solution yet, the residual is f at every t > t0, so
        eps2 = theta_f^T G_ff theta_f does not depend on time; one sweep of the empty elliptic system gives
        sqrt(|eps2| / beta), and the Simpson quadrature of a constant (zero at t0) is the sum of its weights."""
        from rbnics_b200.parabolic import simpson_weights
        eng = self.engine
        Qa, Qf = theta["a"].shape[1], theta["f"].shape[1]
        eng.set_system(np.zeros((Qa, 0, 0)), np.zeros((Qf, 0)))
        eng.set_estimator([-1], [0], [Qf], Qa, Qf, _stack(rp.error_estimation_operator["f", "f"]))
        K = int(round((rp.T - rp.t0) / rp.dt))
        comm = self.rm.training_set.mpi_comm
        Theta = np.ascontiguousarray(np.concatenate([theta["a"], theta["f"]], axis=1)[comm.rank::comm.size])
        r = eng.sweep(Theta, np.ascontiguousarray(beta[comm.rank::comm.size]), kind=L.RB_EST_SQRT_OF_EPS2_OVER_BETA,
                      index_offset=comm.rank, index_stride=comm.size)
        v, i = r["max"] * float(np.sqrt(simpson_weights(K, rp.dt)[1:].sum())), r["argmax"]
        if self.dist is not None:
            v, i = self.dist.maxloc(v, i)
        return v, i

    def _max_parabolic(self, rp, N, theta, beta):
        """POD-Greedy: one implicit-Euler trajectory per parameter, estimator per step, Simpson quadrature."""
        from rbnics_b200.parabolic import ParabolicSweep
        key2 = (slice(None, N), slice(None, N))
        M, A = _stack(rp.operator["m"][key2]), _stack(rp.operator["a"][key2])
        F = _stack(rp.operator["f"][key2[0]])
        e = rp.error_estimation_operator
        E = {}
        for (t1, t2) in rp.error_estimation_terms:
            o1, o2 = rp.terms_order[t1], rp.terms_order[t2]
            k = key2 if (o1 == 2 and o2 == 2) else (key2[0] if (o1 == 2 or o2 == 2) else None)
            E[t1, t2] = _stack(e[t1, t2][k] if k is not None else e[t1, t2])
        K = int(round((rp.T - rp.t0) / rp.dt))
        sweep = ParabolicSweep(self.engine, M, A, F, E, rp.dt, K, bc_rows=self._bc_rows(rp))
        comm = self.rm.training_set.mpi_comm
        r = sweep.sweep(theta["m"][comm.rank::comm.size], theta["a"][comm.rank::comm.size],
                        theta["f"][comm.rank::comm.size], beta[comm.rank::comm.size],
                        index_offset=comm.rank, index_stride=comm.size)
        v, i = r["max"], r["argmax"]
        if self.dist is not None:
            v, i = self.dist.maxloc(v, i)
        return v, i

    def close(self):
        self.engine.close()
        if self.dist is not None:
            self.dist.close()


def _device_count():
    import ctypes
    n = ctypes.c_int(0)
    try:
        ctypes.CDLL("libcudart.so").cudaGetDeviceCount(ctypes.byref(n))
    except OSError:
        return 1
    return max(1, n.value)


# ----------------------------------------------------------------------------------------------------------------------
# the decorators
# ----------------------------------------------------------------------------------------------------------------------
def B200Sweep(**decorator_kwargs):
    @ProblemDecoratorFor(B200Sweep)
    def B200Sweep_Decorator(ParametrizedDifferentialProblem_DerivedClass):
        return ParametrizedDifferentialProblem_DerivedClass  # marker only: the truth problem is untouched
    return B200Sweep_Decorator


@ReducedProblemDecoratorFor(B200Sweep)
def B200SweepDecoratedReducedProblem(ParametrizedReducedDifferentialProblem_DerivedClass):
    return ParametrizedReducedDifferentialProblem_DerivedClass  # the reduced problem is untouched as well


@ReductionMethodDecoratorFor(B200Sweep)
def B200SweepDecoratedReductionMethod(DifferentialProblemReductionMethod_DerivedClass):

    @PreserveClassName
    class B200SweepDecoratedReductionMethod_Class_Base(DifferentialProblemReductionMethod_DerivedClass):
        def __init__(self, truth_problem, **kwargs):
            DifferentialProblemReductionMethod_DerivedClass.__init__(self, truth_problem, **kwargs)
            self._b200_greedy_sweep = None

    if hasattr(DifferentialProblemReductionMethod_DerivedClass, "greedy"):  # RB reduction (not POD-Galerkin)
        @PreserveClassName
        class B200SweepDecoratedReductionMethod_Class(B200SweepDecoratedReductionMethod_Class_Base):
            def _greedy(self):
                if _total(self.reduced_problem.N) > 0:  # same consistency prints as the reference (rb_reduction.py:167-170)
                    print("absolute error for current mu =", self.reduced_problem.compute_error())
                    print("absolute error estimator for current mu =", self.reduced_problem.estimate_error())
                print("find initial mu" if _total(self.reduced_problem.N) == 0 else "find next mu")
                if self._b200_greedy_sweep is None:
                    self._b200_greedy_sweep = GreedySweep(self)
                return self._b200_greedy_sweep.max(self.reduced_problem)

        return B200SweepDecoratedReductionMethod_Class
    return B200SweepDecoratedReductionMethod_Class_Base

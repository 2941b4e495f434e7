#!/usr/bin/env python
"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN FUNCTION BODIES for the hot path.

The reference (RBniCS) cannot be imported as a package in the build container (multipledispatch, mpi4py, dolfin are
absent), so this script loads the individual source files that hold the arithmetic of the path from
``/root/reference`` and executes them with the missing infrastructure stubbed out:

  * ``rbnics.utils.decorators.overload`` -> a tiny isinstance-based dispatcher (the reference's is multipledispatch)
  * ``mpi4py.MPI.COMM_WORLD``            -> an in-process fake communicator (threads + barrier) for 1..4 "ranks"
  * online Matrix / Vector wrappers      -> plain numpy arrays (the wrappers forward ``*``/``+=`` to
                                            ``ndarray.__rmul__/__iadd__``, rbnics/backends/online/basic/matrix.py:216-235)

Executed reference code (unmodified source text, read at run time):
  rbnics/backends/online/basic/product.py            (order-1 / order-2 affine products: the loops and zero skipping)
  rbnics/backends/online/basic/sum.py
  rbnics/backends/online/basic/wrapping/DirichletBC.py (lifting rows)
  rbnics/problems/elliptic/elliptic_rb_reduced_problem.py      get_residual_norm_squared / estimate_error (source of
  rbnics/problems/elliptic/elliptic_coercive_compliant_rb_reduced_problem.py   the two methods, run on a stub ``self``)
  rbnics/sampling/parameter_space_subset.py          ParameterSpaceSubset.max (loop, cyclic shard, numpy.argmax)
  rbnics/utils/mpi/parallel_max.py                   (max-loc over ranks)
``numpy.linalg.solve`` is the reference's linear solver itself (rbnics/backends/online/numpy/linear_solver.py:29-31).

Output: tests/golden/reference_sweep_*.npz -- inputs and outputs; committed.  Run:  python tests/golden/make_golden.py
"""
from __future__ import annotations

import ast
import inspect
import os
import sys
import textwrap
import threading
import types
from math import sqrt
from numbers import Number

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------------------------------------------------------------
# stubs
# ----------------------------------------------------------------------------------------------------------------------
class _TupleOfNumbers(tuple):
    pass


ThetaType = (tuple,)
DictOfThetaType = (dict,)


def _matches(arg, spec):
    if spec is None:
        return arg is None
    if isinstance(spec, tuple):
        return any(_matches(arg, s) for s in spec)
    if spec is object:
        return True
    return isinstance(arg, spec)


_REGISTRY = {}


def overload(*types_or_fn):
    """isinstance-based stand-in for rbnics.utils.decorators.overload (both ``@overload(T1, T2)`` and bare
    ``@overload`` with annotations)."""

    def register(fn, types):
        key = (fn.__module__, fn.__qualname__)
        entries = _REGISTRY.setdefault(key, [])
        entries.append((types, fn))
        is_method = "." in fn.__qualname__ and "<locals>" not in fn.__qualname__.split(".")[-2:]

        def dispatcher(*args, **kwargs):
            call_args = args[1:] if _looks_like_method(fn) else args
            for tys, f in entries:
                padded = list(call_args) + [None] * (len(tys) - len(call_args))
                if len(call_args) <= len(tys) and all(_matches(a, t) for a, t in zip(padded, tys)):
                    return f(*args, **kwargs)
            raise TypeError("no overload of %s for %r" % (fn.__qualname__, [type(a) for a in call_args]))

        dispatcher.__name__ = fn.__name__
        dispatcher.__qualname__ = fn.__qualname__
        return dispatcher

    if len(types_or_fn) == 1 and inspect.isfunction(types_or_fn[0]):
        fn = types_or_fn[0]
        params = list(inspect.signature(fn).parameters.values())
        if _looks_like_method(fn):
            params = params[1:]
        tys = tuple(object if p.annotation is inspect.Parameter.empty else p.annotation for p in params)
        return register(fn, tys)
    return lambda fn: register(fn, tuple(types_or_fn))


def _looks_like_method(fn):
    params = list(inspect.signature(fn).parameters)
    return bool(params) and params[0] == "self"


class Storage:
    """numpy-object-array stand-in for OnlineAffineExpansionStorage (order 1 or 2, int / slice indexing)."""

    def __init__(self, content):
        self._content = np.empty(np.shape(content)[:self._order_of(content)], dtype=object)
        if self._content.ndim == 1:
            for i, c in enumerate(content):
                self._content[i] = c
        else:
            for i, row in enumerate(content):
                for j, c in enumerate(row):
                    self._content[i, j] = c

    @staticmethod
    def _order_of(content):
        first = content[0]
        if isinstance(first, (list, tuple)):
            return 2
        return 1

    def order(self):
        return self._content.ndim

    def __len__(self):
        assert self.order() == 1
        return self._content.shape[0]

    def __getitem__(self, key):
        if isinstance(key, int) or (isinstance(key, tuple) and all(isinstance(k, int) for k in key)):
            return self._content[key]
        # slicing of every item (affine_expansion_storage.py:329-349)
        out = Storage.__new__(Storage)
        out._content = np.empty(self._content.shape, dtype=object)
        for idx in np.ndindex(self._content.shape):
            item = self._content[idx]
            out._content[idx] = item if isinstance(item, Number) else np.array(item[key])
        return out


class NonAffineStorage:
    pass


class Solution(np.ndarray):
    """reduced solution vector with the ``N`` attribute the estimator reads (``self._solution.N``)."""

    @property
    def N(self):
        return self.shape[0]


class Transpose:
    """transpose(u) * M * v == dot(u, M @ v); transpose(u) * w == dot(u, w)
    (rbnics/backends/basic/transpose.py:126-237 with the numpy wrapping functions vector_mul.py / matrix.py:31-41)."""

    def __init__(self, u, M=None):
        self.u, self.M = np.asarray(u), M

    def __mul__(self, other):
        other = np.asarray(other)
        if self.M is not None:
            return float(np.dot(self.u, self.M @ other))
        if other.ndim == 2:
            return Transpose(self.u, other)
        return float(np.dot(self.u, other))


def transpose(u):
    return Transpose(u)


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _exec_reference_file(relpath, modname):
    src = open(os.path.join(REF, relpath)).read()
    mod = types.ModuleType(modname)
    mod.__file__ = os.path.join(REF, relpath)
    sys.modules[modname] = mod
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod


def _method_source(relpath, name):
    """Source text of a method of the reference, de-indented to a plain function."""
    src = open(os.path.join(REF, relpath)).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            lines = src.splitlines()[node.lineno - 1:node.end_lineno]
            return textwrap.dedent("\n".join(lines))
    raise KeyError(name)


# ----------------------------------------------------------------------------------------------------------------------
# fake MPI (threads) for parallel_max / ParameterSpaceSubset.max
# ----------------------------------------------------------------------------------------------------------------------
class FakeComm:
    def __init__(self, rank, size, shared):
        self.rank, self.size, self._s = rank, size, shared

    def _exchange(self, value):
        s = self._s
        s["slots"][self.rank] = value
        s["barrier"].wait()
        vals = list(s["slots"])
        s["barrier"].wait()
        return vals

    def allreduce(self, value, op=None):
        vals = self._exchange(value)
        assert op == "MAX"
        return max(vals)

    def bcast(self, value, root=0):
        return self._exchange(value)[root]


def run_ranks(size, fn):
    shared = {"slots": [None] * size, "barrier": threading.Barrier(size)}
    out = [None] * size

    def work(r):
        out[r] = fn(FakeComm(r, size, shared))

    ts = [threading.Thread(target=work, args=(r,)) for r in range(size)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return out


# ----------------------------------------------------------------------------------------------------------------------
def load_reference():
    ndarray = np.ndarray

    class _TypeHolder:
        def __init__(self, t):
            self._t = t

        def Type(self):
            return self._t

    backend = types.SimpleNamespace(AffineExpansionStorage=Storage, NonAffineExpansionStorage=NonAffineStorage,
                                    Matrix=_TypeHolder(ndarray), Vector=_TypeHolder(ndarray),
                                    Function=_TypeHolder(ndarray))
    _module("rbnics")
    _module("rbnics.backends")
    _module("rbnics.backends.abstract", ParametrizedTensorFactory=type("PTF", (), {}))
    _module("rbnics.backends.basic")
    _module("rbnics.backends.basic.wrapping", DelayedTranspose=type("DelayedTranspose", (), {}))
    _module("rbnics.utils")
    _module("rbnics.utils.decorators", overload=overload, ThetaType=ThetaType, DictOfThetaType=DictOfThetaType)
    _module("rbnics.utils.io", ComponentNameToBasisComponentIndexDict=dict, OnlineSizeDict=dict,
            ExportableList=type("ExportableList", (), {"__init__": lambda self, kind: setattr(self, "_list", []),
                                                       "__len__": lambda self: len(self._list)}))
    MPI = _module("mpi4py.MPI", COMM_WORLD=FakeComm(0, 1, {"slots": [None], "barrier": threading.Barrier(1)}),
                  MAX="MAX")
    _module("mpi4py", MPI=MPI)
    prod_mod = _exec_reference_file("rbnics/backends/online/basic/product.py", "ref_product")
    sum_mod = _exec_reference_file("rbnics/backends/online/basic/sum.py", "ref_sum")
    bc_mod = _exec_reference_file("rbnics/backends/online/basic/wrapping/DirichletBC.py", "ref_dirichlet")
    pmax_mod = _exec_reference_file("rbnics/utils/mpi/parallel_max.py", "ref_parallel_max")
    _module("rbnics.utils.mpi", parallel_io=lambda f, comm=None: f(), parallel_max=pmax_mod.parallel_max)
    _module("rbnics.sampling")
    _module("rbnics.sampling.distributions", CompositeDistribution=object, UniformDistribution=object)
    pss_mod = _exec_reference_file("rbnics/sampling/parameter_space_subset.py", "ref_pss")

    wrapping = types.SimpleNamespace(DelayedTransposeWithArithmetic=type("DTWA", (), {}))
    product_base, ProductOutput = prod_mod.product(backend, wrapping)

    def product(thetas, operators, thetas2=None):  # rbnics/backends/online/numpy/product.py:23-25
        return product_base(thetas, operators, thetas2)

    ns = {"product": product, "sum": sum_mod.sum, "transpose": transpose, "sqrt": sqrt, "isclose": np.isclose,
          "abs": abs}
    exec(_method_source("rbnics/problems/elliptic/elliptic_rb_reduced_problem.py", "get_residual_norm_squared"), ns)
    exec(_method_source("rbnics/problems/elliptic/elliptic_rb_reduced_problem.py", "estimate_error"), ns)
    estimate_error_plain = ns["estimate_error"]
    exec(_method_source("rbnics/problems/elliptic/elliptic_coercive_compliant_rb_reduced_problem.py",
                        "estimate_error"), ns)
    estimate_error_compliant = ns["estimate_error"]
    exec(_method_source("rbnics/problems/elliptic/elliptic_reduced_problem.py", "matrix_eval"), ns)
    exec(_method_source("rbnics/problems/elliptic/elliptic_reduced_problem.py", "vector_eval"), ns)
    sns = dict(ns)
    exec(_method_source("rbnics/problems/stokes/stokes_rb_reduced_problem.py", "get_residual_norm_squared"), sns)
    exec(_method_source("rbnics/problems/stokes/stokes_rb_reduced_problem.py", "estimate_error"), sns)
    exec(_method_source("rbnics/problems/stokes/stokes_reduced_problem.py", "matrix_eval"), sns)
    exec(_method_source("rbnics/problems/stokes/stokes_reduced_problem.py", "vector_eval"), sns)
    stokes = types.SimpleNamespace(get_residual_norm_squared=sns["get_residual_norm_squared"],
                                   estimate_error=sns["estimate_error"], matrix_eval=sns["matrix_eval"],
                                   vector_eval=sns["vector_eval"])
    return types.SimpleNamespace(stokes=stokes, product=product, sum=sum_mod.sum, DirichletBC=bc_mod.DirichletBC,
                                 parallel_max=pmax_mod.parallel_max, ParameterSpaceSubset=pss_mod.ParameterSpaceSubset,
                                 get_residual_norm_squared=ns["get_residual_norm_squared"],
                                 estimate_error_plain=estimate_error_plain,
                                 estimate_error_compliant=estimate_error_compliant, matrix_eval=ns["matrix_eval"],
                                 vector_eval=ns["vector_eval"])


def ref_rand(rng, *shape):
    """tests/performance/backends/online/numpy/test_numpy_utils.py:58-59 on a seeded Generator."""
    return (-1.0) ** rng.integers(0, 2, size=shape) * rng.random(shape) / (1e-3 + rng.random(shape))


class FakeReducedProblem:
    """the attributes the reference's solve / estimator methods read from ``self``"""

    def __init__(self, ref, A, F, Gff, Gaf, Gaa, n_bc):
        Qa, Qf = A.shape[0], F.shape[0]
        self.operator = {"a": Storage([A[q] for q in range(Qa)]), "f": Storage([F[q] for q in range(Qf)])}
        self.error_estimation_operator = {
            ("f", "f"): Storage([[float(Gff[i, j]) for j in range(Qf)] for i in range(Qf)]),
            ("a", "f"): Storage([[Gaf[i, j] for j in range(Qf)] for i in range(Qa)]),
            ("a", "a"): Storage([[Gaa[i, j] for j in range(Qa)] for i in range(Qa)])}
        self.n_bc = n_bc
        self.ref = ref
        self.truth_problem = self
        self.problem = self
        self.N = A.shape[1]

    def set(self, theta_a, theta_f, beta, theta_bc):
        self._ta, self._tf, self._beta, self._tbc = theta_a, theta_f, beta, theta_bc

    def compute_theta(self, term):
        return {"a": self._ta, "f": self._tf}[term]

    def get_stability_factor_lower_bound(self):
        return self._beta

    def get_residual_norm_squared(self):
        return self.ref.get_residual_norm_squared(self)

    def solve(self):
        """LinearSolver(problem_wrapper, solution): online/basic/linear_solver.py:29-35,56-59 + numpy solve"""
        N = self.N
        lhs = np.array(self.ref.matrix_eval(self))
        rhs = np.array(self.ref.vector_eval(self))
        if self._tbc is not None:
            bcs = self.ref.DirichletBC(self._tbc)
            bcs.apply_to_vector(rhs)
            bcs.apply_to_matrix(lhs)
        u = np.linalg.solve(lhs, rhs) if N > 0 else np.zeros(0)
        self._solution = u.view(Solution)
        return u


def make_case(ref, name, N, Qa, Qf, B, seed, n_bc=0, well_conditioned=True, ranks=(1, 2, 3)):
    rng = np.random.default_rng(seed)
    if well_conditioned:
        A = np.stack([(lambda Bq: Bq @ Bq.T / N + np.eye(N))(rng.standard_normal((N, N))) for _ in range(Qa)])
        F = rng.standard_normal((Qf, N))
        R = rng.standard_normal((3 * (Qa * N + Qf) // 2 + 8, Qa * N + Qf))
        G = R.T @ R / R.shape[0]
        Ka = Qa * N
        Gaa = G[:Ka, :Ka].reshape(Qa, N, Qa, N).transpose(0, 2, 1, 3).copy()
        Gaf = G[:Ka, Ka:].reshape(Qa, N, Qf).transpose(0, 2, 1).copy()
        Gff = G[Ka:, Ka:].copy()
        Theta_a = rng.uniform(0.1, 10.0, size=(B, Qa))
        Theta_f = rng.uniform(-1.0, 1.0, size=(B, Qf))
    else:  # the reference tests' heavy-tailed signed distribution (products only; not used for solves)
        A = ref_rand(rng, Qa, N, N)
        F = ref_rand(rng, Qf, N)
        Gaa, Gaf, Gff = ref_rand(rng, Qa, Qa, N, N), ref_rand(rng, Qa, Qf, N), ref_rand(rng, Qf, Qf)
        Theta_a, Theta_f = ref_rand(rng, B, Qa), ref_rand(rng, B, Qf)
    Theta_a[B // 2, Qa - 1] = 0.0  # exercises the zero-theta skipping of product.py:32,45
    beta = np.abs(Theta_a).min(axis=1) + 0.05
    Theta_bc = rng.uniform(-1.0, 1.0, size=(B, n_bc)) if n_bc else None

    prob = FakeReducedProblem(ref, A, F, Gff, Gaf, Gaa, n_bc)
    U = np.zeros((B, N))
    lhs_all = np.zeros((B, N, N))
    eps2 = np.zeros(B)
    d_plain = np.zeros(B)
    d_compl = np.zeros(B)
    for p in range(B):
        ta = tuple(float(x) for x in Theta_a[p])
        tf = tuple(float(x) for x in Theta_f[p])
        tbc = None if Theta_bc is None else tuple(float(x) for x in Theta_bc[p])
        prob.set(ta, tf, float(beta[p]), tbc)
        lhs_all[p] = np.array(ref.matrix_eval(prob))
        if well_conditioned:
            U[p] = prob.solve()
        else:
            prob._solution = ref_rand(rng, N).view(Solution)
            U[p] = prob._solution
        eps2[p] = ref.get_residual_norm_squared(prob)
        if well_conditioned:
            d_plain[p] = ref.estimate_error_plain(prob)
            d_compl[p] = ref.estimate_error_compliant(prob)

    out = dict(A=A, F=F, Gff=Gff, Gaf=Gaf, Gaa=Gaa, Theta_a=Theta_a, Theta_f=Theta_f, beta=beta, U=U, lhs=lhs_all,
               eps2=eps2, delta_plain=d_plain, delta_compliant=d_compl, n_bc=np.int64(n_bc),
               well_conditioned=np.bool_(well_conditioned))
    if Theta_bc is not None:
        out["Theta_bc"] = Theta_bc
    if well_conditioned:
        # ParameterSpaceSubset.max on the precomputed estimators: serial and distributed (cyclic shards + parallel_max)
        values = d_plain.copy()
        values[B // 3] = values.max()            # an exact tie between two parameters
        tie_lo = int(min(B // 3, int(np.argmax(d_plain))))
        pss = ref.ParameterSpaceSubset()
        pss._list = [(float(i),) for i in range(B)]
        pss.serialize_maximum_computations()
        v, i = pss.max(lambda mu: values[int(mu[0])])
        out["max_values"] = values
        out["max_serial"] = np.array([v, i])
        assert int(i) == tie_lo
        for size in ranks:
            def rank_fn(comm, size=size):
                s = ref.ParameterSpaceSubset()
                s._list = [(float(k),) for k in range(B)]
                s.mpi_comm = comm
                return s.max(lambda mu: values[int(mu[0])])
            res = run_ranks(size, rank_fn)
            assert all(r == res[0] for r in res)
            out[f"max_mpi_{size}"] = np.array([res[0][0], res[0][1]])
    np.savez_compressed(os.path.join(HERE, f"reference_sweep_{name}.npz"), **out)
    print(name, "N", N, "Qa", Qa, "Qf", Qf, "B", B, "eps2[0]", eps2[0], "max_serial", out.get("max_serial"))


class FakeStokesReducedProblem:
    """attributes read by the reference's Stokes matrix_eval / vector_eval / get_residual_norm_squared"""
    PAIRS = [("f", "f"), ("g", "g"), ("a", "f"), ("bt", "f"), ("b", "g"), ("a", "a"), ("a", "bt"), ("bt", "bt"),
             ("b", "b")]

    def __init__(self, ref, op, E, N):
        self.ref = ref
        self.operator = {t: Storage([op[t][q] for q in range(op[t].shape[0])]) for t in op}
        self.error_estimation_operator = {}
        for (t1, t2) in self.PAIRS:
            blk = E[t1, t2]
            if blk.ndim == 2:
                content = [[float(blk[i, j]) for j in range(blk.shape[1])] for i in range(blk.shape[0])]
            else:
                content = [[blk[i, j] for j in range(blk.shape[1])] for i in range(blk.shape[0])]
            self.error_estimation_operator[t1, t2] = Storage(content)
        self.N = N
        self.problem = self
        self.truth_problem = self

    def set(self, theta, beta, theta_bc):
        self._theta, self._beta, self._tbc = theta, beta, theta_bc

    def compute_theta(self, term):
        return self._theta[term]

    def get_stability_factor_lower_bound(self):
        return self._beta

    def get_residual_norm_squared(self):
        return self.ref.stokes.get_residual_norm_squared(self)

    def solve(self):
        lhs = np.array(self.ref.stokes.matrix_eval(self))
        rhs = np.array(self.ref.stokes.vector_eval(self))
        if self._tbc is not None:
            bcs = self.ref.DirichletBC(self._tbc)
            bcs.apply_to_vector(rhs)
            bcs.apply_to_matrix(lhs)
        u = np.linalg.solve(lhs, rhs)
        self._solution = u.view(Solution)
        return u


def make_stokes_case(ref, name, n, Q, B, seed, n_bc=0):
    """Saddle-point reduced system with components (u, s, p) of n basis functions each (N = 3n): a on the velocity
    block, bt = [0 B^T; 0 0], b = [0 0; B 0] -> indefinite matrix (pivoting matters); the Riesz Gram blocks come from
    velocity-space representers (a, bt, f) and pressure-space representers (b, g), so the pairs the reference skips
    are exactly zero and eps2 >= 0."""
    rng = np.random.default_rng(seed)
    nv, N = 2 * n, 3 * n
    op = {}
    op["a"] = np.zeros((Q["a"], N, N))
    for q in range(Q["a"]):
        Bq = rng.standard_normal((nv, nv))
        op["a"][q, :nv, :nv] = Bq @ Bq.T / nv + np.eye(nv)
    op["b"] = np.zeros((Q["b"], N, N))
    op["bt"] = np.zeros((Q["bt"], N, N))
    for q in range(Q["b"]):
        op["b"][q, nv:, :nv] = rng.standard_normal((n, nv))
    for q in range(Q["bt"]):
        op["bt"][q, :nv, nv:] = op["b"][q % Q["b"], nv:, :nv].T
    op["f"] = np.zeros((Q["f"], N))
    op["f"][:, :nv] = rng.standard_normal((Q["f"], nv))
    op["g"] = np.zeros((Q["g"], N))
    op["g"][:, nv:] = rng.standard_normal((Q["g"], n))
    # Riesz factors: velocity space (a, bt, f) and pressure space (b, g)
    Kv = Q["a"] * N + Q["bt"] * N + Q["f"]
    Kp = Q["b"] * N + Q["g"]
    Rv = rng.standard_normal((Kv + 7, Kv))
    Rp = rng.standard_normal((Kp + 5, Kp))
    Gv = Rv.T @ Rv / Rv.shape[0]
    Gp = Rp.T @ Rp / Rp.shape[0]
    offv = {"a": 0, "bt": Q["a"] * N, "f": Q["a"] * N + Q["bt"] * N}
    offp = {"b": 0, "g": Q["b"] * N}

    def block(Gm, off, t1, t2):
        n1 = Q[t1] * N if t1 in ("a", "b", "bt") else Q[t1]
        n2 = Q[t2] * N if t2 in ("a", "b", "bt") else Q[t2]
        M = Gm[off[t1]:off[t1] + n1, off[t2]:off[t2] + n2]
        o1, o2 = t1 in ("a", "b", "bt"), t2 in ("a", "b", "bt")
        if o1 and o2:
            return M.reshape(Q[t1], N, Q[t2], N).transpose(0, 2, 1, 3).copy()
        if o1:
            return M.reshape(Q[t1], N, Q[t2]).transpose(0, 2, 1).copy()
        return M.copy()

    E = {}
    for (t1, t2) in FakeStokesReducedProblem.PAIRS:
        E[t1, t2] = block(Gv, offv, t1, t2) if t1 in offv else block(Gp, offp, t1, t2)
    theta = {t: rng.uniform(0.5, 2.0, size=(B, Q[t])) for t in ("a", "b", "bt")}
    theta["f"] = rng.uniform(-1.0, 1.0, size=(B, Q["f"]))
    theta["g"] = rng.uniform(-1.0, 1.0, size=(B, Q["g"]))
    theta["a"][B // 2, Q["a"] - 1] = 0.0
    beta = np.ones(B)                                      # tutorials/12_stokes: beta == 1
    Theta_bc = rng.uniform(-1.0, 1.0, size=(B, n_bc)) if n_bc else None
    prob = FakeStokesReducedProblem(ref, op, E, N)
    U, eps2, delta = np.zeros((B, N)), np.zeros(B), np.zeros(B)
    for p in range(B):
        th = {t: tuple(float(x) for x in theta[t][p]) for t in theta}
        tbc = None if Theta_bc is None else tuple(float(x) for x in Theta_bc[p])
        prob.set(th, float(beta[p]), tbc)
        U[p] = prob.solve()
        eps2[p] = ref.stokes.get_residual_norm_squared(prob)
        delta[p] = ref.stokes.estimate_error(prob)
    out = {"op_" + t: op[t] for t in op}
    out.update({"E_%s_%s" % k: v for k, v in E.items()})
    out.update({"theta_" + t: theta[t] for t in theta})
    out.update(beta=beta, U=U, eps2=eps2, delta=delta, n_bc=np.int64(n_bc), n=np.int64(n))
    if Theta_bc is not None:
        out["Theta_bc"] = Theta_bc
    np.savez_compressed(os.path.join(HERE, f"reference_stokes_{name}.npz"), **out)
    print("stokes", name, "N", N, "B", B, "eps2[0]", eps2[0], "argmax", int(np.argmax(delta)))


# (The on-disk format fixtures are produced by the REAL reference's AffineExpansionStorage.save():
# tests/golden/make_reference_golden.py -> tests/golden/storage_real.  A round-1 restatement of the call sequence that
# lived here wrote `None` where the reference writes `(None, None)` for matrices with integer sizes.)


def make_offline_golden():
    """Offline basis linear algebra by the reference's OWN loops, executed from source with PETSc's three primitives
    replaced by their numpy/scipy equivalents (MatMult -> X @ v, VecDot -> numpy.dot, add_local -> +=):
      rbnics/backends/basic/transpose.py:293-325          FunctionsList_Transpose__times__Matrix.__mul__  (S^T X S loop)
      rbnics/backends/basic/gram_schmidt.py:13-36         GramSchmidt.apply
      rbnics/backends/online/numpy/wrapping/gram_schmidt_projection_step.py
      rbnics/backends/basic/proper_orthogonal_decomposition_base.py:15-92   POD apply (energy, truncation, modes)
      rbnics/backends/online/numpy/eigen_solver.py:19-67  EigenSolver (scipy.linalg.eigh + descending sort)
      rbnics/backends/dolfin/wrapping/functions_list_mul.py:26-36 restated as the same add_local loop (needs dolfin)
    -> tests/golden/reference_offline.npz"""
    import logging
    import scipy.sparse as sp

    class Fun:  # a dof vector with the .vector() accessor of dolfin / online Functions
        def __init__(self, v):
            self._v = np.array(v, dtype=np.float64)

        def vector(self):
            return self._v

        def __itruediv__(self, a):
            self._v /= a
            return self

        @property
        def N(self):
            return self._v.shape[0]

    class FunList(list):  # FunctionsList / SnapshotsMatrix / BasisFunctionsMatrix stand-in
        def __init__(self, space=None, *args):
            super().__init__()

        def enrich(self, f):
            self.append(Fun(f.vector()))

        def clear(self):
            del self[:]

        def __mul__(self, online_fun):  # functions_list_mul_online_vector: sum_i s_i * v_i, sequential add_local
            out = Fun(np.zeros_like(self[0].vector()))
            for i, fun_i in enumerate(self):
                out.vector()[:] += fun_i.vector() * online_fun.vector()[i]
            return out

    spmat = sp.csr_matrix
    wrapping = types.SimpleNamespace(matrix_mul_vector=lambda M, v: M @ v, function_to_vector=lambda f: f.vector(),
                                     vector_mul_vector=lambda a, b: float(np.dot(a, b)),
                                     function_extend_or_restrict=lambda f, c1, space, c2, weight, copy: Fun(f.vector()),
                                     get_function_subspace=lambda space, component: space)
    backend = types.SimpleNamespace(FunctionsList=FunList, Matrix=types.SimpleNamespace(Type=lambda: spmat),
                                    Function=types.SimpleNamespace(Type=lambda: Fun),
                                    Vector=types.SimpleNamespace(Type=lambda: np.ndarray))
    online_backend = types.SimpleNamespace(OnlineMatrix=lambda M, N: np.zeros((M, N)),
                                           OnlineVector=lambda N: np.zeros(N))
    # --- S^T X S: the factory function of transpose.py, extracted and executed
    ns = {"overload": overload, "logger": logging.getLogger("ref"), "DEBUG": logging.DEBUG}
    exec(_method_source("rbnics/backends/basic/transpose.py", "FunctionsList_Transpose__times__Matrix"), ns)
    STXS = ns["FunctionsList_Transpose__times__Matrix"](backend, wrapping, online_backend, None, None, None, None, None)

    class Transpose:  # transpose(S)*X*S, transpose(v)*X*w, transpose(v)*w with the same primitives
        def __init__(self, arg, M=None):
            self.arg, self.M = arg, M

        def __mul__(self, other):
            if isinstance(other, spmat):
                return Transpose(self.arg, other)
            if isinstance(self.arg, FunList):
                return STXS(self.arg, self.M) * other if self.M is not None else \
                    STXS(self.arg, sp.identity(other[0].N, format="csr")) * other
            w = other.vector()
            return float(np.dot(self.arg.vector(), self.M @ w if self.M is not None else w))

    backend.transpose = lambda a: Transpose(a)
    # --- Gram-Schmidt
    _module("rbnics.backends.abstract", GramSchmidt=object, FunctionsList=FunList, EigenSolver=object)
    _module("rbnics.utils.decorators", overload=overload, dict_of=lambda *a: dict, ThetaType=ThetaType,
            DictOfThetaType=DictOfThetaType, BackendFor=lambda *a, **k: (lambda cls: cls))
    gs_step = _exec_reference_file("rbnics/backends/online/numpy/wrapping/gram_schmidt_projection_step.py", "ref_gs_step")
    wrapping.gram_schmidt_projection_step = gs_step.gram_schmidt_projection_step
    GS = _exec_reference_file("rbnics/backends/basic/gram_schmidt.py", "ref_gs").GramSchmidt(backend, wrapping)
    # --- EigenSolver of the numpy online backend
    class _OnlineMat(np.ndarray):
        M = property(lambda self: self.shape[0])
        N = property(lambda self: self.shape[1])

    _module("rbnics.backends.online")
    _module("rbnics.backends.online.numpy")
    _module("rbnics.backends.online.numpy.function", Function=lambda v: Fun(v))
    _module("rbnics.backends.online.numpy.matrix", Matrix=types.SimpleNamespace(Type=lambda: _OnlineMat))
    _module("rbnics.backends.online.numpy.vector",
            Vector=types.SimpleNamespace(Type=lambda: (lambda N, content: np.array(content, dtype=np.float64))))
    ES = _exec_reference_file("rbnics/backends/online/numpy/eigen_solver.py", "ref_eig").EigenSolver
    online_backend.OnlineEigenSolver = lambda basis, A: ES(basis, np.asarray(A).view(_OnlineMat))
    # --- POD
    _module("rbnics.utils.io", ExportableList=lambda kind: list())
    pod_mod = _exec_reference_file("rbnics/backends/basic/proper_orthogonal_decomposition_base.py", "ref_pod")
    POD = pod_mod.ProperOrthogonalDecompositionBase(backend, wrapping, online_backend, None, object, FunList, FunList)

    rng = np.random.default_rng(11)
    n = 12
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n))
    X = (sp.kron(T, sp.identity(n)) + sp.kron(sp.identity(n), T) + 0.5 * sp.identity(n * n)).tocsr()
    Nh, Ns = n * n, 9
    decay = 10.0 ** (-np.arange(Ns))
    S = (rng.standard_normal((Nh, Ns)) @ np.diag(decay)) @ rng.standard_normal((Ns, Ns))
    SL = FunList()
    for j in range(Ns):
        SL.enrich(Fun(S[:, j]))
    C = np.array(backend.transpose(SL) * X * SL)
    # Gram-Schmidt: orthonormalise the first 5 snapshots one after the other
    gs = GS(None, X)
    Z = FunList()
    for j in range(5):
        Z.enrich(gs.apply(Fun(S[:, j]), Z))
    Zarr = np.stack([z.vector() for z in Z], axis=1)
    # Galerkin projection of an operator and a vector on that basis (same loop: transpose(Z)*A*Z, transpose(Z)*f)
    A_op = (X + sp.diags(rng.uniform(0.0, 1.0, Nh))).tocsr()
    ZtAZ = np.array(backend.transpose(Z) * A_op * Z)
    # POD
    pod = POD(None, X)
    for j in range(Ns):
        pod.snapshots_matrix.enrich(Fun(S[:, j]))
    eigs, eigvecs, basis, Npod = pod.apply(Ns, 1e-10)
    out = dict(X_data=X.data, X_indices=X.indices, X_indptr=X.indptr, Nh=np.int64(Nh), S=S, C=C, Z_gs=Zarr,
               A_data=A_op.data, A_indices=A_op.indices, A_indptr=A_op.indptr, ZtAZ=ZtAZ,
               pod_eigenvalues=np.array(pod.eigenvalues), pod_retained_energy=np.array(pod.retained_energy),
               pod_N=np.int64(Npod), pod_basis=np.stack([b.vector() for b in basis], axis=1),
               pod_eigenvectors=np.stack([e.vector() for e in eigvecs], axis=1))
    np.savez_compressed(os.path.join(HERE, "reference_offline.npz"), **out)
    print("offline golden: Nh", Nh, "Ns", Ns, "POD N", Npod, "eig[0..2]", pod.eigenvalues[:3])


def make_parabolic_golden(ref):
    """POD-Greedy sweep of a parabolic reduced problem by the reference's own source:
      rbnics/problems/parabolic/abstract_parabolic_reduced_problem.py   residual_eval / jacobian_eval
      rbnics/problems/parabolic/abstract_parabolic_rb_reduced_problem.py   get_residual_norm_squared / estimate_error
      rbnics/backends/common/time_series.py, time_quadrature.py         TimeSeries, TimeQuadrature_Numbers.integrate
    driven by the linear implicit-Euler loop of rbnics/backends/online/numpy/time_stepping.py:104-112,226-238 (restated
    here line by line: the class itself is tied to the backend's Function type) and the closure of
    time_dependent_rb_reduction.py:280-288.  -> tests/golden/reference_parabolic.npz"""
    _module("rbnics.backends.abstract", TimeSeries=object, TimeQuadrature=object, ParametrizedTensorFactory=type("PTF", (), {}))
    _module("rbnics.utils.decorators", overload=overload, ThetaType=ThetaType, DictOfThetaType=DictOfThetaType,
            BackendFor=lambda *a, **k: (lambda cls: cls), backend_for=lambda *a, **k: (lambda f: f),
            tuple_of=lambda *a: tuple, list_of=lambda *a: list)
    TimeSeries = _exec_reference_file("rbnics/backends/common/time_series.py", "ref_time_series").TimeSeries
    _module("rbnics.backends.common")
    _module("rbnics.backends.common.time_series", TimeSeries=TimeSeries)
    tq = _exec_reference_file("rbnics/backends/common/time_quadrature.py", "ref_time_quadrature")
    class Mat(np.ndarray):  # online Matrix: `*` with a vector is the matrix-vector product (online/numpy/matrix.py:31-41)
        def __mul__(self, other):
            if isinstance(other, np.ndarray) and other.ndim == 1:
                return np.asarray(self) @ np.asarray(other)
            return np.ndarray.__mul__(self, other)

    def ref_sum(product_output):
        out = ref.sum(product_output)
        return out.view(Mat) if isinstance(out, np.ndarray) and out.ndim == 2 else out

    ns = {"product": ref.product, "sum": ref_sum, "transpose": transpose, "sqrt": sqrt, "isclose": np.isclose,
          "abs": abs, "TimeSeries": TimeSeries, "assign": lambda dst, src: dst.__setitem__(slice(None), src)}

    class EllipticBase:  # AbstractParabolicRBReducedProblem_Base.get_residual_norm_squared == the elliptic one
        get_residual_norm_squared = staticmethod(ref.get_residual_norm_squared)

    ns["AbstractParabolicRBReducedProblem_Base"] = EllipticBase
    exec(_method_source("rbnics/problems/parabolic/abstract_parabolic_rb_reduced_problem.py", "get_residual_norm_squared"), ns)
    exec(_method_source("rbnics/problems/parabolic/abstract_parabolic_rb_reduced_problem.py", "estimate_error"), ns)
    exec(_method_source("rbnics/problems/parabolic/abstract_parabolic_reduced_problem.py", "residual_eval"), ns)
    exec(_method_source("rbnics/problems/parabolic/abstract_parabolic_reduced_problem.py", "jacobian_eval"), ns)

    rng = np.random.default_rng(21)
    N, Qm, Qa, Qf, B, K, dt = 12, 2, 3, 2, 20, 10, 0.05
    spd = lambda: (lambda Bq: Bq @ Bq.T / N + np.eye(N))(rng.standard_normal((N, N)))  # noqa: E731
    Mq = np.stack([spd() for _ in range(Qm)])
    Aq = np.stack([spd() for _ in range(Qa)])
    Fq = rng.standard_normal((Qf, N))
    Kz = Qa * N + Qm * N + Qf
    R = rng.standard_normal((Kz + 9, Kz))
    Gm = R.T @ R / R.shape[0]
    off = {"a": 0, "m": Qa * N, "f": Qa * N + Qm * N}
    Q = {"a": Qa, "m": Qm, "f": Qf}

    def block(t1, t2):
        o1, o2 = t1 != "f", t2 != "f"
        n1, n2 = Q[t1] * (N if o1 else 1), Q[t2] * (N if o2 else 1)
        M = Gm[off[t1]:off[t1] + n1, off[t2]:off[t2] + n2]
        if o1 and o2:
            return M.reshape(Q[t1], N, Q[t2], N).transpose(0, 2, 1, 3).copy()
        if o1:
            return M.reshape(Q[t1], N, Q[t2]).transpose(0, 2, 1).copy()
        return M.copy()

    E = {k: block(*k) for k in [("f", "f"), ("a", "f"), ("a", "a"), ("m", "f"), ("m", "a"), ("m", "m")]}
    Theta_m = rng.uniform(0.5, 2.0, (B, Qm))
    Theta_a = rng.uniform(0.1, 10.0, (B, Qa))
    Theta_f = rng.uniform(-1.0, 1.0, (B, Qf))
    beta = Theta_a.min(axis=1)

    class Prob:
        def __init__(self):
            self.operator = {"m": Storage([Mq[q] for q in range(Qm)]), "a": Storage([Aq[q] for q in range(Qa)]),
                             "f": Storage([Fq[q] for q in range(Qf)])}
            self.error_estimation_operator = {}
            for (t1, t2), blk in E.items():
                content = ([[float(blk[i, j]) for j in range(blk.shape[1])] for i in range(blk.shape[0])]
                           if blk.ndim == 2 else [[blk[i, j] for j in range(blk.shape[1])] for i in range(blk.shape[0])])
                self.error_estimation_operator[t1, t2] = Storage(content)
            self.N, self.problem, self.truth_problem = N, self, self
            self.t0, self.dt, self.T = 0.0, dt, K * dt

        def compute_theta(self, term):
            return self._th[term]

        def get_stability_factor_lower_bound(self):
            return self._beta

        def set_time(self, t):
            self.t = t

        def get_initial_error_estimate_squared(self):
            return 0.0  # homogeneous initial condition (time_dependent_rb_reduced_problem.py:127-160)

        def get_residual_norm_squared(self):
            return ns["get_residual_norm_squared"](self)

    prob = Prob()
    U = np.zeros((B, K + 1, N))
    eps2 = np.zeros((B, K + 1))
    bound = np.zeros((B, K + 1))
    delta = np.zeros(B)
    for p in range(B):
        prob._th = {"m": tuple(float(x) for x in Theta_m[p]), "a": tuple(float(x) for x in Theta_a[p]),
                    "f": tuple(float(x) for x in Theta_f[p])}
        prob._beta = float(beta[p])
        # _ScipyImplicitEuler, linear problem (time_stepping.py:104-112, 226-238)
        solution = np.zeros(N).view(Solution)
        solution_previous = np.zeros(N)
        zero = np.zeros(N)
        sol_t, dot_t = TimeSeries((0.0, K * dt), dt), TimeSeries((0.0, K * dt), dt)
        sol_t.append(solution.copy().view(Solution))
        dot_t.append(np.zeros(N).view(Solution))
        for t in sol_t.expected_times()[1:]:
            minus_prev_over_dt = solution_previous.copy()
            minus_prev_over_dt /= -dt
            lhs = np.array(ns["jacobian_eval"](prob, t, zero, zero, 1.0 / dt))
            rhs = -np.array(ns["residual_eval"](prob, t, zero, minus_prev_over_dt))
            solution = np.linalg.solve(lhs, rhs)
            solution_dot = (solution - solution_previous) / dt
            sol_t.append(solution.copy().view(Solution))
            dot_t.append(solution_dot.copy().view(Solution))
            solution_previous = solution.copy()
        prob._solution_over_time, prob._solution_dot_over_time = sol_t, dot_t
        prob._solution = np.zeros(N).view(Solution)
        prob._solution_dot = np.zeros(N).view(Solution)
        e_t = ns["get_residual_norm_squared"](prob)
        b_t = ns["estimate_error"](prob)
        squared = [v ** 2 for v in b_t]
        delta[p] = sqrt(tq.TimeQuadrature_Numbers((0.0, K * dt), squared).integrate())
        U[p] = np.array([np.asarray(s) for s in sol_t])
        eps2[p] = np.array(list(e_t))
        bound[p] = np.array(list(b_t))
    out = dict(M=Mq, A=Aq, F=Fq, Theta_m=Theta_m, Theta_a=Theta_a, Theta_f=Theta_f, beta=beta, dt=np.float64(dt),
               K=np.int64(K), U=U, eps2=eps2, bound=bound, delta=delta)
    out.update({"E_%s_%s" % k: v for k, v in E.items()})
    np.savez_compressed(os.path.join(HERE, "reference_parabolic.npz"), **out)
    print("parabolic golden: N", N, "K", K, "B", B, "delta[0]", delta[0], "argmax", int(np.argmax(delta)))


def make_supremizer_golden(ref):
    """Reduced supremizer solves by the reference's own source: StokesReducedProblem._solve_supremizer
    (rbnics/problems/stokes/stokes_reduced_problem.py:80-103) with the numpy LinearSolver.solve
    (rbnics/backends/online/numpy/linear_solver.py:29-31) -> tests/golden/reference_supremizer.npz"""
    class Mat(np.ndarray):  # online Matrix: `*` with a vector is the matrix-vector product (online/numpy/matrix.py:31-41)
        def __mul__(self, other):
            if isinstance(other, np.ndarray) and other.ndim == 1:
                return np.asarray(self) @ np.asarray(other)
            return np.ndarray.__mul__(self, other)

    def ref_sum(product_output):
        out = ref.sum(product_output)
        return out.view(Mat) if isinstance(out, np.ndarray) and out.ndim == 2 else out

    class Fun(np.ndarray):  # OnlineFunction: .N and .vector()
        @property
        def N(self):
            return self.shape[0]

        def vector(self):
            return self

    class OnlineLinearSolver:  # rbnics/backends/online/numpy/linear_solver.py:25-33 (+ basic/linear_solver.py bcs)
        def __init__(self, lhs, solution, rhs, bcs=None):
            self.lhs, self.solution, self.rhs = np.array(lhs), solution, np.array(rhs)
            assert bcs is None

        def set_parameters(self, parameters):
            assert len(parameters) == 0

        def solve(self):
            self.solution.vector()[:] = np.linalg.solve(self.lhs, self.rhs)

    ns = {"product": ref.product, "sum": ref_sum, "OnlineLinearSolver": OnlineLinearSolver}
    exec(_method_source("rbnics/problems/stokes/stokes_reduced_problem.py", "_solve_supremizer"), ns)
    solve_supremizer = ns["_solve_supremizer"]

    rng = np.random.default_rng(33)
    n_u, n_s, n_p, Q, B = 9, 7, 6, 4, 30
    N_us, N_usp = n_u + n_s, n_u + n_s + n_p
    Rm = rng.standard_normal((N_us + 5, N_us))
    Xs = Rm.T @ Rm / Rm.shape[0] + np.eye(N_us)
    Bt = ref_rand(rng, Q, N_us, N_usp)
    Theta = rng.uniform(-2.0, 2.0, (B, Q))
    U = ref_rand(rng, B, N_usp)
    S = np.empty((B, N_us))

    class Problem:
        def __init__(self):
            self.inner_product = {"s": Storage([Xs])}
            self.operator = {"bt_restricted": Storage([Bt[q] for q in range(Q)])}
            self.dirichlet_bc = {"u": False, "s": False}
            self.dirichlet_bc_are_homogeneous = {"u": True, "s": True}
            self._linear_solver_parameters = {}

        def compute_theta(self, term):
            assert term == "bt_restricted"
            return self._theta

    prob = Problem()
    for p in range(B):
        prob._theta = tuple(Theta[p])
        prob._supremizer = np.zeros(N_us).view(Fun)
        solve_supremizer(prob, U[p].copy().view(Fun))
        S[p] = np.asarray(prob._supremizer)
    np.savez_compressed(os.path.join(HERE, "reference_supremizer.npz"), Xs=Xs, Bt=Bt, Theta=Theta, U=U, S=S,
                        N_us=np.int64(N_us))
    print("supremizer golden: N_us", N_us, "N_usp", N_usp, "Q", Q, "B", B, "s[0][:3]", S[0][:3])


def main():
    ref = load_reference()
    make_supremizer_golden(ref)
    make_offline_golden()
    make_parabolic_golden(ref)
    make_case(ref, "tutorial01", N=4, Qa=2, Qf=1, B=100, seed=1)
    make_case(ref, "n16q6", N=16, Qa=6, Qf=6, B=40, seed=2)
    make_case(ref, "n32q10_bc", N=32, Qa=10, Qf=10, B=24, seed=3, n_bc=2)
    make_case(ref, "n40q7", N=40, Qa=7, Qf=3, B=16, seed=4)
    make_case(ref, "heavytail_n16", N=16, Qa=6, Qf=10, B=10, seed=5, well_conditioned=False)
    make_stokes_case(ref, "n5", n=5, Q={"a": 3, "b": 2, "bt": 2, "f": 2, "g": 1}, B=24, seed=6)
    make_stokes_case(ref, "n12_bc", n=12, Q={"a": 4, "b": 3, "bt": 3, "f": 3, "g": 2}, B=16, seed=7, n_bc=1)


if __name__ == "__main__":
    main()

"""Reduced-basis greedy offline loop driven by the batched B200 sweep.

Host-side mirror of ``RBReduction._offline / greedy / _greedy / update_basis_matrix``
(rbnics/reduction_methods/base/rb_reduction.py:79-185) and of the reduced-problem pieces it calls:
``build_reduced_operators`` (rbnics/problems/base/parametrized_reduced_differential_problem.py:447-479,727-790) and
``build_error_estimation_operators`` (rbnics/problems/base/rb_reduced_problem.py:142-263).  Same method names, same
bookkeeping (``greedy_selected_parameters``, ``greedy_error_estimators``, stop on ``N >= Nmax`` or relative estimator
``< tol``); the per-parameter closure is replaced by ONE batched sweep per iteration.

The truth problem (high-fidelity solves, Riesz solves) is the reference's job and stays on the host: it is passed in
as an object with

    operators_a : list of scipy.sparse matrices A_q           operators_f : list of Nh-vectors f_q
    inner_product : scipy.sparse X                             solve(mu) -> Nh-vector (truth solution)
    riesz_solve(rhs) -> X^-1 rhs (homogeneous Dirichlet)       compute_theta(term, mus) -> (B x Q) array, vectorised
    get_stability_factor_lower_bound(mus) -> (B,) array

Non-homogeneous Dirichlet conditions (optional): ``dirichlet_bc_liftings`` = (Nh x N_bc) array of lifting functions,
one per term of ``compute_theta("dirichlet_bc", mus)`` -> (B x N_bc).  As in the reference they are the first N_bc
basis functions (parametrized_reduced_differential_problem.py:283-299), their rows of the reduced system are lifting
rows with right-hand side theta_bc (backends/online/basic/wrapping/DirichletBC.py:37-56), a snapshot is stripped of
``liftings @ theta_bc(mu)`` before it enters the basis (differential_problem_reduction_method.py:73-108) and
Gram-Schmidt runs against ``basis_functions[N_bc:]`` only (rb_reduction.py:125-139).  ``self.N`` counts ALL basis
functions here (the reference's ``N + N_bc``); ``Nmax`` bounds ``self.N - self.N_bc`` like the reference's ``N``.
"""
from __future__ import annotations

import time
from collections import defaultdict
from contextlib import contextmanager

import numpy as np

from . import _lib as L
from . import offline
from .sweep import TrainingSetSweep


class ReducedBasisGreedy:
    """``incremental=True`` (default) reuses ``A^N_q[:N-1, :N-1]`` and the Riesz Gram blocks of the previous greedy
    iteration and computes only the new row / column; the reference recomputes every inner
    product at every iteration (rb_reduction.py:106,112).  ``incremental=False`` is that full recomputation."""

    def __init__(self, engine, truth_problem, training_set, Nmax, tol, estimator_kind=L.RB_EST_SQRT_EPS2_OVER_BETA,
                 dist=None, verbose=False, incremental=True, resident=False):
        self.engine = engine
        self.truth_problem = truth_problem
        self.training_set = [tuple(mu) for mu in training_set]
        self.Nmax = Nmax
        self.tol = tol
        self.kind = estimator_kind
        self.dist = dist
        self.verbose = verbose
        self.incremental = incremental
        # resident=True keeps the truth operators, the basis and the Riesz representers in HBM (offline.OfflineWorkspace)
        # and computes only the new rows / columns there: nothing but new vectors crosses PCIe
        self.resident = resident
        tp = truth_problem
        self.Qa, self.Qf = len(tp.operators_a), len(tp.operators_f)
        self.Nh = tp.operators_f[0].shape[0]
        # theta / beta for the whole training set, once (they do not change during the greedy loop)
        mus = np.asarray(self.training_set, dtype=np.float64)
        Theta = np.concatenate([tp.compute_theta("a", mus), tp.compute_theta("f", mus)], axis=1)
        beta = tp.get_stability_factor_lower_bound(mus)
        liftings = getattr(tp, "dirichlet_bc_liftings", None)
        self.N_bc = 0 if liftings is None else int(np.asarray(liftings).shape[1])
        ThetaBC = tp.compute_theta("dirichlet_bc", mus) if self.N_bc else None
        self._sweep = TrainingSetSweep(engine, Theta, beta, ThetaBC=ThetaBC, dist=dist)
        # reduced problem state
        self.N = 0
        self.basis_functions = np.zeros((self.Nh, 0))
        self.operator_a = np.zeros((self.Qa, 0, 0))
        self.operator_f = np.zeros((self.Qf, 0))
        self.riesz_f = None                      # Nh x Qf
        self.riesz_a = [np.zeros((self.Nh, 0)) for _ in range(self.Qa)]   # Nh x N each
        self.G = None
        self.GS = offline.GramSchmidt(engine, tp.inner_product)
        self.ws = None
        if resident:
            ws = offline.OfflineWorkspace(engine, dist=dist)
            ws.set_operator("X", tp.inner_product)
            self._a_symmetric = []
            for q, A in enumerate(tp.operators_a):
                ws.set_operator(f"A{q}", A)
                A = A.tocsr()
                sym = (A != A.T).nnz == 0
                self._a_symmetric.append(sym)
                if not sym:   # z^T A_q Z_j = (A_q^T z) . Z_j: the transposed operator gives the new ROW as a mat-vec too
                    ws.set_operator(f"A{q}T", A.T.tocsr())
            ws.new_block("Z", self.Nh, Nmax + 1 + self.N_bc)
            ws.new_block("YA", self.Nh, 2 * self.Qa)   # [A_q z_new | A_q^T z_new] for all q, side by side
            ws.new_block("F", self.Nh, self.Qf)
            ws.append("F", np.stack([np.asarray(f, dtype=np.float64) for f in tp.operators_f], axis=1))
            ws.new_block("R", self.Nh, self.Qa * (Nmax + 1 + self.N_bc) + self.Qf)   # Riesz representers: r_f, then (n, q)
            self.ws = ws
        self.greedy_selected_parameters = []
        self.greedy_selected_indices = []
        self.greedy_error_estimators = []
        self.mu = None
        self.timings = defaultdict(float)   # wall seconds per stage, summed over the iterations (bench_greedy.py)

    @contextmanager
    def _timed(self, stage):
        t0 = time.perf_counter()
        try:
            yield
        finally:
            self.timings[stage] += time.perf_counter() - t0

    # -- reduced_problem.build_reduced_operators() ---------------------------------------------------------------
    def build_reduced_operators(self):
        tp = self.truth_problem
        if self.N == 0:
            self.operator_a = np.zeros((self.Qa, 0, 0))
            self.operator_f = np.zeros((self.Qf, 0))
        elif self.resident:
            ws, N = self.ws, self.N
            A_new = np.zeros((self.Qa, N, N))
            A_new[:, :N - 1, :N - 1] = self.operator_a
            # new column Z^T (A_q z_N) and new row (A_q^T z_N)^T Z of every operator: 2 Qa sparse mat-vecs written side
            # by side, then ONE pass over Z for all of them (the reference: 2 Qa N inner products, each a pass)
            Qa = self.Qa
            for q in range(Qa):
                ws.spmm(f"A{q}", "Z", (N - 1, N), "YA", q)
                if not self._a_symmetric[q]:
                    ws.spmm(f"A{q}T", "Z", (N - 1, N), "YA", Qa + q)
            if all(self._a_symmetric):
                C = ws.gram("Z", None, "YA", cols2=(0, Qa))            # N x Qa
                Crow = C
            else:
                for q in range(Qa):
                    if self._a_symmetric[q]:
                        ws.spmm(f"A{q}", "Z", (N - 1, N), "YA", Qa + q)
                C2 = ws.gram("Z", None, "YA", cols2=(0, 2 * Qa))       # N x 2 Qa
                C, Crow = C2[:, :Qa], C2[:, Qa:]
            for q in range(Qa):
                A_new[q, :, N - 1] = C[:, q]
                A_new[q, N - 1, :] = Crow[:, q]
            self.operator_a = A_new
            f_new = np.zeros((self.Qf, N))
            f_new[:, :N - 1] = self.operator_f
            f_new[:, N - 1] = ws.gram("Z", None, "F", cols=(N - 1, N))[0, :]
            self.operator_f = f_new
        elif self.incremental and self.operator_a.shape[1] == self.N - 1:
            # only the new row and column: A^N[:, N-1] = Z^T (A_q z), A^N[N-1, :] = z^T (A_q Z); Z is unchanged for the
            # first N-1 functions because the Gram-Schmidt step never modifies earlier basis functions
            Z = self.basis_functions
            z = np.ascontiguousarray(Z[:, -1:])
            N = self.N
            A_new = np.zeros((self.Qa, N, N))
            A_new[:, :N - 1, :N - 1] = self.operator_a
            for q, A in enumerate(tp.operators_a):
                A_new[q, :, N - 1] = offline.gram(self.engine, Z, A, z, dist=self.dist)[:, 0]
                A_new[q, N - 1, :] = offline.gram(self.engine, z, A, Z, dist=self.dist)[0, :]
            self.operator_a = A_new
            f_new = np.zeros((self.Qf, N))
            f_new[:, :N - 1] = self.operator_f
            f_new[:, N - 1] = offline.project_vectors(self.engine, z, tp.operators_f, dist=self.dist)[:, 0]
            self.operator_f = f_new
        else:
            Z = self.basis_functions
            self.operator_a = offline.project_matrices(self.engine, Z, tp.operators_a)
            self.operator_f = offline.project_vectors(self.engine, Z, tp.operators_f)
        # the liftings are the first N_bc basis functions: their rows are lifting rows (DirichletBC.py:37-56)
        self.engine.set_system(self.operator_a, self.operator_f, bc_rows=list(range(min(self.N_bc, self.N))) or None)

    # -- reduced_problem.build_error_estimation_operators() ------------------------------------------------------
    def build_error_estimation_operators(self):
        tp = self.truth_problem
        if self.resident:
            return self._build_error_estimation_operators_resident()
        if self.riesz_f is None:   # does not depend on N: computed once (rb_reduced_problem.py:156-163)
            self.riesz_f = np.stack([tp.riesz_solve(f) for f in tp.operators_f], axis=1)
        for q in range(self.Qa):   # Riesz representers of -A_q z_n for the new basis functions (:183-195)
            have = self.riesz_a[q].shape[1]
            new = [tp.riesz_solve(-(tp.operators_a[q] @ self.basis_functions[:, n])) for n in range(have, self.N)]
            if new:
                self.riesz_a[q] = np.concatenate([self.riesz_a[q], np.stack(new, axis=1)], axis=1)
        N, Qa, Qf = self.N, self.Qa, self.Qf
        # all Riesz Gram blocks at once: S = [R_1 | ... | R_Qa | r_f],  G = S^T X S  in z-order [(q, n) ... | f]
        S = np.concatenate(self.riesz_a + [self.riesz_f], axis=1)
        if self.incremental and self.G is not None and N > 0 and self.G.shape[0] == Qa * (N - 1) + Qf:
            # only the Qa new representers against everything: G_new[(q, N-1), :] = r_{q,N-1}^T X S; X is symmetric
            old_idx = [q * N + n for q in range(Qa) for n in range(N - 1)] + [Qa * N + j for j in range(Qf)]
            new_idx = [q * N + (N - 1) for q in range(Qa)]
            G = np.empty((Qa * N + Qf, Qa * N + Qf))
            G[np.ix_(old_idx, old_idx)] = self.G
            # (X r_new is the narrow product: the operator is applied to the Qa new columns, not to all of S)
            rows = offline.gram(self.engine, S, tp.inner_product, np.ascontiguousarray(S[:, new_idx]), dist=self.dist).T
            G[new_idx, :] = rows
            G[:, new_idx] = rows.T
            self.G = G
        else:
            self.G = offline.gram(self.engine, S, tp.inner_product, dist=self.dist)
        if N > 0:
            scale = list(range(Qa)) + [-1]
            src = [0] * Qa + [N]
            length = [N] * Qa + [Qf]
        else:
            scale, src, length = [-1], [0], [Qf]
        self.engine.set_estimator(scale, src, length, Qa, Qf, self.G)

    def _build_error_estimation_operators_resident(self):
        """Same Gram blocks from the device-resident Riesz block R = [r_f | (n=0: q=0..Qa-1) | (n=1: ...) | ...]: only
        the Qa new representers are uploaded, only their rows of the Gram matrix are computed."""
        tp, ws = self.truth_problem, self.ws
        N, Qa, Qf = self.N, self.Qa, self.Qf
        if ws.ncols("R") == 0:
            ws.append("R", np.stack([tp.riesz_solve(f) for f in tp.operators_f], axis=1))
            self._Gr = ws.gram("R", "X")                      # (Qf x Qf) in R-order
        have = (ws.ncols("R") - Qf) // Qa
        for n in range(have, N):
            zn = self.basis_functions[:, n]
            ws.append("R", np.stack([tp.riesz_solve(-(tp.operators_a[q] @ zn)) for q in range(Qa)], axis=1))
            c0 = Qf + n * Qa
            # new representers against everything so far; X is symmetric, so X is applied to the Qa new columns only
            # (R_all^T (X R_new))^T instead of R_new^T (X R_all)
            rows = ws.gram("R", "X", "R", cols2=(c0, c0 + Qa)).T
            G = np.empty((c0 + Qa, c0 + Qa))
            G[:c0, :c0] = self._Gr
            G[c0:, :] = rows
            G[:, c0:] = rows.T
            self._Gr = G
        # R-order -> z-order [(q, n) ... | f]
        perm = [Qf + n * Qa + q for q in range(Qa) for n in range(N)] + list(range(Qf))
        self.G = self._Gr[np.ix_(perm, perm)]
        if N > 0:
            scale = list(range(Qa)) + [-1]
            src = [0] * Qa + [N]
            length = [N] * Qa + [Qf]
        else:
            scale, src, length = [-1], [0], [Qf]
        self.engine.set_estimator(scale, src, length, Qa, Qf, self.G)

    # -- RBReduction.greedy / _greedy ----------------------------------------------------------------------------
    def _greedy(self):
        with self._timed("gpu_sweep"):
            return self._sweep.max(kind=self.kind)

    def greedy(self):
        (error_estimator_max, error_estimator_argmax) = self._greedy()
        self.mu = self.training_set[error_estimator_argmax]
        self.greedy_selected_parameters.append(self.mu)
        self.greedy_selected_indices.append(int(error_estimator_argmax))
        self.greedy_error_estimators.append(error_estimator_max)
        return (error_estimator_max, error_estimator_max / self.greedy_error_estimators[0])

    def postprocess_snapshot(self, snapshot):
        """Remove the non-homogeneous Dirichlet part (differential_problem_reduction_method.py:73-108)."""
        if self.N_bc == 0:
            return snapshot
        theta_bc = self.truth_problem.compute_theta("dirichlet_bc", [self.mu])[0]
        return snapshot - self.basis_functions[:, :self.N_bc] @ theta_bc

    def update_basis_matrix(self, snapshot):
        if self.resident:
            new_basis_function, _ = self.ws.gram_schmidt("Z", "X", snapshot, append=True, first_col=self.N_bc)
        else:
            new_basis_function = self.GS.apply(snapshot, self.basis_functions[:, self.N_bc:])
        self.basis_functions = np.concatenate([self.basis_functions, new_basis_function[:, None]], axis=1)
        self.N += 1

    def _init_lifting_functions(self):
        """The lifting functions enter the basis as they are (no orthonormalisation), one at a time so that the
        incremental operator updates see them like any other new basis function."""
        liftings = np.asarray(self.truth_problem.dirichlet_bc_liftings, dtype=np.float64)
        for k in range(self.N_bc):
            if self.resident:
                self.ws.append("Z", liftings[:, k])
            self.basis_functions = np.concatenate([self.basis_functions, liftings[:, k:k + 1]], axis=1)
            self.N += 1
            if k + 1 < self.N_bc:   # (the last one is followed by the builds of offline())
                self.build_reduced_operators()
                self.build_error_estimation_operators()

    # -- RBReduction._offline ------------------------------------------------------------------------------------
    def offline(self):
        if self.N_bc and self.N == 0:
            self._init_lifting_functions()
        self.build_reduced_operators()
        self.build_error_estimation_operators()
        (absolute, relative) = self.greedy()
        if self.verbose:
            print("initial maximum absolute error estimator over training set =", absolute)
        while self.N - self.N_bc < self.Nmax and relative >= self.tol:
            with self._timed("host_truth_solve"):
                snapshot = self.postprocess_snapshot(self.truth_problem.solve(self.mu))
            with self._timed("gpu_gram_schmidt"):
                self.update_basis_matrix(snapshot)
            with self._timed("gpu_reduced_operators"):
                self.build_reduced_operators()
            with self._timed("riesz_solves_and_gram"):
                self.build_error_estimation_operators()
            (absolute, relative) = self.greedy()
            if self.verbose:
                print("N =", self.N, "max estimator", absolute, "relative", relative)
        return self

"""POD-Greedy offline loop of a parabolic problem driven by the B200 kernels (BASELINE.json configs[3]:
tutorials/06_thermal_block_unsteady, RB variant).

Host-side mirror of ``TimeDependentRBReduction`` (rbnics/reduction_methods/base/time_dependent_rb_reduction.py):
``_offline`` is inherited from ``RBReduction`` (rb_reduction.py:79-123), ``update_basis_matrix`` with the default
``POD_greedy_basis_extension = "orthogonal"`` (:150-164), ``_POD_greedy_orthogonalize_snapshot`` (:178-189: the truth
trajectory minus its projection on the current basis, ``reduced_problem.project`` = the reduced linear system
``(Z^T X Z) c = Z^T X s``), ``_POD_greedy_compute_basis_extension_with_orthogonal_snapshot`` (:191-211: POD of the K+1
orthogonalised time snapshots, N1 modes), the closure of ``_greedy`` (:262-295).  Reduced-problem side:
``build_reduced_operators`` for the terms m, a, f and ``build_error_estimation_operators`` with the Riesz terms f, a, m
and the six term pairs of rbnics/problems/parabolic/abstract_parabolic_rb_reduced_problem.py:20-27.

Where the work runs: the trajectory Gram matrix ``S^T X S`` (K+1 = 61 columns: ``rb_gram_fused_kernel``), the
projections ``Z^T M_q Z``, ``Z^T A_q Z``, ``Z^T f_q``, ``Z^T X S``, the Riesz Gram matrix and the mode construction are
the offline kernels of ``rbnics_b200.offline``; the sweep over the training set is ``parabolic.ParabolicSweep`` (one
fused kernel).  Truth trajectories, Riesz solves and the (K+1) x (K+1) eigenproblem stay on the host, as in the
reference.

The truth problem is passed in as an object with

    operators_m, operators_a : lists of scipy.sparse matrices      operators_f : list of Nh-vectors
    inner_product : scipy.sparse X                                  dt, K : time step and number of steps (T = K dt)
    solve_over_time(mu) -> (K+1, Nh) array, u^0 = 0                 riesz_solve(rhs) -> X^-1 rhs
    compute_theta(term, mus) -> (B x Q) array for "m", "a", "f"     get_stability_factor_lower_bound(mus) -> (B,)
Limitations (those of ``ParabolicSweep``): thetas constant in time, homogeneous initial condition in this loop.
"""
from __future__ import annotations

import time
from collections import defaultdict
from contextlib import contextmanager

import numpy as np

from . import _lib as L
from . import offline
from .parabolic import ParabolicSweep, simpson_weights


class PODGreedyReducedBasis:
    def __init__(self, engine, truth_problem, training_set, Nmax, tol, N1, tol1=0.0, dist=None, verbose=False,
                 resident=False):
        self.engine = engine
        self.truth_problem = tp = truth_problem
        self.training_set = [tuple(mu) for mu in training_set]
        self.Nmax, self.tol, self.N1, self.tol1 = Nmax, tol, N1, tol1
        self.dist = dist
        self.verbose = verbose
        self.Qm, self.Qa, self.Qf = len(tp.operators_m), len(tp.operators_a), len(tp.operators_f)
        self.Nh = tp.operators_f[0].shape[0]
        self.dt, self.K = float(tp.dt), int(tp.K)
        mus = np.asarray(self.training_set, dtype=np.float64)
        rank, size = (dist.rank, dist.size) if dist is not None else (0, 1)
        self._rank, self._size = rank, size
        sl = slice(rank, None, size)                       # cyclic shard (parameter_space_subset.py:57)
        self._Tm = np.ascontiguousarray(tp.compute_theta("m", mus)[sl])
        self._Ta = np.ascontiguousarray(tp.compute_theta("a", mus)[sl])
        self._Tf = np.ascontiguousarray(tp.compute_theta("f", mus)[sl])
        self._beta = np.ascontiguousarray(tp.get_stability_factor_lower_bound(mus)[sl])
        # reduced problem state
        self.N = 0
        self.basis_functions = np.zeros((self.Nh, 0))
        self.operator_m = np.zeros((self.Qm, 0, 0))
        self.operator_a = np.zeros((self.Qa, 0, 0))
        self.operator_f = np.zeros((self.Qf, 0))
        self.riesz_f = None
        self.riesz_a = [np.zeros((self.Nh, 0)) for _ in range(self.Qa)]
        self.riesz_m = [np.zeros((self.Nh, 0)) for _ in range(self.Qm)]
        self.error_estimation_operator = {}
        self.POD_time_trajectory = offline.ProperOrthogonalDecomposition(engine, tp.inner_product, dist=None)
        self.greedy_selected_parameters = []
        self.greedy_selected_indices = []
        self.greedy_error_estimators = []
        self.mu = None
        self.timings = defaultdict(float)   # wall seconds per stage, summed over the iterations (bench_pod_greedy.py)
        # resident=True keeps the truth operators, the basis, the trajectory of the iteration and the Riesz representers
        # in HBM (offline.OfflineWorkspace): a trajectory crosses PCIe once, the modes and the representers once
        self.resident = resident
        self.ws = None
        if resident:
            ws = offline.OfflineWorkspace(engine)
            ws.set_operator("X", tp.inner_product)
            for q, M in enumerate(tp.operators_m):
                ws.set_operator(f"M{q}", M)
            for q, A in enumerate(tp.operators_a):
                ws.set_operator(f"A{q}", A)
            cap = Nmax + N1
            ws.new_block("Z", self.Nh, cap)
            ws.new_block("Y", self.Nh, cap)
            ws.new_block("F", self.Nh, self.Qf)
            ws.append("F", np.stack([np.asarray(f, dtype=np.float64) for f in tp.operators_f], axis=1))
            ws.new_block("R", self.Nh, (self.Qa + self.Qm) * cap + self.Qf)   # r_f, then per basis function: a_q, m_q
            self.ws = ws

    @contextmanager
    def _timed(self, stage):
        t0 = time.perf_counter()
        try:
            yield
        finally:
            self.timings[stage] += time.perf_counter() - t0

    # -- reduced_problem.build_reduced_operators() ---------------------------------------------------------------
    def build_reduced_operators(self):
        tp, Z = self.truth_problem, self.basis_functions
        if self.N == 0:
            return
        if self.resident:
            ws, N = self.ws, self.N
            with self._timed("gpu_projections"):
                def project(name):   # Z^T (A Z) on resident operands: one SpMM into the scratch block, one Gram pass
                    ws.spmm(name, "Z", (0, N), "Y", 0)
                    return ws.gram("Z", None, "Y", cols2=(0, N))
                self.operator_m = np.stack([project(f"M{q}") for q in range(self.Qm)])
                self.operator_a = np.stack([project(f"A{q}") for q in range(self.Qa)])
                self.operator_f = ws.gram("Z", None, "F").T.copy()
            return
        with self._timed("gpu_projections"):
            self.operator_m = offline.project_matrices(self.engine, Z, tp.operators_m)
            self.operator_a = offline.project_matrices(self.engine, Z, tp.operators_a)
            self.operator_f = offline.project_vectors(self.engine, Z, tp.operators_f)

    # -- reduced_problem.build_error_estimation_operators() ------------------------------------------------------
    def build_error_estimation_operators(self):
        tp, N, Qa, Qm, Qf = self.truth_problem, self.N, self.Qa, self.Qm, self.Qf
        with self._timed("host_riesz_solves"):
            if self.riesz_f is None:   # order-1 term: once (rb_reduced_problem.py:156-163)
                self.riesz_f = np.stack([tp.riesz_solve(f) for f in tp.operators_f], axis=1)
            for ops, store in ((tp.operators_a, self.riesz_a), (tp.operators_m, self.riesz_m)):
                for q, A in enumerate(ops):   # representers of -A_q z_n for the new basis functions (:183-195)
                    new = [tp.riesz_solve(-(A @ self.basis_functions[:, n])) for n in range(store[q].shape[1], N)]
                    if new:
                        store[q] = np.concatenate([store[q], np.stack(new, axis=1)], axis=1)
        if self.resident:
            # device block R = [r_f | (n = 0: a_0 .. a_Qa-1, m_0 .. m_Qm-1) | (n = 1: ...) | ...]: only the representers of
            # the new basis functions are uploaded
            ws, Qt = self.ws, Qa + Qm
            with self._timed("gpu_riesz_gram"):
                if ws.ncols("R") == 0:
                    ws.append("R", self.riesz_f)
                for n in range((ws.ncols("R") - Qf) // Qt, N):
                    ws.append("R", np.stack([self.riesz_a[q][:, n] for q in range(Qa)]
                                            + [self.riesz_m[q][:, n] for q in range(Qm)], axis=1))
                GR = ws.gram("R", "X")
            perm = ([Qf + n * Qt + q for q in range(Qa) for n in range(N)]
                    + [Qf + n * Qt + Qa + q for q in range(Qm) for n in range(N)] + list(range(Qf)))
            G = GR[np.ix_(perm, perm)]
        else:
            # all Riesz Gram blocks in one product: S = [R_a (q-major) | R_m (q-major) | r_f]
            S = np.ascontiguousarray(np.concatenate(self.riesz_a + self.riesz_m + [self.riesz_f], axis=1))
            with self._timed("gpu_riesz_gram"):
                G = offline.gram(self.engine, S, tp.inner_product)
        off = {"a": 0, "m": Qa * N, "f": (Qa + Qm) * N}
        Q = {"a": Qa, "m": Qm, "f": Qf}

        def block(t1, t2):
            o1, o2 = t1 != "f", t2 != "f"
            n1, n2 = Q[t1] * (N if o1 else 1), Q[t2] * (N if o2 else 1)
            B = G[off[t1]:off[t1] + n1, off[t2]:off[t2] + n2]
            if o1 and o2:
                return B.reshape(Q[t1], N, Q[t2], N).transpose(0, 2, 1, 3).copy()
            if o1:
                return B.reshape(Q[t1], N, Q[t2]).transpose(0, 2, 1).copy()
            return B.copy()

        self.error_estimation_operator = {k: block(*k) for k in
                                          [("f", "f"), ("a", "f"), ("a", "a"), ("m", "f"), ("m", "a"), ("m", "m")]}

    # -- TimeDependentRBReduction._greedy ------------------------------------------------------------------------
    def _greedy(self):
        with self._timed("gpu_sweep"):
            return self._greedy_sweep()

    def _greedy_sweep(self):
        E = self.error_estimation_operator
        if self.N == 0:
            # no reduced solution yet: the residual is f for every t > 0, eps2 = theta_f^T G_ff theta_f; one elliptic
            # sweep of an empty system gives sqrt(|eps2| / beta), the time quadrature of a constant is its weight sum
            eng = self.engine
            eng.set_system(np.zeros((self.Qa, 0, 0)), np.zeros((self.Qf, 0)))
            eng.set_estimator([-1], [0], [self.Qf], self.Qa, self.Qf, E["f", "f"])
            Theta = np.ascontiguousarray(np.concatenate([self._Ta, self._Tf], axis=1))
            r = eng.sweep(Theta, self._beta, kind=L.RB_EST_SQRT_OF_EPS2_OVER_BETA, index_offset=self._rank,
                          index_stride=self._size)
            scale = float(np.sqrt(simpson_weights(self.K, self.dt)[1:].sum()))
            mx, am = r["max"] * scale, r["argmax"]
        else:
            sw = ParabolicSweep(self.engine, self.operator_m, self.operator_a, self.operator_f, E, dt=self.dt, K=self.K)
            r = sw.sweep(self._Tm, self._Ta, self._Tf, self._beta, index_offset=self._rank, index_stride=self._size)
            mx, am = r["max"], r["argmax"]
        if self.dist is not None and self._size > 1:
            mx, am = self.dist.maxloc(mx, am)
        return mx, am

    def greedy(self):
        (error_estimator_max, error_estimator_argmax) = self._greedy()
        self.mu = self.training_set[error_estimator_argmax]
        self.greedy_selected_parameters.append(self.mu)
        self.greedy_selected_indices.append(int(error_estimator_argmax))
        self.greedy_error_estimators.append(error_estimator_max)
        return (error_estimator_max, error_estimator_max / self.greedy_error_estimators[0])

    # -- TimeDependentRBReduction.update_basis_matrix ("orthogonal" extension) ------------------------------------
    def _POD_greedy_orthogonalize_snapshot(self, snapshot_over_time):
        S = np.ascontiguousarray(np.asarray(snapshot_over_time, dtype=np.float64).T)      # Nh x (K+1)
        if self.N == 0:
            return S
        Z, X = self.basis_functions, self.truth_problem.inner_product
        with self._timed("gpu_orthogonalize"):
            inner_product_N = offline.gram(self.engine, Z, X)                             # Z^T X Z
            rhs = offline.gram(self.engine, Z, X, S)                                      # Z^T X s_k for every k
            projected = np.linalg.solve(inner_product_N, rhs)                             # reduced_problem.project
            return S - offline.lincomb(self.engine, Z, projected)

    def _POD_greedy_compute_basis_extension_with_orthogonal_snapshot(self, orthogonal_snapshot_over_time):
        pod = self.POD_time_trajectory
        pod.clear()
        for k in range(orthogonal_snapshot_over_time.shape[1]):
            pod.store_snapshot(orthogonal_snapshot_over_time[:, k])
        with self._timed("gpu_trajectory_pod"):
            (_, _, basis_functions1, N1) = pod.apply(self.N1, self.tol1)
        return basis_functions1, N1

    def _update_basis_matrix_resident(self, snapshot_over_time):
        """The same two steps on device-resident data: the trajectory is uploaded once into block S, orthogonalised in
        place (S -= Z c, rb_off_lincomb_add), its Gram matrix and the modes Z[:, N:N+N1] = S V are computed there, and
        only the N1 new basis functions come back (the host needs them for the Riesz solves)."""
        import scipy.linalg
        ws, N = self.ws, self.N
        S = np.ascontiguousarray(np.asarray(snapshot_over_time, dtype=np.float64).T)      # Nh x (K+1)
        with self._timed("gpu_orthogonalize"):
            ws.new_block("S", self.Nh, S.shape[1])
            ws.append("S", S)
            if N > 0:
                inner_product_N = ws.gram("Z", "X", cols=(0, N))
                rhs = ws.gram("Z", "X", "S", cols=(0, N))
                projected = np.linalg.solve(inner_product_N, rhs)
                ws.lincomb("Z", -projected, "S", col0=0, cols=(0, N), accumulate=True)
        with self._timed("gpu_trajectory_pod"):
            correlation = ws.gram("S", "X")
            eigs, eigv = scipy.linalg.eigh(correlation)          # online/numpy/eigen_solver.py:39
            idx = eigs.argsort()[::-1]
            eigs, eigv = eigs[idx], eigv[:, idx]
            abs_e = np.abs(eigs)
            total = abs_e.sum()
            retained = np.cumsum(abs_e) / total if total > 0.0 else np.ones(len(eigs))
            N1 = 0
            for n in range(min(self.N1, S.shape[1])):            # proper_orthogonal_decomposition_base.py:68-92
                N1 = n + 1
                if self.tol1 > 0.0 and retained[n] > 1.0 - self.tol1:
                    break
            ws.lincomb("S", np.ascontiguousarray(eigv[:, :N1]), "Z", col0=N)
            norms2 = ws.column_norms2("Z", "X", cols=(N, N + N1))
            ws.scale_columns("Z", np.array([1.0 / np.sqrt(v) if v > 0.0 else 1.0 for v in norms2]), col0=N)
            basis_functions1 = ws.read("Z", N, N1)
        self.basis_functions = np.ascontiguousarray(np.concatenate([self.basis_functions, basis_functions1], axis=1))
        self.N += N1

    def update_basis_matrix(self, snapshot_over_time):
        if self.resident:
            return self._update_basis_matrix_resident(snapshot_over_time)
        orthogonal = self._POD_greedy_orthogonalize_snapshot(snapshot_over_time)
        (basis_functions1, N1) = self._POD_greedy_compute_basis_extension_with_orthogonal_snapshot(orthogonal)
        self.basis_functions = np.ascontiguousarray(np.concatenate([self.basis_functions, basis_functions1], axis=1))
        self.N += N1

    # -- RBReduction._offline ------------------------------------------------------------------------------------
    def offline(self):
        self.build_reduced_operators()
        self.build_error_estimation_operators()
        (absolute, relative) = self.greedy()
        if self.verbose:
            print("initial maximum absolute error estimator over training set =", absolute)
        while self.N < self.Nmax and relative >= self.tol:
            with self._timed("host_truth_trajectory"):
                snapshot_over_time = self.truth_problem.solve_over_time(self.mu)
            self.update_basis_matrix(snapshot_over_time)
            self.build_reduced_operators()
            self.build_error_estimation_operators()
            (absolute, relative) = self.greedy()
            if self.verbose:
                print("N =", self.N, "max estimator", absolute, "relative", relative)
        return self

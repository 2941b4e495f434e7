"""POD-Galerkin offline stage driven by the B200 offline kernels (BASELINE.json configs 2 / 4).

Host-side mirror of ``PODGalerkinReduction`` (rbnics/reduction_methods/base/pod_galerkin_reduction.py:108-186):
``_offline`` = one truth solve per parameter of the training set, ``update_snapshots_matrix``, then
``compute_basis_functions`` (``POD.apply(Nmax, tol)``: Gram matrix S^T X S, eigen-decomposition, modes, normalisation --
rbnics/backends/basic/proper_orthogonal_decomposition_base.py:39-92) and ``reduced_problem.build_reduced_operators``.
Same method names and bookkeeping (``N``, ``basis_functions``, ``POD.eigenvalues``, ``POD.retained_energy``).

``resident=True`` (default): snapshots go into a device-resident block as they are produced (each crosses PCIe once);
the Gram matrix, the modes ``S V``, their norms and the projections ``Z^T A_q Z`` / ``Z^T f_q`` run on resident data
(``offline.ResidentProperOrthogonalDecomposition``, ``OfflineWorkspace``); only the basis comes back.  The truth solves
are the reference's input generator and stay on the host; the truth problem is an object with ``operators_a``,
``operators_f``, ``inner_product`` and ``solve(mu)`` (see ``rbnics_b200.greedy``)."""
from __future__ import annotations

import numpy as np

from . import offline


class PODGalerkin:
    def __init__(self, engine, truth_problem, training_set, Nmax, tol=0.0, resident=True):
        self.engine = engine
        self.truth_problem = tp = truth_problem
        self.training_set = [tuple(mu) for mu in training_set]
        self.Nmax, self.tol = Nmax, tol
        self.resident = resident
        self.Qa, self.Qf = len(tp.operators_a), len(tp.operators_f)
        self.Nh = tp.operators_f[0].shape[0]
        self.N = 0
        self.basis_functions = np.zeros((self.Nh, 0))
        self.operator_a = np.zeros((self.Qa, 0, 0))
        self.operator_f = np.zeros((self.Qf, 0))
        if resident:
            self.ws = ws = offline.OfflineWorkspace(engine)
            ws.set_operator("X", tp.inner_product)
            for q, A in enumerate(tp.operators_a):
                ws.set_operator(f"A{q}", A)
            self.POD = offline.ResidentProperOrthogonalDecomposition(ws, "X", self.Nh, max(len(self.training_set), 2),
                                                                    snapshots="pod_snapshots", basis="Z")
        else:
            self.ws = None
            self.POD = offline.ProperOrthogonalDecomposition(engine, tp.inner_product)

    def update_snapshots_matrix(self, snapshot):
        self.POD.store_snapshot(snapshot)

    def compute_basis_functions(self):
        (_, _, basis_functions, N) = self.POD.apply(self.Nmax, self.tol)
        self.basis_functions = np.ascontiguousarray(basis_functions)
        self.N += N

    def build_reduced_operators(self):
        tp, N = self.truth_problem, self.N
        if self.resident:
            ws = self.ws
            ws.new_block("Y", self.Nh, max(N, 2))
            ws.new_block("F", self.Nh, max(self.Qf, 2))
            ws.append("F", np.stack([np.asarray(f, dtype=np.float64) for f in tp.operators_f], axis=1))

            def project(name):   # Z^T (A Z) on resident operands
                ws.spmm(name, "Z", (0, N), "Y", 0)
                return ws.gram("Z", None, "Y", cols=(0, N), cols2=(0, N))
            self.operator_a = np.stack([project(f"A{q}") for q in range(self.Qa)])
            self.operator_f = ws.gram("Z", None, "F", cols=(0, N), cols2=(0, self.Qf)).T.copy()
        else:
            Z = self.basis_functions
            self.operator_a = offline.project_matrices(self.engine, Z, tp.operators_a)
            self.operator_f = offline.project_vectors(self.engine, Z, tp.operators_f)

    def offline(self):
        for mu in self.training_set:
            snapshot = self.truth_problem.solve(mu)
            self.update_snapshots_matrix(snapshot)
        self.compute_basis_functions()
        self.build_reduced_operators()
        return self

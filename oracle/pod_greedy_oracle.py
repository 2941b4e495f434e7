"""CPU restatement of the reference's POD-Greedy offline loop for parabolic problems (TEST INFRASTRUCTURE ONLY).

rbnics/reduction_methods/base/time_dependent_rb_reduction.py: `_offline` (inherited, rb_reduction.py:79-123),
`update_basis_matrix` with the "orthogonal" extension (:150-164), `_POD_greedy_orthogonalize_snapshot` (:178-189),
`_POD_greedy_compute_basis_extension_with_orthogonal_snapshot` (:191-211), the closure of `_greedy` (:262-295); the
reduced problem's operators and the six error-estimation storages (problems/base/rb_reduced_problem.py:142-263 with
the terms of problems/parabolic/abstract_parabolic_rb_reduced_problem.py:20-27), built from the loop-faithful pieces of
``rb_oracle``."""
from __future__ import annotations

import numpy as np

from . import rb_oracle as O


def _storages(Z, X, tp):
    """Riesz representers r_f, -X^-1 A_q z_n, -X^-1 M_q z_n and their Gram blocks as nested lists (the layout of
    rb_oracle.parabolic_residual_norm_squared)."""
    N = Z.shape[1]
    r_f = [tp.riesz_solve(f) for f in tp.operators_f]

    def reps(ops):
        return [np.stack([-tp.riesz_solve(A @ Z[:, n]) for n in range(N)], axis=1) if N > 0 else np.zeros((Z.shape[0], 0))
                for A in ops]

    R_a, R_m = reps(tp.operators_a), reps(tp.operators_m)
    ff = [[float(np.dot(a, X @ b)) for b in r_f] for a in r_f]
    vec = lambda R: [[Ri.T @ (X @ b) for b in r_f] for Ri in R]                      # noqa: E731
    mat = lambda R1, R2: [[O.gram_blas(Ri, X, Rj) for Rj in R2] for Ri in R1]        # noqa: E731
    return {"ff": ff, "af": vec(R_a), "aa": mat(R_a, R_a), "mf": vec(R_m), "ma": mat(R_m, R_a), "mm": mat(R_m, R_m)}


def pod_greedy_offline(tp, training_set, Nmax, tol, N1, tol1=0.0):
    Nh = tp.operators_f[0].shape[0]
    X = tp.inner_product
    mus = np.asarray(training_set, dtype=np.float64)
    Tm, Ta, Tf = tp.compute_theta("m", mus), tp.compute_theta("a", mus), tp.compute_theta("f", mus)
    beta = tp.get_stability_factor_lower_bound(mus)
    dt, K = float(tp.dt), int(tp.K)
    Z = np.zeros((Nh, 0))
    selected, estimators = [], []

    def sweep():
        N = Z.shape[1]
        M_q = [O.gram_blas(Z, M) for M in tp.operators_m]
        A_q, f_q = O.build_reduced_operators(Z, tp.operators_a, tp.operators_f)
        G = _storages(Z, X, tp)
        d, i, _, _ = O.parabolic_sweep_loop(Tm, Ta, Tf, beta, M_q, A_q, f_q, G, N, dt, K)
        return i, float(d[i])

    i, v = sweep()
    selected.append(i)
    estimators.append(v)
    while Z.shape[1] < Nmax and estimators[-1] / estimators[0] >= tol:
        S = np.asarray(tp.solve_over_time(training_set[selected[-1]]), dtype=np.float64).T       # Nh x (K+1)
        if Z.shape[1] > 0:   # snapshot - basis_functions * project(snapshot): (Z^T X Z) c = Z^T X s
            c = np.linalg.solve(O.gram_blas(Z, X), Z.T @ (X @ S))
            S = S - Z @ c
        _, _, Z1, n1, _ = O.pod(S, X, N1, tol1)
        Z = np.concatenate([Z, Z1], axis=1)
        i, v = sweep()
        selected.append(i)
        estimators.append(v)
    return {"selected": selected, "estimators": estimators, "Z": Z}

"""CPU restatement of the reference's RB greedy offline loop (TEST INFRASTRUCTURE ONLY).

rbnics/reduction_methods/base/rb_reduction.py:79-185 with the per-parameter closure of :173-178, built from the
loop-faithful pieces of ``rb_oracle``."""
from __future__ import annotations

import numpy as np

from . import rb_oracle as O


def greedy_offline(tp, training_set, Nmax, tol, kind="sqrt_eps2_over_beta", vectorised=False):
    Qa, Qf = len(tp.operators_a), len(tp.operators_f)
    Nh = tp.operators_f[0].shape[0]
    X = tp.inner_product
    mus = np.asarray(training_set, dtype=np.float64)
    Theta_a = tp.compute_theta("a", mus)
    Theta_f = tp.compute_theta("f", mus)
    beta = tp.get_stability_factor_lower_bound(mus)
    # non-homogeneous Dirichlet conditions: the lifting functions are the first N_bc basis functions
    # (rbnics/problems/base/parametrized_reduced_differential_problem.py:283-299; N counts the others, :329)
    liftings = getattr(tp, "dirichlet_bc_liftings", None)
    Z = np.zeros((Nh, 0)) if liftings is None else np.array(liftings, dtype=np.float64)
    N_bc = Z.shape[1]
    Theta_bc = tp.compute_theta("dirichlet_bc", mus) if N_bc else None
    selected, estimators, all_deltas = [], [], []

    def sweep(N):
        A_q, f_q = O.build_reduced_operators(Z, tp.operators_a, tp.operators_f)
        G_ff, G_af, G_aa, _, _ = O.build_error_estimation_operators(Z, X, tp.operators_a, tp.operators_f, tp.riesz_solve)
        if vectorised:
            A, F, Gff, Gaf, Gaa = O.stack_operators(A_q, f_q, G_ff, G_af, G_aa, N + N_bc)
            d, U, i, v = O.sweep_batched(Theta_a, Theta_f, beta, A, F, Gff, Gaf, Gaa, Theta_bc=Theta_bc, kind=kind)
        else:
            d, U, i, v = O.sweep_loop(Theta_a, Theta_f, beta, A_q, f_q, G_ff, G_af, G_aa, N + N_bc, Theta_bc=Theta_bc,
                                      kind=kind)
        return d, i, v

    d, i, v = sweep(0)
    selected.append(i)
    estimators.append(v)
    all_deltas.append(d)
    N = 0
    while N < Nmax and estimators[-1] / estimators[0] >= tol:
        mu = training_set[selected[-1]]
        snapshot = tp.solve(mu)
        if N_bc:  # postprocess_snapshot (differential_problem_reduction_method.py:73-108); GS against the basis
            # functions after the liftings only (rb_reduction.py:125-139)
            snapshot = snapshot - Z[:, :N_bc] @ tp.compute_theta("dirichlet_bc", [mu])[0]
        z = O.gram_schmidt(snapshot, Z[:, N_bc:], X)
        Z = np.concatenate([Z, z[:, None]], axis=1)
        N += 1
        d, i, v = sweep(N)
        selected.append(i)
        estimators.append(v)
        all_deltas.append(d)
    return {"selected": selected, "estimators": estimators, "Z": Z, "deltas": all_deltas}

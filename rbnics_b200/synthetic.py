"""Seeded synthetic reduced problems of the shapes BASELINE.json names.  Input generators only."""
from __future__ import annotations

import numpy as np


def elliptic_reduced_problem(N, Qa, Qf, seed=0, riesz_rank=512, n_bc=0):
    """Well-conditioned affine elliptic reduced system + consistent estimator operators (eps2 >= 0).

    A_q = B_q B_q^T / N + I (SPD), f_q ~ N(0,1); G = R^T R / riesz_rank partitioned into Gff, Gaf, Gaa with
    R in R^{K x (Qf + Qa N)} -- the Gram matrix of K-dimensional "Riesz representers", so the reference's
    ``assert eps2 >= 0`` (elliptic_rb_reduced_problem.py:38) holds for every parameter."""
    rng = np.random.default_rng(seed)
    A = np.empty((Qa, N, N))
    for q in range(Qa):
        Bq = rng.standard_normal((N, N))
        A[q] = Bq @ Bq.T / max(N, 1) + np.eye(N)
    F = rng.standard_normal((Qf, N))
    Ka = Qa * N
    R = rng.standard_normal((riesz_rank, Ka + Qf))
    G = R.T @ R / riesz_rank
    Gaa = G[:Ka, :Ka].reshape(Qa, N, Qa, N).transpose(0, 2, 1, 3).copy()
    Gaf = G[:Ka, Ka:].reshape(Qa, N, Qf).transpose(0, 2, 1).copy()
    Gff = G[Ka:, Ka:].copy()
    return {"A": A, "F": F, "Gff": Gff, "Gaf": Gaf, "Gaa": Gaa, "N": N, "Qa": Qa, "Qf": Qf, "n_bc": n_bc}


def training_set(B, Qa, Qf, seed=1, n_bc=0, out_theta=None, out_beta=None):
    """theta_a ~ U[0.1, 10], theta_f ~ U[-1, 1], beta = min_q theta_a (tutorial 01's stability factor)."""
    rng = np.random.default_rng(seed)
    Theta = out_theta if out_theta is not None else np.empty((B, Qa + Qf))
    Theta[:, :Qa] = rng.uniform(0.1, 10.0, size=(B, Qa))
    Theta[:, Qa:] = rng.uniform(-1.0, 1.0, size=(B, Qf))
    beta = out_beta if out_beta is not None else np.empty(B)
    beta[:] = Theta[:, :Qa].min(axis=1)
    ThetaBC = rng.uniform(-1.0, 1.0, size=(B, n_bc)) if n_bc else None
    return Theta, beta, ThetaBC

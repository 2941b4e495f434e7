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


def stokes_reduced_problem(n, Q, seed=0):
    """Saddle-point reduced system with components (u, s, p) of n basis functions each (N = 3n), the shape of
    tutorials/12_stokes (BASELINE.json configs[4]): a on the velocity block, bt = [0 B^T; 0 0], b = [0 0; B 0]
    (indefinite: pivoting matters); Riesz Gram blocks from velocity-space representers (a, bt, f) and pressure-space
    representers (b, g), so the pairs the reference skips are exactly zero and eps2 >= 0.
    Q = {"a", "b", "bt", "f", "g"} term counts.  Returns (op, E): op[term] stacked operators, E[(t1, t2)] blocks."""
    rng = np.random.default_rng(seed)
    nv, N = 2 * n, 3 * n
    op = {"a": np.zeros((Q["a"], N, N)), "b": np.zeros((Q["b"], N, N)), "bt": np.zeros((Q["bt"], N, N)),
          "f": np.zeros((Q["f"], N)), "g": np.zeros((Q["g"], N))}
    for q in range(Q["a"]):
        Bq = rng.standard_normal((nv, nv))
        op["a"][q, :nv, :nv] = Bq @ Bq.T / nv + np.eye(nv)
    for q in range(Q["b"]):
        op["b"][q, nv:, :nv] = rng.standard_normal((n, nv))
    for q in range(Q["bt"]):
        op["bt"][q, :nv, nv:] = op["b"][q % Q["b"], nv:, :nv].T
    op["f"][:, :nv] = rng.standard_normal((Q["f"], nv))
    op["g"][:, nv:] = rng.standard_normal((Q["g"], n))
    Kv = Q["a"] * N + Q["bt"] * N + Q["f"]
    Kp = Q["b"] * N + Q["g"]
    Rv = rng.standard_normal((min(Kv, 256) + 7, Kv))
    Rp = rng.standard_normal((min(Kp, 256) + 5, Kp))
    Gv = Rv.T @ Rv / Rv.shape[0]
    Gp = Rp.T @ Rp / Rp.shape[0]
    offv = {"a": 0, "bt": Q["a"] * N, "f": Q["a"] * N + Q["bt"] * N}
    offp = {"b": 0, "g": Q["b"] * N}
    ops = ("a", "b", "bt")

    def block(Gm, off, t1, t2):
        n1 = Q[t1] * N if t1 in ops else Q[t1]
        n2 = Q[t2] * N if t2 in ops else Q[t2]
        M = Gm[off[t1]:off[t1] + n1, off[t2]:off[t2] + n2]
        if t1 in ops and t2 in ops:
            return M.reshape(Q[t1], N, Q[t2], N).transpose(0, 2, 1, 3).copy()
        if t1 in ops:
            return M.reshape(Q[t1], N, Q[t2]).transpose(0, 2, 1).copy()
        return M.copy()

    pairs = [("f", "f"), ("g", "g"), ("a", "f"), ("bt", "f"), ("b", "g"), ("a", "a"), ("a", "bt"), ("bt", "bt"),
             ("b", "b")]
    E = {(t1, t2): (block(Gv, offv, t1, t2) if t1 in offv else block(Gp, offp, t1, t2)) for (t1, t2) in pairs}
    return op, E

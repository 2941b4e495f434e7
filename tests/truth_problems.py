"""FEniCS-free stand-ins for the reference's truth problems (INPUT GENERATORS for the tests; the high-fidelity
solves are out of scope and stay on the host).

ThermalBlock mimics tutorials/01_thermal_block: a P1-like 5-point diffusion operator on an n x n grid split into
``nblocks`` vertical sub-domain strips with conductivities (mu_0, ..., 1), homogeneous Dirichlet boundary folded in,
one or more load terms scaled by further parameters; X = sum_q A_q (the H^1_0 inner product at mu = 1)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def _strip_operator(n, x_lo, x_hi):
    """5-point diffusion stencil with conductivity 1 on edges whose midpoint lies in [x_lo, x_hi), 0 elsewhere."""
    h = 1.0 / (n + 1)
    idx = lambda i, j: i * n + j  # noqa: E731
    rows, cols, vals = [], [], []

    def add(a, b, v):
        rows.append(a)
        cols.append(b)
        vals.append(v)

    for i in range(n):          # x index
        for j in range(n):      # y index
            p = idx(i, j)
            # edges in x direction: between (i-1,j)-(i,j) and (i,j)-(i+1,j); midpoint x = (i+0.5)h or (i+1.5)h
            for di in (-1, 1):
                xm = (i + 1 + 0.5 * di) * h
                k = 1.0 if x_lo <= xm < x_hi else 0.0
                if k:
                    add(p, p, k)
                    if 0 <= i + di < n:
                        add(p, idx(i + di, j), -k)
            # edges in y direction lie at x = (i+1)h
            xm = (i + 1) * h
            k = 1.0 if x_lo <= xm < x_hi else 0.0
            if k:
                for dj in (-1, 1):
                    add(p, p, k)
                    if 0 <= j + dj < n:
                        add(p, idx(i, j + dj), -k)
    return sp.csr_matrix((vals, (rows, cols)), shape=(n * n, n * n))


class ThermalBlock:
    def __init__(self, n=12, nblocks=2, nloads=1, mu_range=None, seed=0):
        self.n = n
        self.Nh = n * n
        edges = np.linspace(0.0, 1.0 + 1e-12, nblocks + 1)
        self.operators_a = [_strip_operator(n, edges[b], edges[b + 1]) for b in range(nblocks)]
        rng = np.random.default_rng(seed)
        xs = (np.arange(n) + 1.0) / (n + 1)
        Xg, Yg = np.meshgrid(xs, xs, indexing="ij")
        self.operators_f = []
        for k in range(nloads):
            f = np.sin((k + 1) * np.pi * Xg) * np.cos(k * np.pi * Yg) + 0.1 * rng.standard_normal((n, n))
            self.operators_f.append((f / (n + 1) ** 2).ravel())
        self.inner_product = sum(self.operators_a[1:], self.operators_a[0]).tocsr()
        self._lu = spla.splu(self.inner_product.tocsc())
        self.nblocks, self.nloads = nblocks, nloads
        self.mu_range = mu_range or ([(0.1, 10.0)] * (nblocks - 1) + [(-1.0, 1.0)] * nloads)

    # theta_a = (mu_0, ..., mu_{nb-2}, 1), theta_f = (mu_{nb-1}, ...)   (tutorial_thermal_block.ipynb:165-169)
    def compute_theta(self, term, mus):
        mus = np.atleast_2d(np.asarray(mus, dtype=np.float64))
        nb = self.nblocks
        if term == "a":
            return np.concatenate([mus[:, :nb - 1], np.ones((mus.shape[0], 1))], axis=1)
        if term == "f":
            return mus[:, nb - 1:nb - 1 + self.nloads].copy()
        raise ValueError(term)

    def get_stability_factor_lower_bound(self, mus):
        return self.compute_theta("a", mus).min(axis=1)          # min(theta_a) (:158-159)

    def solve(self, mu):
        tha = self.compute_theta("a", [mu])[0]
        thf = self.compute_theta("f", [mu])[0]
        A = sum(t * Aq for t, Aq in zip(tha, self.operators_a))
        f = sum(t * fq for t, fq in zip(thf, self.operators_f))
        return spla.spsolve(A.tocsc(), f)

    def riesz_solve(self, rhs):
        return self._lu.solve(np.asarray(rhs, dtype=np.float64))

    def sample_training_set(self, ntrain, seed=0):
        """numpy.random.uniform per coordinate, sample-major (rbnics/sampling/distributions/uniform_distribution.py:11-19)."""
        rs = np.random.RandomState(seed)
        out = []
        for _ in range(ntrain):
            out.append(tuple(float(rs.uniform(lo, hi)) for (lo, hi) in self.mu_range))
        return out

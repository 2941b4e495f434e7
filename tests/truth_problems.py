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


class ThermalBlockUnsteady(ThermalBlock):
    """tutorials/06_thermal_block_unsteady as a FEniCS-free stand-in: the thermal block above with a mass term,
    m(u, v) = int u v (lumped: identity / (n+1)^2), theta_m = (1,), implicit Euler with step dt up to T = K dt from a
    homogeneous initial condition (tutorial_thermal_block_unsteady.ipynb: dt = 0.05, T = 3)."""

    def __init__(self, n=12, nblocks=2, nloads=1, dt=0.05, K=60, **kw):
        ThermalBlock.__init__(self, n=n, nblocks=nblocks, nloads=nloads, **kw)
        self.operators_m = [(sp.identity(self.Nh) / (n + 1) ** 2).tocsr()]
        self.dt, self.K = float(dt), int(K)

    def compute_theta(self, term, mus):
        if term == "m":
            return np.ones((np.atleast_2d(np.asarray(mus)).shape[0], 1))
        return ThermalBlock.compute_theta(self, term, mus)

    def solve_over_time(self, mu):
        """(K+1) x Nh: u^0 = 0, (M/dt + A) u^k = M u^{k-1}/dt + f (time_stepping.py:104-112 at truth level)."""
        thm = self.compute_theta("m", [mu])[0]
        tha = self.compute_theta("a", [mu])[0]
        thf = self.compute_theta("f", [mu])[0]
        M = sum(t * Mq for t, Mq in zip(thm, self.operators_m))
        A = sum(t * Aq for t, Aq in zip(tha, self.operators_a))
        f = sum(t * fq for t, fq in zip(thf, self.operators_f))
        lu = spla.splu((M / self.dt + A).tocsc())
        U = np.zeros((self.K + 1, self.Nh))
        for k in range(1, self.K + 1):
            U[k] = lu.solve(M @ U[k - 1] / self.dt + f)
        return U


# ----------------------------------------------------------------------------------------------------------------------
# The reference's OWN tutorial meshes (tests/golden/mesh_*.npz, extracted from the DOLFIN XML files by
# tests/golden/make_mesh_fixtures.py) with a plain-numpy P1 assembler: same geometry, sub-domains, boundary markers,
# affine terms and parameter ranges as tutorials/01_thermal_block and tutorials/02_elastic_block.  What differs from
# the reference is only who assembles (numpy here, FEniCS there) and how the homogeneous Dirichlet condition enters
# (constrained dofs eliminated here; identity rows in DOLFIN) -- the reduced problem sees the same operators restricted
# to the free dofs.
# ----------------------------------------------------------------------------------------------------------------------
import os as _os

_GOLD = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden")


class P1Mesh:
    def __init__(self, name):
        m = np.load(_os.path.join(_GOLD, "mesh_%s.npz" % name))
        self.vertices, self.cells = m["vertices"], m["cells"]
        self.cell_markers, self.marked_facets = m["cell_markers"], m["marked_facets"]
        self.nv = self.vertices.shape[0]
        x = self.vertices[self.cells]                                  # nc x 3 x 2
        e1, e2 = x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]
        self.area = 0.5 * np.abs(e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0])
        # gradients of the three hat functions on every cell: grad(phi_i) = rot90(x_k - x_j) / (2 area), (i, j, k) cyclic
        det = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
        g = np.zeros((self.cells.shape[0], 3, 2))
        for i in range(3):
            j, k = (i + 1) % 3, (i + 2) % 3
            d = x[:, j] - x[:, k]
            g[:, i, 0] = d[:, 1] / det
            g[:, i, 1] = -d[:, 0] / det
        self.grad = g

    def facet_vertices(self, marker):
        """(n x 2) vertex indices of the boundary edges carrying `marker` (local facet i is opposite local vertex i)."""
        rows = self.marked_facets[self.marked_facets[:, 2] == marker]
        c, lf = rows[:, 0], rows[:, 1]
        return np.stack([self.cells[c, (lf + 1) % 3], self.cells[c, (lf + 2) % 3]], axis=1)

    def vertices_on(self, marker):
        return np.unique(self.facet_vertices(marker))

    def stiffness(self, cells_mask):
        """int_{cells} grad(u) . grad(v): nv x nv CSR."""
        idx = np.where(cells_mask)[0]
        K = np.einsum("cid,cjd->cij", self.grad[idx], self.grad[idx]) * self.area[idx, None, None]
        r = np.repeat(self.cells[idx], 3, axis=1).ravel()
        c = np.tile(self.cells[idx], (1, 3)).ravel()
        return sp.csr_matrix((K.ravel(), (r, c)), shape=(self.nv, self.nv))

    def mass(self):
        Mloc = (np.ones((3, 3)) + np.eye(3)) / 12.0
        M = self.area[:, None, None] * Mloc[None]
        r = np.repeat(self.cells, 3, axis=1).ravel()
        c = np.tile(self.cells, (1, 3)).ravel()
        return sp.csr_matrix((M.ravel(), (r, c)), shape=(self.nv, self.nv))

    def boundary_load(self, marker):
        """int_{Gamma_marker} v ds for the hat functions: half the edge length to each end point."""
        f = np.zeros(self.nv)
        ev = self.facet_vertices(marker)
        ln = np.linalg.norm(self.vertices[ev[:, 0]] - self.vertices[ev[:, 1]], axis=1)
        np.add.at(f, ev[:, 0], 0.5 * ln)
        np.add.at(f, ev[:, 1], 0.5 * ln)
        return f


class _AffineTruth:
    """Common part: operators on the free dofs, truth / Riesz solves, sampling (same interface as ThermalBlock)."""

    def _finish(self, ops_a, ops_f, X, free, mu_range):
        self.free = free
        self.operators_a = [A[free][:, free].tocsr() for A in ops_a]
        self.operators_f = [np.asarray(f)[free] for f in ops_f]
        self.inner_product = X[free][:, free].tocsr()
        self._lu = spla.splu(self.inner_product.tocsc())
        self.Nh = int(free.size)
        self.mu_range = mu_range

    def get_stability_factor_lower_bound(self, mus):
        return self.compute_theta("a", mus).min(axis=1)

    def solve(self, mu):
        tha = self.compute_theta("a", [mu])[0]
        thf = self.compute_theta("f", [mu])[0]
        A = sum(t * Aq for t, Aq in zip(tha, self.operators_a))
        f = sum(t * fq for t, fq in zip(thf, self.operators_f))
        return spla.spsolve(A.tocsc(), f)

    def riesz_solve(self, rhs):
        return self._lu.solve(np.asarray(rhs, dtype=np.float64))

    sample_training_set = ThermalBlock.sample_training_set


class ThermalBlockTutorialMesh(_AffineTruth):
    """tutorials/01_thermal_block/tutorial_thermal_block.ipynb:143-203,240-250: P1 on data/thermal_block.xml (304
    vertices), a_0 = int_{Omega_1} grad u . grad v, a_1 = int_{Omega_2} ..., f_0 = int_{Gamma_1} v ds, u = 0 on Gamma_3,
    X = int grad u . grad v; theta_a = (mu_0, 1), theta_f = (mu_1,), mu in [0.1, 10] x [-1, 1]."""
    nblocks, nloads = 2, 1

    def __init__(self):
        m = P1Mesh("thermal_block")
        free = np.setdiff1d(np.arange(m.nv), m.vertices_on(3))
        ops_a = [m.stiffness(m.cell_markers == 1), m.stiffness(m.cell_markers == 2)]
        self._finish(ops_a, [m.boundary_load(1)], m.stiffness(np.ones_like(m.cell_markers, dtype=bool)), free,
                     [(0.1, 10.0), (-1.0, 1.0)])
        self.mesh = m

    def compute_theta(self, term, mus):
        mus = np.atleast_2d(np.asarray(mus, dtype=np.float64))
        if term == "a":
            return np.stack([mus[:, 0], np.ones(mus.shape[0])], axis=1)
        if term == "f":
            return mus[:, 1:2].copy()
        raise ValueError(term)


class ThermalBlockNonHomogeneousDirichlet:
    """The thermal block of tutorial 01 on its own mesh with NON-homogeneous Dirichlet data: u = mu_2 on Gamma_3 (top)
    and u = mu_3 on Gamma_1 (bottom), flux mu_1 on Gamma_2 (sides) -- the situation of the reference's lifting
    machinery (tutorials 03/04: one lifting function per theta_bc term, parametrized_reduced_differential_problem.py:
    283-299).  All vertices are dofs; the operators are assembled WITHOUT boundary conditions (as DOLFIN's affine
    terms are), the inner product and the Riesz solves carry the homogeneous condition (identity rows/columns on the
    Dirichlet vertices), `solve` imposes the parametrised values, and the lifting functions are the discrete harmonic
    extensions of the indicator of each Dirichlet part at theta_a = (1, 1)."""
    nblocks, nloads = 2, 1

    def __init__(self):
        m = P1Mesh("thermal_block")
        self.mesh = m
        self.Nh = m.nv
        self.operators_a = [m.stiffness(m.cell_markers == 1).tocsr(), m.stiffness(m.cell_markers == 2).tocsr()]
        self.operators_f = [m.boundary_load(2)]
        top = m.vertices_on(3)
        bottom = np.setdiff1d(m.vertices_on(1), top)
        self.dirichlet_parts = [top, bottom]
        self.dirichlet = np.concatenate(self.dirichlet_parts)
        self.free = np.setdiff1d(np.arange(m.nv), self.dirichlet)
        K = (self.operators_a[0] + self.operators_a[1]).tolil()
        K[self.dirichlet, :] = 0.0
        K[:, self.dirichlet] = 0.0
        K[self.dirichlet, self.dirichlet] = 1.0
        self.inner_product = K.tocsr()
        self._lu = spla.splu(self.inner_product.tocsc())
        self.mu_range = [(0.1, 10.0), (-1.0, 1.0), (0.5, 2.0), (-0.5, 0.5)]
        ones = (1.0, 1.0)
        self.dirichlet_bc_liftings = np.stack(
            [self._solve_with_values(ones, np.zeros(m.nv), [(part, 1.0)]) for part in self.dirichlet_parts], axis=1)

    def compute_theta(self, term, mus):
        mus = np.atleast_2d(np.asarray(mus, dtype=np.float64))
        if term == "a":
            return np.stack([mus[:, 0], np.ones(mus.shape[0])], axis=1)
        if term == "f":
            return mus[:, 1:2].copy()
        if term == "dirichlet_bc":
            return mus[:, 2:4].copy()
        raise ValueError(term)

    def get_stability_factor_lower_bound(self, mus):
        return self.compute_theta("a", mus).min(axis=1)

    def _solve_with_values(self, theta_a, f, values):
        A = sum(t * Aq for t, Aq in zip(theta_a, self.operators_a)).tocsr()
        u = np.zeros(self.Nh)
        for part, v in values:
            u[part] = v
        rhs = (f - A @ u)[self.free]
        u[self.free] = spla.spsolve(A[self.free][:, self.free].tocsc(), rhs)
        return u

    def solve(self, mu):
        tha = self.compute_theta("a", [mu])[0]
        thf = self.compute_theta("f", [mu])[0]
        thbc = self.compute_theta("dirichlet_bc", [mu])[0]
        f = sum(t * fq for t, fq in zip(thf, self.operators_f))
        return self._solve_with_values(tha, f, list(zip(self.dirichlet_parts, thbc)))

    def riesz_solve(self, rhs):
        r = np.array(rhs, dtype=np.float64)
        r[self.dirichlet] = 0.0          # the residual is a functional on the functions that vanish on the Dirichlet part
        return self._lu.solve(r)

    sample_training_set = ThermalBlock.sample_training_set


class ElasticBlockTutorialMesh(_AffineTruth):
    """tutorials/02_elastic_block/tutorial_elastic_block.ipynb:170-262,301-313: vector P1 on data/elastic_block.xml
    (1346 vertices, 2692 dofs), a_q = int_{Omega_q} 2 lambda_2 sym grad u : sym grad v + lambda_1 div u div v
    (E = 1, nu = 0.3) for the 9 blocks, f_q = int_{Gamma_{2,3,4}} (1, 0) . v ds, u = 0 on Gamma_6,
    X = int u . v + grad u : grad v; theta_a = (mu_0..mu_7, 1), theta_f = (mu_8, mu_9, mu_10)."""
    nblocks, nloads = 9, 3

    def __init__(self):
        m = P1Mesh("elastic_block")
        E, nu = 1.0, 0.3
        l1 = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))
        l2 = E / (2.0 * (1.0 + nu))
        nv = m.nv
        # dof (vertex v, component c) -> 2 v + c; strain of phi_i e_c: B[:, i, c] = (eps_xx, eps_yy, 2 eps_xy)
        gx, gy = m.grad[:, :, 0], m.grad[:, :, 1]
        Bm = np.zeros((m.cells.shape[0], 3, 6))
        Bm[:, 0, 0::2] = gx
        Bm[:, 1, 1::2] = gy
        Bm[:, 2, 0::2] = gy
        Bm[:, 2, 1::2] = gx
        D = np.array([[2 * l2 + l1, l1, 0.0], [l1, 2 * l2 + l1, 0.0], [0.0, 0.0, l2]])
        Kloc = np.einsum("cai,ab,cbj->cij", Bm, D, Bm) * m.area[:, None, None]
        dofs = np.stack([2 * m.cells, 2 * m.cells + 1], axis=2).reshape(-1, 6)    # (v0x, v0y, v1x, v1y, v2x, v2y)

        def block(mask):
            idx = np.where(mask)[0]
            r = np.repeat(dofs[idx], 6, axis=1).ravel()
            c = np.tile(dofs[idx], (1, 6)).ravel()
            return sp.csr_matrix((Kloc[idx].ravel(), (r, c)), shape=(2 * nv, 2 * nv))

        ops_a = [block(m.cell_markers == q) for q in range(1, 10)]
        ops_f = []
        for marker in (2, 3, 4):
            f = np.zeros(2 * nv)
            f[0::2] = m.boundary_load(marker)        # traction (1, 0)
            ops_f.append(f)
        K1, M1 = m.stiffness(np.ones_like(m.cell_markers, dtype=bool)), m.mass()
        X = sp.kron(K1 + M1, sp.identity(2)).tocsr()
        clamped = m.vertices_on(6)
        free = np.setdiff1d(np.arange(2 * nv), np.concatenate([2 * clamped, 2 * clamped + 1]))
        self._finish(ops_a, ops_f, X, free, [(1.0, 100.0)] * 8 + [(-1.0, 1.0)] * 3)
        self.mesh = m

    def compute_theta(self, term, mus):
        mus = np.atleast_2d(np.asarray(mus, dtype=np.float64))
        if term == "a":
            return np.concatenate([mus[:, :8], np.ones((mus.shape[0], 1))], axis=1)
        if term == "f":
            return mus[:, 8:11].copy()
        raise ValueError(term)

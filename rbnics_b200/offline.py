"""Offline basis linear algebra on the B200: POD correlation (Gram) matrix, Galerkin projections, POD, Gram-Schmidt.

Mirrors (same operation names / argument meaning):
  transpose(S) * X * S            rbnics/backends/basic/transpose.py:304-313      -> :func:`gram`
  transpose(Z) * A_q * Z, Z^T f   rbnics/backends/basic/transpose.py:441-468,365-400 -> :func:`project_matrices`, :func:`project_vectors`
  S * eigenvector                 rbnics/backends/dolfin/wrapping/functions_list_mul.py:10-36 -> :func:`lincomb`
  ProperOrthogonalDecomposition   rbnics/backends/basic/proper_orthogonal_decomposition_base.py:39-92 -> :class:`ProperOrthogonalDecomposition`
  GramSchmidt                     rbnics/backends/basic/gram_schmidt.py:22-36     -> :class:`GramSchmidt`

Snapshots / basis functions are dof-major dense arrays (Nh x Ns, one column per function); inner-product and
operator matrices are scipy.sparse CSR.  Multi-GPU: the dof rows are partitioned contiguously over the ranks, every
rank computes a full-size partial and one SUM all-reduce (NCCL) replaces the reference's Ns^2 scalar VecDot
all-reduces.  The eigen-decomposition of the small Ns x Ns matrix stays on the host
(scipy.linalg.eigh, exactly the reference's call, rbnics/backends/online/numpy/eigen_solver.py:36-56): replicas only.
"""
from __future__ import annotations

import ctypes as C
from math import sqrt

import numpy as np
import scipy.linalg

from . import _lib as L


def _csr_arrays(X):
    X = X.tocsr()
    X.sort_indices()
    return (np.ascontiguousarray(X.indptr, dtype=np.int64), np.ascontiguousarray(X.indices, dtype=np.int32),
            np.ascontiguousarray(X.data, dtype=np.float64))


def row_range(Nh, rank, size):
    """Contiguous dof-row partition of rank ``rank`` out of ``size`` (PETSc-style ownership ranges)."""
    per = (Nh + size - 1) // size
    return min(Nh, rank * per), min(Nh, (rank + 1) * per)


def _dist_range(Nh, dist):
    if dist is None or dist.size == 1:
        return 0, Nh
    return row_range(Nh, dist.rank, dist.size)


def install_device_allreduce(engine, dist):
    """Multi-GPU offline path: hand the library an all-reduce that works on its device buffers (rb_set_allreduce), so
    the partial Gram / projection / Gram-Schmidt results are summed over the ranks in HBM (NCCL) before they are copied
    to the host once.  Returns True if installed (NCCL, more than one rank); with gloo (CPU tests) the host-side
    all-reduce of the returned arrays is used instead."""
    if dist is None or dist.size == 1 or dist.backend != "nccl":
        if getattr(engine, "_allreduce_on", False):   # a single-rank call on an engine that also serves partitioned ones
            engine._check(engine._lib.rb_set_allreduce(engine._h, None, None))
            engine._allreduce_on = False
        return False
    if getattr(engine, "_allreduce_cb", None) is None:
        engine._allreduce_cb = dist.device_allreduce_callback()   # keep the ctypes thunk alive
    if not getattr(engine, "_allreduce_on", False):
        engine._check(engine._lib.rb_set_allreduce(engine._h, C.cast(engine._allreduce_cb, C.c_void_p), None))
        engine._allreduce_on = True
    return True


def gram(engine, S, X=None, S2=None, dist=None):
    """C[i, j] = s_i . (X s2_j)  (Ns x Ns2)."""
    S = np.ascontiguousarray(S, dtype=np.float64)
    Nh, Ns = S.shape
    S2c = None
    Ns2 = Ns
    if S2 is not None:
        S2c = np.ascontiguousarray(S2, dtype=np.float64)
        assert S2c.shape[0] == Nh
        Ns2 = S2c.shape[1]
    Cm = np.empty((Ns, Ns2))
    rb, re = _dist_range(Nh, dist)
    if X is not None:
        assert X.shape == (Nh, Nh)
        ip, ix, dt = _csr_arrays(X)
        args = (ip.ctypes.data_as(L.c_int64_p), ix.ctypes.data_as(L.c_int32_p), L.dptr(dt))
    else:
        args = (None, None, None)
    on_device = install_device_allreduce(engine, dist)
    engine._check(engine._lib.rb_gram(engine._h, Nh, Ns, Ns2, L.dptr(S), L.dptr(S2c), *args, rb, re, L.dptr(Cm)))
    if dist is not None and dist.size > 1 and not on_device:
        dist.allreduce_sum_(Cm)
    return Cm


def project_matrices(engine, Z, operators, dist=None):
    """[Z^T A_q Z for q]  -> (Q, N, N).  Single GPU: one batched call; multi GPU: per-operator partial Grams."""
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    Nh, N = Z.shape
    Q = len(operators)
    if dist is not None and dist.size > 1:
        return np.stack([gram(engine, Z, A, dist=dist) for A in operators])
    ips, ixs, dts, off = [], [], [], [0]
    for A in operators:
        ip, ix, dt = _csr_arrays(A)
        ips.append(ip)
        ixs.append(ix)
        dts.append(dt)
        off.append(off[-1] + len(dt))
    ip = np.concatenate(ips)
    ix = np.concatenate(ixs) if off[-1] else np.zeros(1, dtype=np.int32)
    dt = np.concatenate(dts) if off[-1] else np.zeros(1)
    off = np.asarray(off, dtype=np.int64)
    out = np.empty((Q, N, N))
    engine._check(engine._lib.rb_project(engine._h, Nh, N, Q, L.dptr(Z), ip.ctypes.data_as(L.c_int64_p),
                                         ix.ctypes.data_as(L.c_int32_p), L.dptr(dt), off.ctypes.data_as(L.c_int64_p),
                                         L.dptr(out)))
    return out


def project_vectors(engine, Z, vectors, dist=None):
    """[Z^T f_q for q] -> (Q, N): one Gram with S2 = [f_1 ... f_Q]."""
    Fm = np.ascontiguousarray(np.stack([np.asarray(f, dtype=np.float64) for f in vectors], axis=1))
    return gram(engine, Z, None, Fm, dist=dist).T.copy()


def lincomb(engine, S, V):
    """Z = S V  (functions_list_mul_online_matrix / _vector)."""
    S = np.ascontiguousarray(S, dtype=np.float64)
    V = np.ascontiguousarray(V, dtype=np.float64)
    Nh, Ns = S.shape
    assert V.shape[0] == Ns
    n = V.shape[1]
    Z = np.empty((Nh, n))
    engine._check(engine._lib.rb_lincomb(engine._h, Nh, Ns, n, L.dptr(S), L.dptr(V), L.dptr(Z)))
    return Z


class ProperOrthogonalDecomposition:
    """POD by the method of snapshots; same interface as the reference's backend class
    (rbnics/backends/abstract/proper_orthogonal_decomposition.py:11-44)."""

    def __init__(self, engine, inner_product=None, dist=None):
        self.engine = engine
        self.inner_product = inner_product
        self.dist = dist
        self.clear()

    def clear(self):
        self._snapshots = []
        self.eigenvalues = []
        self.retained_energy = []

    def store_snapshot(self, snapshot, weight=None):
        s = np.asarray(snapshot, dtype=np.float64)
        self._snapshots.append(s if weight is None else s * weight)

    def snapshots_matrix(self):
        return np.ascontiguousarray(np.stack(self._snapshots, axis=1))

    def apply(self, Nmax, tol):
        """-> (eigenvalues[:N], eigenvectors, basis (Nh x N), N)."""
        S = self.snapshots_matrix()
        X = self.inner_product
        correlation = gram(self.engine, S, X, dist=self.dist)
        eigs, eigv = scipy.linalg.eigh(correlation)          # online/numpy/eigen_solver.py:39
        idx = eigs.argsort()[::-1]                           # "largest real" (:44-46)
        eigs, eigv = eigs[idx], eigv[:, idx]
        Neigs = S.shape[1]
        Nmax = min(Nmax, Neigs)
        self.eigenvalues = [float(e) for e in eigs]
        abs_e = np.abs(eigs)
        total = abs_e.sum()
        self.retained_energy = list(np.cumsum(abs_e) / total) if total > 0.0 else [1.0] * Neigs
        N = 0
        for n in range(Nmax):
            N = n + 1
            if tol > 0.0 and self.retained_energy[n] > 1.0 - tol:
                break
        V = np.ascontiguousarray(eigv[:, :N])
        Z = lincomb(self.engine, S, V)
        # norm_b = sqrt(b^T X b) from the modes themselves (proper_orthogonal_decomposition_base.py:80-87)
        norms2 = np.diag(gram(self.engine, Z, X, dist=self.dist))
        scale = np.array([1.0 / sqrt(v) if v > 0.0 else 1.0 for v in norms2])
        Z = lincomb(self.engine, S, V * scale[None, :])
        return self.eigenvalues[:N], [eigv[:, n].copy() for n in range(N)], Z, N


class GramSchmidt:
    """Single-pass modified Gram-Schmidt in the X inner product (rbnics/backends/basic/gram_schmidt.py:22-36)."""

    def __init__(self, engine, inner_product):
        self.engine = engine
        self.inner_product = inner_product
        self._csr = _csr_arrays(inner_product) if inner_product is not None else None

    def apply(self, new_basis_function, basis_functions):
        z = np.ascontiguousarray(new_basis_function, dtype=np.float64).copy()
        Nh = z.shape[0]
        if basis_functions is None:
            Zb = np.zeros((Nh, 0))
        else:
            Zb = np.ascontiguousarray(basis_functions, dtype=np.float64)
        n = Zb.shape[1]
        if self._csr is not None:
            ip, ix, dt = self._csr
            args = (ip.ctypes.data_as(L.c_int64_p), ix.ctypes.data_as(L.c_int32_p), L.dptr(dt))
        else:
            args = (None, None, None)
        norm = C.c_double()
        self.engine._check(self.engine._lib.rb_gram_schmidt(self.engine._h, Nh, n, L.dptr(Zb) if n else None, *args,
                                                            L.dptr(z), C.byref(norm)))
        return z


class OfflineWorkspace:
    """Device-resident truth-space data of one greedy run (C ABI ``rb_off_*``): named sparse operators and named dense
    column blocks (basis functions, snapshots, Riesz representers) stay in HBM; every product below runs on them
    without re-uploading anything but new columns.  The reference rebuilds all reduced operators from host/PETSc
    vectors at every iteration (rbnics/reduction_methods/base/rb_reduction.py:106,112)."""

    SLOTS = 128

    def __init__(self, engine, dist=None, row_local=True):
        self.engine = engine
        self.dist = dist
        self._csr = {}
        self._blk = {}
        self.Nh = None
        self._device_allreduce = install_device_allreduce(engine, dist)
        # dof-row partition: a rank uploads only the rows its products read -- its own rows and the rows the operator
        # rows it owns reference (the ghost entries PETSc keeps); HBM keeps the full-size arrays, PCIe carries 1/size
        self.row_local = bool(row_local) and dist is not None and dist.size > 1
        self._halo = {}     # operator name -> (lo, hi): rows referenced by the rank's operator rows
        self._valid = {}    # block name -> (lo, hi): rows of the block that hold data on this rank

    def _operator_halo(self, X):
        rb, re = _dist_range(X.shape[0], self.dist)
        if re <= rb:
            return rb, rb
        X = X.tocsr()
        idx = X.indices[X.indptr[rb]:X.indptr[re]]
        if idx.size == 0:
            return rb, re
        return min(rb, int(idx.min())), max(re, int(idx.max()) + 1)

    def upload_range(self):
        """Rows a new block column must hold on this rank: own rows + halo of every operator registered so far."""
        lo, hi = _dist_range(self.Nh, self.dist)
        for (a, b) in self._halo.values():
            lo, hi = min(lo, a), max(hi, b)
        return lo, hi

    def _narrow(self, block, rows):
        a, b = self._valid.get(block, (0, self.Nh))
        self._valid[block] = (max(a, rows[0]), min(b, rows[1]))

    def _require(self, block, operator=None):
        """A product over the rank's rows needs them (and, under an operator, its halo) to hold data."""
        if not self.row_local:
            return
        need = self._halo[operator] if operator is not None else _dist_range(self.Nh, self.dist)
        have = self._valid.get(block, (0, self.Nh))
        if need[0] < have[0] or need[1] > have[1]:
            raise RuntimeError(f"OfflineWorkspace: block {block!r} holds rows {have} on this rank but {need} are "
                               f"needed (register the operators before appending to the block)")

    def _slot(self, table, name):
        if name not in table:
            if len(table) >= self.SLOTS:
                raise ValueError("OfflineWorkspace: out of slots")
            table[name] = len(table)
        return table[name]

    def set_operator(self, name, X):
        ip, ix, dt = _csr_arrays(X)
        e = self.engine
        e._check(e._lib.rb_off_csr(e._h, self._slot(self._csr, name), X.shape[0], ip.ctypes.data_as(L.c_int64_p),
                                   ix.ctypes.data_as(L.c_int32_p), L.dptr(dt)))
        self.Nh = X.shape[0]
        if self.row_local:
            self._halo[name] = self._operator_halo(X)

    def new_block(self, name, Nh, capacity):
        e = self.engine
        e._check(e._lib.rb_off_block(e._h, self._slot(self._blk, name), int(Nh), int(capacity)))
        self.Nh = int(Nh)

    def ncols(self, name):
        e = self.engine
        nh, n = C.c_int64(), C.c_int()
        e._check(e._lib.rb_off_size(e._h, self._blk[name], C.byref(nh), C.byref(n)))
        return n.value

    def append(self, name, cols):
        cols = np.asarray(cols, dtype=np.float64)
        if cols.ndim == 1:
            cols = cols[:, None]
        cols = np.ascontiguousarray(cols)
        e = self.engine
        if self.row_local:
            lo, hi = self.upload_range()
            e._check(e._lib.rb_off_append_rows(e._h, self._blk[name], cols.shape[1], L.dptr(cols), lo, hi))
            self._narrow(name, (lo, hi))
        else:
            e._check(e._lib.rb_off_append(e._h, self._blk[name], cols.shape[1], L.dptr(cols)))

    def read(self, name, c0=0, n=None):
        """Columns of a block on the host, complete on every rank (row-local blocks: the ranks' own rows are summed)."""
        n = self.ncols(name) - c0 if n is None else n
        out = np.empty((self.Nh, n))
        e = self.engine
        e._check(e._lib.rb_off_read(e._h, self._blk[name], int(c0), int(n), L.dptr(out)))
        if self.row_local and self._valid.get(name, (0, self.Nh)) != (0, self.Nh):
            self._require(name)
            rb, re = _dist_range(self.Nh, self.dist)
            out[:rb] = 0.0
            out[re:] = 0.0
            self.dist.allreduce_sum_(out)
        return out

    def spmm(self, A, S, cols, Y, col0):
        """Y[:, col0:col0+n] = A * S[:, cols[0]:cols[1]] on resident operands (operator and block names)."""
        e = self.engine
        self._require(S, A)
        e._check(e._lib.rb_off_spmm(e._h, self._csr[A], self._blk[S], int(cols[0]), int(cols[1] - cols[0]),
                                    self._blk[Y], int(col0)))
        if self.row_local and self._valid.get(S, (0, self.Nh)) != (0, self.Nh):
            self._narrow(Y, _dist_range(self.Nh, self.dist))   # Y = A S is right where S held the rows A referenced

    def gram(self, S, X=None, S2=None, cols=None, cols2=None):
        """C = S[:, cols]^T X S2[:, cols2]; S, S2 block names, X operator name or None (identity); cols = (begin, end)."""
        S2 = S if S2 is None else S2
        s0, s1 = cols if cols is not None else (0, self.ncols(S))
        t0, t1 = cols2 if cols2 is not None else (0, self.ncols(S2))
        Cm = np.empty((s1 - s0, t1 - t0))
        rb, re = _dist_range(self.Nh, self.dist)
        e = self.engine
        self._require(S)
        self._require(S2, X)
        self._device_allreduce = install_device_allreduce(e, self.dist)
        e._check(e._lib.rb_off_gram(e._h, self._blk[S], int(s0), int(s1 - s0), self._csr[X] if X is not None else -1,
                                    self._blk[S2], int(t0), int(t1 - t0), rb, re, L.dptr(Cm)))
        if self.dist is not None and self.dist.size > 1 and not self._device_allreduce:
            self.dist.allreduce_sum_(Cm)
        return Cm

    def lincomb(self, S, V, Z, col0=0, cols=None, accumulate=False):
        """Z[:, col0:col0+n] = S[:, cols] V on resident blocks (V: host, ns x n): S * eigenvectors of the POD;
        ``accumulate=True`` adds to existing columns instead (snapshot -= basis_functions * projection)."""
        V = np.ascontiguousarray(V, dtype=np.float64)
        if V.ndim == 1:
            V = np.ascontiguousarray(V[:, None])
        s0, s1 = cols if cols is not None else (0, self.ncols(S))
        assert V.shape[0] == s1 - s0
        e = self.engine
        e._check(e._lib.rb_off_lincomb_add(e._h, self._blk[S], int(s0), int(s1 - s0), L.dptr(V), int(V.shape[1]),
                                           self._blk[Z], int(col0), 1 if accumulate else 0))
        if self.row_local:
            self._narrow(Z, self._valid.get(S, (0, self.Nh)))

    def column_norms2(self, Z, X=None, cols=None):
        """[z_j^T X z_j for the columns j of the block]: the diagonal of Z^T X Z without the off-diagonal work."""
        c0, c1 = cols if cols is not None else (0, self.ncols(Z))
        out = np.empty(c1 - c0)
        rb, re = _dist_range(self.Nh, self.dist)
        e = self.engine
        self._require(Z, X)
        self._device_allreduce = install_device_allreduce(e, self.dist)
        for a in range(c0, c1, 128):
            n = min(128, c1 - a)
            part = np.empty(n)
            e._check(e._lib.rb_off_column_norms2(e._h, self._blk[Z], int(a), int(n), self._csr[X] if X is not None else -1,
                                                 rb, re, L.dptr(part)))
            out[a - c0:a - c0 + n] = part
        if self.dist is not None and self.dist.size > 1 and not self._device_allreduce:
            self.dist.allreduce_sum_(out)
        return out

    def scale_columns(self, Z, scale, col0=0):
        scale = np.ascontiguousarray(scale, dtype=np.float64)
        e = self.engine
        e._check(e._lib.rb_off_scale_columns(e._h, self._blk[Z], int(col0), int(scale.shape[0]), L.dptr(scale)))

    def gram_schmidt(self, Z, X, new_function, append=True, first_col=0):
        """Orthonormalise ``new_function`` against the columns ``first_col:`` of block Z (X inner product); optionally
        append it.  ``first_col = N_bc`` leaves the lifting functions out (rb_reduction.py:125-139).
        With more than one rank the dof rows are partitioned like the Gram products: every projection coefficient is
        one all-reduce (rbnics/backends/basic/gram_schmidt.py:29-33 under PETSc's row partition: one VecDot each), the
        order of the reference's modified Gram-Schmidt is kept."""
        z = np.ascontiguousarray(new_function, dtype=np.float64).copy()
        norm = C.c_double()
        e = self.engine
        self._device_allreduce = install_device_allreduce(e, self.dist)
        rb, re = _dist_range(self.Nh, self.dist) if self._device_allreduce else (0, self.Nh)
        if self._device_allreduce:
            self._require(Z, X)
        elif self.row_local and self._valid.get(Z, (0, self.Nh)) != (0, self.Nh):
            raise RuntimeError("OfflineWorkspace.gram_schmidt: a row-local block needs the device all-reduce (NCCL)")
        e._check(e._lib.rb_off_gram_schmidt_from(e._h, self._blk[Z], self._csr[X] if X is not None else -1, L.dptr(z),
                                                 1 if append else 0, int(first_col), rb, re, C.byref(norm)))
        return z, norm.value


class ResidentProperOrthogonalDecomposition:
    """POD by the method of snapshots on DEVICE-RESIDENT snapshots (rbnics/backends/basic/
    proper_orthogonal_decomposition_base.py:39-92; interface of rbnics/backends/abstract/
    proper_orthogonal_decomposition.py:11-44).  Snapshots are appended to a block of the workspace as they are produced
    (each crosses PCIe once); ``apply`` is ONE Gram matrix S^T X S, the host eigen-decomposition of the small matrix
    (scipy.linalg.eigh, the reference's own call, rbnics/backends/online/numpy/eigen_solver.py:36-56), ONE tall-skinny
    product for the modes, their X-norms from the diagonal only and a column scaling -- nothing of size Nh x Ns goes back
    to the host unless the caller reads the basis."""

    def __init__(self, workspace, inner_product, Nh, capacity, snapshots="pod_snapshots", basis="pod_basis"):
        self.ws = workspace
        self.X = inner_product          # operator name in the workspace, or None for the Euclidean product
        self.snapshots, self.basis = snapshots, basis
        self.Nh, self.capacity = int(Nh), int(capacity)
        self.clear()

    def clear(self):
        self.ws.new_block(self.snapshots, self.Nh, self.capacity)
        self.eigenvalues = []
        self.retained_energy = []

    def store_snapshot(self, snapshot, weight=None):
        s = np.asarray(snapshot, dtype=np.float64)
        self.ws.append(self.snapshots, s if weight is None else s * weight)

    def apply(self, Nmax, tol, read_basis=True):
        """-> (eigenvalues[:N], eigenvectors, basis (Nh x N array, or the block name if not read_basis), N)."""
        ws = self.ws
        Ns = ws.ncols(self.snapshots)
        correlation = ws.gram(self.snapshots, self.X)
        eigs, eigv = scipy.linalg.eigh(correlation)
        idx = eigs.argsort()[::-1]
        eigs, eigv = eigs[idx], eigv[:, idx]
        Nmax = min(Nmax, Ns)
        self.eigenvalues = [float(e) for e in eigs]
        abs_e = np.abs(eigs)
        total = abs_e.sum()
        self.retained_energy = list(np.cumsum(abs_e) / total) if total > 0.0 else [1.0] * Ns
        N = 0
        for n in range(Nmax):
            N = n + 1
            if tol > 0.0 and self.retained_energy[n] > 1.0 - tol:
                break
        V = np.ascontiguousarray(eigv[:, :N])
        ws.new_block(self.basis, self.Nh, max(N, 2))
        ws.lincomb(self.snapshots, V, self.basis)
        norms2 = ws.column_norms2(self.basis, self.X)          # b^T X b of every mode (:80-87), diagonal only
        ws.scale_columns(self.basis, np.array([1.0 / sqrt(v) if v > 0.0 else 1.0 for v in norms2]))
        Z = ws.read(self.basis) if read_basis else self.basis
        return self.eigenvalues[:N], [eigv[:, n].copy() for n in range(N)], Z, N

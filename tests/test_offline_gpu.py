"""GPU parity of the offline basis linear algebra (Gram, projections, POD, Gram-Schmidt) against the CPU oracle.

Tolerances: Gram / projection entries within 1e-10 relative to max|C| (parity defined on the same CSR
matrix + dense vectors); POD eigenvalues within 1e-10 * lambda_max (north_star)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import rb_oracle as O
from rbnics_b200 import offline
from truth_problems import ThermalBlock

pytestmark = pytest.mark.gpu


def _lap2d(n):
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n))
    I = sp.identity(n)
    return (sp.kron(T, I) + sp.kron(I, T) + 0.5 * sp.identity(n * n)).tocsr()


@pytest.mark.parametrize("n,Ns", [(9, 1), (20, 7), (31, 61), (40, 64), (17, 65), (22, 97), (20, 100), (25, 128),
                                  (33, 130), (50, 200)])
def test_gram_matches_loop_oracle(engine, n, Ns):
    rng = np.random.default_rng(n + Ns)
    X = _lap2d(n)
    S = rng.standard_normal((n * n, Ns))
    C = offline.gram(engine, S, X)
    ref = O.gram_loop(S, X) if Ns <= 64 else O.gram_blas(S, X)
    assert np.max(np.abs(C - ref)) <= 1e-10 * np.max(np.abs(ref))


@pytest.mark.parametrize("Ns", [129, 200, 257, 400])
@pytest.mark.parametrize("symmetric", [True, False])
def test_gram_multi_tile(engine, Ns, symmetric):
    """Several 128-column tiles: a symmetric X lets the kernel mirror the tiles below the diagonal; a general operator
    (S^T A S of a Galerkin projection) and the identity inner product take the full tile grid."""
    rng = np.random.default_rng(Ns)
    n = 24
    X = _lap2d(n)
    if not symmetric:
        X = (X + sp.diags([0.3], [1], shape=X.shape)).tocsr()
    S = rng.standard_normal((n * n, Ns))
    C = offline.gram(engine, S, X)
    ref = O.gram_blas(S, X)
    assert np.max(np.abs(C - ref)) <= 1e-10 * np.max(np.abs(ref))
    if symmetric:
        assert np.max(np.abs(C - C.T)) <= 1e-12 * np.max(np.abs(ref))
    C0 = offline.gram(engine, S, None)
    ref0 = S.T @ S
    assert np.max(np.abs(C0 - ref0)) <= 1e-10 * np.max(np.abs(ref0))


@pytest.mark.parametrize("Ns,Ns2", [(70, 3), (3, 70), (100, 100), (65, 128), (128, 9), (104, 105)])
def test_gram_single_wide_tile_shapes(engine, Ns, Ns2):
    """65..128 columns on either side: one 128-wide tile whose 8-column sub-tiles are dealt evenly to the 16 warps
    (also when one side has fewer sub-tiles than warp rows)."""
    rng = np.random.default_rng(Ns * 131 + Ns2)
    n = 19
    X = _lap2d(n)
    S = rng.standard_normal((n * n, Ns))
    S2 = rng.standard_normal((n * n, Ns2))
    for Xop in (X, None):
        C = offline.gram(engine, S, Xop, S2)
        ref = S.T @ ((X @ S2) if Xop is not None else S2)
        assert np.max(np.abs(C - ref)) <= 1e-10 * np.max(np.abs(ref))


def test_gram_rectangular_multi_tile(engine):
    rng = np.random.default_rng(8)
    n = 20
    X = _lap2d(n)
    S = rng.standard_normal((n * n, 150))
    S2 = rng.standard_normal((n * n, 301))
    C = offline.gram(engine, S, X, S2)
    ref = S.T @ (X @ S2)
    assert np.max(np.abs(C - ref)) <= 1e-10 * np.max(np.abs(ref))


def test_workspace_gram_odd_column_offsets(engine):
    """Column ranges that start at odd columns of a resident block are not 16-byte aligned: they are copied once into
    aligned scratch (rb_copy_cols_kernel) before the bulk-copy kernels run; results equal the aligned path."""
    rng = np.random.default_rng(9)
    n = 18
    X = _lap2d(n)
    S = np.ascontiguousarray(rng.standard_normal((n * n, 37)))
    ws = offline.OfflineWorkspace(engine)
    ws.set_operator("X", X)
    ws.new_block("S", n * n, 37)
    ws.append("S", S)
    # (same first column with different widths: the right operand may need columns the staged left one lacks)
    for (s0, ns, t0, ns2) in [(0, 37, 0, 37), (1, 20, 3, 11), (2, 35, 1, 36), (5, 1, 0, 37), (0, 5, 0, 30), (0, 30, 0, 5),
                              (4, 2, 4, 33), (4, 1, 4, 1)]:
        C = ws.gram("S", "X", cols=(s0, s0 + ns), cols2=(t0, t0 + ns2))
        ref = S[:, s0:s0 + ns].T @ (X @ S[:, t0:t0 + ns2])
        assert np.max(np.abs(C - ref)) <= 1e-10 * np.max(np.abs(ref))


@pytest.mark.parametrize("Nh,Ns,density", [(1000, 17, 0.012), (2053, 64, 0.02), (4099, 33, 0.004), (50, 5, 0.5)])
def test_fused_gram_dense_rows_and_tails(engine, Nh, Ns, density):
    """Fused kernel (<= 64 columns): rows with far more entries than one per lane (the register-resident window of 32
    entries per four rows overflows into the slow path), dof counts that are not multiples of the 16-row chunk,
    unsymmetric operators, columns both inside and outside a chunk."""
    rng = np.random.default_rng(Nh + Ns)
    A = sp.random(Nh, Nh, density=density, format="csr", random_state=Nh, data_rvs=rng.standard_normal)
    A = (A + sp.diags(rng.standard_normal(Nh)) + sp.diags(rng.standard_normal(Nh - 1), 1)).tocsr()
    A.sort_indices()
    S = rng.standard_normal((Nh, Ns))
    C = offline.gram(engine, S, A)
    ref = S.T @ (A @ S)
    assert np.max(np.abs(C - ref)) <= 1e-10 * np.max(np.abs(ref))
    # different right operand (no reuse of the staged S rows), rectangular
    S2 = rng.standard_normal((Nh, max(1, Ns // 2)))
    C2 = offline.gram(engine, S, A, S2)
    ref2 = S.T @ (A @ S2)
    assert np.max(np.abs(C2 - ref2)) <= 1e-10 * np.max(np.abs(ref2))


def test_fused_gram_row_partition_and_empty_rows(engine):
    """Row partitions of the fused path (any split point, also inside a chunk) add up; empty rows are fine."""
    import ctypes as C
    from rbnics_b200 import _lib as L
    rng = np.random.default_rng(12)
    Nh, Ns = 777, 40
    A = sp.random(Nh, Nh, density=0.01, format="lil", random_state=3)
    A[100:140, :] = 0.0  # a band of empty rows
    A = A.tocsr()
    A.eliminate_zeros()
    S = np.ascontiguousarray(rng.standard_normal((Nh, Ns)))
    ip, ix, dt = offline._csr_arrays(A)
    total = np.zeros((Ns, Ns))
    for (rb, re) in [(0, 1), (1, 250), (250, 250), (250, 601), (601, Nh)]:
        part = np.empty((Ns, Ns))
        engine._check(engine._lib.rb_gram(engine._h, Nh, Ns, Ns, L.dptr(S), None, ip.ctypes.data_as(L.c_int64_p),
                                          ix.ctypes.data_as(L.c_int32_p), L.dptr(dt), rb, re, L.dptr(part)))
        total += part
    ref = S.T @ (A @ S)
    assert np.max(np.abs(total - ref)) <= 1e-10 * np.max(np.abs(ref))


def test_workspace_spmm_then_single_gram(engine):
    """rb_off_spmm: operators applied to one new column each, written side by side; one identity Gram then yields
    Z^T A_q z for all q (and with the transposed operators z^T A_q Z)."""
    rng = np.random.default_rng(31)
    Nh, N, Q = 900, 12, 5
    ops = [(sp.random(Nh, Nh, density=0.01, format="csr", random_state=q) + sp.identity(Nh)).tocsr() for q in range(Q)]
    Z = np.ascontiguousarray(rng.standard_normal((Nh, N)))
    ws = offline.OfflineWorkspace(engine)
    ws.new_block("Z", Nh, N)
    ws.append("Z", Z)
    ws.new_block("Y", Nh, 2 * Q)
    for q, A in enumerate(ops):
        ws.set_operator(f"A{q}", A)
        ws.set_operator(f"A{q}T", A.T.tocsr())
    for q in range(Q):   # out of order on purpose: columns never written stay zero
        ws.spmm(f"A{q}T", "Z", (N - 1, N), "Y", Q + q)
        ws.spmm(f"A{q}", "Z", (N - 1, N), "Y", q)
    C = ws.gram("Z", None, "Y")
    assert C.shape == (N, 2 * Q)
    for q, A in enumerate(ops):
        col = Z.T @ (A @ Z[:, -1])
        row = (Z[:, -1] @ A) @ Z
        assert np.max(np.abs(C[:, q] - col)) <= 1e-10 * np.max(np.abs(col))
        assert np.max(np.abs(C[:, Q + q] - row)) <= 1e-10 * np.max(np.abs(row))
    # wide product through the warp-per-row kernel (n > 8)
    ws.new_block("W", Nh, N)
    ws.spmm("A0", "Z", (0, N), "W", 0)
    assert np.max(np.abs(ws.read("W") - ops[0] @ Z)) <= 1e-10 * np.max(np.abs(ops[0] @ Z))
    with pytest.raises(ValueError):
        ws.spmm("A0", "Z", (0, N), "Y", 0)   # 12 columns do not fit a block of capacity 10


def test_gram_identity_inner_product_and_rectangular(engine):
    rng = np.random.default_rng(5)
    S = rng.standard_normal((1000, 13))
    S2 = rng.standard_normal((1000, 70))
    C = offline.gram(engine, S, None, S2)
    ref = S.T @ S2
    assert C.shape == (13, 70)
    assert np.max(np.abs(C - ref)) <= 1e-10 * np.max(np.abs(ref))


def test_gram_reference_distribution(engine):
    """Heavy-tailed signed inputs of the reference's own performance tests (test_numpy_utils.py:58-59)."""
    rng = np.random.default_rng(0)
    X = _lap2d(16)
    S = O.reference_rand(rng, 256, 10)
    C = offline.gram(engine, S, X)
    ref = O.gram_loop(S, X)
    assert np.max(np.abs(C - ref)) <= 1e-10 * np.max(np.abs(ref))


def test_gram_row_partition_sums_to_full(engine):
    """dof-row partition: the per-rank partials add up to the full Gram matrix."""
    import ctypes as C
    from rbnics_b200 import _lib as L
    rng = np.random.default_rng(2)
    n = 30
    X = _lap2d(n)
    S = np.ascontiguousarray(rng.standard_normal((n * n, 20)))
    ip, ix, dt = offline._csr_arrays(X)
    total = np.zeros((20, 20))
    for rank in range(3):
        rb, re = offline.row_range(n * n, rank, 3)
        part = np.empty((20, 20))
        engine._check(engine._lib.rb_gram(engine._h, n * n, 20, 20, L.dptr(S), None, ip.ctypes.data_as(L.c_int64_p),
                                          ix.ctypes.data_as(L.c_int32_p), L.dptr(dt), rb, re, L.dptr(part)))
        total += part
    ref = O.gram_blas(S, X)
    assert np.max(np.abs(total - ref)) <= 1e-10 * np.max(np.abs(ref))


def test_gram_row_partition_multi_tile_symmetric(engine):
    """Several tiles + symmetric X + row partition: each rank mirrors the upper tiles of ITS partial; the sum over the
    ranks is the full (symmetric) Gram matrix."""
    from rbnics_b200 import _lib as L
    rng = np.random.default_rng(21)
    n, Ns = 26, 200
    X = _lap2d(n)
    S = np.ascontiguousarray(rng.standard_normal((n * n, Ns)))
    ip, ix, dt = offline._csr_arrays(X)
    total = np.zeros((Ns, Ns))
    for rank in range(3):
        rb, re = offline.row_range(n * n, rank, 3)
        part = np.empty((Ns, Ns))
        engine._check(engine._lib.rb_gram(engine._h, n * n, Ns, Ns, L.dptr(S), None, ip.ctypes.data_as(L.c_int64_p),
                                          ix.ctypes.data_as(L.c_int32_p), L.dptr(dt), rb, re, L.dptr(part)))
        total += part
    ref = O.gram_blas(S, X)
    assert np.max(np.abs(total - ref)) <= 1e-10 * np.max(np.abs(ref))


def test_projection_of_operators_and_vectors(engine):
    tp = ThermalBlock(n=14, nblocks=3, nloads=2)
    rng = np.random.default_rng(1)
    Z = rng.standard_normal((tp.Nh, 9))
    A_N = offline.project_matrices(engine, Z, tp.operators_a)
    f_N = offline.project_vectors(engine, Z, tp.operators_f)
    A_ref, f_ref = O.build_reduced_operators(Z, tp.operators_a, tp.operators_f)
    for q in range(3):
        assert np.max(np.abs(A_N[q] - O.gram_loop(Z, tp.operators_a[q]))) <= 1e-10 * np.max(np.abs(A_ref[q]))
    for q in range(2):
        assert np.max(np.abs(f_N[q] - O.project_vector_loop(Z, tp.operators_f[q]))) <= 1e-10 * np.max(np.abs(f_ref[q]))


def test_lincomb(engine):
    rng = np.random.default_rng(3)
    S = rng.standard_normal((777, 45))
    V = rng.standard_normal((45, 70))
    Z = offline.lincomb(engine, S, V)
    ref = S @ V
    assert np.max(np.abs(Z - ref)) <= 1e-10 * np.max(np.abs(ref))


@pytest.mark.parametrize("Nh,Ns,n", [(1, 1, 1), (127, 3, 2), (129, 32, 32), (1000, 33, 31), (4097, 100, 20),
                                     (513, 61, 65), (2000, 130, 7)])
def test_lincomb_shapes(engine, Nh, Ns, n):
    """Row-block tails, partial last column chunks of S, odd n (scalar stores), both tile widths."""
    rng = np.random.default_rng(Nh + Ns + n)
    S = rng.standard_normal((Nh, Ns))
    V = rng.standard_normal((Ns, n))
    Z = offline.lincomb(engine, S, V)
    ref = S @ V
    assert Z.shape == ref.shape
    assert np.max(np.abs(Z - ref)) <= 1e-10 * max(np.max(np.abs(ref)), 1e-300)


def test_gram_schmidt_step(engine):
    tp = ThermalBlock(n=12, nblocks=2)
    X = tp.inner_product
    rng = np.random.default_rng(4)
    Z = np.zeros((tp.Nh, 0))
    gs = offline.GramSchmidt(engine, X)
    for k in range(6):
        snap = rng.standard_normal(tp.Nh)
        z_gpu = gs.apply(snap, Z)
        z_ref = O.gram_schmidt(snap, Z, X)
        assert np.max(np.abs(z_gpu - z_ref)) <= 1e-10 * np.max(np.abs(z_ref))
        Z = np.concatenate([Z, z_ref[:, None]], axis=1)
    # X-orthonormal basis
    assert np.max(np.abs(O.gram_blas(Z, X) - np.eye(6))) < 1e-10


def test_pod_eigenvalues_and_modes(engine):
    tp = ThermalBlock(n=16, nblocks=3, nloads=1)
    X = tp.inner_product
    mus = tp.sample_training_set(30, seed=3)
    snaps = [tp.solve(mu) for mu in mus]
    pod = offline.ProperOrthogonalDecomposition(engine, X)
    for s in snaps:
        pod.store_snapshot(s)
    eigs, vecs, Z, N = pod.apply(10, 1e-9)
    S = np.stack(snaps, axis=1)
    eigs_ref, vecs_ref, Z_ref, N_ref, _ = O.pod(S, X, 10, 1e-9)
    assert N == N_ref
    lam_max = abs(eigs_ref[0])
    assert np.max(np.abs(np.array(eigs) - eigs_ref)) <= 1e-10 * lam_max
    # modes agree up to sign wherever the eigenvalue is well separated from round-off
    for n in range(N):
        if eigs_ref[n] > 1e-8 * lam_max:
            s = np.sign(np.dot(Z[:, n], X @ Z_ref[:, n]))
            assert np.max(np.abs(s * Z[:, n] - Z_ref[:, n])) <= 1e-6 * np.max(np.abs(Z_ref[:, n]))
    G = O.gram_blas(Z, X)
    assert np.max(np.abs(np.diag(G) - 1.0)) < 1e-10


def test_gram_schmidt_mirrors_follow_the_operator_and_the_block(engine):
    """rb_off_gram_schmidt caches X z_k per block.  Replacing the matrix in the SAME operator slot, or
    overwriting mirrored columns with rb_off_spmm, must invalidate that cache (generation counter / gs_n reset);
    every step is compared with the host-path GramSchmidt on the current data."""
    import scipy.sparse as sp
    rng = np.random.default_rng(11)
    Nh = 3000
    X1 = (sp.diags([-1.0, 2.5, -1.0], [-1, 0, 1], shape=(Nh, Nh))).tocsr()
    X2 = (sp.diags([-0.5, 4.0, -0.5], [-1, 0, 1], shape=(Nh, Nh)) + sp.diags([0.3, 0.3], [-7, 7], shape=(Nh, Nh))).tocsr()
    ws = offline.OfflineWorkspace(engine)
    ws.set_operator("X", X1)
    ws.new_block("Z", Nh, 16)
    basis = []
    gs1 = offline.GramSchmidt(engine, X1)
    for k in range(4):
        v = rng.standard_normal(Nh)
        z, _ = ws.gram_schmidt("Z", "X", v)
        zh = gs1.apply(v, np.stack(basis, axis=1) if basis else None)
        assert np.max(np.abs(z - zh)) <= 1e-12 * np.max(np.abs(zh))
        basis.append(z)
    # (1) a NEW matrix registered under the same name / slot: the mirrors X z_k are stale
    ws.set_operator("X", X2)
    v = rng.standard_normal(Nh)
    z, _ = ws.gram_schmidt("Z", "X", v, append=False)
    zh = offline.GramSchmidt(engine, X2).apply(v, np.stack(basis, axis=1))
    assert np.max(np.abs(z - zh)) <= 1e-12 * np.max(np.abs(zh))
    # (2) rb_off_spmm overwrites mirrored columns 1..2 of the block
    ws.new_block("S", Nh, 4)
    ws.append("S", rng.standard_normal((Nh, 2)))
    ws.set_operator("A", X1)
    ws.spmm("A", "S", (0, 2), "Z", 1)
    Znow = ws.read("Z")
    z, _ = ws.gram_schmidt("Z", "X", v, append=False)
    zh = offline.GramSchmidt(engine, X2).apply(v, Znow)
    assert np.max(np.abs(z - zh)) <= 1e-12 * np.max(np.abs(zh))


def test_resident_pod_equals_host_path_and_golden(engine):
    """ResidentProperOrthogonalDecomposition (snapshots resident, one Gram, modes by one product, norms from the
    diagonal) gives the eigenvalues / modes of the upload-every-call class and of the reference's loop (golden)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_offline.npz"))
    Nh = int(g["Nh"])
    X = sp.csr_matrix((g["X_data"], g["X_indices"], g["X_indptr"]), shape=(Nh, Nh))
    S = g["S"]
    Ns = S.shape[1]
    ws = offline.OfflineWorkspace(engine)
    ws.set_operator("X", X)
    pod = offline.ResidentProperOrthogonalDecomposition(ws, "X", Nh, Ns + 3)
    host = offline.ProperOrthogonalDecomposition(engine, X)
    for k in range(Ns):
        pod.store_snapshot(S[:, k])
        host.store_snapshot(S[:, k])
    e1, v1, Z1, N1 = pod.apply(Ns, 1e-10)
    e2, v2, Z2, N2 = host.apply(Ns, 1e-10)
    assert N1 == N2 == int(g["pod_N"])
    lam = g["pod_eigenvalues"]
    assert np.max(np.abs(np.array(pod.eigenvalues) - lam)) <= 1e-10 * lam[0]       # north_star tolerance
    for k in range(min(N1, 3)):       # well-separated leading modes, up to sign
        ref = g["pod_basis"][:, k]
        sgn = np.sign(Z1[:, k] @ (X @ ref))
        assert np.max(np.abs(sgn * Z1[:, k] - ref)) <= 1e-6 * np.max(np.abs(ref))
    for k in range(N1):               # same eigenvectors -> same modes as the host-path class
        sgn2 = np.sign(Z1[:, k] @ (X @ Z2[:, k]))
        assert np.max(np.abs(sgn2 * Z1[:, k] - Z2[:, k])) <= 1e-9 * np.max(np.abs(Z2[:, k]))
    n2 = ws.column_norms2("pod_basis", "X")
    assert np.max(np.abs(n2 - 1.0)) <= 1e-12
    Gs = np.sum(S * S, axis=0)
    assert np.max(np.abs(ws.column_norms2("pod_snapshots", None) - Gs)) <= 1e-12 * np.max(Gs)


class _EmulatedRank:
    """One rank of a dof-row partition without a process group: the test sums the ranks' partials itself."""
    backend = "emulated"

    def __init__(self, rank, size):
        self.rank, self.size = rank, size

    def allreduce_sum_(self, a):
        return a


def test_row_local_uploads_of_a_partitioned_workspace():
    """SURVEY.md 8e row 2 / VERDICT r1 item 7: a rank uploads only its own dof rows plus the rows its operator rows
    reference (rb_off_append_rows); the partials over the ranks' rows still sum to the full products, Y = A S is right on
    the own rows, `read` reassembles the block, and a product whose halo was not uploaded fails loudly."""
    from rbnics_b200.engine import SweepEngine
    side = 60
    Nh, Ns, size = side * side, 37, 3
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(side, side))
    I = sp.identity(side)
    X = (sp.kron(T, I) + sp.kron(I, T) + 0.5 * sp.identity(Nh)).tocsr()          # 5-point stencil: halo of `side` rows
    A = (X + sp.diags([0.3], [2], shape=(Nh, Nh))).tocsr()                       # unsymmetric, halo side / side + ...
    rng = np.random.default_rng(11)
    S = rng.standard_normal((Nh, Ns))
    C_full, CI_full, CA_full = S.T @ (X @ S), S.T @ S, S.T @ (A @ S)
    C_sum, CI_sum, CA_sum, S_sum = 0.0, 0.0, 0.0, 0.0
    for rank in range(size):
        eng = SweepEngine(0)
        try:
            ws = offline.OfflineWorkspace(eng, dist=_EmulatedRank(rank, size))
            assert ws.row_local
            ws.set_operator("X", X)
            ws.set_operator("A", A)
            ws.new_block("S", Nh, Ns)
            ws.new_block("Y", Nh, Ns)
            ws.append("S", S)
            rb, re = offline.row_range(Nh, rank, size)
            lo, hi = ws.upload_range()
            assert (lo, hi) == (max(0, rb - side), min(Nh, re + side)) and hi - lo < Nh
            C_sum = C_sum + ws.gram("S", "X")
            CI_sum = CI_sum + ws.gram("S")
            ws.spmm("A", "S", (0, Ns), "Y", 0)
            CA_sum = CA_sum + ws.gram("S", None, "Y")
            with pytest.raises(RuntimeError, match="are needed"):
                ws.gram("S", "X", "Y")                       # X Y would need halo rows of Y = A S, which are not there
            S_sum = S_sum + ws.read("S")                     # own rows, zeros elsewhere (the all-reduce is emulated)
            wide = (X + sp.diags([1e-3], [3 * side], shape=(Nh, Nh))).tocsr()
            ws.set_operator("W", wide)
            if rank < size - 1:
                with pytest.raises(RuntimeError, match="register the operators before"):
                    ws.gram("S", "W")
        finally:
            eng.close()
    assert np.max(np.abs(C_sum - C_full)) <= 1e-12 * np.max(np.abs(C_full))
    assert np.max(np.abs(CI_sum - CI_full)) <= 1e-12 * np.max(np.abs(CI_full))
    assert np.max(np.abs(CA_sum - CA_full)) <= 1e-12 * np.max(np.abs(CA_full))
    assert np.array_equal(S_sum, S)

"""GPU parity of the batched LU solve (A7: LinearSolver + lifting rows) against numpy.linalg.solve (= LAPACK dgesv, the
reference's solver, rbnics/backends/online/numpy/linear_solver.py:29-31) over the whole size range of every kernel
variant: warp-per-system register kernel (N <= 32), left-looking L2-streamed kernel (32 < N <= 128), shared-memory
fallback (N <= 160); matrices that NEED partial pivoting at every step (row-permuted, indefinite saddle points)."""
import numpy as np
import pytest

from rbnics_b200 import online

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _systems(N, B, Qa, Qf, seed, kind):
    rng = np.random.default_rng(seed)
    A = np.empty((Qa, N, N))
    for q in range(Qa):
        if kind == "permuted":          # diagonally dominant, rows permuted: the natural pivot is never on the diagonal
            M = np.diag(rng.uniform(1.0, 2.0, N)) * N ** 0.5 + 0.3 * rng.standard_normal((N, N))
            A[q] = M
        elif kind == "saddle":           # [[K, B^T], [B, 0]]: indefinite, zero diagonal block
            nv = (2 * N) // 3
            K = rng.standard_normal((nv, nv))
            K = K @ K.T / nv + np.eye(nv)
            Bm = rng.standard_normal((N - nv, nv))
            A[q] = np.block([[K, Bm.T], [Bm, np.zeros((N - nv, N - nv))]])
        else:
            M = rng.standard_normal((N, N))
            A[q] = M @ M.T / N + np.eye(N)
    if kind == "permuted":
        perm = rng.permutation(N)
        A = A[:, perm, :]                # the same permutation for every term keeps the sum well conditioned
    F = rng.standard_normal((Qf, N))
    Theta = np.concatenate([rng.uniform(0.5, 2.0, (B, Qa)), rng.uniform(-1.0, 1.0, (B, Qf))], axis=1)
    return A, F, Theta


@pytest.mark.parametrize("kind", ["spd", "permuted", "saddle"])
@pytest.mark.parametrize("N", [5, 17, 32, 33, 40, 50, 64, 72, 97, 100, 104, 120, 128, 140])
def test_batched_solve_matches_dgesv(engine, N, kind):
    B, Qa, Qf = 150, 3, 2
    A, F, Theta = _systems(N, B, Qa, Qf, seed=N, kind=kind)
    engine.set_system(A, F)
    engine.set_estimator([-1], [N], [Qf], Qa, Qf, np.eye(Qf))      # trivial estimator: only u is under test
    engine.set_training_set(Theta, np.ones(B))
    u = engine.sweep_resident(want_u=True)["u"]
    lhs = np.einsum("pq,qij->pij", Theta[:, :Qa], A)
    rhs = Theta[:, Qa:] @ F
    ref = np.linalg.solve(lhs, rhs[..., None])[..., 0]
    cond = np.linalg.cond(lhs[0])
    err = np.max(np.abs(u - ref), axis=1) / np.max(np.abs(ref), axis=1)
    assert err.max() <= max(RTOL, 1e-14 * cond), f"N={N} {kind}: rel err {err.max():.3e} (cond {cond:.1e})"


@pytest.mark.parametrize("N,nbc", [(40, 3), (100, 2), (128, 5)])
def test_batched_solve_with_lifting_rows(engine, N, nbc):
    """rows i < N_bc replaced by identity rows with rhs theta_bc (DirichletBC.py:41-56), plus a lifting row in the
    middle of the matrix (block problems: one per component, slice_to_array.py:27-45)."""
    B, Qa, Qf = 64, 2, 2
    A, F, Theta = _systems(N, B, Qa, Qf, seed=7 * N, kind="spd")
    rng = np.random.default_rng(1)
    rows = list(range(nbc - 1)) + [N // 2]
    ThetaBC = rng.uniform(-1, 1, (B, nbc))
    engine.set_system(A, F, bc_rows=rows)
    engine.set_estimator([-1], [N], [Qf], Qa, Qf, np.eye(Qf))
    engine.set_training_set(Theta, np.ones(B), ThetaBC)
    u = engine.sweep_resident(want_u=True)["u"]
    lhs = np.einsum("pq,qij->pij", Theta[:, :Qa], A)
    rhs = Theta[:, Qa:] @ F
    for t, r in enumerate(rows):
        lhs[:, r, :] = 0.0
        lhs[:, r, r] = 1.0
        rhs[:, r] = ThetaBC[:, t]
    ref = np.linalg.solve(lhs, rhs[..., None])[..., 0]
    assert np.max(np.abs(u - ref)) <= RTOL * np.max(np.abs(ref))
    assert np.allclose(u[:, rows], ThetaBC, rtol=0, atol=1e-13)


@pytest.mark.parametrize("N", [12, 48, 100])
def test_single_solve_and_singular_matrix(engine, N):
    rng = np.random.default_rng(N)
    M = rng.standard_normal((N, N)) + N ** 0.5 * np.eye(N)
    rhs = rng.standard_normal(N)
    x = online.solve_dense(engine, M, rhs)
    ref = np.linalg.solve(M, rhs)
    assert np.max(np.abs(x - ref)) <= RTOL * np.max(np.abs(ref))
    M[:, 3] = 0.0                                       # exactly singular (zero pivot): numpy raises LinAlgError
    with pytest.raises(np.linalg.LinAlgError):
        np.linalg.solve(M, rhs)
    from rbnics_b200._lib import SingularReducedMatrixError
    with pytest.raises(SingularReducedMatrixError):
        online.solve_dense(engine, M, rhs)

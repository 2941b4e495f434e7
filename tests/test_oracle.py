"""CPU tests: the oracle against the reference's own known-answer formulas and its internal variants."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import rb_oracle as O


def _ref_problem(rng, N, Qa, Qf):
    """Inputs drawn like the reference's performance tests (test_numpy_utils.py:58-59)."""
    Gaa = O.reference_rand(rng, Qa, Qa, N, N)
    Gaf = O.reference_rand(rng, Qa, Qf, N)
    Gff = O.reference_rand(rng, Qf, Qf)
    return Gaa, Gaf, Gff


@pytest.mark.parametrize("N", [16, 32])
@pytest.mark.parametrize("Qa", [6, 10])
@pytest.mark.parametrize("Qf", [6, 10])
def test_order2_products_match_reference_einsum_kat(N, Qa, Qf):
    """test_numpy_greedy_prototype.py:72-116: backend expression vs closed-form einsum, mean rel err <= 1e-10, Ntrain=10."""
    rng = np.random.default_rng(N * 100 + Qa * 10 + Qf)
    Gaa, Gaf, Gff = _ref_problem(rng, N, Qa, Qf)
    rel = []
    for _ in range(10):
        tha = tuple(float(x) for x in O.reference_rand(rng, Qa))
        thf = tuple(float(x) for x in O.reference_rand(rng, Qf))
        u = O.reference_rand(rng, N)
        v = O.reference_rand(rng, N)
        builtin = O.estimator_builtin_einsum(u, v, np.array(tha), np.array(thf), Gaa, Gaf, Gff)
        aa = O.product_order2(tha, [[Gaa[i, j] for j in range(Qa)] for i in range(Qa)], tha)
        af = O.product_order2(tha, [[Gaf[i, j] for j in range(Qf)] for i in range(Qa)], thf)
        ff = O.product_order2(thf, [[float(Gff[i, j]) for j in range(Qf)] for i in range(Qf)], thf)
        backend = np.dot(u, aa @ v) + np.dot(u, af) + ff
        rel.append(abs(builtin - backend) / abs(builtin))
    assert np.isclose(sum(rel) / 10, 0.0, atol=1e-10)


@pytest.mark.parametrize("N,Q", [(4, 2), (16, 6), (64, 9)])
def test_order1_assembly_matches_builtin_sum(N, Q):
    """test_numpy_matrix_assembly.py:36-48 / test_numpy_vector_assembly.py: sequential sum theta_q A_q, rel err <= 1e-12."""
    rng = np.random.default_rng(N + Q)
    A = O.reference_rand(rng, Q, N, N)
    F = O.reference_rand(rng, Q, N)
    th = tuple(float(x) for x in O.reference_rand(rng, Q))
    out = O.product_order1(th, [A[q] for q in range(Q)])
    builtin = th[0] * A[0]
    for q in range(1, Q):
        builtin = builtin + th[q] * A[q]
    assert np.linalg.norm(out - builtin) / np.linalg.norm(builtin) <= 1e-12
    outf = O.product_order1(th, [F[q] for q in range(Q)])
    assert np.linalg.norm(outf - np.einsum("q,qn->n", th, F)) / np.linalg.norm(outf) <= 1e-12


def test_zero_theta_terms_are_skipped_but_first_is_not():
    """product.py:29-32: index 0 always multiplies (even by 0), later zero thetas are skipped -> NaN operators
    behind a zero theta do not poison the sum unless they are the first."""
    A = [np.ones((2, 2)), np.full((2, 2), np.nan), 2 * np.ones((2, 2))]
    out = O.product_order1((1.0, 0.0, 3.0), A)
    assert np.array_equal(out, 7 * np.ones((2, 2)))
    out = O.product_order1((0.0, 1.0), [np.full((2, 2), np.nan), np.ones((2, 2))])
    assert np.all(np.isnan(out))


def test_lifting_rows():
    """DirichletBC.py:41-56 via linear_solver.py:56-59."""
    lhs = np.arange(16.0).reshape(4, 4) + 10 * np.eye(4)
    rhs = np.ones(4)
    O.apply_dirichlet_rows(lhs, rhs, (0.5, -2.0))
    assert np.array_equal(lhs[0], [1, 0, 0, 0]) and np.array_equal(lhs[1], [0, 1, 0, 0])
    assert rhs[0] == 0.5 and rhs[1] == -2.0 and rhs[2] == 1.0
    u = np.linalg.solve(lhs, rhs)
    assert u[0] == 0.5 and u[1] == -2.0


def test_argmax_and_parallel_max_rules():
    """numpy.argmax: first maximum, NaN wins (parameter_space_subset.py:67); MPI rule: highest rank wins ties
    (parallel_max.py:18-33)."""
    assert O.argmax_first(np.array([1.0, 3.0, 3.0, 2.0])) == 1
    assert O.argmax_first(np.array([1.0, np.nan, 5.0, np.nan])) == 1
    assert O.parallel_max([3.0, 5.0, 5.0, 1.0], [7, 1, 2, 3]) == (5.0, 2)
    vals = np.array([0.3, 0.9, 0.1, 0.9, 0.5, 0.2, 0.9])
    ts = list(range(7))
    locs = [O.training_set_max(lambda m: vals[m], ts, rank=r, size=3) for r in range(3)]
    assert locs == [(0.9, 3), (0.9, 1), (0.2, 5)]  # rank 0 owns 0,3,6; rank 1 owns 1,4; rank 2 owns 2,5


@pytest.mark.parametrize("N,Qa,Qf", [(0, 2, 3), (5, 2, 1), (12, 4, 3)])
def test_loop_and_batched_sweeps_agree(N, Qa, Qf):
    from rbnics_b200 import synthetic
    prob = synthetic.elliptic_reduced_problem(N, Qa, Qf, seed=1, riesz_rank=16)
    Theta, beta, _ = synthetic.training_set(25, Qa, Qf, seed=2)
    A_q = [prob["A"][q] for q in range(Qa)]
    f_q = [prob["F"][q] for q in range(Qf)]
    G_ff = [[prob["Gff"][i, j] for j in range(Qf)] for i in range(Qf)]
    G_af = [[prob["Gaf"][i, j] for j in range(Qf)] for i in range(Qa)]
    G_aa = [[prob["Gaa"][i, j] for j in range(Qa)] for i in range(Qa)]
    for kind in ("sqrt_eps2_over_beta", "sqrt_of_eps2_over_beta"):
        d1, U1, i1, v1 = O.sweep_loop(Theta[:, :Qa], Theta[:, Qa:], beta, A_q, f_q, G_ff, G_af, G_aa, N, kind=kind)
        d2, U2, i2, v2 = O.sweep_batched(Theta[:, :Qa], Theta[:, Qa:], beta, prob["A"], prob["F"], prob["Gff"],
                                         prob["Gaf"], prob["Gaa"], kind=kind)
        assert i1 == i2
        assert np.max(np.abs(d1 - d2) / d1) <= 1e-12
        if N:
            assert np.max(np.abs(U1 - U2)) <= 1e-12 * np.max(np.abs(U1))


def test_gram_loop_matches_builtin_nested_loops():
    """tests/performance/backends/dolfin/test_dolfin_S_T_dot_A_S.py:42-56: S^T A S vs nested-loop builtin, 1e-12."""
    rng = np.random.default_rng(0)
    n = 40
    X = (sp.random(n, n, density=0.2, random_state=1) + sp.identity(n)).tocsr()
    S = rng.standard_normal((n, 6))
    C = O.gram_loop(S, X)
    Xd = X.toarray()
    builtin = np.array([[S[:, i] @ (Xd @ S[:, j]) for j in range(6)] for i in range(6)])
    assert np.linalg.norm(C - builtin) / np.linalg.norm(builtin) <= 1e-12
    assert np.linalg.norm(O.gram_blas(S, X) - builtin) / np.linalg.norm(builtin) <= 1e-12


def test_pod_and_gram_schmidt_properties():
    rng = np.random.default_rng(3)
    n = 50
    T = sp.diags([-1.0, 2.5, -1.0], [-1, 0, 1], shape=(n, n)).tocsr()
    S = rng.standard_normal((n, 8)) @ np.diag(10.0 ** -np.arange(8))
    eigs, vecs, Z, N, retained = O.pod(S, T, 8, 1e-6)
    assert np.all(np.diff(eigs) <= 0)                      # sorted descending (eigen_solver.py:44-46)
    assert retained[N - 1] > 1 - 1e-6 and (N == 1 or retained[N - 2] <= 1 - 1e-6)
    assert np.allclose(np.diag(Z.T @ (T @ Z)), 1.0, atol=1e-12)
    Zb = np.zeros((n, 0))
    for k in range(5):
        z = O.gram_schmidt(rng.standard_normal(n), Zb, T)
        Zb = np.concatenate([Zb, z[:, None]], axis=1)
    assert np.max(np.abs(Zb.T @ (T @ Zb) - np.eye(5))) < 1e-12

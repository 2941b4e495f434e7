"""GPU parity of the greedy sweep (assembly + LU solve + estimator + arg-max) against the CPU oracle.

Tolerances (north_star): reduced solutions and estimators within 1e-10 relative in FP64; identical arg-max index."""
import numpy as np
import pytest

from oracle import rb_oracle as O
from rbnics_b200 import synthetic
from rbnics_b200._lib import RB_EST_SQRT_EPS2_OVER_BETA, RB_EST_SQRT_OF_EPS2_OVER_BETA

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _run(engine, prob, Theta, beta, ThetaBC=None, bc_rows=None, kind=RB_EST_SQRT_EPS2_OVER_BETA):
    engine.set_system(prob["A"], prob["F"], bc_rows=bc_rows)
    engine.set_elliptic_estimator(prob["Gff"], prob["Gaf"], prob["Gaa"])
    engine.set_training_set(Theta, beta, ThetaBC)
    return engine.sweep_resident(kind=kind, want_delta=True, want_u=True, want_eps2=True)


def _check(res, ref, N):
    d_ref, U_ref, i_ref, v_ref, e_ref = ref
    if N > 0:
        err_u = np.max(np.abs(res["u"] - U_ref), axis=1) / np.max(np.abs(U_ref), axis=1)
        assert err_u.max() <= RTOL, f"u rel err {err_u.max():.3e}"
    err_e = np.abs(res["eps2"] - e_ref) / np.abs(e_ref)
    assert err_e.max() <= RTOL, f"eps2 rel err {err_e.max():.3e}"
    err_d = np.abs(res["delta"] - d_ref) / np.abs(d_ref)
    assert err_d.max() <= RTOL, f"delta rel err {err_d.max():.3e}"
    assert res["argmax"] == i_ref
    assert abs(res["max"] - v_ref) <= RTOL * abs(v_ref)


@pytest.mark.parametrize("N,Qa,Qf,B", [(4, 2, 1, 100), (16, 6, 6, 10), (32, 10, 10, 10), (20, 3, 2, 777),
                                        (33, 5, 3, 300), (64, 9, 3, 500), (100, 50, 10, 384)])
def test_sweep_matches_batched_oracle(engine, N, Qa, Qf, B):
    prob = synthetic.elliptic_reduced_problem(N, Qa, Qf, seed=N + Qa, riesz_rank=64)
    Theta, beta, _ = synthetic.training_set(B, Qa, Qf, seed=B)
    res = _run(engine, prob, Theta, beta)
    ref = O.sweep_batched(Theta[:, :Qa], Theta[:, Qa:], beta, prob["A"], prob["F"], prob["Gff"], prob["Gaf"],
                          prob["Gaa"], return_eps2=True)
    _check(res, ref, N)


@pytest.mark.parametrize("N,Qa,Qf,B", [(4, 2, 1, 100), (16, 6, 6, 10), (25, 4, 3, 40)])
def test_sweep_matches_reference_loop_oracle(engine, N, Qa, Qf, B):
    """Against the loop-faithful restatement (sequential sum_q, row-major (i,j) loops, numpy.linalg.solve)."""
    prob = synthetic.elliptic_reduced_problem(N, Qa, Qf, seed=7, riesz_rank=48)
    Theta, beta, _ = synthetic.training_set(B, Qa, Qf, seed=11)
    res = _run(engine, prob, Theta, beta, kind=RB_EST_SQRT_OF_EPS2_OVER_BETA)
    A_q = [prob["A"][q] for q in range(Qa)]
    f_q = [prob["F"][q] for q in range(Qf)]
    G_ff = [[prob["Gff"][i, j] for j in range(Qf)] for i in range(Qf)]
    G_af = [[prob["Gaf"][i, j] for j in range(Qf)] for i in range(Qa)]
    G_aa = [[prob["Gaa"][i, j] for j in range(Qa)] for i in range(Qa)]
    ref = O.sweep_loop(Theta[:, :Qa], Theta[:, Qa:], beta, A_q, f_q, G_ff, G_af, G_aa, N,
                       kind="sqrt_of_eps2_over_beta", return_eps2=True)
    _check(res, ref, N)


def test_sweep_N0_first_greedy_iteration(engine):
    """N == 0: empty solution, estimator = sqrt(|thf^T Gff thf|)/beta (parametrized_reduced_differential_problem.py:328-333)."""
    Qa, Qf, B = 3, 4, 1000
    prob = synthetic.elliptic_reduced_problem(0, Qa, Qf, seed=3, riesz_rank=16)
    Theta, beta, _ = synthetic.training_set(B, Qa, Qf, seed=5)
    res = _run(engine, prob, Theta, beta)
    ref = O.sweep_batched(Theta[:, :Qa], Theta[:, Qa:], beta, prob["A"], prob["F"], prob["Gff"], prob["Gaf"],
                          prob["Gaa"], return_eps2=True)
    _check(res, ref, 0)


def test_sweep_with_lifting_rows(engine):
    """N_bc > 0: rows i < N_bc become identity rows with rhs theta_bc (DirichletBC.py:41-56)."""
    N, Qa, Qf, B, nbc = 12, 3, 2, 200, 2
    prob = synthetic.elliptic_reduced_problem(N, Qa, Qf, seed=2, riesz_rank=32)
    Theta, beta, ThetaBC = synthetic.training_set(B, Qa, Qf, seed=9, n_bc=nbc)
    res = _run(engine, prob, Theta, beta, ThetaBC=ThetaBC, bc_rows=list(range(nbc)))
    ref = O.sweep_batched(Theta[:, :Qa], Theta[:, Qa:], beta, prob["A"], prob["F"], prob["Gff"], prob["Gaf"],
                          prob["Gaa"], Theta_bc=ThetaBC, return_eps2=True)
    _check(res, ref, N)


def test_argmax_ties_take_lowest_index(engine):
    """Duplicate parameters -> equal estimators -> numpy.argmax returns the first (parameter_space_subset.py:67)."""
    N, Qa, Qf = 8, 2, 2
    prob = synthetic.elliptic_reduced_problem(N, Qa, Qf, seed=4, riesz_rank=32)
    Theta, beta, _ = synthetic.training_set(50, Qa, Qf, seed=1)
    Theta = np.concatenate([Theta, Theta, Theta])
    beta = np.concatenate([beta, beta, beta])
    res = _run(engine, prob, Theta, beta)
    ref = O.sweep_batched(Theta[:, :Qa], Theta[:, Qa:], beta, prob["A"], prob["F"], prob["Gff"], prob["Gaf"],
                          prob["Gaa"])
    assert res["argmax"] == ref[2] < 50
    d = res["delta"]
    assert np.array_equal(d[:50], d[50:100]) and np.array_equal(d[:50], d[100:])


def test_global_index_mapping_cyclic_shard(engine):
    """index_offset/stride reproduce the reference's cyclic shard range(rank, len, size) (parameter_space_subset.py:57)."""
    N, Qa, Qf, B = 8, 2, 2, 101
    prob = synthetic.elliptic_reduced_problem(N, Qa, Qf, seed=4, riesz_rank=32)
    Theta, beta, _ = synthetic.training_set(B, Qa, Qf, seed=1)
    engine.set_system(prob["A"], prob["F"])
    engine.set_elliptic_estimator(prob["Gff"], prob["Gaf"], prob["Gaa"])
    full = O.sweep_batched(Theta[:, :Qa], Theta[:, Qa:], beta, prob["A"], prob["F"], prob["Gff"], prob["Gaf"], prob["Gaa"])
    best = (-1.0, -1)
    for rank in range(3):
        sl = slice(rank, B, 3)
        engine.set_training_set(Theta[sl], beta[sl], index_offset=rank, index_stride=3)
        r = engine.sweep_resident()
        if r["max"] > best[0] or (r["max"] == best[0] and r["argmax"] < best[1]):
            best = (r["max"], r["argmax"])
    assert best[1] == full[2]


def test_testing_set_sweep_with_outputs(engine):
    """error_analysis' reduced-order half: u, Delta and the output s_N = u^T sum theta_s s_q for a testing set
    (rb_reduction.py:312-390, elliptic_reduced_problem.py:31-32)."""
    from rbnics_b200.sweep import TestingSetSweep
    N, Qa, Qf, Qs, B = 40, 4, 3, 2, 700
    prob = synthetic.elliptic_reduced_problem(N, Qa, Qf, seed=8, riesz_rank=48)
    Theta, beta, _ = synthetic.training_set(B, Qa, Qf, seed=3)
    rng = np.random.default_rng(4)
    S, ThetaS = rng.standard_normal((Qs, N)), rng.uniform(-1, 1, (B, Qs))
    engine.set_system(prob["A"], prob["F"])
    engine.set_elliptic_estimator(prob["Gff"], prob["Gaf"], prob["Gaa"])
    r = TestingSetSweep(engine, Theta, beta).evaluate(S=S, ThetaS=ThetaS)
    d_ref, U_ref, _, _ = O.sweep_batched(Theta[:, :Qa], Theta[:, Qa:], beta, prob["A"], prob["F"], prob["Gff"],
                                         prob["Gaf"], prob["Gaa"])
    s_ref = np.array([np.dot(U_ref[p], O.product_order1(tuple(ThetaS[p]), [S[q] for q in range(Qs)])) for p in range(B)])
    assert np.max(np.abs(r["u"] - U_ref)) <= RTOL * np.max(np.abs(U_ref))
    assert np.max(np.abs(r["delta"] - d_ref) / d_ref) <= RTOL
    assert np.max(np.abs(r["output"] - s_ref)) <= RTOL * np.max(np.abs(s_ref))


@pytest.mark.parametrize("N,Qa,Qf,B", [(1, 1, 1, 1), (2, 1, 1, 3), (3, 2, 1, 129), (5, 1, 2, 127), (31, 3, 1, 257),
                                        (32, 2, 2, 1), (33, 2, 1, 1), (128, 2, 1, 130)])
def test_sweep_edge_shapes(engine, N, Qa, Qf, B):
    """Smallest / ragged shapes: N = 1..3, a single parameter, batch sizes around the 128-row tile, the kernel
    switch-over points N = 32 / 33 and the largest register-panel size N = 128."""
    prob = synthetic.elliptic_reduced_problem(N, Qa, Qf, seed=N + B, riesz_rank=16)
    Theta, beta, _ = synthetic.training_set(B, Qa, Qf, seed=N)
    res = _run(engine, prob, Theta, beta)
    ref = O.sweep_batched(Theta[:, :Qa], Theta[:, Qa:], beta, prob["A"], prob["F"], prob["Gff"], prob["Gaf"],
                          prob["Gaa"], return_eps2=True)
    _check(res, ref, N)


@pytest.mark.parametrize("N,Ql", [(100, 50), (37, 7), (64, 12), (5, 1)])
def test_record_fed_assembly_equals_the_cp_async_kernel(engine, monkeypatch, N, Ql):
    """rb_assemble_rec_kernel (producer warp + DMMA B-fragment records, the default) and rb_assemble_kernel
    (RB_ASM_KERNEL=old at rb_set_system) accumulate the terms in the same order: identical reduced solutions."""
    rng = np.random.default_rng(N * 100 + Ql)
    Qr, B = 3, 700
    A = np.stack([np.eye(N) * (2.0 + q) + 0.1 * rng.standard_normal((N, N)) for q in range(Ql)])
    F = rng.standard_normal((Qr, N))
    Theta = np.concatenate([rng.uniform(0.5, 2.0, (B, Ql)), rng.uniform(-1, 1, (B, Qr))], axis=1)
    beta = rng.uniform(0.5, 2.0, B)
    G = rng.standard_normal((N + Qr, N + Qr))
    out = {}
    for kind in ("new", "old"):
        monkeypatch.setenv("RB_ASM_KERNEL", kind)
        engine.set_system(A, F)
        engine.set_estimator([0, -1], [0, N], [N, Qr], Ql, Qr, G)
        engine.set_training_set(Theta, beta)
        out[kind] = engine.sweep_resident(want_u=True, raise_on_singular=False)["u"].copy()
    assert np.array_equal(out["new"], out["old"])
    U_ref = np.linalg.solve(np.einsum("pq,qij->pij", Theta[:, :Ql], A), (Theta[:, Ql:] @ F)[:, :, None])[:, :, 0]
    assert np.max(np.abs(out["new"] - U_ref)) <= 1e-10 * np.max(np.abs(U_ref))

"""POD-Greedy offline loop of a parabolic problem (BASELINE.json configs[3], tutorials/06_thermal_block_unsteady, RB
variant) end to end: rbnics_b200.pod_greedy.PODGreedyReducedBasis -- trajectory POD, projections and Riesz Gram blocks on
the offline kernels, the sweep in the fused parabolic kernel -- against the CPU restatement of the reference's loop
(oracle/pod_greedy_oracle.py): identical sequence of selected parameters, estimator maxima, the same basis."""
import numpy as np
import pytest

from oracle import pod_greedy_oracle
from rbnics_b200.pod_greedy import PODGreedyReducedBasis
from truth_problems import ThermalBlockUnsteady

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("n,nblocks,N1,Nmax,K", [(10, 2, 2, 8, 20), (12, 3, 3, 9, 30)])
def test_pod_greedy_matches_the_reference_loop(engine, n, nblocks, N1, Nmax, K, resident):
    tp = ThermalBlockUnsteady(n=n, nblocks=nblocks, nloads=1, dt=0.05, K=K)
    train = tp.sample_training_set(60, seed=n)
    ref = pod_greedy_oracle.pod_greedy_offline(tp, train, Nmax=Nmax, tol=1e-6, N1=N1)
    run = PODGreedyReducedBasis(engine, tp, train, Nmax=Nmax, tol=1e-6, N1=N1, resident=resident).offline()
    assert run.greedy_selected_indices == ref["selected"]
    assert run.N == ref["Z"].shape[1] and run.N >= 2 * N1
    est, est_ref = np.array(run.greedy_error_estimators), np.array(ref["estimators"])
    # the time-integrated estimator cancels down by orders of magnitude over the iterations: parity relative to the
    # first (uncancelled) value, as for the elliptic greedy (tests/test_greedy_gpu.py)
    assert np.max(np.abs(est - est_ref)) <= 1e-7 * est_ref[0]
    assert abs(est[0] - est_ref[0]) <= 1e-12 * est_ref[0]
    # POD modes are defined up to sign; the spans agree and the basis is X-orthonormal
    Z, Zr, X = run.basis_functions, ref["Z"], tp.inner_product
    cross = Zr.T @ (X @ Z)
    assert np.max(np.abs(np.abs(np.diag(cross)) - 1.0)) <= 1e-6
    gram = Z.T @ (X @ Z)
    assert np.max(np.abs(gram - np.eye(run.N))) <= 1e-8


@pytest.mark.parametrize("resident", [False, True])
def test_pod_greedy_reduced_operators_and_storages(engine, resident):
    """After two iterations the reduced operators and the six error-estimation storages equal their definitions."""
    tp = ThermalBlockUnsteady(n=9, nblocks=2, nloads=1, dt=0.05, K=12)
    train = tp.sample_training_set(30, seed=1)
    run = PODGreedyReducedBasis(engine, tp, train, Nmax=4, tol=1e-12, N1=2, resident=resident).offline()
    Z, X, N = run.basis_functions, tp.inner_product, run.N
    assert N == 4
    assert np.max(np.abs(run.operator_m[0] - Z.T @ (tp.operators_m[0] @ Z))) <= 1e-13
    for q, A in enumerate(tp.operators_a):
        assert np.max(np.abs(run.operator_a[q] - Z.T @ (A @ Z))) <= 1e-12 * np.max(np.abs(run.operator_a[q]))
    Ra = [np.stack([-tp.riesz_solve(A @ Z[:, k]) for k in range(N)], axis=1) for A in tp.operators_a]
    Rm = [np.stack([-tp.riesz_solve(M @ Z[:, k]) for k in range(N)], axis=1) for M in tp.operators_m]
    rf = [tp.riesz_solve(f) for f in tp.operators_f]
    E = run.error_estimation_operator
    assert abs(E["f", "f"][0, 0] - rf[0] @ (X @ rf[0])) <= 1e-12 * abs(E["f", "f"][0, 0])
    assert np.max(np.abs(E["m", "a"][0, 1] - Rm[0].T @ (X @ Ra[1]))) <= 1e-11 * np.max(np.abs(E["m", "a"][0, 1]))
    assert np.max(np.abs(E["a", "f"][1, 0] - Ra[1].T @ (X @ rf[0]))) <= 1e-11 * np.max(np.abs(E["a", "f"][1, 0]))
    assert np.max(np.abs(E["m", "m"][0, 0] - Rm[0].T @ (X @ Rm[0]))) <= 1e-11 * np.max(np.abs(E["m", "m"][0, 0]))

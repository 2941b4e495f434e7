"""Parity against the REAL reference package (tests/golden/make_reference_golden.py: RBniCS imported with the
third-party stand-ins of oracle/shims, its own EllipticCoercive[Compliant]RBReducedProblem.solve()/estimate_error(),
ParameterSpaceSubset.max and OnlineAffineExpansionStorage.save produce the fixtures).

CPU part: the loop-faithful oracle is bit-identical to the reference's outputs (solutions, estimators, max, arg-max).
GPU part: the CUDA sweep through the C ABI agrees to 1e-10 relative with an identical arg-max."""
import glob
import os

import numpy as np
import pytest

from oracle import rb_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(os.path.basename(p)[len("reference_real_sweep_"):-len(".npz")]
               for p in glob.glob(os.path.join(HERE, "golden", "reference_real_sweep_*.npz")))
RTOL = 1e-10


def _load(case):
    return dict(np.load(os.path.join(HERE, "golden", f"reference_real_sweep_{case}.npz")))


def test_fixtures_present():
    assert {"n8_q3", "n20_q4", "n33_q6", "n16_q5_compliant"} <= set(CASES)


@pytest.mark.parametrize("case", CASES)
def test_loop_oracle_is_bit_identical_to_the_real_reference(case):
    g = _load(case)
    Qa, Qf, N = int(g["Qa"]), int(g["Qf"]), int(g["N"])
    A_q = [g["A"][q] for q in range(Qa)]
    f_q = [g["F"][q] for q in range(Qf)]
    G_ff = [[float(g["Gff"][i, j]) for j in range(Qf)] for i in range(Qf)]
    G_af = [[g["Gaf"][i, j] for j in range(Qf)] for i in range(Qa)]
    G_aa = [[g["Gaa"][i, j] for j in range(Qa)] for i in range(Qa)]
    kind = "sqrt_of_eps2_over_beta" if int(g["compliant"]) else "sqrt_eps2_over_beta"
    Theta = g["Theta"]
    beta = Theta[:, :Qa].min(axis=1)      # get_stability_factor_lower_bound of the synthetic truth problem
    assert np.array_equal(beta, g["beta"])
    delta, U, imax, vmax = O.sweep_loop(Theta[:, :Qa], Theta[:, Qa:], beta, A_q, f_q, G_ff, G_af, G_aa, N, kind=kind)
    assert np.array_equal(np.asarray(U), g["u"])
    assert np.array_equal(np.asarray(delta), g["delta"])
    assert (float(vmax), int(imax)) == (float(g["max"]), int(g["argmax"]))


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_sweep_matches_the_real_reference(engine, case):
    from rbnics_b200 import _lib as L
    g = _load(case)
    Qa = int(g["Qa"])
    engine.set_system(g["A"], g["F"])
    engine.set_elliptic_estimator(g["Gff"], g["Gaf"], g["Gaa"])
    kind = L.RB_EST_SQRT_OF_EPS2_OVER_BETA if int(g["compliant"]) else L.RB_EST_SQRT_EPS2_OVER_BETA
    res = engine.sweep(g["Theta"], g["Theta"][:, :Qa].min(axis=1), kind=kind, want_delta=True, want_u=True)
    assert res["argmax"] == int(g["argmax"])
    assert abs(res["max"] - float(g["max"])) <= RTOL * float(g["max"])
    assert np.max(np.abs(res["delta"] - g["delta"]) / np.abs(g["delta"])) <= RTOL
    assert np.max(np.abs(res["u"] - g["u"])) <= RTOL * np.max(np.abs(g["u"]))


# ----------------------------------------------------------------------------------------------------------------------
# POD-Greedy (BASELINE configs[3]): the COMPLETE offline stage of the real package on a parabolic problem,
# ReducedBasis(ParabolicCoerciveProblem).offline() with set_Nmax(Nmax, POD_Greedy=N1)
# (tests/golden/reference_real_pod_greedy_<case>.npz, tests/refrun/run_reference.py pod_greedy)
# ----------------------------------------------------------------------------------------------------------------------
POD_GREEDY_CASES = ("unsteady2", "unsteady3")


def _pod_greedy_case(case):
    import sys
    sys.path.insert(0, HERE)
    from truth_problems import ThermalBlockUnsteady
    g = dict(np.load(os.path.join(HERE, "golden", f"reference_real_pod_greedy_{case}.npz")))
    tp = ThermalBlockUnsteady(n=int(g["n"]), nblocks=int(g["nblocks"]), nloads=int(g["nloads"]), dt=float(g["dt"]),
                              K=int(g["K"]))
    train = [tuple(float(x) for x in mu) for mu in g["training_set"]]
    assert train == tp.sample_training_set(int(g["ntrain"]), seed=int(g["seed"]))   # the reference's own sampling
    return g, tp, train


def _same_basis_up_to_signs(Z, Zr, X, tol):
    cross = Zr.T @ (X @ Z)
    assert np.max(np.abs(np.abs(np.diag(cross)) - 1.0)) <= tol
    return np.sign(np.diag(cross))


@pytest.mark.parametrize("case", POD_GREEDY_CASES)
def test_pod_greedy_oracle_matches_the_real_reference(case):
    """The CPU restatement of the POD-Greedy loop selects the real package's parameters, with its estimators and basis."""
    from oracle import pod_greedy_oracle
    g, tp, train = _pod_greedy_case(case)
    ref = pod_greedy_oracle.pod_greedy_offline(tp, train, Nmax=int(g["Nmax"]), tol=float(g["tol"]), N1=int(g["N1"]))
    assert ref["selected"] == [int(i) for i in g["selected_indices"]]
    est, est_ref = np.array(ref["estimators"]), g["error_estimators"]
    assert np.max(np.abs(est - est_ref)) <= 1e-9 * est_ref[0]
    assert ref["Z"].shape[1] == int(g["N"])
    _same_basis_up_to_signs(ref["Z"], g["Z"], tp.inner_product, 1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("case", POD_GREEDY_CASES)
def test_cuda_pod_greedy_matches_the_real_reference(engine, case, resident):
    """rbnics_b200.pod_greedy on the GPU against the real package's complete POD-Greedy offline stage: identical
    sequence of selected parameters, estimator maxima, basis and reduced operators (POD modes up to their signs)."""
    from rbnics_b200.pod_greedy import PODGreedyReducedBasis
    g, tp, train = _pod_greedy_case(case)
    run = PODGreedyReducedBasis(engine, tp, train, Nmax=int(g["Nmax"]), tol=float(g["tol"]), N1=int(g["N1"]),
                                resident=resident).offline()
    assert run.greedy_selected_indices == [int(i) for i in g["selected_indices"]]
    est, est_ref = np.array(run.greedy_error_estimators), g["error_estimators"]
    assert np.max(np.abs(est - est_ref)) <= 1e-7 * est_ref[0]
    assert abs(est[0] - est_ref[0]) <= 1e-12 * est_ref[0]
    assert run.N == int(g["N"])
    sgn = _same_basis_up_to_signs(run.basis_functions, g["Z"], tp.inner_product, 1e-6)
    D = np.outer(sgn, sgn)
    for ours, theirs in ((run.operator_m, g["M"]), (run.operator_a, g["A"])):
        assert np.max(np.abs(ours * D[None] - theirs)) <= 1e-6 * np.max(np.abs(theirs))
    assert np.max(np.abs(run.operator_f * sgn[None] - g["F"])) <= 1e-6 * np.max(np.abs(g["F"]))


# ----------------------------------------------------------------------------------------------------------------------
# POD-Galerkin: the COMPLETE offline stage of the real package, PODGalerkin(problem).offline()
# (tests/golden/reference_real_pod_galerkin.npz, tests/refrun/run_reference.py pod_galerkin)
# ----------------------------------------------------------------------------------------------------------------------
def _pod_galerkin_case():
    import sys
    sys.path.insert(0, HERE)
    from truth_problems import ThermalBlock
    g = dict(np.load(os.path.join(HERE, "golden", "reference_real_pod_galerkin.npz")))
    tp = ThermalBlock(n=int(g["n"]), nblocks=int(g["nblocks"]), nloads=int(g["nloads"]))
    train = [tuple(float(x) for x in mu) for mu in g["training_set"]]
    assert train == tp.sample_training_set(int(g["ntrain"]), seed=int(g["seed"]))
    return g, tp, train


def test_pod_oracle_matches_the_real_reference_pod_galerkin():
    g, tp, train = _pod_galerkin_case()
    S = np.stack([tp.solve(mu) for mu in train], axis=1)
    eigs, _, Z, N, retained = O.pod(S, tp.inner_product, int(g["Nmax"]), float(g["tol"]))
    lam = g["eigenvalues"]
    assert N == int(g["N"])
    assert np.max(np.abs(eigs - lam[:N])) <= 1e-10 * lam[0]          # north_star: 1e-10 relative to the largest
    assert np.max(np.abs(retained - g["retained_energy"])) <= 1e-10
    _same_basis_up_to_signs(Z, g["Z"], tp.inner_product, 1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [True, False])
def test_cuda_pod_galerkin_matches_the_real_reference(engine, resident):
    """rbnics_b200.pod_galerkin.PODGalerkin on the GPU against PODGalerkin(problem).offline() of the real package: POD
    eigenvalues within 1e-10 of the largest (north_star), the same number of modes and retained energy, the same basis
    and reduced operators up to the signs of the modes."""
    from rbnics_b200.pod_galerkin import PODGalerkin
    g, tp, train = _pod_galerkin_case()
    run = PODGalerkin(engine, tp, train, Nmax=int(g["Nmax"]), tol=float(g["tol"]), resident=resident).offline()
    lam = g["eigenvalues"]
    assert run.N == int(g["N"])
    assert np.max(np.abs(np.array(run.POD.eigenvalues) - lam)) <= 1e-10 * lam[0]
    assert np.max(np.abs(np.array(run.POD.retained_energy) - g["retained_energy"])) <= 1e-10
    sgn = _same_basis_up_to_signs(run.basis_functions, g["Z"], tp.inner_product, 1e-6)
    assert np.max(np.abs(run.basis_functions * sgn[None] - g["Z"])) <= 1e-6 * np.max(np.abs(g["Z"]))
    D = np.outer(sgn, sgn)
    assert np.max(np.abs(run.operator_a * D[None] - g["A"])) <= 1e-6 * np.max(np.abs(g["A"]))
    assert np.max(np.abs(run.operator_f * sgn[None] - g["F"])) <= 1e-6 * np.max(np.abs(g["F"]))

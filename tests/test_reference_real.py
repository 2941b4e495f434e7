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

"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE for the path (tests/golden/make_golden.py; the
generator reads /root/reference, the committed .npz fixtures travel).

CPU part (``-m "not gpu"``): pins the oracle -- loop-faithful port must reproduce the reference bit for bit (same
numpy calls in the same order), the vectorised port within 1e-12.
GPU part (``-m gpu``): the CUDA path through the C ABI against the same reference outputs, at the north-star
tolerance (1e-10 relative on solutions and estimators, identical arg-max index)."""
import glob
import os

import numpy as np
import pytest

from oracle import rb_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(os.path.basename(p)[len("reference_sweep_"):-len(".npz")]
               for p in glob.glob(os.path.join(HERE, "golden", "reference_sweep_*.npz")))
SOLVE_CASES = [c for c in CASES if not c.startswith("heavytail")]
RTOL = 1e-10


def _load(case):
    g = dict(np.load(os.path.join(HERE, "golden", f"reference_sweep_{case}.npz")))
    g["n_bc"] = int(g["n_bc"])
    return g


def _lists(g):
    Qa, Qf = g["A"].shape[0], g["F"].shape[0]
    A_q = [g["A"][q] for q in range(Qa)]
    f_q = [g["F"][q] for q in range(Qf)]
    G_ff = [[float(g["Gff"][i, j]) for j in range(Qf)] for i in range(Qf)]
    G_af = [[g["Gaf"][i, j] for j in range(Qf)] for i in range(Qa)]
    G_aa = [[g["Gaa"][i, j] for j in range(Qa)] for i in range(Qa)]
    return A_q, f_q, G_ff, G_af, G_aa


def test_fixtures_present():
    assert {"tutorial01", "n16q6", "n32q10_bc", "n40q7", "heavytail_n16"} <= set(CASES)


# ----------------------------------------------------------------------------------------------------------------------
# CPU: oracle vs the reference's outputs
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", CASES)
def test_oracle_assembly_is_bit_identical_to_reference(case):
    """sum(product(theta_a, operator['a'])) -- rbnics/backends/online/basic/product.py:22-33 incl. zero skipping."""
    g = _load(case)
    A_q = _lists(g)[0]
    for p in range(g["Theta_a"].shape[0]):
        lhs = O.product_order1(tuple(float(x) for x in g["Theta_a"][p]), A_q)
        assert np.array_equal(np.asarray(lhs), g["lhs"][p])


@pytest.mark.parametrize("case", CASES)
def test_oracle_residual_norm_is_bit_identical_to_reference(case):
    """get_residual_norm_squared on the reference's own solutions (elliptic_rb_reduced_problem.py:55-64)."""
    g = _load(case)
    _, _, G_ff, G_af, G_aa = _lists(g)
    N = g["A"].shape[1]
    for p in range(g["Theta_a"].shape[0]):
        e = O.residual_norm_squared(g["U"][p], tuple(float(x) for x in g["Theta_a"][p]),
                                    tuple(float(x) for x in g["Theta_f"][p]), G_ff, G_af, G_aa, N)
        assert e == g["eps2"][p]


@pytest.mark.parametrize("case", SOLVE_CASES)
def test_oracle_sweep_loop_is_bit_identical_to_reference(case):
    g = _load(case)
    A_q, f_q, G_ff, G_af, G_aa = _lists(g)
    N = g["A"].shape[1]
    for kind, key in (("sqrt_eps2_over_beta", "delta_plain"), ("sqrt_of_eps2_over_beta", "delta_compliant")):
        d, U, i, v, e = O.sweep_loop(g["Theta_a"], g["Theta_f"], g["beta"], A_q, f_q, G_ff, G_af, G_aa, N,
                                     Theta_bc=g.get("Theta_bc"), kind=kind, return_eps2=True)
        assert np.array_equal(U, g["U"])
        assert np.array_equal(e, g["eps2"])
        assert np.array_equal(d, g[key])
        assert i == int(np.argmax(g[key]))


@pytest.mark.parametrize("case", SOLVE_CASES)
def test_oracle_sweep_batched_matches_reference(case):
    g = _load(case)
    d, U, i, v, e = O.sweep_batched(g["Theta_a"], g["Theta_f"], g["beta"], g["A"], g["F"], g["Gff"], g["Gaf"],
                                    g["Gaa"], Theta_bc=g.get("Theta_bc"), return_eps2=True)
    assert np.max(np.abs(U - g["U"])) <= 1e-12 * np.max(np.abs(g["U"]))
    assert np.max(np.abs(e - g["eps2"]) / np.abs(g["eps2"])) <= 1e-11
    assert np.max(np.abs(d - g["delta_plain"]) / g["delta_plain"]) <= 1e-11
    assert i == int(np.argmax(g["delta_plain"]))


@pytest.mark.parametrize("case", SOLVE_CASES)
def test_oracle_max_rules_match_reference(case):
    """ParameterSpaceSubset.max serial (lowest index on an exact tie) and under 1..3 MPI ranks (cyclic shards +
    parallel_max: highest rank wins) -- parameter_space_subset.py:52-77, parallel_max.py:11-35."""
    g = _load(case)
    values = g["max_values"]
    B = len(values)
    v_ref, i_ref = g["max_serial"]
    i = O.argmax_first(values)
    assert (values[i], i) == (v_ref, int(i_ref))
    from rbnics_b200.sweep import maxloc_serial_rule
    for size in (1, 2, 3):
        loc = [O.training_set_max(lambda mu: values[mu], list(range(B)), rank=r, size=size) for r in range(size)]
        v, gi = O.parallel_max([x[0] for x in loc], [x[1] for x in loc])
        assert (v, gi) == (g[f"max_mpi_{size}"][0], int(g[f"max_mpi_{size}"][1]))
        # the product's multi-GPU rule is the SERIAL one whatever the number of ranks
        assert maxloc_serial_rule([x[0] for x in loc], [x[1] for x in loc]) == (v_ref, int(i_ref))


def test_heavytail_order2_against_reference_kat():
    """The reference's own known-answer formula (test_numpy_greedy_prototype.py:72-84) on its own input
    distribution; eps2 of the golden file obeys it with the factor 2 on the af term."""
    g = _load("heavytail_n16")
    for p in range(g["Theta_a"].shape[0]):
        u, ta, tf = g["U"][p], g["Theta_a"][p], g["Theta_f"][p]
        kat = (np.einsum("n,i,ijnm,j,m", u, ta, g["Gaa"], ta, u, optimize=True)
               + 2.0 * np.einsum("i,ijn,j,n", ta, g["Gaf"], tf, u, optimize=True)
               + np.einsum("i,ij,j", tf, g["Gff"], tf, optimize=True))
        assert abs(kat - g["eps2"][p]) <= 1e-10 * abs(g["eps2"][p])


# ----------------------------------------------------------------------------------------------------------------------
# GPU: the CUDA path vs the reference's outputs
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", SOLVE_CASES)
def test_cuda_sweep_matches_reference_golden(engine, case):
    from rbnics_b200._lib import RB_EST_SQRT_EPS2_OVER_BETA, RB_EST_SQRT_OF_EPS2_OVER_BETA
    g = _load(case)
    Theta = np.concatenate([g["Theta_a"], g["Theta_f"]], axis=1)
    nbc = g["n_bc"]
    engine.set_system(g["A"], g["F"], bc_rows=list(range(nbc)) if nbc else None)
    engine.set_elliptic_estimator(g["Gff"], g["Gaf"], g["Gaa"])
    engine.set_training_set(Theta, g["beta"], g.get("Theta_bc"))
    for kind, key in ((RB_EST_SQRT_EPS2_OVER_BETA, "delta_plain"), (RB_EST_SQRT_OF_EPS2_OVER_BETA, "delta_compliant")):
        r = engine.sweep_resident(kind=kind, want_delta=True, want_u=True, want_eps2=True)
        err_u = np.max(np.abs(r["u"] - g["U"]), axis=1) / np.max(np.abs(g["U"]), axis=1)
        assert err_u.max() <= RTOL
        assert np.max(np.abs(r["eps2"] - g["eps2"]) / np.abs(g["eps2"])) <= RTOL
        assert np.max(np.abs(r["delta"] - g[key]) / g[key]) <= RTOL
        assert r["argmax"] == int(np.argmax(g[key]))
        assert abs(r["max"] - g[key].max()) <= RTOL * g[key].max()


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_single_parameter_products_are_bit_identical_to_reference(engine, case):
    """rb_affine_combine / rb_affine_combine2 keep the reference's order AND rounding: bitwise equal lhs, and the
    order-2 products recombined as in get_residual_norm_squared give the reference's eps2 to 1e-13."""
    from rbnics_b200 import online
    g = _load(case)
    Qa, N = g["A"].shape[0], g["A"].shape[1]
    Qf = g["F"].shape[0]
    for p in range(min(6, g["Theta_a"].shape[0])):
        ta, tf = g["Theta_a"][p], g["Theta_f"][p]
        lhs = online.affine_combine(engine, ta, g["A"])
        assert np.array_equal(lhs, g["lhs"][p])
        aa = online.affine_combine2(engine, ta, ta, g["Gaa"])
        af = online.affine_combine2(engine, ta, tf, g["Gaf"])
        ff = online.affine_combine2(engine, tf, tf, g["Gff"].reshape(Qf, Qf, 1))[0]
        u = g["U"][p]
        eps2 = ff + 2.0 * online.bilinear(engine, u, None, af) + online.bilinear(engine, u, aa, u)
        assert abs(eps2 - g["eps2"][p]) <= 1e-13 * max(abs(g["eps2"][p]), abs(ff))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["tutorial01", "n32q10_bc"])
def test_cuda_per_parameter_online_api_matches_reference_golden(engine, case):
    """The reference's per-mu call sequence (matrix_eval / vector_eval / LinearSolver / get_residual_norm_squared,
    elliptic_reduced_problem.py:19-28, elliptic_rb_reduced_problem.py:55-64) written against rbnics_b200.online."""
    from rbnics_b200 import online
    online.set_engine(engine)
    g = _load(case)
    Qa, N = g["A"].shape[0], g["A"].shape[1]
    Qf = g["F"].shape[0]
    op_a = online.AffineExpansionStorage(tuple(online.Matrix(g["A"][q]) for q in range(Qa)))
    op_f = online.AffineExpansionStorage(tuple(online.Vector(g["F"][q]) for q in range(Qf)))
    e_ff = online.AffineExpansionStorage(Qf, Qf)
    e_af = online.AffineExpansionStorage(Qa, Qf)
    e_aa = online.AffineExpansionStorage(Qa, Qa)
    for i in range(Qf):
        for j in range(Qf):
            e_ff[i, j] = float(g["Gff"][i, j])
    for i in range(Qa):
        for j in range(Qf):
            e_af[i, j] = g["Gaf"][i, j]
        for j in range(Qa):
            e_aa[i, j] = g["Gaa"][i, j]
    for p in range(min(8, g["Theta_a"].shape[0])):
        ta = tuple(float(x) for x in g["Theta_a"][p])
        tf = tuple(float(x) for x in g["Theta_f"][p])
        bcs = tuple(float(x) for x in g["Theta_bc"][p]) if g["n_bc"] else None
        lhs = online.sum(online.product(ta, op_a[:N, :N]))
        rhs = online.sum(online.product(tf, op_f[:N]))
        assert np.array_equal(lhs, g["lhs"][p])
        solution = online.Function(N)
        solver = online.LinearSolver(lhs, solution, rhs, bcs)
        solver.set_parameters({})
        solver.solve()
        u = solution.vector()
        assert np.max(np.abs(u - g["U"][p])) <= RTOL * np.max(np.abs(g["U"][p]))
        eps2 = (online.sum(online.product(tf, e_ff, tf))
                + 2.0 * (online.transpose(u) * online.sum(online.product(ta, e_af[:N], tf)))
                + online.transpose(u) * online.sum(online.product(ta, e_aa[:N, :N], ta)) * u)
        assert abs(eps2 - g["eps2"][p]) <= RTOL * abs(g["eps2"][p])


# ----------------------------------------------------------------------------------------------------------------------
# Stokes (BASELINE.json configs[4]): block saddle-point system, 9 estimator term pairs
# (rbnics/problems/stokes/stokes_reduced_problem.py:38-53, stokes_rb_reduced_problem.py:35-83)
# ----------------------------------------------------------------------------------------------------------------------
STOKES_CASES = sorted(os.path.basename(p)[len("reference_stokes_"):-len(".npz")]
                      for p in glob.glob(os.path.join(HERE, "golden", "reference_stokes_*.npz")))
STOKES_PAIRS = [("f", "f"), ("g", "g"), ("a", "f"), ("bt", "f"), ("b", "g"), ("a", "a"), ("a", "bt"), ("bt", "bt"),
                ("b", "b")]


def _load_stokes(case):
    g = dict(np.load(os.path.join(HERE, "golden", f"reference_stokes_{case}.npz")))
    op = {t: g["op_" + t] for t in ("a", "b", "bt", "f", "g")}
    E = {(t1, t2): g[f"E_{t1}_{t2}"] for (t1, t2) in STOKES_PAIRS}
    theta = {t: g["theta_" + t] for t in ("a", "b", "bt", "f", "g")}
    return g, op, E, theta


def test_stokes_fixtures_present():
    assert {"n5", "n12_bc"} <= set(STOKES_CASES)


@pytest.mark.parametrize("case", STOKES_CASES)
def test_oracle_stokes_is_bit_identical_to_reference(case):
    g, op, E, theta = _load_stokes(case)
    N = op["a"].shape[1]
    Eo = {k: ([[float(v[i, j]) for j in range(v.shape[1])] for i in range(v.shape[0])] if v.ndim == 2 else
              [[v[i, j] for j in range(v.shape[1])] for i in range(v.shape[0])]) for k, v in E.items()}
    opl = {t: [op[t][q] for q in range(op[t].shape[0])] for t in op}
    for p in range(g["U"].shape[0]):
        th = {t: tuple(float(x) for x in theta[t][p]) for t in theta}
        tbc = tuple(float(x) for x in g["Theta_bc"][p]) if int(g["n_bc"]) else None
        u = O.reduced_solve_stokes(th, opl, N, tbc)
        assert np.array_equal(u, g["U"][p])
        assert O.residual_norm_squared_stokes(u, th, Eo, N) == g["eps2"][p]


@pytest.mark.parametrize("case", STOKES_CASES)
def test_stokes_quadratic_form_packing_matches_reference(case):
    """z^T G z with the packed G of engine.stokes_quadratic_form == the reference's 9-term eps2 (CPU, numpy only)."""
    from rbnics_b200.engine import stokes_quadratic_form
    g, op, E, theta = _load_stokes(case)
    N = op["a"].shape[1]
    Q = {t: op[t].shape[0] for t in op}
    G, (scale, src, length, v_off, v_len) = stokes_quadratic_form(E, Q, N)
    for p in range(g["U"].shape[0]):
        Theta = np.concatenate([theta[t][p] for t in ("a", "b", "bt", "f", "g")])
        V = np.concatenate([g["U"][p], Theta[v_off:v_off + v_len], [1.0]])
        z = np.concatenate([(Theta[sc] if sc >= 0 else 1.0) * V[so:so + ln] for sc, so, ln in zip(scale, src, length)])
        assert abs(z @ G @ z - g["eps2"][p]) <= 1e-12 * abs(g["eps2"][p])


@pytest.mark.gpu
@pytest.mark.parametrize("case", STOKES_CASES)
def test_cuda_stokes_sweep_matches_reference_golden(engine, case):
    from rbnics_b200._lib import RB_EST_SQRT_OF_EPS2_OVER_BETA
    g, op, E, theta = _load_stokes(case)
    N = op["a"].shape[1]
    Q = {t: op[t].shape[0] for t in op}
    nbc = int(g["n_bc"])
    A = np.concatenate([op["a"], op["b"], op["bt"]])            # every lhs term block-embedded in N x N
    F = np.concatenate([op["f"], op["g"]])
    Theta = np.concatenate([theta[t] for t in ("a", "b", "bt", "f", "g")], axis=1)
    engine.set_system(A, F, bc_rows=list(range(nbc)) if nbc else None)
    engine.set_stokes_estimator(E, Q)
    engine.set_training_set(Theta, g["beta"], g.get("Theta_bc"))
    r = engine.sweep_resident(kind=RB_EST_SQRT_OF_EPS2_OVER_BETA, want_delta=True, want_u=True, want_eps2=True)
    err_u = np.max(np.abs(r["u"] - g["U"]), axis=1) / np.max(np.abs(g["U"]), axis=1)
    assert err_u.max() <= RTOL
    assert np.max(np.abs(r["eps2"] - g["eps2"]) / np.abs(g["eps2"])) <= RTOL
    assert np.max(np.abs(r["delta"] - g["delta"]) / g["delta"]) <= RTOL
    assert r["argmax"] == int(np.argmax(g["delta"]))


# ----------------------------------------------------------------------------------------------------------------------
# Reduced supremizer solve: golden vectors from the reference's own _solve_supremizer source
# (tests/golden/make_golden.py:make_supremizer_golden)
# ----------------------------------------------------------------------------------------------------------------------
def _load_supremizer():
    return dict(np.load(os.path.join(HERE, "golden", "reference_supremizer.npz")))


def test_oracle_supremizer_is_bit_identical_to_reference():
    g = _load_supremizer()
    N_us = int(g["N_us"])
    Bt = [g["Bt"][q] for q in range(g["Bt"].shape[0])]
    for p in range(g["U"].shape[0]):
        s = O.supremizer_solve(g["Xs"], Bt, tuple(g["Theta"][p]), g["U"][p], N_us)
        assert np.array_equal(s, g["S"][p])


@pytest.mark.gpu
def test_cuda_supremizers_match_reference_golden(engine):
    g = _load_supremizer()
    S = engine.solve_supremizers(g["Xs"], g["Bt"], g["Theta"], g["U"])
    err = np.max(np.abs(S - g["S"]), axis=1) / np.max(np.abs(g["S"]), axis=1)
    assert err.max() <= RTOL


@pytest.mark.gpu
def test_cuda_supremizers_from_resident_sweep_solutions(engine):
    """U = None: the supremizers are computed from the solutions the Stokes sweep left in HBM; ragged batch, more
    supremizer entries than the 128 threads of a CTA."""
    g, op, E, theta = _load_stokes("n12_bc")
    N = op["a"].shape[1]
    Q = {t: op[t].shape[0] for t in op}
    nbc = int(g["n_bc"])
    A = np.concatenate([op["a"], op["b"], op["bt"]])
    F = np.concatenate([op["f"], op["g"]])
    Theta = np.concatenate([theta[t] for t in ("a", "b", "bt", "f", "g")], axis=1)
    engine.set_system(A, F, bc_rows=list(range(nbc)) if nbc else None)
    engine.set_stokes_estimator(E, Q)
    engine.set_training_set(Theta, g["beta"], g.get("Theta_bc"))
    r = engine.sweep_resident(want_u=True)
    rng = np.random.default_rng(5)
    N_us, Qbt = 150, 3
    Rm = rng.standard_normal((N_us + 3, N_us))
    Xs = Rm.T @ Rm / N_us + np.eye(N_us)
    Bt = rng.standard_normal((Qbt, N_us, N))
    Th = rng.uniform(-1.0, 1.0, (Theta.shape[0], Qbt))
    S = engine.solve_supremizers(Xs, Bt, Th)
    for p in range(Theta.shape[0]):
        ref = O.supremizer_solve(Xs, [Bt[q] for q in range(Qbt)], tuple(Th[p]), r["u"][p], N_us)
        assert np.max(np.abs(S[p] - ref)) <= RTOL * np.max(np.abs(ref))

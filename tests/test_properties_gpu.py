"""Size-independent properties of the sweep at BASELINE.json's full shapes (N=100, Q_a=50, Q_f=10), where the loop
oracle is too slow to run the whole batch: linearity in theta_f, permutation equivariance, shard union == whole,
bit-reproducibility, and an oracle spot check on a random subsample."""
import numpy as np
import pytest

from oracle import rb_oracle as O
from rbnics_b200 import synthetic

pytestmark = pytest.mark.gpu
N, QA, QF, B = 100, 50, 10, 40_000


@pytest.fixture(scope="module")
def setup(engine):
    prob = synthetic.elliptic_reduced_problem(N, QA, QF, seed=0, riesz_rank=512)
    Theta, beta, _ = synthetic.training_set(B, QA, QF, seed=3)
    engine.set_system(prob["A"], prob["F"])
    engine.set_elliptic_estimator(prob["Gff"], prob["Gaf"], prob["Gaa"])
    engine.set_training_set(Theta, beta)
    base = engine.sweep_resident(want_delta=True, want_u=True, want_eps2=True)
    return prob, Theta, beta, base


def test_spot_check_against_oracle(setup):
    prob, Theta, beta, base = setup
    idx = np.random.default_rng(0).choice(B, 256, replace=False)
    d, U, _, _, e = O.sweep_batched(Theta[idx, :QA], Theta[idx, QA:], beta[idx], prob["A"], prob["F"], prob["Gff"],
                                    prob["Gaf"], prob["Gaa"], return_eps2=True)
    assert np.max(np.abs(base["u"][idx] - U) / np.max(np.abs(U), axis=1, keepdims=True)) <= 1e-10
    assert np.max(np.abs(base["eps2"][idx] - e) / np.abs(e)) <= 1e-10
    assert np.max(np.abs(base["delta"][idx] - d) / d) <= 1e-10


def test_argmax_is_first_maximum_of_the_deltas(setup):
    _, _, _, base = setup
    assert base["argmax"] == int(np.argmax(base["delta"]))
    assert base["max"] == base["delta"].max()


def test_bit_reproducible(engine, setup):
    _, _, _, base = setup
    again = engine.sweep_resident(want_delta=True, want_u=True)
    assert np.array_equal(again["delta"], base["delta"]) and np.array_equal(again["u"], base["u"])
    assert again["argmax"] == base["argmax"]


def test_linearity_in_theta_f(engine, setup):
    """u is linear and eps2 quadratic in theta_f: scaling by 2 (exact in binary FP) scales them by exactly 2 and 4."""
    _, Theta, beta, base = setup
    T2 = Theta.copy()
    T2[:, QA:] *= 2.0
    engine.set_training_set(T2, beta)
    r = engine.sweep_resident(want_u=True, want_eps2=True)
    assert np.array_equal(r["u"], 2.0 * base["u"])
    assert np.array_equal(r["eps2"], 4.0 * base["eps2"])
    engine.set_training_set(Theta, beta)


def test_permutation_equivariance_and_shard_union(engine, setup):
    _, Theta, beta, base = setup
    perm = np.random.default_rng(1).permutation(B)
    engine.set_training_set(Theta[perm], beta[perm])
    r = engine.sweep_resident(want_delta=True)
    assert np.array_equal(r["delta"], base["delta"][perm])          # per-parameter results do not depend on the batch
    assert perm[r["argmax"]] == base["argmax"] or base["delta"][perm[r["argmax"]]] == base["max"]
    best = None
    for rank in range(4):                                            # cyclic shards of the reference (rank::size)
        engine.set_training_set(Theta[rank::4], beta[rank::4], index_offset=rank, index_stride=4)
        s = engine.sweep_resident(want_delta=True)
        # the column split of the estimator GEMM depends on the shard size, so the partial sums are grouped
        # differently: equal to rounding, not bitwise
        assert np.max(np.abs(s["delta"] - base["delta"][rank::4]) / base["delta"][rank::4]) <= 1e-13
        cand = (s["max"], -s["argmax"])
        best = cand if best is None or cand > best else best
    assert -best[1] == base["argmax"] and abs(best[0] - base["max"]) <= 1e-13 * base["max"]
    engine.set_training_set(Theta, beta)


def test_config2_elastic_block_shapes_at_1e5(engine):
    """BASELINE.json configs[1]: Q_a = 9, Q_f = 3 (tutorials/02_elastic_block), N = 20, training set scaled to 10^5:
    spot check against the oracle + arg-max consistency."""
    n, qa, qf, b = 20, 9, 3, 100_000
    prob = synthetic.elliptic_reduced_problem(n, qa, qf, seed=2, riesz_rank=128)
    Theta, beta, _ = synthetic.training_set(b, qa, qf, seed=4)
    engine.set_system(prob["A"], prob["F"])
    engine.set_elliptic_estimator(prob["Gff"], prob["Gaf"], prob["Gaa"])
    engine.set_training_set(Theta, beta)
    r = engine.sweep_resident(want_delta=True, want_u=True)
    idx = np.random.default_rng(0).choice(b, 512, replace=False)
    d, U, _, _ = O.sweep_batched(Theta[idx, :qa], Theta[idx, qa:], beta[idx], prob["A"], prob["F"], prob["Gff"],
                                 prob["Gaf"], prob["Gaa"])
    assert np.max(np.abs(r["u"][idx] - U) / np.max(np.abs(U), axis=1, keepdims=True)) <= 1e-10
    assert np.max(np.abs(r["delta"][idx] - d) / d) <= 1e-10
    assert r["argmax"] == int(np.argmax(r["delta"]))


def test_config5_stokes_shapes(engine):
    """BASELINE.json configs[4]: block system (u, s, p) with 25 functions per component (N = 75), 9 estimator term
    pairs, beta == 1: against numpy.linalg.solve + the packed quadratic form (pinned to the reference's 9-term
    formula on the CPU in tests/test_golden.py)."""
    from rbnics_b200._lib import RB_EST_SQRT_OF_EPS2_OVER_BETA
    from rbnics_b200.engine import stokes_quadratic_form
    n, b = 25, 4096
    Q = {"a": 6, "b": 4, "bt": 4, "f": 3, "g": 2}
    op, E = synthetic.stokes_reduced_problem(n, Q, seed=1)
    nn = 3 * n
    rng = np.random.default_rng(5)
    th = {t: rng.uniform(0.5, 2.0, (b, Q[t])) for t in ("a", "b", "bt")}
    th["f"] = rng.uniform(-1, 1, (b, Q["f"]))
    th["g"] = rng.uniform(-1, 1, (b, Q["g"]))
    Theta = np.concatenate([th[t] for t in ("a", "b", "bt", "f", "g")], axis=1)
    engine.set_system(np.concatenate([op["a"], op["b"], op["bt"]]), np.concatenate([op["f"], op["g"]]))
    engine.set_stokes_estimator(E, Q)
    engine.set_training_set(Theta, np.ones(b))
    r = engine.sweep_resident(kind=RB_EST_SQRT_OF_EPS2_OVER_BETA, want_delta=True, want_u=True, want_eps2=True)
    lhs = (np.einsum("pq,qij->pij", th["a"], op["a"]) + np.einsum("pq,qij->pij", th["b"], op["b"])
           + np.einsum("pq,qij->pij", th["bt"], op["bt"]))
    rhs = th["f"] @ op["f"] + th["g"] @ op["g"]
    U = np.linalg.solve(lhs, rhs[..., None])[..., 0]
    assert np.max(np.abs(r["u"] - U) / np.max(np.abs(U), axis=1, keepdims=True)) <= 1e-10
    G, (scale, src, length, v_off, v_len) = stokes_quadratic_form(E, Q, nn)
    V = np.concatenate([U, Theta[:, v_off:v_off + v_len], np.ones((b, 1))], axis=1)
    Z = np.concatenate([(Theta[:, sc:sc + 1] if sc >= 0 else 1.0) * V[:, so:so + ln]
                        for sc, so, ln in zip(scale, src, length)], axis=1)
    eps2 = np.einsum("pk,pk->p", Z @ G, Z)
    assert np.max(np.abs(r["eps2"] - eps2) / np.abs(eps2)) <= 1e-10
    assert r["argmax"] == int(np.argmax(np.sqrt(np.abs(eps2))))

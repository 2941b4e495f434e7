"""POD-Greedy sweep of parabolic reduced problems against outputs of the reference's own source
(tests/golden/reference_parabolic.npz; generator: tests/golden/make_golden.py:make_parabolic_golden).
CPU: the oracle restatement; GPU: rbnics_b200.parabolic through the C ABI."""
import os

import numpy as np
import pytest

from oracle import rb_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = dict(np.load(os.path.join(HERE, "golden", "reference_parabolic.npz")))
K, DT = int(G["K"]), float(G["dt"])
N = G["A"].shape[1]


def _lists():
    def storage(blk):
        if blk.ndim == 2:
            return [[float(blk[i, j]) for j in range(blk.shape[1])] for i in range(blk.shape[0])]
        return [[blk[i, j] for j in range(blk.shape[1])] for i in range(blk.shape[0])]
    Gd = {"ff": storage(G["E_f_f"]), "af": storage(G["E_a_f"]), "aa": storage(G["E_a_a"]), "mf": storage(G["E_m_f"]),
          "ma": storage(G["E_m_a"]), "mm": storage(G["E_m_m"])}
    return ([G["M"][q] for q in range(G["M"].shape[0])], [G["A"][q] for q in range(G["A"].shape[0])],
            [G["F"][q] for q in range(G["F"].shape[0])], Gd)


def test_oracle_parabolic_sweep_matches_reference():
    M_q, A_q, f_q, Gd = _lists()
    d, i, eps2, U = O.parabolic_sweep_loop(G["Theta_m"], G["Theta_a"], G["Theta_f"], G["beta"], M_q, A_q, f_q, Gd, N,
                                           DT, K)
    assert np.array_equal(U, G["U"])                       # same numpy calls in the same order
    assert np.array_equal(eps2, G["eps2"])
    assert np.max(np.abs(d - G["delta"]) / G["delta"]) <= 1e-15
    assert i == int(np.argmax(G["delta"]))


def test_parabolic_quadratic_form_and_simpson_weights_match_reference():
    """CPU, numpy only: z^T G z of the packed form == the reference's eps2 at every step; the Simpson functional."""
    from rbnics_b200.parabolic import parabolic_quadratic_form, simpson_weights
    E = {("f", "f"): G["E_f_f"], ("a", "f"): G["E_a_f"], ("a", "a"): G["E_a_a"], ("m", "f"): G["E_m_f"],
         ("m", "a"): G["E_m_a"], ("m", "m"): G["E_m_m"]}
    Qm, Qa, Qf = G["M"].shape[0], G["A"].shape[0], G["F"].shape[0]
    Gq, (scale, src, length, v_off, v_len) = parabolic_quadratic_form(E, Qa, Qm, Qf, N)
    w = simpson_weights(K, DT)
    for p in range(G["U"].shape[0]):
        Theta = np.concatenate([G["Theta_m"][p], G["Theta_a"][p], G["Theta_f"][p]])
        for k in range(1, K + 1):
            u, udot = G["U"][p, k], (G["U"][p, k] - G["U"][p, k - 1]) / DT
            V = np.concatenate([u, udot, Theta[v_off:v_off + v_len], [1.0]])
            z = np.concatenate([(Theta[sc] if sc >= 0 else 1.0) * V[so:so + ln] for sc, so, ln in zip(scale, src, length)])
            assert abs(z @ Gq @ z - G["eps2"][p, k]) <= 1e-11 * np.max(np.abs(G["eps2"][p]))
        assert abs(np.sqrt(np.dot(w, G["bound"][p] ** 2)) - G["delta"][p]) <= 1e-14 * G["delta"][p]


@pytest.fixture(params=["fused", "per_step_launches"])
def path(request, monkeypatch):
    """Both device paths of rb_sweep_parabolic: the one-kernel time loop (rb_parabolic.cu; NP <= 32, Kz <= 128) and the
    per-step launches that serve every other size."""
    monkeypatch.setenv("RB_PARABOLIC_FUSED", "1" if request.param == "fused" else "0")
    return request.param


@pytest.mark.gpu
def test_cuda_parabolic_sweep_matches_reference(engine, path):
    from rbnics_b200.parabolic import ParabolicSweep
    E = {("f", "f"): G["E_f_f"], ("a", "f"): G["E_a_f"], ("a", "a"): G["E_a_a"], ("m", "f"): G["E_m_f"],
         ("m", "a"): G["E_m_a"], ("m", "m"): G["E_m_m"]}
    sw = ParabolicSweep(engine, G["M"], G["A"], G["F"], E, dt=DT, K=K)
    r = sw.sweep(G["Theta_m"], G["Theta_a"], G["Theta_f"], G["beta"], want_delta=True, want_eps2=True, want_u=True)
    assert np.max(np.abs(r["u"] - G["U"])) <= 1e-10 * np.max(np.abs(G["U"]))
    scale = np.max(np.abs(G["eps2"]), axis=1, keepdims=True)
    assert np.max(np.abs(r["eps2"] - G["eps2"]) / scale) <= 1e-10
    assert np.max(np.abs(r["delta"] - G["delta"]) / G["delta"]) <= 1e-10
    assert r["argmax"] == int(np.argmax(G["delta"]))
    assert abs(r["max"] - G["delta"].max()) <= 1e-10 * G["delta"].max()


@pytest.mark.gpu
def test_cuda_parabolic_sweep_with_initial_condition(engine, path):
    """Non-homogeneous initial condition (SURVEY.md 8f row 2 / VERDICT r1): u^0 of every parameter and the squared
    error estimate at t0 enter the sweep; compared with the loop oracle started from the same u^0
    (rbnics/backends/online/numpy/time_stepping.py:104-112, abstract_parabolic_rb_reduced_problem.py:29-46)."""
    from rbnics_b200.parabolic import ParabolicSweep
    E = {("f", "f"): G["E_f_f"], ("a", "f"): G["E_a_f"], ("a", "a"): G["E_a_a"], ("m", "f"): G["E_m_f"],
         ("m", "a"): G["E_m_a"], ("m", "m"): G["E_m_m"]}
    B = G["Theta_m"].shape[0]
    rng = np.random.default_rng(4)
    U0 = 0.3 * rng.standard_normal((B, N))
    e0 = rng.uniform(0.0, 0.2, B)
    Gd = {"ff": [[float(x) for x in row] for row in G["E_f_f"]],
          "af": [[v for v in row] for row in G["E_a_f"]], "aa": [[v for v in row] for row in G["E_a_a"]],
          "mf": [[v for v in row] for row in G["E_m_f"]], "ma": [[v for v in row] for row in G["E_m_a"]],
          "mm": [[v for v in row] for row in G["E_m_m"]]}
    d, i, eps2, U = O.parabolic_sweep_loop(G["Theta_m"], G["Theta_a"], G["Theta_f"], G["beta"], list(G["M"]), list(G["A"]),
                                           list(G["F"]), Gd, N, DT, K, U0=U0, init_err2=e0)
    sw = ParabolicSweep(engine, G["M"], G["A"], G["F"], E, dt=DT, K=K)
    r = sw.sweep(G["Theta_m"], G["Theta_a"], G["Theta_f"], G["beta"], want_delta=True, want_eps2=True, want_u=True,
                 U0=U0, initial_error_estimate_squared=e0)
    assert np.array_equal(r["u"][:, 0], U0)
    assert np.max(np.abs(r["u"] - U)) <= 1e-10 * np.max(np.abs(U))
    scale = np.max(np.abs(eps2), axis=1, keepdims=True)
    assert np.max(np.abs(r["eps2"][:, 1:] - eps2[:, 1:]) / scale) <= 1e-10
    assert np.max(np.abs(r["delta"] - d) / d) <= 1e-10
    assert r["argmax"] == i
    zero = sw.sweep(G["Theta_m"], G["Theta_a"], G["Theta_f"], G["beta"], want_delta=True)   # and the default is unchanged
    assert np.max(np.abs(zero["delta"] - G["delta"]) / G["delta"]) <= 1e-10


def _random_parabolic(N, Qm, Qa, Qf, B, seed):
    rng = np.random.default_rng(seed)
    spd = lambda: (lambda R: R @ R.T / N + np.eye(N))(rng.standard_normal((N, N)))  # noqa: E731
    M = np.stack([spd() for _ in range(Qm)])
    A = np.stack([spd() + 0.3 * rng.standard_normal((N, N)) for _ in range(Qa)])   # non-symmetric: pivoting matters
    F = rng.standard_normal((Qf, N))
    Kz = (Qa + Qm) * N + Qf
    R = rng.standard_normal((Kz + 5, Kz))
    Gz = R.T @ R / R.shape[0]
    off = {"a": 0, "m": Qa * N, "f": (Qa + Qm) * N}
    Q = {"a": Qa, "m": Qm, "f": Qf}

    def block(t1, t2):
        o1, o2 = t1 != "f", t2 != "f"
        n1, n2 = Q[t1] * (N if o1 else 1), Q[t2] * (N if o2 else 1)
        Mb = Gz[off[t1]:off[t1] + n1, off[t2]:off[t2] + n2]
        if o1 and o2:
            return Mb.reshape(Q[t1], N, Q[t2], N).transpose(0, 2, 1, 3).copy()
        if o1:
            return Mb.reshape(Q[t1], N, Q[t2]).transpose(0, 2, 1).copy()
        return Mb.copy()

    E = {k: block(*k) for k in [("f", "f"), ("a", "f"), ("a", "a"), ("m", "f"), ("m", "a"), ("m", "m")]}
    Tm = rng.uniform(0.5, 2.0, (B, Qm))
    Ta = rng.uniform(0.1, 10.0, (B, Qa))
    Tf = rng.uniform(-1.0, 1.0, (B, Qf))
    return M, A, F, E, Tm, Ta, Tf, Ta.min(axis=1)


def _storages(E):
    st = lambda blk: [[(float(blk[i, j]) if blk.ndim == 2 else blk[i, j]) for j in range(blk.shape[1])]  # noqa: E731
                      for i in range(blk.shape[0])]
    return {a + b: st(E[a, b]) for (a, b) in E}


@pytest.mark.gpu
@pytest.mark.parametrize("N,Qm,Qa,Qf,B", [(3, 1, 2, 1, 5), (7, 2, 2, 3, 17), (16, 1, 2, 1, 8), (20, 1, 2, 1, 33),
                                          (21, 1, 2, 1, 9), (30, 1, 1, 1, 12), (20, 1, 3, 1, 19), (24, 2, 3, 2, 10),
                                          (32, 1, 2, 1, 6), (32, 2, 2, 1, 7)])
def test_cuda_parabolic_sweep_sizes(engine, path, N, Qm, Qa, Qf, B):
    """Every instantiation of the fused kernel (NP = 4 ... 32; z rows of 64 and of 128 entries: Kz = 81, 97, 122),
    partial last CTAs, and a size that only the per-step path serves (N = 32 with Kz = 129), against the loop oracle."""
    from rbnics_b200.parabolic import ParabolicSweep
    K, dt = 9, 0.07
    M, A, F, E, Tm, Ta, Tf, beta = _random_parabolic(N, Qm, Qa, Qf, B, seed=N)
    d, i, eps2, U = O.parabolic_sweep_loop(Tm, Ta, Tf, beta, list(M), list(A), list(F), _storages(E), N, dt, K)
    r = ParabolicSweep(engine, M, A, F, E, dt=dt, K=K).sweep(Tm, Ta, Tf, beta, want_delta=True, want_eps2=True, want_u=True)
    assert np.max(np.abs(r["u"] - U)) <= 1e-10 * np.max(np.abs(U))
    scale = np.max(np.abs(eps2), axis=1, keepdims=True)
    assert np.max(np.abs(r["eps2"] - eps2) / scale) <= 1e-10
    assert np.max(np.abs(r["delta"] - d) / d) <= 1e-10
    assert r["argmax"] == i


@pytest.mark.gpu
def test_cuda_parabolic_fused_equals_per_step_path_with_lifting_rows(engine, monkeypatch):
    """Lifting rows (DirichletBC.py:37-56: unit row, rhs = theta_bc at every step) through both device paths."""
    from rbnics_b200.parabolic import ParabolicSweep
    N, K, dt, B = 14, 12, 0.05, 21
    M, A, F, E, Tm, Ta, Tf, beta = _random_parabolic(N, 1, 2, 1, B, seed=99)
    ThetaBC = np.random.default_rng(5).uniform(0.5, 1.5, (B, 2))
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("RB_PARABOLIC_FUSED", flag)
        sw = ParabolicSweep(engine, M, A, F, E, dt=dt, K=K, bc_rows=[0, 9])
        out[flag] = sw.sweep(Tm, Ta, Tf, beta, ThetaBC=ThetaBC, want_delta=True, want_eps2=True, want_u=True)
    # (a lifting row is a unit row, but partial pivoting may still eliminate it with a larger entry of its column)
    assert np.allclose(out["1"]["u"][:, 1:, 0], np.repeat(ThetaBC[:, :1], K, axis=1), rtol=1e-13, atol=0)
    assert np.allclose(out["1"]["u"][:, 1:, 9], np.repeat(ThetaBC[:, 1:], K, axis=1), rtol=1e-13, atol=0)
    assert np.max(np.abs(out["1"]["u"] - out["0"]["u"])) <= 1e-12 * np.max(np.abs(out["0"]["u"]))
    scale = np.max(np.abs(out["0"]["eps2"]), axis=1, keepdims=True)
    assert np.max(np.abs(out["1"]["eps2"] - out["0"]["eps2"]) / scale) <= 1e-11
    assert np.max(np.abs(out["1"]["delta"] - out["0"]["delta"]) / out["0"]["delta"]) <= 1e-11
    assert out["1"]["argmax"] == out["0"]["argmax"]

"""Randomised block tables for the generic quadratic-form estimator (rb_set_estimator): arbitrary numbers of blocks
per slice class, block lengths that are not multiples of 4 / 16, partial slices of u, several theta blocks, non-symmetric
G.  Checks eps2 = z^T G z against numpy: exercises the folded tiling (16-column ranges, leftover tiles, odd group
sizes, classes with a single member) far from the benchmark shape."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(rng):
    N = int(rng.integers(1, 72))
    n_theta_l = int(rng.integers(1, 8))            # lhs terms (scale columns available)
    n_theta_r = int(rng.integers(1, 6))            # rhs terms = v_theta candidates
    v_len = int(rng.integers(0, n_theta_r + 1))
    blocks = []                                    # (scale col or -1, src, len)
    for _ in range(int(rng.integers(1, 10))):      # full-u blocks: one class
        blocks.append((int(rng.integers(0, n_theta_l)), 0, N))
    if N > 3 and rng.random() < 0.6:               # blocks on a partial slice of u: another class
        s0 = int(rng.integers(0, N - 1))
        ln = int(rng.integers(1, N - s0 + 1))
        for _ in range(int(rng.integers(1, 4))):
            blocks.append((int(rng.integers(0, n_theta_l)), s0, ln))
    if v_len > 0:
        blocks.append((-1, N, v_len))              # [theta_f]
        if rng.random() < 0.5:
            blocks.append((int(rng.integers(0, n_theta_l)), N, v_len))   # scaled theta block, same slice: foldable
    if rng.random() < 0.5:
        blocks.append((-1, N + v_len, 1))          # the constant 1
    order = rng.permutation(len(blocks))
    return N, n_theta_l, n_theta_r, v_len, [blocks[i] for i in order]


@pytest.mark.parametrize("seed", range(24))
def test_random_block_tables(engine, seed):
    rng = np.random.default_rng(1000 + seed)
    N, Ql, Qr, v_len, blocks = _case(rng)
    B = int(rng.integers(1, 400))
    A = np.stack([np.eye(N) * (2.0 + q) + 0.1 * rng.standard_normal((N, N)) for q in range(Ql)])
    F = rng.standard_normal((Qr, N))
    Theta = np.concatenate([rng.uniform(0.5, 2.0, (B, Ql)), rng.uniform(-1, 1, (B, Qr))], axis=1)
    beta = rng.uniform(0.5, 2.0, B)
    scale = [b[0] for b in blocks]
    src = [b[1] for b in blocks]
    length = [b[2] for b in blocks]
    Kz = sum(length)
    G = rng.standard_normal((Kz, Kz))              # deliberately NOT symmetric
    engine.set_system(A, F)
    engine.set_estimator(scale, src, length, Ql, v_len, G)   # v_theta = the first v_len rhs thetas
    engine.set_training_set(Theta, beta)
    r = engine.sweep_resident(want_u=True, want_eps2=True, raise_on_singular=False)
    U = r["u"]
    V = np.concatenate([U, Theta[:, Ql:Ql + v_len], np.ones((B, 1))], axis=1)
    Z = np.concatenate([(Theta[:, sc:sc + 1] if sc >= 0 else 1.0) * V[:, so:so + ln] for sc, so, ln in blocks], axis=1)
    eps2 = np.einsum("pk,pk->p", Z @ G, Z)
    scale_ref = np.einsum("pk,pk->p", np.abs(Z) @ np.abs(G), np.abs(Z))   # cancellation-aware tolerance
    assert np.max(np.abs(r["eps2"] - eps2) / scale_ref) <= 1e-12, (N, blocks)


@pytest.mark.parametrize("seed", range(8))
def test_pair_form_and_folded_kernel_agree(engine, seed, monkeypatch):
    """The two estimator kernels (rb_estimator_pair_kernel, the default, and the folded rb_estimator_kernel behind
    RB_EST_KERNEL=fold) evaluate the same quadratic form: same eps2 to rounding, and the pair form issues fewer
    multiply-adds than the folded one on a table with several blocks per class."""
    rng = np.random.default_rng(5000 + seed)
    N, Ql, Qr, v_len, blocks = _case(rng)
    B = int(rng.integers(100, 700))
    A = np.stack([np.eye(N) * (2.0 + q) + 0.1 * rng.standard_normal((N, N)) for q in range(Ql)])
    F = rng.standard_normal((Qr, N))
    Theta = np.concatenate([rng.uniform(0.5, 2.0, (B, Ql)), rng.uniform(-1, 1, (B, Qr))], axis=1)
    beta = rng.uniform(0.5, 2.0, B)
    scale, src, length = [b[0] for b in blocks], [b[1] for b in blocks], [b[2] for b in blocks]
    Kz = sum(length)
    G = rng.standard_normal((Kz, Kz))
    engine.set_system(A, F)
    engine.set_training_set(Theta, beta)
    out = {}
    for kind in ("pair", "fold"):
        monkeypatch.setenv("RB_EST_KERNEL", kind)
        engine.set_estimator(scale, src, length, Ql, v_len, G)
        r = engine.sweep_resident(want_u=True, want_eps2=True, raise_on_singular=False)
        out[kind] = (r["eps2"].copy(), r["u"].copy(), engine.estimator_work())
    V = np.concatenate([out["pair"][1], Theta[:, Ql:Ql + v_len], np.ones((B, 1))], axis=1)
    Z = np.concatenate([(Theta[:, sc:sc + 1] if sc >= 0 else 1.0) * V[:, so:so + ln] for sc, so, ln in blocks], axis=1)
    scale_ref = np.einsum("pk,pk->p", np.abs(Z) @ np.abs(G), np.abs(Z))
    assert np.array_equal(out["pair"][1], out["fold"][1])
    assert np.max(np.abs(out["pair"][0] - out["fold"][0]) / scale_ref) <= 1e-12
    assert out["pair"][2] != out["fold"][2]       # the two packings really were different kernels


def test_pair_form_work_at_benchmark_shape(engine):
    """N=100, Q_a=50, Q_f=10: the pair form issues the algorithmic multiply-adds plus 0.6 % of padding."""
    from rbnics_b200 import synthetic
    N, Qa, Qf = 100, 50, 10
    prob = synthetic.elliptic_reduced_problem(N, Qa, Qf, seed=0, riesz_rank=64)
    engine.set_system(prob["A"], prob["F"])
    engine.set_elliptic_estimator(prob["Gff"], prob["Gaf"], prob["Gaa"])
    w = engine.estimator_work()
    algorithmic = Qa * (Qa + 1) // 2 * (N * (N + 1) // 2) + Qa * N * Qf + Qf * (Qf + 1) // 2
    assert algorithmic <= w["issued_macs_per_param"] <= 1.007 * algorithmic


def test_wide_v_tile_falls_back_to_the_folded_kernel(engine):
    """N = 160 with 60 theta columns: the V' tile of the pair form (128 x (N + Q + 2) doubles) does not fit in shared
    memory next to its record ring, so rb_set_estimator keeps the folded kernel; the sweep must still be right."""
    rng = np.random.default_rng(77)
    N, Ql, Qr, B = 160, 48, 12, 300
    A = np.stack([np.eye(N) * (2.0 + q) + 0.05 * rng.standard_normal((N, N)) for q in range(Ql)])
    F = rng.standard_normal((Qr, N))
    Theta = np.concatenate([rng.uniform(0.5, 2.0, (B, Ql)), rng.uniform(-1, 1, (B, Qr))], axis=1)
    beta = rng.uniform(0.5, 2.0, B)
    blocks = [(q, 0, N) for q in range(3)] + [(-1, N, Qr)]
    scale, src, length = [b[0] for b in blocks], [b[1] for b in blocks], [b[2] for b in blocks]
    Kz = sum(length)
    G = rng.standard_normal((Kz, Kz))
    engine.set_system(A, F)
    engine.set_estimator(scale, src, length, Ql, Qr, G)
    engine.set_training_set(Theta, beta)
    r = engine.sweep_resident(want_u=True, want_eps2=True, raise_on_singular=False)
    V = np.concatenate([r["u"], Theta[:, Ql:], np.ones((B, 1))], axis=1)
    Z = np.concatenate([(Theta[:, sc:sc + 1] if sc >= 0 else 1.0) * V[:, so:so + ln] for sc, so, ln in blocks], axis=1)
    eps2 = np.einsum("pk,pk->p", Z @ G, Z)
    scale_ref = np.einsum("pk,pk->p", np.abs(Z) @ np.abs(G), np.abs(Z))
    assert np.max(np.abs(r["eps2"] - eps2) / scale_ref) <= 1e-12

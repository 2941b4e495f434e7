"""POD-Greedy sweep of a parabolic reduced problem over a whole training set on the B200.

Replaces, for every mu of the training set at once, the closure of
``TimeDependentRBReduction._greedy`` (rbnics/reduction_methods/base/time_dependent_rb_reduction.py:262-295):
``reduced_problem.solve()`` = K implicit-Euler steps (rbnics/backends/online/numpy/time_stepping.py:104-112,226-238
with the residual / jacobian of rbnics/problems/parabolic/abstract_parabolic_reduced_problem.py:17-36),
``estimate_error()`` over time (abstract_parabolic_rb_reduced_problem.py:29-92) and the Simpson time quadrature
(rbnics/backends/common/time_quadrature.py:22-29), then ``training_set.max``.

Differences in HOW, not WHAT: the reference refactorises ``M/dt + A(mu)`` at every step and evaluates the six Gram
storages term by term; here the assembled matrix is kept per parameter for all K steps and the residual norm of a
step is ONE quadratic form ``z^T G z`` with ``z = [theta_a (x) u^k ; theta_m (x) udot^k ; theta_f]``.
Limitations: zero initial condition, thetas constant in time (tutorial 06).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
from scipy.integrate import simpson

from . import _lib as L

PARABOLIC_TERM_PAIRS = (("f", "f"), ("a", "f"), ("a", "a"), ("m", "f"), ("m", "a"), ("m", "m"))


def simpson_weights(K, dt):
    """scipy.integrate.simpson(y, dx=dt) as a linear functional on K+1 samples: w[k] = simpson(e_k)."""
    eye = np.eye(K + 1)
    return np.array([simpson(eye[k], dx=dt) for k in range(K + 1)])


def parabolic_quadratic_form(E, Qa, Qm, Qf, N):
    """Pack the six error-estimation storages into one quadratic form over
    z = [theta_a (x) u ; theta_m (x) udot ; theta_f],  V = [u (N) | udot (N) | theta_f | 1],
    Theta columns [theta_m | theta_a | theta_f].  Cross pairs go into both triangles (factor 2 of the reference)."""
    off = {"a": 0, "m": Qa * N, "f": Qa * N + Qm * N}
    Q = {"a": Qa, "m": Qm, "f": Qf}
    Kz = Qa * N + Qm * N + Qf
    G = np.zeros((Kz, Kz))

    def sl(t):
        n = Q[t] * N if t != "f" else Qf
        return slice(off[t], off[t] + n)

    for (t1, t2) in PARABOLIC_TERM_PAIRS:
        blk = np.asarray(E[t1, t2], dtype=np.float64)
        o1, o2 = t1 != "f", t2 != "f"
        if o1 and o2:
            M = blk.reshape(Q[t1], Q[t2], N, N).transpose(0, 2, 1, 3).reshape(Q[t1] * N, Q[t2] * N)
        elif o1:
            M = blk.reshape(Q[t1], Q[t2], N).transpose(0, 2, 1).reshape(Q[t1] * N, Q[t2])
        else:
            M = blk.reshape(Q[t1], Q[t2])
        G[sl(t1), sl(t2)] += M
        if t1 != t2:
            G[sl(t2), sl(t1)] += M.T
    scale = [Qm + q for q in range(Qa)] + list(range(Qm)) + [-1]
    src = [0] * Qa + [N] * Qm + [2 * N]
    length = [N] * Qa + [N] * Qm + [Qf]
    return G, (scale, src, length, Qm + Qa, Qf)


class ParabolicSweep:
    """``ParabolicSweep(engine, M, A, F, E, dt, K).sweep(Theta_m, Theta_a, Theta_f, beta)`` -> max / argmax of the
    time-integrated error estimator over the training set (optionally per-mu values, eps2 and solutions over time).

    M: (Qm, N, N), A: (Qa, N, N), F: (Qf, N) reduced operators; E[(t1, t2)] stacked error-estimation storages for the
    pairs of ``PARABOLIC_TERM_PAIRS``; final time T = K * dt."""

    def __init__(self, engine, M, A, F, E, dt, K, bc_rows=None):
        self.engine = engine
        M = np.asarray(M, dtype=np.float64)
        A = np.asarray(A, dtype=np.float64)
        F = np.asarray(F, dtype=np.float64)
        self.Qm, self.Qa, self.Qf, self.N = M.shape[0], A.shape[0], F.shape[0], A.shape[1]
        self.dt, self.K = float(dt), int(K)
        engine.set_system(np.concatenate([M / self.dt, A]), F, bc_rows=bc_rows)
        G, (scale, src, length, voff, vlen) = parabolic_quadratic_form(E, self.Qa, self.Qm, self.Qf, self.N)
        sc = np.ascontiguousarray(scale, dtype=np.int32)
        so = np.ascontiguousarray(src, dtype=np.int32)
        ln = np.ascontiguousarray(length, dtype=np.int32)
        G = np.ascontiguousarray(G)
        engine._check(engine._lib.rb_set_estimator_ex(engine._h, int(sc.size), sc.ctypes.data_as(L.c_int_p),
                                                      so.ctypes.data_as(L.c_int_p), ln.ctypes.data_as(L.c_int_p),
                                                      int(voff), int(vlen), L.dptr(G), 2 * self.N))
        self.weights = np.ascontiguousarray(simpson_weights(self.K, self.dt))

    def sweep(self, Theta_m, Theta_a, Theta_f, beta, ThetaBC=None, index_offset=0, index_stride=1, want_delta=False,
              want_eps2=False, want_u=False, U0=None, initial_error_estimate_squared=None):
        """``U0`` (B x N): reduced initial condition of every parameter (None: homogeneous);
        ``initial_error_estimate_squared`` (B): the squared bound at t0 (None: zero).  Both are cheap host-side batched
        products of theta_ic with the reduced initial-condition vectors."""
        Theta = np.ascontiguousarray(np.concatenate([Theta_m, Theta_a, Theta_f], axis=1), dtype=np.float64)
        eng = self.engine
        eng.set_training_set(Theta, beta, ThetaBC, index_offset=index_offset, index_stride=index_stride)
        B = Theta.shape[0]
        mv, mi = C.c_double(), C.c_int64()
        delta = np.empty(B) if want_delta else None
        eps2 = np.empty((B, self.K + 1)) if want_eps2 else None
        u = np.empty((B, self.K + 1, self.N)) if want_u else None
        if U0 is not None:
            U0 = np.ascontiguousarray(U0, dtype=np.float64)
            assert U0.shape == (B, self.N)
        e0 = None
        if initial_error_estimate_squared is not None:
            e0 = np.ascontiguousarray(initial_error_estimate_squared, dtype=np.float64)
            assert e0.shape == (B,)
        eng._check(eng._lib.rb_sweep_parabolic_ic(eng._h, self.Qm, self.K, self.dt, L.dptr(self.weights), L.dptr(U0),
                                                  L.dptr(e0), C.byref(mv), C.byref(mi), L.dptr(delta), L.dptr(eps2),
                                                  L.dptr(u)))
        return {"max": mv.value, "argmax": mi.value, "delta": delta, "eps2": eps2, "u": u}

"""SweepEngine: one handle of the CUDA library = one GPU.  Thin, typed wrapper over the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), f"expected shape {shape}, got {a.shape}"
    return a


class SweepEngine:
    """Greedy sweep over a training set on one B200 (include/rbnics_b200.h)."""

    def __init__(self, device: int = 0):
        self._lib = L.load()
        h = C.c_void_p()
        rc = self._lib.rb_ctx_create(int(device), C.byref(h))
        if rc != L.RB_OK:
            raise L.B200LibraryError(
                f"rb_ctx_create(device={device}) failed with code {rc}: no sm_100 GPU visible? "
                "(rbnics_b200 has no CPU fallback)")
        self._h = h
        self.device = device
        self.N = self.Ql = self.Qr = 0
        self.B = 0

    # -- plumbing ----------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.rb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc == L.RB_OK:
            return
        msg = self._lib.rb_last_error(self._h).decode()
        if rc == L.RB_ERR_SINGULAR:
            raise L.SingularReducedMatrixError(msg)
        if rc == L.RB_ERR_ARG:
            raise ValueError(msg)
        raise L.B200LibraryError(f"[{rc}] {msg}")

    def device_info(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._check(self._lib.rb_device_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"cc": (a.value, b.value), "sm_count": c.value}

    # -- operators ---------------------------------------------------------------------------------------------
    def set_system(self, A, F, bc_rows=None):
        """A: (Ql, N, N), F: (Qr, N); bc_rows: indices of the lifting rows."""
        A = np.asarray(A, dtype=np.float64)
        F = np.asarray(F, dtype=np.float64)
        if A.ndim != 3 or A.shape[1] != A.shape[2]:
            raise ValueError("A must be (Ql, N, N)")
        Ql, N = A.shape[0], A.shape[1]
        if F.ndim != 2 or (N > 0 and F.shape[1] != N):
            raise ValueError("F must be (Qr, N)")
        Qr = F.shape[0]
        A = np.ascontiguousarray(A)
        F = np.ascontiguousarray(F)
        bc = np.ascontiguousarray(bc_rows if bc_rows is not None else [], dtype=np.int32)
        self._check(self._lib.rb_set_system(self._h, N, Ql, Qr, L.dptr(A), L.dptr(F), int(bc.size),
                                            bc.ctypes.data_as(L.c_int_p)))
        self.N, self.Ql, self.Qr, self.n_bc = N, Ql, Qr, int(bc.size)

    def set_estimator(self, blk_scale_col, blk_src_off, blk_len, v_theta_off, v_theta_len, G):
        sc = np.ascontiguousarray(blk_scale_col, dtype=np.int32)
        so = np.ascontiguousarray(blk_src_off, dtype=np.int32)
        ln = np.ascontiguousarray(blk_len, dtype=np.int32)
        Kz = int(ln.sum())
        G = _f64(G, (Kz, Kz))
        self._check(self._lib.rb_set_estimator(self._h, int(sc.size), sc.ctypes.data_as(L.c_int_p),
                                               so.ctypes.data_as(L.c_int_p), ln.ctypes.data_as(L.c_int_p),
                                               int(v_theta_off), int(v_theta_len), L.dptr(G)))

    def set_elliptic_estimator(self, Gff, Gaf, Gaa):
        """Residual-norm estimator of rbnics/problems/elliptic/elliptic_rb_reduced_problem.py:55-64.

        Gff: (Qf, Qf), Gaf: (Qa, Qf, N), Gaa: (Qa, Qa, N, N).  z = [theta_a (x) u ; theta_f],
        G = [[Gaa, Gaf], [Gaf^T, Gff]]  so that z^T G z = aa + 2 af + ff."""
        G, desc = elliptic_quadratic_form(Gff, Gaf, Gaa, self.Ql, self.Qr, self.N)
        self.set_estimator(*desc, G)

    def set_stokes_estimator(self, E, Q):
        """Residual-norm estimator of rbnics/problems/stokes/stokes_rb_reduced_problem.py:53-83 (9 term pairs).

        The system must have been set with lhs terms ordered [a | b | bt] and rhs terms [f | g];
        Q = {"a": Qa, "b": Qb, "bt": Qbt, "f": Qf, "g": Qg}; E[(t1, t2)] stacked arrays (Q1, Q2[, N[, N]])."""
        G, desc = stokes_quadratic_form(E, Q, self.N)
        self.set_estimator(*desc, G)

    # -- training set + sweep ----------------------------------------------------------------------------------
    def set_training_set(self, Theta, beta, ThetaBC=None, index_offset=0, index_stride=1):
        Theta = _f64(Theta)
        beta = _f64(beta)
        B, QT = Theta.shape
        assert beta.shape == (B,)
        nbc = 0
        if ThetaBC is not None:
            ThetaBC = _f64(ThetaBC)
            nbc = ThetaBC.shape[1]
        self._check(self._lib.rb_set_training_set(self._h, B, QT, L.dptr(Theta), nbc, L.dptr(ThetaBC), L.dptr(beta),
                                                  int(index_offset), int(index_stride)))
        self.B = B

    def sweep_resident(self, kind=L.RB_EST_SQRT_EPS2_OVER_BETA, want_delta=False, want_u=False, want_eps2=False,
                       raise_on_singular=True):
        mv, mi = C.c_double(), C.c_int64()
        delta = np.empty(self.B) if want_delta else None
        u = np.empty((self.B, self.N)) if want_u else None
        eps2 = np.empty(self.B) if want_eps2 else None
        rc = self._lib.rb_sweep_resident(self._h, int(kind), C.byref(mv), C.byref(mi), L.dptr(delta), L.dptr(u),
                                         L.dptr(eps2))
        if rc != L.RB_ERR_SINGULAR or raise_on_singular:
            self._check(rc)
        return {"max": mv.value, "argmax": mi.value, "delta": delta, "u": u, "eps2": eps2}

    def sweep(self, Theta, beta, ThetaBC=None, kind=L.RB_EST_SQRT_EPS2_OVER_BETA, index_offset=0, index_stride=1,
              want_delta=False, want_u=False, want_eps2=False):
        """End-to-end entry: host buffers in, (max, argmax) out, copies inside."""
        Theta = _f64(Theta)
        beta = _f64(beta)
        B, QT = Theta.shape
        nbc = 0
        if ThetaBC is not None:
            ThetaBC = _f64(ThetaBC)
            nbc = ThetaBC.shape[1]
        mv, mi = C.c_double(), C.c_int64()
        delta = np.empty(B) if want_delta else None
        u = np.empty((B, self.N)) if want_u else None
        eps2 = np.empty(B) if want_eps2 else None
        self.B = B
        self._check(self._lib.rb_sweep(self._h, B, QT, L.dptr(Theta), nbc, L.dptr(ThetaBC), L.dptr(beta),
                                       int(index_offset), int(index_stride), int(kind), C.byref(mv), C.byref(mi),
                                       L.dptr(delta), L.dptr(u), L.dptr(eps2)))
        return {"max": mv.value, "argmax": mi.value, "delta": delta, "u": u, "eps2": eps2}

    def last_timings(self):
        ms = (C.c_float * 5)()
        n = C.c_int()
        self._check(self._lib.rb_last_timings(self._h, ms, C.byref(n)))
        return {"assemble_lu_ms": ms[0], "estimator_ms": ms[2], "finalize_ms": ms[3], "total_ms": ms[4],
                "launches": n.value}

    def compute_outputs(self, S, ThetaS):
        """s_N(mu) = u(mu)^T sum_q theta_s_q(mu) s_q for the solutions of the last sweep (elliptic_reduced_problem.py:31-32)."""
        S = _f64(S)
        ThetaS = _f64(ThetaS, (self.B, S.shape[0]))
        assert S.shape[1] == self.N
        out = np.empty(self.B)
        self._check(self._lib.rb_compute_outputs(self._h, S.shape[0], L.dptr(S), L.dptr(ThetaS), L.dptr(out)))
        return out

    def solve_supremizers(self, inner_product_s, operator_bt_restricted, Theta_bt, U=None, N_us=None):
        """Reduced supremizers of a whole batch, ``StokesReducedProblem._solve_supremizer`` for every mu at once
        (rbnics/problems/stokes/stokes_reduced_problem.py:80-103): ``X_s s = (sum_q theta_q Bt_q) u``.

        inner_product_s: the single matrix of ``inner_product["s"]``; operator_bt_restricted: (Q, N_us, N_usp);
        Theta_bt: (B, Q); U: (B, N_usp) reduced solutions, or None for those left in HBM by the last sweep.
        X_s does not depend on mu: it is factorised once on the host (``W_q = X_s^-1 Bt_q``)."""
        Bt = _f64(operator_bt_restricted)
        Q, n_us, n_usp = Bt.shape
        N_us = n_us if N_us is None else int(N_us)
        Xs = _f64(inner_product_s)[:N_us, :N_us]
        W = np.ascontiguousarray(np.stack([np.linalg.solve(Xs, Bt[q, :N_us, :]) for q in range(Q)]))
        Theta_bt = _f64(Theta_bt)
        B = Theta_bt.shape[0]
        assert Theta_bt.shape == (B, Q)
        Uc = None
        if U is not None:
            Uc = _f64(U, (B, n_usp))
        S = np.empty((B, N_us))
        self._check(self._lib.rb_supremizers(self._h, N_us, n_usp, Q, L.dptr(W), B, L.dptr(Theta_bt), L.dptr(Uc),
                                             L.dptr(S)))
        return S

    def estimator_work(self):
        """Multiply-adds per parameter issued by the estimator kernel (padding included) and its tile count."""
        m, n = C.c_int64(), C.c_int()
        self._check(self._lib.rb_estimator_work(self._h, C.byref(m), C.byref(n)))
        return {"issued_macs_per_param": m.value, "column_tiles": n.value}

    def result_device_ptrs(self):
        p = C.c_void_p()
        d = L.c_double_p()
        self._check(self._lib.rb_result_device_ptrs(self._h, C.byref(p), C.byref(d)))
        return p.value, C.cast(d, C.c_void_p).value


def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by page-locked host memory (cudaMallocHost) for the end-to-end path."""
    lib = L.load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib.rb_host_alloc(max(n, 8))
    if not p:
        raise L.B200LibraryError("rb_host_alloc failed")
    buf = (C.c_char * max(n, 8)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    # keep the allocation alive with the array; freed explicitly by pinned_free
    arr.flags.writeable = True
    return arr, p


def pinned_free(p):
    L.load().rb_host_free(C.c_void_p(p))


def elliptic_quadratic_form(Gff, Gaf, Gaa, Qa, Qf, N):
    """Pack (Gff, Gaf, Gaa) into the dense quadratic form + block descriptors of rb_set_estimator.

    Returns (G, (blk_scale_col, blk_src_off, blk_len, v_theta_off, v_theta_len)).
    Theta columns are [theta_a (Qa) | theta_f (Qf)]; V = [u (N) | theta_f (Qf) | 1]."""
    Gff = np.asarray(Gff, dtype=np.float64).reshape(Qf, Qf)
    if N > 0:
        Gaf = np.asarray(Gaf, dtype=np.float64).reshape(Qa, Qf, N)
        Gaa = np.asarray(Gaa, dtype=np.float64).reshape(Qa, Qa, N, N)
        Ka = Qa * N
        G = np.empty((Ka + Qf, Ka + Qf))
        G[:Ka, :Ka] = Gaa.transpose(0, 2, 1, 3).reshape(Ka, Ka)
        Gaf2 = Gaf.transpose(0, 2, 1).reshape(Ka, Qf)
        G[:Ka, Ka:] = Gaf2
        G[Ka:, :Ka] = Gaf2.T
        G[Ka:, Ka:] = Gff
        scale = list(range(Qa)) + [-1]
        src = [0] * Qa + [N]
        length = [N] * Qa + [Qf]
    else:
        G = Gff.copy()
        scale, src, length = [-1], [0], [Qf]
    return G, (scale, src, length, Qa, Qf)


STOKES_TERM_PAIRS = (("f", "f"), ("g", "g"), ("a", "f"), ("bt", "f"), ("b", "g"),
                     ("a", "a"), ("a", "bt"), ("bt", "bt"), ("b", "b"))


def stokes_quadratic_form(E, Q, N):
    """Pack the 9 Stokes error-estimation storages into ONE quadratic form z^T G z (rb_set_estimator).

    Theta columns: [a | b | bt | f | g];  V = [u (N) | theta_f | theta_g | 1];
    z = [theta_a (x) u ; theta_b (x) u ; theta_bt (x) u ; theta_f ; theta_g].  Cross pairs (t1 != t2) are placed in
    both triangles so that z^T G z carries the reference's factor 2; pairs the reference skips ("useless Riesz
    products", stokes_rb_reduced_problem.py:27-32) are zero blocks."""
    Qa, Qb, Qbt, Qf, Qg = Q["a"], Q["b"], Q["bt"], Q["f"], Q["g"]
    col = {"a": 0, "b": Qa, "bt": Qa + Qb, "f": Qa + Qb + Qbt, "g": Qa + Qb + Qbt + Qf}
    # z offsets
    zoff, o = {}, 0
    for t in ("a", "b", "bt"):
        zoff[t] = o
        o += Q[t] * N
    zoff["f"] = o
    o += Qf
    zoff["g"] = o
    o += Qg
    Kz = o
    G = np.zeros((Kz, Kz))

    def zslice(t):
        n = Q[t] * N if t in ("a", "b", "bt") else Q[t]
        return slice(zoff[t], zoff[t] + n)

    for (t1, t2) in STOKES_TERM_PAIRS:
        blk = np.asarray(E[t1, t2], dtype=np.float64)
        q1, q2 = Q[t1], Q[t2]
        op1, op2 = t1 in ("a", "b", "bt"), t2 in ("a", "b", "bt")
        if op1 and op2:
            M = blk.reshape(q1, q2, N, N).transpose(0, 2, 1, 3).reshape(q1 * N, q2 * N)
        elif op1:
            M = blk.reshape(q1, q2, N).transpose(0, 2, 1).reshape(q1 * N, q2)
        else:
            M = blk.reshape(q1, q2)
        if N == 0 and (op1 or op2):
            continue
        G[zslice(t1), zslice(t2)] += M
        if t1 != t2:
            G[zslice(t2), zslice(t1)] += M.T
    scale, src, length = [], [], []
    if N > 0:
        for t in ("a", "b", "bt"):
            for q in range(Q[t]):
                scale.append(col[t] + q)
                src.append(0)
                length.append(N)
    else:  # empty solution: only the (f,f) and (g,g) terms remain
        G = G[zoff["f"]:, zoff["f"]:].copy()
    if Qf:
        scale.append(-1)
        src.append(N)
        length.append(Qf)
    if Qg:
        scale.append(-1)
        src.append(N + Qf)
        length.append(Qg)
    return G, (scale, src, length, col["f"], Qf + Qg)

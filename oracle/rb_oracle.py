"""Numpy restatement of the reference's greedy-sweep hot path and offline basis linear algebra.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): this is the checker, never the product.

Every function cites the reference file:line (relative to the RBniCS tree) whose arithmetic it
restates.  The *loop* variants keep the reference's operation order (sequential sum over q, row-major
(i, j) double loop, zero-theta skipping, ``numpy.linalg.solve``, first-index arg-max); the *batched*
variants are the vectorised "strong CPU" comparator and are themselves checked
against the loop variants in ``tests/test_oracle.py``.

Pinning status (DESIGN.md "Oracle"): the reference cannot be imported in the build container
(``multipledispatch``/``mpi4py``/``dolfin`` are absent), and it ships no stored numeric vectors for this
path.  The oracle is pinned (a) against the closed-form einsum known-answer formulas the reference's own
tests use (tests/performance/backends/online/numpy/test_numpy_greedy_prototype.py:72-84 and siblings) and
(b) against golden vectors produced by executing the reference's own function bodies (product.py,
parameter_space_subset.py, parallel_max.py, DirichletBC.py, numpy linear_solver) under stubbed imports
(``tests/golden/make_golden.py``).
"""
from __future__ import annotations

from math import sqrt

import numpy as np
import scipy.linalg
import scipy.sparse as sp

# ----------------------------------------------------------------------------------------------------------------------
# Online sweep, reference-faithful (one parameter at a time)
# ----------------------------------------------------------------------------------------------------------------------


def product_order1(thetas, operators):
    """sum_q theta_q * op_q, sequential, skipping theta_q == 0 for q > 0.

    rbnics/backends/online/basic/product.py:22-33 (``output = theta*operator`` then ``output += ...``)."""
    assert len(thetas) == len(operators)
    output = None
    for index, (theta, operator) in enumerate(zip(thetas, operators)):
        if index == 0:
            output = theta * operator
        elif theta != 0.0:
            output += theta * operator
    return output


def product_order2(thetas, operators, thetas2):
    """sum_ij theta_i * op[i][j] * theta2_j in row-major (i, j) order with zero skipping.

    rbnics/backends/online/basic/product.py:34-46."""
    output = None
    for i in range(len(thetas)):
        for j in range(len(thetas2)):
            if i == 0 and j == 0:
                output = thetas[0] * operators[0][0] * thetas2[0]
            elif thetas[i] != 0.0 and thetas2[j] != 0.0:
                output += thetas[i] * operators[i][j] * thetas2[j]
    return output


def apply_dirichlet_rows(lhs, rhs, theta_bc):
    """Lifting rows: A[i,:]=0, A[i,i]=1, f[i]=theta_bc[i] for i < N_bc.

    rbnics/backends/online/basic/linear_solver.py:56-59 and wrapping/DirichletBC.py:37-56."""
    if theta_bc is None:
        return
    for i, bc_i in enumerate(theta_bc):
        rhs[i] = bc_i
    for i, _ in enumerate(theta_bc):
        lhs[i, :] = 0.0
        lhs[i, i] = 1.0


def reduced_solve(theta_a, theta_f, A_q, f_q, N, theta_bc=None):
    """One reduced solve.  A_q: list of (Nmax x Nmax) arrays, f_q: list of Nmax vectors; uses [:N] slices.

    rbnics/problems/base/parametrized_reduced_differential_problem.py:320-342 (N == 0 -> empty solution),
    rbnics/problems/elliptic/elliptic_reduced_problem.py:19-28 (matrix_eval / vector_eval),
    rbnics/backends/online/numpy/linear_solver.py:29-31 (numpy.linalg.solve == LAPACK dgesv)."""
    if N == 0:
        return np.zeros(0)
    lhs = product_order1(theta_a, [A[:N, :N] for A in A_q])
    rhs = product_order1(theta_f, [f[:N] for f in f_q])
    lhs = np.array(lhs, dtype=np.float64)
    rhs = np.array(rhs, dtype=np.float64)
    apply_dirichlet_rows(lhs, rhs, theta_bc)
    return np.linalg.solve(lhs, rhs)


def residual_norm_squared(u, theta_a, theta_f, G_ff, G_af, G_aa, N):
    """eps2 = thf^T Gff thf + 2 u^T (sum_ij tha_i thf_j gaf_ij) + u^T (sum_ij tha_i tha_j Gaa_ij) u.

    G_ff[i][j] floats; G_af[i][j] vectors (Nmax); G_aa[i][j] matrices (Nmax x Nmax).
    rbnics/problems/elliptic/elliptic_rb_reduced_problem.py:55-64; ``transpose(u)*M*u`` is
    ``dot(u, M @ u)`` (rbnics/backends/basic/transpose.py:208-220, online/numpy/matrix.py:31-41)."""
    ff = product_order2(theta_f, G_ff, theta_f)
    if N == 0:
        # solution is empty: the af / aa products are over empty slices and contribute 0
        return float(ff)
    af = product_order2(theta_a, [[g[:N] for g in row] for row in G_af], theta_f)
    aa = product_order2(theta_a, [[g[:N, :N] for g in row] for row in G_aa], theta_a)
    return float(ff + 2.0 * np.dot(u, af) + np.dot(u, aa @ u))


def estimate_error(eps2, beta, kind="sqrt_eps2_over_beta"):
    """Delta = sqrt(|eps2|)/beta  (rbnics/problems/elliptic/elliptic_rb_reduced_problem.py:35-40) or
    sqrt(|eps2|/beta) (elliptic_coercive_compliant_rb_reduced_problem.py:22-27)."""
    assert eps2 >= 0.0 or np.isclose(eps2, 0.0)
    assert beta >= 0.0
    if kind == "sqrt_eps2_over_beta":
        return sqrt(abs(eps2)) / beta
    elif kind == "sqrt_of_eps2_over_beta":
        return sqrt(abs(eps2) / beta)
    raise ValueError(kind)


def argmax_first(values):
    """numpy.argmax semantics: first index among equal maxima, NaN wins.
    rbnics/sampling/parameter_space_subset.py:67,75."""
    return int(np.argmax(values))


def training_set_max(generator, training_set, rank=0, size=1):
    """ParameterSpaceSubset.max: cyclic shard + local arg-max.  Returns (local_value_max, global_index).

    rbnics/sampling/parameter_space_subset.py:52-77 (``range(rank, len, size)``)."""
    local_idx = list(range(rank, len(training_set), size))
    values = np.zeros(len(local_idx))
    for i, gi in enumerate(local_idx):
        values[i] = generator(training_set[gi])
    li = int(np.argmax(values))
    return float(values[li]), local_idx[li]


def parallel_max(local_values, local_indices):
    """MPI max-loc rule: global max value; among ranks holding it the HIGHEST rank wins.

    rbnics/utils/mpi/parallel_max.py:18-33.  (The serial rule -- lowest index -- is ``argmax_first``.)"""
    gmax = max(local_values)
    owner = max(r for r, v in enumerate(local_values) if v == gmax)
    return local_values[owner], local_indices[owner]


def sweep_loop(Theta_a, Theta_f, beta, A_q, f_q, G_ff, G_af, G_aa, N, Theta_bc=None,
               kind="sqrt_eps2_over_beta", return_eps2=False):
    """The reference's per-parameter greedy closure over a whole training set (serial).

    rbnics/reduction_methods/base/rb_reduction.py:160-185 + everything above.
    Returns (delta[B], u[B, N], argmax index, max value)."""
    B = Theta_a.shape[0]
    deltas = np.zeros(B)
    eps2s = np.zeros(B)
    U = np.zeros((B, N))
    for p in range(B):
        tha = tuple(float(x) for x in Theta_a[p])
        thf = tuple(float(x) for x in Theta_f[p])
        thbc = None if Theta_bc is None else tuple(float(x) for x in Theta_bc[p])
        u = reduced_solve(tha, thf, A_q, f_q, N, thbc)
        eps2 = residual_norm_squared(u, tha, thf, G_ff, G_af, G_aa, N)
        eps2s[p] = eps2
        b = float(beta[p])
        if kind == "sqrt_eps2_over_beta":
            deltas[p] = sqrt(abs(eps2)) / b
        else:
            deltas[p] = sqrt(abs(eps2) / b)
        U[p] = u
    i = argmax_first(deltas)
    if return_eps2:
        return deltas, U, i, float(deltas[i]), eps2s
    return deltas, U, i, float(deltas[i])


# ----------------------------------------------------------------------------------------------------------------------
# Online sweep, vectorised (strong CPU baseline; same mathematics, different summation order)
# ----------------------------------------------------------------------------------------------------------------------

def stack_operators(A_q, f_q, G_ff, G_af, G_aa, N):
    """Flatten the reference's object arrays into contiguous tensors (what the C-ABI takes)."""
    Qa, Qf = len(A_q), len(f_q)
    A = np.stack([np.asarray(a)[:N, :N] for a in A_q]) if N > 0 else np.zeros((Qa, 0, 0))
    F = np.stack([np.asarray(f)[:N] for f in f_q]) if N > 0 else np.zeros((Qf, 0))
    Gff = np.array([[float(G_ff[i][j]) for j in range(Qf)] for i in range(Qf)])
    if N > 0:
        Gaf = np.stack([np.stack([np.asarray(G_af[i][j])[:N] for j in range(Qf)]) for i in range(Qa)])
        Gaa = np.stack([np.stack([np.asarray(G_aa[i][j])[:N, :N] for j in range(Qa)]) for i in range(Qa)])
    else:
        Gaf = np.zeros((Qa, Qf, 0))
        Gaa = np.zeros((Qa, Qa, 0, 0))
    return A, F, Gff, Gaf, Gaa


def sweep_batched(Theta_a, Theta_f, beta, A, F, Gff, Gaf, Gaa, Theta_bc=None,
                  kind="sqrt_eps2_over_beta", chunk=2048, return_eps2=False):
    """Chunked einsum / batched-solve version of ``sweep_loop`` on stacked tensors.

    A: (Qa,N,N)  F: (Qf,N)  Gff: (Qf,Qf)  Gaf: (Qa,Qf,N)  Gaa: (Qa,Qa,N,N)."""
    B = Theta_a.shape[0]
    Qa, N = A.shape[0], A.shape[1]
    Qf = F.shape[0]
    deltas = np.zeros(B)
    eps2s = np.zeros(B)
    U = np.zeros((B, N))
    Aflat = A.reshape(Qa, N * N)
    # G as one (Qa*N) x (Qa*N) matrix: rows (i,n), cols (j,m)
    Gbig = Gaa.transpose(0, 2, 1, 3).reshape(Qa * N, Qa * N) if N > 0 else None
    Gaf2 = Gaf.transpose(0, 2, 1).reshape(Qa * N, Qf) if N > 0 else None
    for s in range(0, B, chunk):
        e = min(B, s + chunk)
        ta, tf = Theta_a[s:e], Theta_f[s:e]
        ff = np.einsum("pi,ij,pj->p", tf, Gff, tf)
        if N > 0:
            lhs = (ta @ Aflat).reshape(-1, N, N)
            rhs = tf @ F
            if Theta_bc is not None:
                nbc = Theta_bc.shape[1]
                for i in range(nbc):
                    lhs[:, i, :] = 0.0
                    lhs[:, i, i] = 1.0
                    rhs[:, i] = Theta_bc[s:e, i]
            u = np.linalg.solve(lhs, rhs[..., None])[..., 0]
            z = (ta[:, :, None] * u[:, None, :]).reshape(e - s, Qa * N)
            aa = np.einsum("pk,pk->p", z @ Gbig, z)
            af = np.einsum("pk,pk->p", z @ Gaf2, tf)
            eps2 = ff + 2.0 * af + aa
            U[s:e] = u
        else:
            eps2 = ff
        eps2s[s:e] = eps2
        if kind == "sqrt_eps2_over_beta":
            deltas[s:e] = np.sqrt(np.abs(eps2)) / beta[s:e]
        else:
            deltas[s:e] = np.sqrt(np.abs(eps2) / beta[s:e])
    i = argmax_first(deltas)
    if return_eps2:
        return deltas, U, i, float(deltas[i]), eps2s
    return deltas, U, i, float(deltas[i])


def estimator_builtin_einsum(u, v, theta_a, theta_f, Gaa, Gaf, Gff):
    """The closed-form known-answer formula of the reference's own test (note: no factor 2 on the af term and
    distinct u, v -- it pins the *operations*): tests/performance/backends/online/numpy/
    test_numpy_greedy_prototype.py:72-84."""
    return (np.einsum("n,i,ijnm,j,m", u, theta_a, Gaa, theta_a, v, optimize=True)
            + np.einsum("i,ijn,j,n", theta_a, Gaf, theta_f, u, optimize=True)
            + np.einsum("i,ij,j", theta_f, Gff, theta_f, optimize=True))


def reference_rand(rng, *shape):
    """The reference tests' heavy-tailed signed generator ``(-1)**randint(2) * random()/(1e-3+random())``
    (tests/performance/backends/online/numpy/test_numpy_utils.py:58-59), on a seeded Generator."""
    return (-1.0) ** rng.integers(0, 2, size=shape) * rng.random(shape) / (1e-3 + rng.random(shape))


# ----------------------------------------------------------------------------------------------------------------------
# Offline basis linear algebra
# ----------------------------------------------------------------------------------------------------------------------

def gram_loop(S, X=None, S2=None):
    """C[i,j] = s_i . (X s2_j): loop j -> one mat-vec, loop i -> one dot.

    rbnics/backends/basic/transpose.py:304-313 (S^T X S) and :441-468 (Z^T A Z), with
    ``matrix_mul_vector`` = PETSc MatMult and ``vector_mul_vector`` = VecDot
    (rbnics/backends/dolfin/wrapping/matrix_mul.py:10-11, vector_mul.py:8-9).
    S: (Nh, Ns) columns are snapshots / basis functions; X: scipy.sparse CSR or None (identity)."""
    S2 = S if S2 is None else S2
    cols = [np.ascontiguousarray(S[:, i]) for i in range(S.shape[1])]  # the reference's vectors are contiguous
    C = np.zeros((S.shape[1], S2.shape[1]))
    for j in range(S2.shape[1]):
        s2j = np.ascontiguousarray(S2[:, j])
        y = s2j if X is None else X @ s2j
        for i in range(S.shape[1]):
            C[i, j] = np.dot(cols[i], y)
    return C


def gram_blas(S, X=None, S2=None):
    """Vectorised S^T (X S2)."""
    S2 = S if S2 is None else S2
    Y = S2 if X is None else X @ S2
    return S.T @ Y


def project_vector_loop(Z, f):
    """f_N[i] = z_i . f   (rbnics/backends/basic/transpose.py:365-400)."""
    out = np.zeros(Z.shape[1])
    for i in range(Z.shape[1]):
        out[i] = np.dot(Z[:, i], f)
    return out


def eigh_descending(C):
    """scipy.linalg.eigh + argsort()[::-1]  (rbnics/backends/online/numpy/eigen_solver.py:36-56)."""
    eigs, eigv = scipy.linalg.eigh(C)
    idx = eigs.argsort()[::-1]
    return eigs[idx], eigv[:, idx]


def pod(S, X, Nmax, tol):
    """POD by the method of snapshots.

    rbnics/backends/basic/proper_orthogonal_decomposition_base.py:39-92: correlation = S^T X S; eigh sorted
    descending; retained energy = cumsum|lambda|/sum|lambda|; b_n = S v_n / sqrt(b^T X b); stop after the
    first n with retained_energy[n] > 1 - tol (only when tol > 0).
    Returns (eigenvalues[:N], eigenvectors (list), Z (Nh x N), N, retained_energy)."""
    C = gram_loop(S, X) if S.shape[1] <= 64 else gram_blas(S, X)
    eigs, eigv = eigh_descending(C)
    Neigs = S.shape[1]
    Nmax = min(Nmax, Neigs)
    abs_e = np.abs(eigs)
    total = abs_e.sum()
    retained = np.cumsum(abs_e) / total if total > 0.0 else np.ones(Neigs)
    Z = []
    vecs = []
    N = 0
    for n in range(Nmax):
        v = eigv[:, n]
        vecs.append(v)
        b = S @ v  # functions_list_mul_online_vector: sum_k v_k s_k (dolfin/wrapping/functions_list_mul.py:26-36)
        Xb = b if X is None else X @ b
        norm_b = sqrt(np.dot(b, Xb))
        if norm_b != 0.0:
            b = b / norm_b
        Z.append(b)
        N = n + 1
        if tol > 0.0 and retained[n] > 1.0 - tol:
            break
    return eigs[:N], vecs, np.stack(Z, axis=1), N, retained


def gram_schmidt(new, Z, X):
    """One (modified, single pass) Gram-Schmidt step of a new function against the columns of Z.

    rbnics/backends/basic/gram_schmidt.py:22-36; projection step
    rbnics/backends/dolfin/wrapping/gram_schmidt_projection_step.py:8-11: z -= (z^T X b) b, sequentially."""
    z = np.array(new, dtype=np.float64)
    for k in range(Z.shape[1]):
        b = Z[:, k]
        Xb = b if X is None else X @ b
        z = z - np.dot(z, Xb) * b
    Xz = z if X is None else X @ z
    nrm = sqrt(np.dot(z, Xz))
    if nrm != 0.0:
        z = z / nrm
    return z


def build_reduced_operators(Z, A_ops, f_ops):
    """A^N_q = Z^T A_q Z, f^N_q = Z^T f_q.
    rbnics/problems/base/parametrized_reduced_differential_problem.py:727-790."""
    return [gram_blas(Z, A) for A in A_ops], [Z.T @ f for f in f_ops]


def build_error_estimation_operators(Z, X, A_ops, f_ops, solve_X):
    """Riesz representers and their Gram blocks.

    rbnics/problems/base/rb_reduced_problem.py:172-263: r_f[q] = X^-1 f_q; R_a[q][:, n] = -X^-1 A_q z_n;
    G_ff[q,q'] = r_q^T X r_q'; G_af[q,q'] = R_q^T X r_q' (N-vector); G_aa[q,q'] = R_q^T X R_q' (N x N).
    ``solve_X`` is the truth Riesz solve (out of scope; supplied by the test's input generator)."""
    N = Z.shape[1]
    r_f = [solve_X(f) for f in f_ops]
    R_a = [np.stack([-solve_X(A @ Z[:, n]) for n in range(N)], axis=1) if N > 0 else np.zeros((Z.shape[0], 0))
           for A in A_ops]
    Qa, Qf = len(A_ops), len(f_ops)
    G_ff = [[float(np.dot(r_f[i], X @ r_f[j])) for j in range(Qf)] for i in range(Qf)]
    G_af = [[R_a[i].T @ (X @ r_f[j]) for j in range(Qf)] for i in range(Qa)]
    G_aa = [[gram_blas(R_a[i], X, R_a[j]) for j in range(Qa)] for i in range(Qa)]
    return G_ff, G_af, G_aa, r_f, R_a


# ----------------------------------------------------------------------------------------------------------------------
# Stokes (saddle point, components u, s, p): multi-term assembly and the 9-term residual norm
# ----------------------------------------------------------------------------------------------------------------------

def reduced_solve_stokes(theta, op, N, theta_bc=None):
    """lhs = sum_a + sum_b + sum_bt, rhs = sum_f + sum_g, each an order-1 product; then the lifting rows and dgesv.

    theta / op: dicts over the terms "a", "b", "bt", "f", "g" (op[term] = list of N x N matrices / N-vectors).
    rbnics/problems/stokes/stokes_reduced_problem.py:38-53 (matrix_eval / vector_eval)."""
    assembled = {}
    for term in ("a", "b", "bt"):
        assembled[term] = product_order1(theta[term], [A[:N, :N] for A in op[term]])
    lhs = np.array(assembled["a"] + assembled["b"] + assembled["bt"], dtype=np.float64)
    for term in ("f", "g"):
        assembled[term] = product_order1(theta[term], [f[:N] for f in op[term]])
    rhs = np.array(assembled["f"] + assembled["g"], dtype=np.float64)
    apply_dirichlet_rows(lhs, rhs, theta_bc)
    return np.linalg.solve(lhs, rhs)


def supremizer_solve(inner_product_s, operator_bt_restricted, theta_bt, solution, N_us, theta_bc_s=None):
    """Reduced supremizer of one parameter: X_s[:N_us, :N_us] s = (sum_q theta_q Bt_q[:N_us, :N_usp]) u, then the
    lifting rows (homogeneous for the supremizer component) and dgesv.

    rbnics/problems/stokes/stokes_reduced_problem.py:80-103 (_solve_supremizer): lhs is the single inner-product
    matrix of the component "s", the right-hand side the assembled "bt_restricted" operator times the reduced
    solution [u; s; p] of the same parameter."""
    N_usp = solution.shape[0]
    lhs = np.array(inner_product_s[:N_us, :N_us], dtype=np.float64)
    bt = product_order1(theta_bt, [B[:N_us, :N_usp] for B in operator_bt_restricted])
    rhs = np.array(bt @ solution, dtype=np.float64)
    apply_dirichlet_rows(lhs, rhs, theta_bc_s)
    return np.linalg.solve(lhs, rhs)


STOKES_ERROR_ESTIMATION_TERMS = [("f", "f"), ("g", "g"), ("a", "f"), ("bt", "f"), ("b", "g"),
                                 ("a", "a"), ("a", "bt"), ("bt", "bt"), ("b", "b")]


def residual_norm_squared_stokes(u, theta, E, N):
    """The 9 term pairs of rbnics/problems/stokes/stokes_rb_reduced_problem.py:53-83, in the reference's order.

    E[(t1, t2)][i][j]: scalar (f,f / g,g), N-vector (x,f / x,g) or N x N matrix."""
    def o2(t1, t2, slicer):
        return product_order2(theta[t1], [[slicer(x) for x in row] for row in E[t1, t2]], theta[t2])

    sc = lambda x: x            # noqa: E731
    vec = lambda x: x[:N]       # noqa: E731
    mat = lambda x: x[:N, :N]   # noqa: E731
    return float(
        o2("f", "f", sc)
        + o2("g", "g", sc)
        + 2.0 * np.dot(u, o2("a", "f", vec))
        + 2.0 * np.dot(u, o2("bt", "f", vec))
        + 2.0 * np.dot(u, o2("b", "g", vec))
        + np.dot(u, o2("a", "a", mat) @ u)
        + 2.0 * np.dot(u, o2("a", "bt", mat) @ u)
        + np.dot(u, o2("bt", "bt", mat) @ u)
        + np.dot(u, o2("b", "b", mat) @ u))


# ----------------------------------------------------------------------------------------------------------------------
# Parabolic problems: POD-Greedy sweep
# ----------------------------------------------------------------------------------------------------------------------

def parabolic_solve(theta_m, theta_a, theta_f, M_q, A_q, f_q, N, dt, K, u0=None):
    """Linear implicit Euler from the initial condition u0 (zero if None), K steps.

    rbnics/problems/parabolic/abstract_parabolic_reduced_problem.py:17-36 (residual = M udot + A u - f,
    jacobian = M * coef + A) driven by rbnics/backends/online/numpy/time_stepping.py:104-112 (lhs = jacobian(1/dt),
    rhs = -residual(0, -u_prev/dt)) and :226-238 (udot = (u - u_prev)/dt).  The reference refactorises every step.
    Returns (solution_over_time[K+1][N], solution_dot_over_time[K+1][N])."""
    M = np.array(product_order1(theta_m, [m[:N, :N] for m in M_q]))
    A = np.array(product_order1(theta_a, [a[:N, :N] for a in A_q]))
    f = np.array(product_order1(theta_f, [v[:N] for v in f_q]))
    zero = np.zeros(N)
    u_prev = np.zeros(N) if u0 is None else np.array(u0, dtype=np.float64)
    sols, dots = [u_prev.copy()], [np.zeros(N)]
    for _ in range(K):
        minus_prev_over_dt = u_prev.copy()
        minus_prev_over_dt /= -dt
        lhs = M * (1.0 / dt) + A
        rhs = -(M @ minus_prev_over_dt + A @ zero - f)
        u = np.linalg.solve(lhs, rhs)
        sols.append(u)
        dots.append((u - u_prev) / dt)
        u_prev = u.copy()
    return np.array(sols), np.array(dots)


def parabolic_residual_norm_squared(u, udot, theta_m, theta_a, theta_f, G_ff, G_af, G_aa, G_mf, G_ma, G_mm, N):
    """Elliptic residual + the three time-derivative term pairs (m,f), (m,a), (m,m).

    rbnics/problems/parabolic/abstract_parabolic_rb_reduced_problem.py:62-92."""
    elliptic = residual_norm_squared(u, theta_a, theta_f, G_ff, G_af, G_aa, N)
    mf = product_order2(theta_m, [[g[:N] for g in row] for row in G_mf], theta_f)
    ma = product_order2(theta_m, [[g[:N, :N] for g in row] for row in G_ma], theta_a)
    mm = product_order2(theta_m, [[g[:N, :N] for g in row] for row in G_mm], theta_m)
    return float(elliptic + 2.0 * np.dot(udot, mf) + 2.0 * np.dot(udot, ma @ u) + np.dot(udot, mm @ udot))


def parabolic_estimate_error(eps2_over_time, beta, dt, initial_error_estimate_squared=0.0):
    """Error bound over time (0 at t0 for a homogeneous initial condition, sqrt(|eps2|/beta) afterwards) and its time
    quadrature sqrt(simpson(bound^2, dx=dt)).

    rbnics/problems/parabolic/abstract_parabolic_rb_reduced_problem.py:29-46,
    rbnics/reduction_methods/base/time_dependent_rb_reduction.py:280-288,
    rbnics/backends/common/time_quadrature.py:22-29 (scipy.integrate.simpson)."""
    from scipy.integrate import simpson
    bound = [sqrt(abs(initial_error_estimate_squared))] + [sqrt(abs(e) / beta) for e in eps2_over_time[1:]]
    return sqrt(simpson([v ** 2 for v in bound], dx=dt)), bound


def parabolic_sweep_loop(Theta_m, Theta_a, Theta_f, beta, M_q, A_q, f_q, G, N, dt, K, U0=None, init_err2=None):
    """The reference's POD-Greedy closure over a training set (serial).  G = dict of the six Gram storages
    ("ff", "af", "aa", "mf", "ma", "mm").  Returns (delta[B], argmax, eps2[B][K+1], U[B][K+1][N])."""
    B = Theta_a.shape[0]
    deltas = np.zeros(B)
    eps2_all = np.zeros((B, K + 1))
    U_all = np.zeros((B, K + 1, N))
    for p in range(B):
        tm = tuple(float(x) for x in Theta_m[p])
        ta = tuple(float(x) for x in Theta_a[p])
        tf = tuple(float(x) for x in Theta_f[p])
        sols, dots = parabolic_solve(tm, ta, tf, M_q, A_q, f_q, N, dt, K, None if U0 is None else U0[p])
        for k in range(1, K + 1):
            eps2_all[p, k] = parabolic_residual_norm_squared(sols[k], dots[k], tm, ta, tf, G["ff"], G["af"], G["aa"],
                                                             G["mf"], G["ma"], G["mm"], N)
        deltas[p], _ = parabolic_estimate_error(eps2_all[p], float(beta[p]), dt,
                                                0.0 if init_err2 is None else float(init_err2[p]))
        U_all[p] = sols
    return deltas, argmax_first(deltas), eps2_all, U_all


def gram_schmidt_row_partitioned(new, Z, X, row_begin, row_end, allreduce_sum, gather_rows):
    """What rb_off_gram_schmidt_rows does on one rank, in numpy (TEST INFRASTRUCTURE: the gloo test of the multi-rank
    orchestration runs this on CPU ranks).  Same order as gram_schmidt above (rbnics/backends/basic/gram_schmidt.py:22-36):
    for every old basis function one partial dot over the rank's dof rows + one all-reduce, then the projected rows are
    gathered and the norm is one more partial dot + all-reduce.  `allreduce_sum(x)` sums a float over the ranks,
    `gather_rows(v)` sums zero-padded vectors over the ranks."""
    z = np.array(new, dtype=np.float64)
    rows = slice(row_begin, row_end)
    Xr = X.tocsr()[rows] if X is not None else None
    for k in range(Z.shape[1]):
        b = Z[:, k]
        xb = Xr @ b if Xr is not None else b[rows]
        d = allreduce_sum(float(z[rows] @ xb))
        z[rows] = z[rows] - d * b[rows]
    piece = np.zeros_like(z)
    piece[rows] = z[rows]
    z = gather_rows(piece)
    xz = Xr @ z if Xr is not None else z[rows]
    nrm2 = allreduce_sum(float(z[rows] @ xz))
    nrm = sqrt(nrm2)
    if nrm != 0.0:
        z = z / nrm
    return z, nrm

/*
 * rbnics_b200.h -- C ABI of the B200-native greedy-sweep engine (drop-in boundary).
 *
 * Plain pointers and sizes only; the caller (the Python online backend `rbnics_b200`, via ctypes) owns every
 * host buffer, the library owns all device memory inside the handle.  One handle per GPU / process, not
 * thread-safe per handle (the reference is single-threaded per MPI rank).  Every function returns 0 on
 * success and a negative RB_ERR_* code otherwise; rb_last_error(ctx) then holds a message (the reference
 * raises Python exceptions at the same places; a singular reduced matrix raises numpy.linalg.LinAlgError in
 * rbnics/backends/online/numpy/linear_solver.py:29-31 and is RB_ERR_SINGULAR here).
 *
 * All matrices are FP64, row-major, dense.
 */
#ifndef RBNICS_B200_H
#define RBNICS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rb_ctx rb_ctx;

#define RB_OK 0
#define RB_ERR_CUDA (-1)
#define RB_ERR_ARG (-2)
#define RB_ERR_STATE (-3)
#define RB_ERR_SINGULAR (-4)
#define RB_ERR_UNSUPPORTED (-5)

/* Error-estimator normalisations.
 * RB_EST_SQRT_EPS2_OVER_BETA : sqrt(|eps2|)/beta   rbnics/problems/elliptic/elliptic_rb_reduced_problem.py:35-40
 * RB_EST_SQRT_OF_EPS2_OVER_BETA : sqrt(|eps2|/beta) rbnics/problems/elliptic/elliptic_coercive_compliant_rb_reduced_problem.py:22-27 */
#define RB_EST_SQRT_EPS2_OVER_BETA 0
#define RB_EST_SQRT_OF_EPS2_OVER_BETA 1

/* ---- handle ---------------------------------------------------------------------------------------------------- */
int rb_ctx_create(int device, rb_ctx** out);
void rb_ctx_destroy(rb_ctx* ctx);
const char* rb_last_error(const rb_ctx* ctx);
/* "major.minor" of the device the handle is bound to, and SM count (for bench bookkeeping). */
int rb_device_info(const rb_ctx* ctx, int* cc_major, int* cc_minor, int* sm_count);

/* ---- reduced system: replaces sum(product(theta, operator[term][:N,:N])) + LinearSolver -------------------------
 * rbnics/problems/elliptic/elliptic_reduced_problem.py:19-28 (matrix_eval / vector_eval),
 * rbnics/backends/online/basic/product.py:22-33 (order-1 product), rbnics/backends/online/basic/linear_solver.py:56-71
 * + wrapping/DirichletBC.py:37-98 (lifting rows), rbnics/backends/online/numpy/linear_solver.py:29-31 (dgesv).
 *
 *   lhs(mu) = sum_{q<Ql} Theta[q]      * A[q]   (A: Ql x N x N)
 *   rhs(mu) = sum_{q<Qr} Theta[Ql + q] * F[q]   (F: Qr x N)
 *   rows bc_rows[i] (i < n_bc): lhs[row,:] = 0, lhs[row,row] = 1, rhs[row] = ThetaBC[i]
 * Multi-term / block problems (Stokes a, b, bt; f, g) pass every term's block-embedded N x N matrix as one A[q].
 * N == 0 is legal (first greedy sweep: empty solution).                                                        */
int rb_set_system(rb_ctx* ctx, int N, int Ql, int Qr, const double* A, const double* F, int n_bc, const int* bc_rows);

/* ---- estimator: replaces get_residual_norm_squared / estimate_error ---------------------------------------------
 * rbnics/problems/elliptic/elliptic_rb_reduced_problem.py:55-64 (3 terms), rbnics/problems/stokes/
 * stokes_rb_reduced_problem.py:35-83 (9 terms), rbnics/backends/online/basic/product.py:34-46 (order-2 product).
 *
 * Every such residual norm is one quadratic form  eps2(mu) = z^T G z  with z the concatenation of nblk blocks
 *   z_b = scale_b * V[src_b : src_b + len_b],   V = [ u(0..N-1) | Theta[v_theta_off .. +v_theta_len) | 1.0 ],
 *   scale_b = Theta[blk_scale_col[b]]   (blk_scale_col[b] < 0  ->  scale_b = 1)
 * and G the dense (Kz x Kz, Kz = sum len_b) matrix of Riesz Gram blocks (need not be symmetric: only its symmetric
 * part contributes and the kernel uses G + G^T).  Elliptic: blocks b < Qa = (scale theta_a[b], src 0, len N) and one
 * block (scale 1, src N, len Qf) with v_theta = theta_f;  G = [[Gaa, Gaf], [Gaf^T, Gff]].
 * The form is evaluated as one dense GEMM over factor pairs per pair of block classes (rb_est_pair.cu); when its
 * shared-memory tile of V does not fit (N + Ql + Qr > ~210), or with RB_EST_KERNEL=fold in the environment at this
 * call, by the folded triangular kernel of rb_sweep.cu.  Both give the same eps2 to rounding.                        */
int rb_set_estimator(rb_ctx* ctx, int nblk, const int* blk_scale_col, const int* blk_src_off, const int* blk_len,
                     int v_theta_off, int v_theta_len, const double* G);

/* Same with an explicit length of the solution part of V (rb_set_estimator uses N): the parabolic estimator works on
 * V = [ u^k (N) | udot^k (N) | Theta[v_theta_off ..) | 1 ], solution_len = 2 N
 * (rbnics/problems/parabolic/abstract_parabolic_rb_reduced_problem.py:62-92).                                      */
int rb_set_estimator_ex(rb_ctx* ctx, int nblk, const int* blk_scale_col, const int* blk_src_off, const int* blk_len,
                        int v_theta_off, int v_theta_len, const double* G, int solution_len);

/* ---- training set: replaces ParameterSpaceSubset + compute_theta / get_stability_factor_lower_bound --------------
 * rbnics/sampling/parameter_space_subset.py:52-77.  Theta: B x (Ql+Qr) (theta of every lhs then rhs term, already
 * evaluated for every mu of this rank's shard), ThetaBC: B x n_bc or NULL, beta: B.
 * The local row i has global training-set index  index_offset + i*index_stride  (cyclic sharding of the reference:
 * offset = rank, stride = size, parameter_space_subset.py:57).  Copies host -> device.                            */
int rb_set_training_set(rb_ctx* ctx, int64_t B, int n_theta, const double* Theta, int n_bc, const double* ThetaBC,
                        const double* beta, int64_t index_offset, int64_t index_stride);

/* ---- the sweep: replaces training_set.max(solve_and_estimate_error) ----------------------------------------------
 * rbnics/reduction_methods/base/rb_reduction.py:160-185.  Runs assembly + LU solve + estimator + arg-max over the
 * resident training set.  max_idx is the GLOBAL index, ties -> lowest index, NaN wins (numpy.argmax,
 * parameter_space_subset.py:67,75).  delta_out (B) / u_out (B x N) / eps2_out (B) may be NULL.
 * Returns RB_ERR_SINGULAR if any estimator is non-finite (outputs are still written).                            */
int rb_sweep_resident(rb_ctx* ctx, int estimator_kind, double* max_val, int64_t* max_idx,
                      double* delta_out, double* u_out, double* eps2_out);
/* Same, from host buffers in one call (host->device copies inside): the end-to-end entry. */
int rb_sweep(rb_ctx* ctx, int64_t B, int n_theta, const double* Theta, int n_bc, const double* ThetaBC,
             const double* beta, int64_t index_offset, int64_t index_stride, int estimator_kind,
             double* max_val, int64_t* max_idx, double* delta_out, double* u_out, double* eps2_out);
/* ---- POD-Greedy sweep of a parabolic problem: replaces training_set.max(solve_and_estimate_error) of
 * rbnics/reduction_methods/base/time_dependent_rb_reduction.py:262-295 for the whole resident training set.
 * Per parameter: K implicit-Euler steps from a zero initial condition
 *     (M(mu)/dt + A(mu)) u^k = M(mu) u^{k-1}/dt + f(mu)       udot^k = (u^k - u^{k-1}) / dt
 * (rbnics/backends/online/numpy/time_stepping.py:104-112,226-238; abstract_parabolic_reduced_problem.py:17-36),
 * eps2_k = z_k^T G z_k with the estimator set by rb_set_estimator_ex(solution_len = 2N), bound_k = sqrt(|eps2_k|/beta)
 * (bound_0 = 0), Delta(mu) = sqrt(sum_k weights[k] * bound_k^2): weights = the time quadrature of
 * rbnics/backends/common/time_quadrature.py:22-29 (Simpson) as a linear functional, K + 1 entries.
 * Conventions: rb_set_system was called with the lhs terms [ M_q / dt (Qm of them) | A_q ] and the rhs terms f_q;
 * Theta columns [theta_m | theta_a | theta_f].  thetas are constant in time.
 * delta_out: B, eps2_out: B x (K+1), u_out: B x (K+1) x N (any may be NULL).                                      */
int rb_sweep_parabolic(rb_ctx* ctx, int Qm, int K, double dt, const double* weights, double* max_val,
                       int64_t* max_idx, double* delta_out, double* eps2_out, double* u_out);
/* The same with a non-homogeneous initial condition: U0 (B x N, row-major, may be NULL = zero) holds the reduced initial
 * condition of every parameter (rbnics/problems/base/time_dependent_reduced_problem.py:277-302 and the projection of
 * rbnics/backends/online/numpy/time_stepping.py:104-112), init_err2 (B, may be NULL = zero) its squared error estimate
 * get_initial_error_estimate_squared(), which is the bound at t0 (abstract_parabolic_rb_reduced_problem.py:41-44; no
 * division by beta) and enters the time quadrature with weights[0].                                                 */
int rb_sweep_parabolic_ic(rb_ctx* ctx, int Qm, int K, double dt, const double* weights, const double* U0,
                          const double* init_err2, double* max_val, int64_t* max_idx, double* delta_out,
                          double* eps2_out, double* u_out);

/* ---- output functional of the solutions of the last sweep (testing-set sweeps of error_analysis / speedup_analysis,
 * rbnics/reduction_methods/base/rb_reduction.py:312-390):  out[p] = u_p^T ( sum_q ThetaS[p][q] * S[q] ),
 * rbnics/problems/elliptic/elliptic_reduced_problem.py:31-32 (_compute_output).  S: Qs x N, ThetaS: B x Qs (host),
 * out: B (host).  Uses the solutions left on the device by rb_sweep / rb_sweep_resident.                          */
int rb_compute_outputs(rb_ctx* ctx, int Qs, const double* S, const double* ThetaS, double* out);
/* Reduced supremizers of B parameters at once (StokesReducedProblem._solve_supremizer,
 * rbnics/problems/stokes/stokes_reduced_problem.py:80-103): s_p = X_s^{-1} (sum_q Theta[p][q] Bt_q) u_p.  The caller
 * passes W[q] = X_s^{-1} Bt_q (Q x Ns x Nu, row-major; X_s does not depend on the parameter, so one host LU serves
 * all of them), Theta (B x Q), the reduced solutions U (B x Nu, or NULL for the solutions left in HBM by the last
 * sweep, which must have B parameters and N == Nu) and receives S (B x Ns).  Homogeneous lifting only.            */
int rb_supremizers(rb_ctx* ctx, int Ns, int Nu, int Q, const double* W, int64_t B, const double* Theta,
                   const double* U, double* S);

/* Device pointer to the 16-byte {double max_val; int64 max_idx} result of the last sweep (for a NCCL all-gather
 * issued by the host layer without a host round trip), and to the per-parameter estimators. */
int rb_result_device_ptrs(rb_ctx* ctx, void** maxloc16, double** delta);
/* Work of the estimator kernel per parameter as packed by rb_set_estimator: multiply-adds issued on the tensor pipe
 * (k-step records x 256, padding included) and number of 64-column tiles, of the kernel that serves the sweeps. */
int rb_estimator_work(rb_ctx* ctx, int64_t* issued_macs_per_param, int* n_column_tiles);
/* Kernel-only timings (ms, CUDA events on the handle's stream) of the last sweep:
 * [0]=assembly GEMM, [1]=LU solve, [2]=estimator GEMM, [3]=finalize+argmax, [4]=total; launches counted. */
int rb_last_timings(rb_ctx* ctx, float* ms5, int* n_launches);

/* ---- single-parameter entries: the per-mu API of the reference's online backend ---------------------------------
 * out = sum_q c[q] * ops[q]  for one parameter, in the reference's order AND rounding (multiply, then add, no FMA;
 * the first term always, later terms skipped when c[q] == 0): rbnics/backends/online/basic/product.py:22-33.
 * ops: Q x numel.  Bit-identical to the numpy backend.                                                          */
int rb_affine_combine(rb_ctx* ctx, int Q, int64_t numel, const double* coeffs, const double* ops, double* out);
/* out = sum_{i,j} th1[i] * ops[i][j] * th2[j], row-major (i, j) order, ((th1*op)*th2) rounding, skipping terms with a
 * zero theta except (0,0): rbnics/backends/online/basic/product.py:34-46.  ops: Q1 x Q2 x numel.                 */
int rb_affine_combine2(rb_ctx* ctx, int Q1, int Q2, int64_t numel, const double* th1, const double* th2,
                       const double* ops, double* out);
/* x = lhs^-1 rhs for ONE dense N x N system after applying the lifting rows (lhs[row,:] = e_row, rhs[row] = bc_vals):
 * LinearSolver.solve(), rbnics/backends/online/numpy/linear_solver.py:29-33 + online/basic/linear_solver.py:56-71.
 * RB_ERR_SINGULAR if the solution is not finite.                                                                  */
int rb_solve_dense(rb_ctx* ctx, int N, const double* lhs, const double* rhs, int n_bc, const int* bc_rows,
                   const double* bc_vals, double* x);
/* *out = u^T (M v), M: m x n row-major (M == NULL: u^T v with m == n): transpose(u)*M*v / transpose(u)*v,
 * rbnics/backends/basic/transpose.py:126-237.                                                                     */
int rb_bilinear(rb_ctx* ctx, int m, int n, const double* u, const double* M, const double* v, double* out);

/* Pinned host memory helpers for the end-to-end path. */
void* rb_host_alloc(int64_t bytes);
void rb_host_free(void* p);

/* ---- offline basis linear algebra -------------------------------------------------------------------------------
 * C = S^T X S2: rbnics/backends/basic/transpose.py:304-313 (POD correlation), :441-468 (Z^T A_q Z), :365-400 (Z^T f).
 * S: Nh x Ns, S2: Nh x Ns2 row-major (dof-major: row = dof, column = snapshot), X: CSR (Nh x Nh) or indptr == NULL
 * for the identity.  C: Ns x Ns2 row-major.  If S2 == NULL, S2 = S.  Rows [row_begin, row_end) of X/S are this
 * rank's dof partition (the partial result is all-reduced by the host layer); pass 0, Nh for a single GPU.
 * For S2 == S with more than 128 columns and X == X^T (checked on the device) only the 128-column tiles on and above
 * the diagonal are computed and mirrored: a rank's partial then holds the transposed upper tiles of ITS partial
 * below the diagonal tiles -- the sum over the ranks is the exact (symmetric) product, a single partial is not the
 * rank's true contribution there.                                                                                 */
int rb_gram(rb_ctx* ctx, int64_t Nh, int Ns, int Ns2, const double* S, const double* S2, const int64_t* indptr,
            const int32_t* indices, const double* data, int64_t row_begin, int64_t row_end, double* C);
/* Batched projection of Q operators sharing one basis: out[q] = Z^T A_q Z (Q x N x N); CSR arrays concatenated,
 * indptr_all has Q*(Nh+1) entries each starting at 0 relative to its own nnz_off[q].                              */
int rb_project(rb_ctx* ctx, int64_t Nh, int N, int Q, const double* Z, const int64_t* indptr_all,
               const int32_t* indices_all, const double* data_all, const int64_t* nnz_off, double* out);

/* Z = S V : S Nh x Ns, V Ns x n, Z Nh x n (row-major).  Mode construction b_n = S * eigenvector_n of the POD
 * (rbnics/backends/basic/proper_orthogonal_decomposition_base.py:77-79, dolfin/wrapping/functions_list_mul.py:10-36)
 * and reconstruction Z u_N (basis_functions_matrix_mul.py:10-46).                                                 */
int rb_lincomb(rb_ctx* ctx, int64_t Nh, int Ns, int n, const double* S, const double* V, double* Z);
/* One single-pass modified Gram-Schmidt step in the X inner product (identity if indptr == NULL):
 *   for k < n: z -= (z^T X b_k) b_k ;  norm = sqrt(z^T X z) ;  if norm != 0: z /= norm
 * rbnics/backends/basic/gram_schmidt.py:22-36, dolfin/wrapping/gram_schmidt_projection_step.py:8-11.
 * Zb: Nh x n row-major (may be NULL when n == 0); z: Nh, updated in place; *norm_out = the norm before scaling.  */
int rb_gram_schmidt(rb_ctx* ctx, int64_t Nh, int n, const double* Zb, const int64_t* indptr, const int32_t* indices,
                    const double* data, double* z, double* norm_out);

/* ---- device-resident offline workspace -------------------------------------------------------------------------
 * The reference rebuilds every reduced operator from the truth vectors at every greedy iteration
 * (rbnics/reduction_methods/base/rb_reduction.py:106,112).  Here the sparse operators (X, A_q) and the dense column
 * blocks (basis Z, snapshots S, Riesz representers) stay in HBM in numbered slots (0..127 each kind); only new columns
 * cross PCIe.  Blocks are dof-major: Nh rows, up to `capacity` columns.                                            */
int rb_off_csr(rb_ctx* ctx, int slot, int64_t Nh, const int64_t* indptr, const int32_t* indices, const double* data);
int rb_off_block(rb_ctx* ctx, int slot, int64_t Nh, int capacity);             /* (re)create an empty block            */
int rb_off_append(rb_ctx* ctx, int slot, int n, const double* cols);
/* The same for a rank of the dof-row partitioned offline path: only the rows [row_lo, row_hi) of `cols` (still the
 * address of the FULL Nh x n array) are transferred -- the rank's own rows plus the halo its operator rows reference
 * (PETSc keeps exactly those ghost entries for MatMult / VecDot on the local rows of the dolfin backend);
 * the other rows of the new columns are zero on this device.                                                          */
int rb_off_append_rows(rb_ctx* ctx, int slot, int n, const double* cols, int64_t row_lo, int64_t row_hi);           /* cols: Nh x n row-major on the host   */
int rb_off_size(rb_ctx* ctx, int slot, int64_t* Nh, int* ncols);
int rb_off_read(rb_ctx* ctx, int slot, int c0, int n, double* out);            /* out: Nh x n row-major                */
/* C (ns x ns2, host) = S[:, s0:s0+ns]^T X S2[:, t0:t0+ns2] over the dof rows [row_begin, row_end); csrX < 0: identity
 * (rbnics/backends/basic/transpose.py:304-313, 441-468, 365-400).                                                  */
int rb_off_gram(rb_ctx* ctx, int blkS, int s0, int ns, int csrX, int blkS2, int t0, int ns2, int64_t row_begin,
                int64_t row_end, double* C);
/* Y[:, y0:y0+n] = A * S[:, s0:s0+n] on resident operands (block blkY grows to y0+n columns if shorter): the operator
 * applied to a few new columns, e.g. A_q z_new for all q side by side, so that ONE rb_off_gram(Z, identity, Y) gives
 * the new column of every reduced operator Z^T A_q z_new (parametrized_reduced_differential_problem.py:727-790 computes
 * them one inner product at a time).                                                                              */
int rb_off_spmm(rb_ctx* ctx, int csrA, int blkS, int s0, int n, int blkY, int y0);
/* z (host, Nh, in/out) orthonormalised against ALL columns of block blkZ like rb_gram_schmidt; append != 0 also
 * appends the result to the block (RBReduction.update_basis_matrix, rb_reduction.py:130-142).                       */
int rb_off_gram_schmidt(rb_ctx* ctx, int blkZ, int csrX, double* z, int append, double* norm_out);

/* Resident POD (rbnics/backends/basic/proper_orthogonal_decomposition_base.py:39-92): modes Z[:, z0:z0+n] =
 * S[:, s0:s0+ns] V (V: ns x n row-major on the host; dolfin/wrapping/functions_list_mul.py:26-36), their squared
 * X-norms z_j^T X z_j from the DIAGONAL only (:80-87; csrX < 0: identity; dof rows [row_begin, row_end), summed over the
 * ranks when rb_set_allreduce is set), and the column scaling z_j *= scale[j].                                        */
int rb_off_lincomb(rb_ctx* ctx, int blkS, int s0, int ns, const double* V, int n, int blkZ, int z0);
/* accumulate != 0: Z[:, z0:z0+n] += S[:, s0:s0+ns] V on existing columns of Z -- `snapshot - basis_functions *
 * projected_snapshot_N` of the POD-Greedy orthogonalisation (time_dependent_rb_reduction.py:178-189) on resident data */
int rb_off_lincomb_add(rb_ctx* ctx, int blkS, int s0, int ns, const double* V, int n, int blkZ, int z0, int accumulate);
int rb_off_column_norms2(rb_ctx* ctx, int blkZ, int z0, int n, int csrX, int64_t row_begin, int64_t row_end, double* out);
int rb_off_scale_columns(rb_ctx* ctx, int blkZ, int z0, int n, const double* scale);

/* ---- multi-GPU offline path: dof rows partitioned over the ranks ---------------------------------
 * The reference runs these products under PETSc's row partition: every VecDot is an MPI all-reduce
 * (rbnics/backends/dolfin/wrapping/vector_mul.py:8-9; N^2 of them per Gram matrix, one per old basis function in
 * Gram-Schmidt, rbnics/backends/basic/gram_schmidt.py:29-33).  Here every rank computes its partial over
 * [row_begin, row_end) and the library calls back ONCE per Gram matrix / ONCE per Gram-Schmidt projection to sum a
 * DEVICE buffer over the ranks in place (the host side does it with NCCL on that very pointer: no staging copy).
 * fn == NULL (default): single rank.  fn returns 0 on success; it is called with the context's stream idle.        */
typedef int (*rb_allreduce_fn)(void* user, double* device_buffer, int64_t count);
int rb_set_allreduce(rb_ctx* ctx, rb_allreduce_fn fn, void* user);
/* rb_off_gram_schmidt on the rank's dof rows: the projections z -= (z^T X b_k) b_k keep the reference's order (one
 * all-reduce each, of the per-CTA partial sums, so every rank subtracts the bit-identical coefficient); the rows of z
 * owned by other ranks are then gathered (sum of zero-padded pieces) for the norm sqrt(z^T X z).  z (host, full length)
 * is returned complete on every rank.                                                                             */
int rb_off_gram_schmidt_rows(rb_ctx* ctx, int blkZ, int csrX, double* z, int append, int64_t row_begin,
                             int64_t row_end, double* norm_out);
/* The same with the first `first_col` columns of the block left out of the projection: lifting functions of
 * non-homogeneous Dirichlet conditions stay in the basis but are not orthogonalised against --
 * GS.apply(new, basis_functions[N_bc:]), rbnics/reduction_methods/base/rb_reduction.py:125-139.                     */
int rb_off_gram_schmidt_from(rb_ctx* ctx, int blkZ, int csrX, double* z, int append, int first_col, int64_t row_begin,
                             int64_t row_end, double* norm_out);

#ifdef __cplusplus
}
#endif
#endif

// rb_online.cu -- single-parameter entries of the online backend: affine sums with the reference's exact rounding,
// one dense solve, bilinear forms.  These are the per-mu calls of rbnics/backends/online/numpy (product, sum,
// LinearSolver, transpose); the greedy loop itself uses the batched sweep of rb_sweep.cu instead.
#include <algorithm>
#include <vector>

#include "rb_lu_kernel.cuh"

int rb_launch_lu_args(rb_ctx* ctx, const LuArgs& a, int NP);

namespace {
struct Dev {
  void* p = nullptr;
  ~Dev() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
  double* d() { return reinterpret_cast<double*>(p); }
};
}  // namespace

// out = c0*op0 (+ c_q*op_q for c_q != 0), separately rounded multiply and add like numpy's  out += theta * operator
__global__ void rb_affine_combine_kernel(int Q, long long numel, const double* __restrict__ c,
                                         const double* __restrict__ ops, double* __restrict__ out) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < numel; e += (long long)gridDim.x * blockDim.x) {
    double acc = __dmul_rn(c[0], ops[e]);
    for (int q = 1; q < Q; ++q) {
      const double cq = c[q];
      if (cq != 0.0) acc = __dadd_rn(acc, __dmul_rn(cq, ops[(long long)q * numel + e]));
    }
    out[e] = acc;
  }
}

__global__ void rb_affine_combine2_kernel(int Q1, int Q2, long long numel, const double* __restrict__ t1,
                                          const double* __restrict__ t2, const double* __restrict__ ops,
                                          double* __restrict__ out) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < numel; e += (long long)gridDim.x * blockDim.x) {
    double acc = __dmul_rn(__dmul_rn(t1[0], ops[e]), t2[0]);
    for (int i = 0; i < Q1; ++i)
      for (int j = 0; j < Q2; ++j) {
        if (i == 0 && j == 0) continue;
        if (t1[i] != 0.0 && t2[j] != 0.0)
          acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(t1[i], ops[((long long)i * Q2 + j) * numel + e]), t2[j]));
      }
    out[e] = acc;
  }
}

// one block: y = M v (row per thread, strided), then u^T y with a fixed-order tree
__global__ void __launch_bounds__(256) rb_bilinear_kernel(int m, int n, const double* __restrict__ u,
                                                          const double* __restrict__ M, const double* __restrict__ v,
                                                          double* __restrict__ out) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < m; i += 256) {
    double y;
    if (M) {
      y = 0.0;
      for (int j = 0; j < n; ++j) y = fma(M[(long long)i * n + j], v[j], y);
    } else {
      y = v[i];
    }
    s = fma(u[i], y, s);
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

extern "C" int rb_affine_combine(rb_ctx* ctx, int Q, int64_t numel, const double* coeffs, const double* ops,
                                 double* out) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (Q <= 0 || numel < 0 || !coeffs || (numel > 0 && (!ops || !out))) RB_FAIL(RB_ERR_ARG, "rb_affine_combine: bad arguments");
  if (numel == 0) return RB_OK;
  cudaStream_t st = ctx->stream;
  Dev dc, dop, dout;
  RB_CUDA(dc.alloc(Q * sizeof(double)));
  RB_CUDA(dop.alloc((size_t)Q * numel * sizeof(double)));
  RB_CUDA(dout.alloc((size_t)numel * sizeof(double)));
  RB_CUDA(cudaMemcpyAsync(dc.p, coeffs, Q * sizeof(double), cudaMemcpyHostToDevice, st));
  RB_CUDA(cudaMemcpyAsync(dop.p, ops, (size_t)Q * numel * sizeof(double), cudaMemcpyHostToDevice, st));
  const int blocks = (int)std::min<long long>((numel + 255) / 256, 148 * 8);
  rb_affine_combine_kernel<<<blocks, 256, 0, st>>>(Q, numel, dc.d(), dop.d(), dout.d());
  RB_CUDA(cudaGetLastError());
  RB_CUDA(cudaMemcpyAsync(out, dout.p, (size_t)numel * sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaStreamSynchronize(st));
  ctx->last_launches = 1;
  return RB_OK;
}

extern "C" int rb_affine_combine2(rb_ctx* ctx, int Q1, int Q2, int64_t numel, const double* th1, const double* th2,
                                  const double* ops, double* out) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (Q1 <= 0 || Q2 <= 0 || numel < 0 || !th1 || !th2 || (numel > 0 && (!ops || !out)))
    RB_FAIL(RB_ERR_ARG, "rb_affine_combine2: bad arguments");
  if (numel == 0) return RB_OK;
  cudaStream_t st = ctx->stream;
  Dev d1, d2, dop, dout;
  RB_CUDA(d1.alloc(Q1 * sizeof(double)));
  RB_CUDA(d2.alloc(Q2 * sizeof(double)));
  RB_CUDA(dop.alloc((size_t)Q1 * Q2 * numel * sizeof(double)));
  RB_CUDA(dout.alloc((size_t)numel * sizeof(double)));
  RB_CUDA(cudaMemcpyAsync(d1.p, th1, Q1 * sizeof(double), cudaMemcpyHostToDevice, st));
  RB_CUDA(cudaMemcpyAsync(d2.p, th2, Q2 * sizeof(double), cudaMemcpyHostToDevice, st));
  RB_CUDA(cudaMemcpyAsync(dop.p, ops, (size_t)Q1 * Q2 * numel * sizeof(double), cudaMemcpyHostToDevice, st));
  const int blocks = (int)std::min<long long>((numel + 255) / 256, 148 * 8);
  rb_affine_combine2_kernel<<<blocks, 256, 0, st>>>(Q1, Q2, numel, d1.d(), d2.d(), dop.d(), dout.d());
  RB_CUDA(cudaGetLastError());
  RB_CUDA(cudaMemcpyAsync(out, dout.p, (size_t)numel * sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaStreamSynchronize(st));
  ctx->last_launches = 1;
  return RB_OK;
}

extern "C" int rb_solve_dense(rb_ctx* ctx, int N, const double* lhs, const double* rhs, int n_bc, const int* bc_rows,
                              const double* bc_vals, double* x) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (N < 0 || n_bc < 0 || n_bc > N) RB_FAIL(RB_ERR_ARG, "rb_solve_dense: bad sizes");
  if (N == 0) return RB_OK;
  if (!lhs || !rhs || !x || (n_bc > 0 && (!bc_rows || !bc_vals))) RB_FAIL(RB_ERR_ARG, "rb_solve_dense: null pointer");
  const int NP = (N + 3) / 4 * 4;
  if (NP > LU_MAX_NP) RB_FAIL(RB_ERR_UNSUPPORTED, "rb_solve_dense: N exceeds the supported maximum (160)");
  const int EP = (NP * NP + ASM_BE - 1) / ASM_BE * ASM_BE;
  std::vector<double> Ab((size_t)EP, 0.0), Fp((size_t)NP, 0.0);
  for (int i = 0; i < N; ++i) {
    for (int j = 0; j < N; ++j) Ab[(size_t)j * NP + i] = lhs[(size_t)i * N + j];
    Fp[i] = rhs[i];
  }
  for (int t = 0; t < n_bc; ++t)
    if (bc_rows[t] < 0 || bc_rows[t] >= N) RB_FAIL(RB_ERR_ARG, "rb_solve_dense: bc row out of range");
  const double one = 1.0;
  cudaStream_t st = ctx->stream;
  Dev dAb, dF, dTh, dBcR, dBcV, dU;
  RB_CUDA(dAb.alloc(Ab.size() * sizeof(double)));
  RB_CUDA(dF.alloc(Fp.size() * sizeof(double)));
  RB_CUDA(dTh.alloc(sizeof(double)));
  RB_CUDA(dBcR.alloc(std::max(n_bc, 1) * sizeof(int)));
  RB_CUDA(dBcV.alloc(std::max(n_bc, 1) * sizeof(double)));
  RB_CUDA(dU.alloc(N * sizeof(double)));
  RB_CUDA(cudaMemcpyAsync(dAb.p, Ab.data(), Ab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  RB_CUDA(cudaMemcpyAsync(dF.p, Fp.data(), Fp.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  RB_CUDA(cudaMemcpyAsync(dTh.p, &one, sizeof(double), cudaMemcpyHostToDevice, st));
  if (n_bc) {
    RB_CUDA(cudaMemcpyAsync(dBcR.p, bc_rows, n_bc * sizeof(int), cudaMemcpyHostToDevice, st));
    RB_CUDA(cudaMemcpyAsync(dBcV.p, bc_vals, n_bc * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  LuArgs a;
  a.Ab = dAb.d();
  a.theta = dTh.d();
  a.thetaBC = dBcV.d();
  a.F = dF.d();
  a.bc_rows = reinterpret_cast<const int*>(dBcR.p);
  a.U = dU.d();
  a.count = 1;
  a.N = N;
  a.EP = EP;
  a.QT = 1;
  a.Ql = 0;
  a.Qr = 1;
  a.n_bc = n_bc;
  a.ldu = N;
  if (int rc = rb_launch_lu_args(ctx, a, NP)) return rc;
  RB_CUDA(cudaMemcpyAsync(x, dU.p, N * sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaStreamSynchronize(st));
  ctx->last_launches = 1;
  for (int i = 0; i < N; ++i)
    if (!(x[i] - x[i] == 0.0)) RB_FAIL(RB_ERR_SINGULAR, "rb_solve_dense: singular matrix (numpy.linalg.LinAlgError in the reference)");
  return RB_OK;
}

extern "C" int rb_bilinear(rb_ctx* ctx, int m, int n, const double* u, const double* M, const double* v, double* out) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (m < 0 || n < 0 || !out || (m > 0 && !u) || (n > 0 && !v) || (!M && m != n)) RB_FAIL(RB_ERR_ARG, "rb_bilinear: bad arguments");
  if (m == 0 || n == 0) {
    *out = 0.0;
    return RB_OK;
  }
  cudaStream_t st = ctx->stream;
  Dev du, dM, dv, dout;
  RB_CUDA(du.alloc(m * sizeof(double)));
  RB_CUDA(dv.alloc(n * sizeof(double)));
  RB_CUDA(dout.alloc(sizeof(double)));
  RB_CUDA(cudaMemcpyAsync(du.p, u, m * sizeof(double), cudaMemcpyHostToDevice, st));
  RB_CUDA(cudaMemcpyAsync(dv.p, v, n * sizeof(double), cudaMemcpyHostToDevice, st));
  if (M) {
    RB_CUDA(dM.alloc((size_t)m * n * sizeof(double)));
    RB_CUDA(cudaMemcpyAsync(dM.p, M, (size_t)m * n * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  rb_bilinear_kernel<<<1, 256, 0, st>>>(m, n, du.d(), M ? dM.d() : nullptr, dv.d(), dout.d());
  RB_CUDA(cudaGetLastError());
  RB_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaStreamSynchronize(st));
  ctx->last_launches = 1;
  return RB_OK;
}

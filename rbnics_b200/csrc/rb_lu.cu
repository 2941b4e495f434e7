// rb_lu.cu -- dispatcher of the batched LU solve + shared-memory fallback for NP > LU_MAX_NP_REG.
#include <stdlib.h>

#include "rb_lu_kernel.cuh"

int rb_lu_launch_unit0(int NP, const LuArgs& args, cudaStream_t stream);
int rb_lu_mw_launch(const LuArgs& args, double* scratch, cudaStream_t stream);
long long rb_lu_mw_scratch_doubles(int N, long long count);
#ifdef RB_AB_VARIANTS
int rb_lu_launch_unit1(int NP, const LuArgs& args, cudaStream_t stream);
int rb_lu_launch_unit2(int NP, const LuArgs& args, cudaStream_t stream);
int rb_lu_launch_unit3(int NP, const LuArgs& args, cudaStream_t stream);
int rb_lu_dmma_dispatch(int N, const LuArgs& args, cudaStream_t stream);
int rb_lu_left_launch(const LuArgs& args, double* scratch, cudaStream_t stream);
long long rb_lu_left_scratch_doubles(int N);
#endif

// Which kernel serves 32 < N <= 128: "mw" (default; up to four warps per system, left-looking, persistent CTAs whose
// L records stay in L2, rb_lu_mw.cu).  Builds with -DRB_AB_VARIANTS (make AB=1) also carry the round-1 kernels for
// A/B profiling, selected by RB_LU_KERNEL: "left" (one warp per system, left-looking, 8 systems per SM),
// "dmma" (register-resident, 2 systems per SM), "rowreg" (row per thread).
// Two right-looking variants that kept the trailing matrix in global memory (thread-per-row and warp-per-system with
// DMMA updates) were measured at the same 0.27 us/system as "dmma" (read-modify-write latency through L2) and removed;
// so was a two-warps-per-system variant of "left" (rows / tile rows split between the warps, one block barrier per pivot
// step): 7 % slower than one warp per system, the barriers cost more than the shorter chain saves.
int rb_lu_variant() {
  static const int v = [] {
#ifdef RB_AB_VARIANTS
    const char* e = getenv("RB_LU_KERNEL");
    if (e && e[0] == 'd') return 1;
    if (e && e[0] == 'r') return 2;
    if (e && e[0] == 'l') return 3;
#endif
    return 0;
  }();
  return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// Fallback: matrix in shared memory ([NP][NP+1], odd pitch), one CTA of 128 threads per system, explicit row swaps.
// Same arithmetic (idamax pivot, reciprocal-scaled multipliers, rank-1 updates); used only for NP > 104.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) rb_lu_smem_kernel(LuArgs args, int NP) {
  extern __shared__ __align__(16) unsigned char lu_smem_raw[];
  const int P = NP + 1;
  double* A = reinterpret_cast<double*>(lu_smem_raw);  // [NP][P]
  double* bs = A + (size_t)NP * P;                     // [NP]
  double* xs = bs + NP;                                // [NP]
  __shared__ unsigned long long wkey[4];
  __shared__ int wrow[4];
  const long long sys = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int N = args.N;
  const double* Ab = args.Ab + sys * (long long)args.EP;
  for (int e = tid; e < NP * NP; e += 128) {
    const int j = e / NP, i = e % NP;
    A[i * P + j] = Ab[e];
  }
  const double* th = args.theta + sys * (long long)args.QT + args.Ql;
  for (int i = tid; i < NP; i += 128) {
    double b = 0.0;
    if (i < N)
    {
      if (args.rhs_in)
        b = args.rhs_in[sys * (long long)N + i];
      else
        for (int q = 0; q < args.Qr; ++q) b = fma(th[q], args.F[q * NP + i], b);
    }
    bs[i] = b;
  }
  __syncthreads();
  for (int i = tid; i < NP; i += 128) {
    bool unit_row = (i >= N);
    for (int t = 0; t < args.n_bc; ++t)
      if (args.bc_rows[t] == i) {
        unit_row = true;
        bs[i] = args.thetaBC[sys * (long long)args.n_bc + t];
      }
    if (unit_row)
      for (int j = 0; j < NP; ++j) A[i * P + j] = (j == i) ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int k = 0; k < NP; ++k) {
    // pivot search over rows k..NP-1 (lowest row on ties)
    unsigned long long key = 0ull;
    int row = 0x7fffffff;
    for (int r = k + tid; r < NP; r += 128) {
      const unsigned long long kr = (unsigned long long)__double_as_longlong(fabs(A[r * P + k])) + 1ull;
      if (kr > key) {
        key = kr;
        row = r;
      }
    }
    for (int off = 16; off > 0; off >>= 1) {
      const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, off);
      const int r2 = __shfl_xor_sync(0xffffffffu, row, off);
      if (k2 > key || (k2 == key && r2 < row)) {
        key = k2;
        row = r2;
      }
    }
    if (lane == 0) {
      wkey[w] = key;
      wrow[w] = row;
    }
    __syncthreads();
    key = wkey[0];
    row = wrow[0];
    for (int w2 = 1; w2 < 4; ++w2)
      if (wkey[w2] > key || (wkey[w2] == key && wrow[w2] < row)) {
        key = wkey[w2];
        row = wrow[w2];
      }
    if (row != k) {
      for (int j = tid; j < NP; j += 128) {
        const double t0 = A[k * P + j];
        A[k * P + j] = A[row * P + j];
        A[row * P + j] = t0;
      }
      if (tid == 0) {
        const double t0 = bs[k];
        bs[k] = bs[row];
        bs[row] = t0;
      }
    }
    __syncthreads();
    const double rinv = 1.0 / A[k * P + k];
    const double bk = bs[k];
    for (int r = k + 1 + tid; r < NP; r += 128) {
      const double l = A[r * P + k] * rinv;
      for (int j = k + 1; j < NP; ++j) A[r * P + j] = fma(-l, A[k * P + j], A[r * P + j]);
      bs[r] = fma(-l, bk, bs[r]);
    }
    __syncthreads();
  }
  if (w == 0) {
    for (int k = NP - 1; k >= 0; --k) {
      double xk = 0.0;
      if (lane == 0) {
        xk = bs[k] / A[k * P + k];
        xs[k] = xk;
      }
      xk = __shfl_sync(0xffffffffu, xk, 0);
      for (int r = lane; r < k; r += 32) bs[r] = fma(-A[r * P + k], xk, bs[r]);
      __syncwarp();
    }
    for (int r = lane; r < N; r += 32) args.U[sys * (long long)args.ldu + r] = xs[r];
  }
}

int rb_launch_lu_args(rb_ctx* ctx, const LuArgs& a, int NP) {
  int rc;
  const int variant = rb_lu_variant();
  auto ensure_scratch = [&](long long need) -> int {
    if (need > ctx->lu_scratch_cap) {
      if (ctx->dLuScratch) cudaFree(ctx->dLuScratch);
      ctx->dLuScratch = nullptr;
      ctx->lu_scratch_cap = 0;
      cudaError_t e = cudaMalloc((void**)&ctx->dLuScratch, (size_t)need * sizeof(double));
      if (e != cudaSuccess) {
        ctx->err = std::string("LU scratch allocation: ") + cudaGetErrorString(e);
        return RB_ERR_CUDA;
      }
      ctx->lu_scratch_cap = need;
    }
    return RB_OK;
  };
  if (NP > 32 && a.N <= 128 && variant == 0) {
    if (int rs = ensure_scratch(rb_lu_mw_scratch_doubles(a.N, a.count))) return rs;
    rc = rb_lu_mw_launch(a, ctx->dLuScratch, ctx->stream);
  }
#ifdef RB_AB_VARIANTS
  else if (NP > 32 && a.N <= 128 && variant == 3) {
    if (int rs = ensure_scratch(a.count * rb_lu_left_scratch_doubles(a.N))) return rs;
    rc = rb_lu_left_launch(a, ctx->dLuScratch, ctx->stream);
  } else if (NP > 32 && NP <= LU_MAX_NP_REG && variant == 1)
    rc = rb_lu_dmma_dispatch(a.N, a, ctx->stream);
  else if (NP > 32 && NP <= 52)
    rc = rb_lu_launch_unit0(NP, a, ctx->stream);
  else if (NP > 32 && NP <= 80)
    rc = rb_lu_launch_unit1(NP, a, ctx->stream);
  else if (NP > 32 && NP <= 96)
    rc = rb_lu_launch_unit2(NP, a, ctx->stream);
  else if (NP > 32 && NP <= LU_MAX_NP_REG)
    rc = rb_lu_launch_unit3(NP, a, ctx->stream);
#endif
  else if (NP <= 32)
    rc = rb_lu_launch_unit0(NP, a, ctx->stream);
  else {
    const size_t bytes = ((size_t)NP * (NP + 1) + 2 * (size_t)NP) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(rb_lu_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      ctx->err = std::string("rb_lu_smem_kernel smem attribute: ") + cudaGetErrorString(e);
      return RB_ERR_CUDA;
    }
    rb_lu_smem_kernel<<<(unsigned)a.count, 128, bytes, ctx->stream>>>(a, NP);
    rc = (int)cudaGetLastError();
  }
  if (rc != 0) {
    ctx->err = std::string("LU launch failed: ") +
               (rc == -1000 ? "no instantiation for this NP" : cudaGetErrorString((cudaError_t)rc));
    return RB_ERR_CUDA;
  }
  return RB_OK;
}

// same with a per-system right-hand side (time stepping) and an explicit pitch of the solution array
int rb_launch_lu_rhs(rb_ctx* ctx, int64_t count, const double* Ab, const double* theta, const double* thetaBC,
                     const double* rhs_in, double* u, int ldu) {
  LuArgs a;
  a.Ab = Ab;
  a.theta = theta;
  a.thetaBC = thetaBC;
  a.F = ctx->dF;
  a.bc_rows = ctx->dBcRows;
  a.U = u;
  a.rhs_in = rhs_in;
  a.count = count;
  a.N = ctx->N;
  a.EP = ctx->EP;
  a.QT = ctx->QT;
  a.Ql = ctx->Ql;
  a.Qr = ctx->Qr;
  a.n_bc = ctx->n_bc;
  a.ldu = ldu;
  return rb_launch_lu_args(ctx, a, ctx->NP);
}

int rb_launch_lu(rb_ctx* ctx, int64_t count, const double* Ab, const double* theta, const double* thetaBC, double* u) {
  LuArgs a;
  a.Ab = Ab;
  a.theta = theta;
  a.thetaBC = thetaBC;
  a.F = ctx->dF;
  a.bc_rows = ctx->dBcRows;
  a.U = u;
  a.count = count;
  a.N = ctx->N;
  a.EP = ctx->EP;
  a.QT = ctx->QT;
  a.Ql = ctx->Ql;
  a.Qr = ctx->Qr;
  a.n_bc = ctx->n_bc;
  a.ldu = ctx->ldu;
  return rb_launch_lu_args(ctx, a, ctx->NP);
}

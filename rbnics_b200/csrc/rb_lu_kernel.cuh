// rb_lu_kernel.cuh -- batched dense LU (partial pivoting) + triangular solves for the reduced systems.
//
// Replaces numpy.linalg.solve (LAPACK dgesv) of rbnics/backends/online/numpy/linear_solver.py:29-31 and the lifting
// rows of rbnics/backends/online/basic/wrapping/DirichletBC.py:37-56, for a whole chunk of parameters at once.
//
// Register-resident, row-per-thread: thread i of a system owns row i of the (padded) NP x NP matrix and its rhs entry
// in registers; the k loop is fully unrolled so every register index is static.  Pivoting is implicit (no row is ever
// moved): the pivot row of step k is the not-yet-used row with the largest |a[i][k]|, lowest row index on ties
// (idamax semantics).  Pivot search = two REDUX.MAX + one ballot per warp on the IEEE bit pattern of |a|, the (up to 4)
// warp winners publish their row to shared memory, ONE named barrier per step, then every remaining row eliminates
// with the overall winner's row (broadcast LDS).  The used pivot rows are kept in shared memory in pivot order (that
// is U of PA = LU, with the forward-eliminated rhs), and one warp does the back-substitution with shuffles.
// NP <= 32: one warp per system, 4 systems per CTA, only __syncwarp() barriers.
#pragma once
#include "rb_common.cuh"

struct LuArgs {
  const double* Ab;       // [count][EP], system s entry (i,j) at s*EP + j*NP + i
  const double* theta;    // [count][QT]
  const double* thetaBC;  // [count][n_bc] or null
  const double* F;        // [Qr][NP]
  const int* bc_rows;     // [n_bc]
  double* U;              // [count][ldu]
  const double* rhs_in = nullptr;  // optional [count][N]: per-system right-hand side (replaces sum theta_f F)
  long long count;
  int N, EP, QT, Ql, Qr, n_bc, ldu;
};

template <int NP, int T>
struct LuSmem {
  static constexpr int NW = T / 32;
  static constexpr int CAND_PITCH = NP + 2;
  static constexpr int U_PITCH = NP + 1;
  static constexpr int CAND_DOUBLES = 2 * NW * CAND_PITCH;
  static constexpr int KEY_DOUBLES = 2 * NW;  // u64 keys
  static constexpr int DOUBLES = CAND_DOUBLES + KEY_DOUBLES + NP * U_PITCH + 2 * NP;
  static constexpr int BYTES_PER_SYS = ((DOUBLES * 8 + 15) / 16) * 16;
  static constexpr int SPC = 128 / T;
  static constexpr int BYTES = BYTES_PER_SYS * SPC;
};

template <int T>
__device__ __forceinline__ void lu_sys_sync(int sys_in_cta) {
  if (T == 32) {
    __syncwarp();
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"(sys_in_cta + 1), "r"(T) : "memory");
  }
}

// One elimination step K as a template so that every register index is a compile-time constant (a #pragma unroll of
// the k loop exceeds the front-end's full-unroll budget for NP > 32 and the row would fall into local memory).
template <int NP, int T, int K>
struct LuStep {
  using L = LuSmem<NP, T>;
  static __device__ __forceinline__ void run(double (&a)[NP], double& b, bool& done, double* cand,
                                             unsigned long long* ckey, double* Urows, double* ys, double* rinvs,
                                             const int w, const int lane, const int sys_in_cta) {
    constexpr int NW = L::NW;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int buf = K & 1;
    const double ak = a[K];
    // key: bit pattern of |a| (+1 so that a live row with a zero entry still beats a used row)
    const unsigned long long key = done ? 0ull : ((unsigned long long)__double_as_longlong(fabs(ak)) + 1ull);
    const double rinv = 1.0 / ak;  // speculative: only the winner's is used
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(FULL, hi);
    const unsigned lo2 = (hi == mhi) ? lo : 0u;
    const unsigned mlo = __reduce_max_sync(FULL, lo2);
    const unsigned ball = __ballot_sync(FULL, hi == mhi && lo == mlo);
    const int wl = __ffs(ball) - 1;  // lowest lane = lowest row index on ties
    double* crow = cand + (buf * NW + w) * L::CAND_PITCH;
    if (lane == wl) {
#pragma unroll
      for (int j = K + 1; j < NP; ++j) crow[j] = a[j];
      crow[NP] = b;
      crow[NP + 1] = rinv;
      ckey[buf * NW + w] = key;
    }
    lu_sys_sync<T>(sys_in_cta);
    int wbest = 0;
    if (NW > 1) {
      unsigned long long kb = ckey[buf * NW];
#pragma unroll
      for (int w2 = 1; w2 < NW; ++w2) {
        const unsigned long long k2 = ckey[buf * NW + w2];
        if (k2 > kb) {  // strict: lowest warp (= lowest row) wins ties
          kb = k2;
          wbest = w2;
        }
      }
    }
    const double* prow = cand + (buf * NW + wbest) * L::CAND_PITCH;
    const bool iwin = (!done) && (w == wbest) && (lane == wl);
    if (iwin) {
      done = true;
#pragma unroll
      for (int j = K; j < NP; ++j) Urows[K * L::U_PITCH + j] = a[j];
      ys[K] = b;
      rinvs[K] = rinv;
    } else if (!done) {
      const double l = ak * prow[NP + 1];
#pragma unroll
      for (int j = K + 1; j < NP; ++j) a[j] = fma(-l, prow[j], a[j]);
      b = fma(-l, prow[NP], b);
    }
    LuStep<NP, T, K + 1>::run(a, b, done, cand, ckey, Urows, ys, rinvs, w, lane, sys_in_cta);
  }
};
template <int NP, int T>
struct LuStep<NP, T, NP> {
  static __device__ __forceinline__ void run(double (&)[NP], double&, bool&, double*, unsigned long long*, double*,
                                             double*, double*, const int, const int, const int) {}
};

template <int NP, int T>
__global__ void __launch_bounds__(128, (NP > 48 ? 2 : (NP > 32 ? 3 : 4))) rb_lu_kernel(LuArgs args) {
  using L = LuSmem<NP, T>;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char lu_smem_raw[];

  const int sys_in_cta = threadIdx.x / T;
  const int i = threadIdx.x % T;
  const int w = i >> 5, lane = i & 31;
  const long long sys = (long long)blockIdx.x * L::SPC + sys_in_cta;
  if (sys >= args.count) return;  // whole system leaves together (barriers are per system)

  double* cand = reinterpret_cast<double*>(lu_smem_raw + (size_t)sys_in_cta * L::BYTES_PER_SYS);
  unsigned long long* ckey = reinterpret_cast<unsigned long long*>(cand + L::CAND_DOUBLES);
  double* Urows = cand + L::CAND_DOUBLES + L::KEY_DOUBLES;
  double* ys = Urows + NP * L::U_PITCH;
  double* rinvs = ys + NP;

  const int N = args.N;
  double a[NP];
  double b = 0.0;
  {
    const double* Ab = args.Ab + sys * (long long)args.EP;
    if (i < NP) {
#pragma unroll
      for (int j = 0; j < NP; ++j) a[j] = __ldcs(Ab + j * NP + i);
    } else {
#pragma unroll
      for (int j = 0; j < NP; ++j) a[j] = 0.0;
    }
    const double* th = args.theta + sys * (long long)args.QT + args.Ql;
    if (i < N) {
      if (args.rhs_in)
        b = args.rhs_in[sys * (long long)N + i];
      else
        for (int q = 0; q < args.Qr; ++q) b = fma(__ldg(th + q), __ldg(args.F + q * NP + i), b);
    }
    // lifting rows (DirichletBC.py:41-56) and identity padding rows N <= i < NP
    bool unit_row = (i >= N && i < NP);
    for (int t = 0; t < args.n_bc; ++t) {
      if (__ldg(args.bc_rows + t) == i) {
        unit_row = true;
        b = __ldg(args.thetaBC + sys * (long long)args.n_bc + t);
      }
    }
    if (unit_row) {
#pragma unroll
      for (int j = 0; j < NP; ++j) a[j] = (j == i) ? 1.0 : 0.0;
    }
  }

  bool done = (i >= NP);
  LuStep<NP, T, 0>::run(a, b, done, cand, ckey, Urows, ys, rinvs, w, lane, sys_in_cta);
  lu_sys_sync<T>(sys_in_cta);

  // back-substitution U x = y by warp 0 of the system; lane owns rows lane, lane+32, ...
  if (w == 0) {
    constexpr int R = (NP + 31) / 32;
    double acc[R], ri[R], xo[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = r * 32 + lane;
      acc[r] = row < NP ? ys[row] : 0.0;
      ri[r] = row < NP ? rinvs[row] : 0.0;
      xo[r] = 0.0;
    }
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
      const int kk_hi = (NP - r * 32 < 32) ? (NP - r * 32 - 1) : 31;
      for (int kk = kk_hi; kk >= 0; --kk) {
        const int k = r * 32 + kk;
        const double xk = __shfl_sync(FULL, acc[r] * ri[r], kk);
        if (lane == kk) xo[r] = xk;
#pragma unroll
        for (int r2 = 0; r2 <= r; ++r2) {
          const int row = r2 * 32 + lane;
          if (row < k) acc[r2] = fma(-Urows[row * L::U_PITCH + k], xk, acc[r2]);
        }
      }
    }
    double* u = args.U + sys * (long long)args.ldu;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = r * 32 + lane;
      if (row < N) u[row] = xo[r];
    }
  }
}

template <int NP>
int rb_lu_launch_np(const LuArgs& args, cudaStream_t stream) {
  constexpr int T = NP <= 32 ? 32 : (NP <= 64 ? 64 : 128);
  using L = LuSmem<NP, T>;
  auto kern = rb_lu_kernel<NP, T>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES);
  if (e != cudaSuccess) return (int)e;
  const long long nblocks = (args.count + L::SPC - 1) / L::SPC;
  kern<<<(unsigned)nblocks, 128, L::BYTES, stream>>>(args);
  return (int)cudaGetLastError();
}

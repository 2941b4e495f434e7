// rb_lu_stream.cu -- batched LU (partial pivoting) + solves for 32 < N <= 128 with MANY systems in flight per SM.
//
// Replaces numpy.linalg.solve (LAPACK dgesv; rbnics/backends/online/numpy/linear_solver.py:29-31) and the lifting rows
// (rbnics/backends/online/basic/wrapping/DirichletBC.py:37-56) for a chunk of parameters.
//
// The pivot chain of one system is sequential (N dependent steps), so throughput comes from concurrency, and concurrency
// is limited by where the 80 KB matrix lives.  Here it lives in GLOBAL memory (the assembly scratch, sized to stay
// L2-resident) and only a 16-column panel is on chip:
//   CTA = one system, 128 threads, thread i = row i (rows never move: implicit pivoting, idamax rule).
//   per panel (16 columns):
//     F  16 pivot steps on the register panel: REDUX.MAX arg-max per warp, 4 warp winners through shared memory, ONE
//        barrier per step; the rhs entry is carried along (forward substitution is part of the factorisation)
//     G  thread j gathers trailing column j of the 16 pivot rows from global memory, solves the 16x16 unit-lower
//        system in registers -> final U rows of the panel (to shared memory for the update, to global for the
//        back-substitution)
//     U  thread i (live rows) streams its row's trailing entries through registers: A[i][j] -= sum_t l[i][t] U[t][j]
//   back-substitution by panels in reverse: a 16x16 triangular solve by one warp, then every earlier pivot row
//   subtracts its 16 entries times the new unknowns.
// ~64 registers/thread and ~19 KB of shared memory per CTA -> 8 systems per SM (the register-resident DMMA kernel of
// rb_lu_dmma.cuh holds 2).
#include <stdlib.h>

#include "rb_lu_kernel.cuh"

constexpr int LUS_PW = 16;       // panel width
constexpr int LUS_T = 128;       // threads = max rows
constexpr int LUS_UP = 128;      // pitch of the U-panel buffer (>= 128 - 16 + LUS_CB, multiple of 2)
constexpr int LUS_CB = 8;        // trailing columns per trip of the update loop
constexpr int LUS_CP = LUS_PW + 2;

struct LusSmem {
  double cand[2][4][LUS_CP];            // candidate pivot rows of the 4 warps (double buffered): panel tail, rhs, 1/pivot
  unsigned long long ckey[2][4];
  double Lpp[LUS_PW][LUS_PW];           // multipliers of the panel's pivot rows w.r.t. earlier pivots of the panel
  double Ubar[LUS_PW][LUS_UP];          // final U rows of the panel, trailing columns
  double xs[LUS_T];                     // solution in pivot order == unknown order
  double yd[LUS_PW], rd[LUS_PW];        // rhs / reciprocal pivots of the panel being back-substituted
  double Ud[LUS_PW][LUS_PW + 1];        // diagonal block of U of that panel
  int pivrow[LUS_PW];
};

template <int K>
__device__ __forceinline__ void lus_step(double (&a)[LUS_PW], double& b, bool& done, int& mystep, double& myrinv,
                                         LusSmem& sm, const int pw, const int w, const int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int buf = K & 1;
  if (K < pw) {  // uniform
    const double ak = a[K];
    const unsigned long long key = done ? 0ull : ((unsigned long long)__double_as_longlong(fabs(ak)) + 1ull);
    const double rinv = 1.0 / ak;  // speculative: only the winner's is used
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(FULL, hi);
    const unsigned lo2 = (hi == mhi) ? lo : 0u;
    const unsigned mlo = __reduce_max_sync(FULL, lo2);
    const unsigned ball = __ballot_sync(FULL, hi == mhi && lo == mlo);
    const int wl = __ffs(ball) - 1;  // lowest lane = lowest row of the warp on ties
    if (lane == wl) {
      double* crow = sm.cand[buf][w];
#pragma unroll
      for (int j = K + 1; j < LUS_PW; ++j) crow[j] = a[j];
      crow[LUS_PW] = b;
      crow[LUS_PW + 1] = rinv;
      sm.ckey[buf][w] = key;
    }
    __syncthreads();
    int wbest = 0;
    unsigned long long kb = sm.ckey[buf][0];
#pragma unroll
    for (int w2 = 1; w2 < 4; ++w2) {
      const unsigned long long k2 = sm.ckey[buf][w2];
      if (k2 > kb) {  // strict: lowest warp (= lowest row) wins ties  -> idamax
        kb = k2;
        wbest = w2;
      }
    }
    const double* prow = sm.cand[buf][wbest];
    if (!done) {
      if (w == wbest && lane == wl) {
        done = true;
        mystep = K;
        myrinv = rinv;
      } else {
        const double l = ak * prow[LUS_PW + 1];
        a[K] = l;
#pragma unroll
        for (int j = K + 1; j < LUS_PW; ++j) a[j] = fma(-l, prow[j], a[j]);
        b = fma(-l, prow[LUS_PW], b);
      }
    }
  }
}

template <int MINB>
__global__ void __launch_bounds__(LUS_T, MINB) rb_lu_stream_kernel(LuArgs args, double* __restrict__ Awork) {
  __shared__ LusSmem sm;
  const int i = threadIdx.x, w = i >> 5, lane = i & 31;
  const long long sys = blockIdx.x;
  const int N = args.N;
  const int NP = (N + 3) / 4 * 4;  // leading dimension of the assembled matrices (column-major)
  double* A = Awork + sys * (long long)args.EP;
  const bool row = i < N;

  // ---- rhs and lifting rows
  double b = 0.0;
  bool bc_row = false;
  if (row) {
    const double* th = args.theta + sys * (long long)args.QT + args.Ql;
    for (int q = 0; q < args.Qr; ++q) b = fma(__ldg(th + q), __ldg(args.F + q * NP + i), b);
    for (int t = 0; t < args.n_bc; ++t)
      if (__ldg(args.bc_rows + t) == i) {
        bc_row = true;
        b = __ldg(args.thetaBC + sys * (long long)args.n_bc + t);
      }
    if (bc_row)
      for (int j = 0; j < N; ++j) A[(long long)j * NP + i] = (j == i) ? 1.0 : 0.0;
  }

  bool done = !row;   // rows >= N never compete
  int mystep = -1;    // step within the panel at which this row became a pivot
  int myk = -1;       // global pivot index of this row
  double myrinv = 0.0;
  const int npan = (N + LUS_PW - 1) / LUS_PW;

  for (int p = 0; p < npan; ++p) {
    const int c0 = p * LUS_PW;
    const int pw = min(LUS_PW, N - c0);
    const bool live = !done;
    // ---- load the panel (own row)
    double a[LUS_PW];
#pragma unroll
    for (int s = 0; s < LUS_PW; ++s) a[s] = (live && s < pw) ? __ldcg(A + (long long)(c0 + s) * NP + i) : 0.0;
    // ---- F: pivot steps
    mystep = -1;
    lus_step<0>(a, b, done, mystep, myrinv, sm, pw, w, lane);
    lus_step<1>(a, b, done, mystep, myrinv, sm, pw, w, lane);
    lus_step<2>(a, b, done, mystep, myrinv, sm, pw, w, lane);
    lus_step<3>(a, b, done, mystep, myrinv, sm, pw, w, lane);
    lus_step<4>(a, b, done, mystep, myrinv, sm, pw, w, lane);
    lus_step<5>(a, b, done, mystep, myrinv, sm, pw, w, lane);
    lus_step<6>(a, b, done, mystep, myrinv, sm, pw, w, lane);
    lus_step<7>(a, b, done, mystep, myrinv, sm, pw, w, lane);
    lus_step<8>(a, b, done, mystep, myrinv, sm, pw, w, lane);
    lus_step<9>(a, b, done, mystep, myrinv, sm, pw, w, lane);
    lus_step<10>(a, b, done, mystep, myrinv, sm, pw, w, lane);
    lus_step<11>(a, b, done, mystep, myrinv, sm, pw, w, lane);
    lus_step<12>(a, b, done, mystep, myrinv, sm, pw, w, lane);
    lus_step<13>(a, b, done, mystep, myrinv, sm, pw, w, lane);
    lus_step<14>(a, b, done, mystep, myrinv, sm, pw, w, lane);
    lus_step<15>(a, b, done, mystep, myrinv, sm, pw, w, lane);
    const int ntrail = N - (c0 + pw);
    if (live && mystep >= 0) {  // this row became pivot c0 + mystep: publish its multipliers and keep its U entries
      myk = c0 + mystep;
      sm.pivrow[mystep] = i;
#pragma unroll
      for (int s = 0; s < LUS_PW; ++s) {
        if (s < mystep) sm.Lpp[mystep][s] = a[s];
        if (s > mystep && s < pw) __stcg(A + (long long)(c0 + s) * NP + i, a[s]);  // U entries inside the panel (back-subst.)
      }
    }
    if (ntrail > 0) {
      __syncthreads();
      // ---- G: final U rows of the panel for the trailing columns; thread j <-> column c0 + pw + j
      if (i < ntrail) {
        const long long col = (long long)(c0 + pw + i) * NP;
        double v[LUS_PW];
#pragma unroll
        for (int t = 0; t < LUS_PW; ++t) v[t] = (t < pw) ? __ldcg(A + col + sm.pivrow[t]) : 0.0;
#pragma unroll
        for (int t = 1; t < LUS_PW; ++t) {
          if (t < pw) {
            double x = v[t];
#pragma unroll
            for (int s = 0; s < t; ++s) x = fma(-sm.Lpp[t][s], v[s], x);
            v[t] = x;
            __stcg(A + col + sm.pivrow[t], x);
          }
        }
#pragma unroll
        for (int t = 0; t < LUS_PW; ++t) sm.Ubar[t][i] = v[t];
      }
      __syncthreads();
      // ---- U: trailing update of the live rows (not pivoted so far), LUS_CB columns per trip: all loads of a trip
      // are issued before the first FMA (the loop-carried store -> load order would otherwise serialise L2 latencies)
      if (!done) {
        double* prow_i = A + (long long)(c0 + pw) * NP + i;
        for (int j = 0; j < ntrail; j += LUS_CB) {
          double x[LUS_CB];
#pragma unroll
          for (int c = 0; c < LUS_CB; ++c) x[c] = (j + c < ntrail) ? __ldcg(prow_i + (long long)(j + c) * NP) : 0.0;
#pragma unroll
          for (int t = 0; t < LUS_PW; ++t) {
#pragma unroll
            for (int c = 0; c < LUS_CB; c += 2) {
              const double2 u = *reinterpret_cast<const double2*>(&sm.Ubar[t][j + c]);
              x[c] = fma(-a[t], u.x, x[c]);
              x[c + 1] = fma(-a[t], u.y, x[c + 1]);
            }
          }
#pragma unroll
          for (int c = 0; c < LUS_CB; ++c)
            if (j + c < ntrail) __stcg(prow_i + (long long)(j + c) * NP, x[c]);
        }
      }
    }
    __syncthreads();  // global writes of this panel are visible to the CTA; pivrow / Lpp / Ubar may be reused
  }

  // ---- back-substitution U x = y by panels in reverse.  Row i is pivot myk; b is y[myk]; U[myk][j] = A[i][j], j > myk.
  for (int p = npan - 1; p >= 0; --p) {
    const int c0 = p * LUS_PW;
    const int pw = min(LUS_PW, N - c0);
    if (row && myk >= c0) {  // (myk < c0 + pw always holds here: later panels are finished)
      const int t = myk - c0;
      if (t >= 0 && t < pw) {
        sm.yd[t] = b;
        sm.rd[t] = myrinv;
#pragma unroll
        for (int s = 1; s < LUS_PW; ++s)
          if (s > t && s < pw) sm.Ud[t][s] = __ldcg(A + (long long)(c0 + s) * NP + i);
      }
    }
    __syncthreads();
    if (w == 0) {
      double y = lane < pw ? sm.yd[lane] : 0.0;
      const double ri = lane < pw ? sm.rd[lane] : 0.0;
      double urow[LUS_PW];
#pragma unroll
      for (int s = 0; s < LUS_PW; ++s) urow[s] = (lane < pw && s > lane && s < pw) ? sm.Ud[lane][s] : 0.0;
      double xo = 0.0;
#pragma unroll
      for (int t = LUS_PW - 1; t >= 0; --t) {
        if (t < pw) {
          const double xt = __shfl_sync(0xffffffffu, y * ri, t);
          if (lane == t) xo = xt;
          y = fma(-urow[t], xt, y);
        }
      }
      if (lane < pw) sm.xs[c0 + lane] = xo;
    }
    __syncthreads();
    if (row && myk >= 0 && myk < c0) {
      double uv[LUS_PW];
#pragma unroll
      for (int s = 0; s < LUS_PW; ++s) uv[s] = (s < pw) ? __ldcg(A + (long long)(c0 + s) * NP + i) : 0.0;
      double acc = b;
#pragma unroll
      for (int s = 0; s < LUS_PW; ++s)
        if (s < pw) acc = fma(-uv[s], sm.xs[c0 + s], acc);
      b = acc;
    }
  }
  if (row) args.U[sys * (long long)args.ldu + i] = sm.xs[i];
}

// systems resident per SM (register budget 65536 / (128 * MINB)): 8 -> 64 registers, 6 -> 80, 5 -> 96, 4 -> 128
int rb_lu_stream_occupancy() {
  static const int occ = [] {
    const char* e = getenv("RB_LU_OCC");
    const int v = e ? atoi(e) : 6;
    return (v == 8 || v == 6 || v == 5 || v == 4) ? v : 6;
  }();
  return occ;
}

int rb_lu_stream_launch(const LuArgs& args, double* Awork, cudaStream_t stream) {
  const unsigned grid = (unsigned)args.count;
  switch (rb_lu_stream_occupancy()) {
    case 8: rb_lu_stream_kernel<8><<<grid, LUS_T, 0, stream>>>(args, Awork); break;
    case 5: rb_lu_stream_kernel<5><<<grid, LUS_T, 0, stream>>>(args, Awork); break;
    case 4: rb_lu_stream_kernel<4><<<grid, LUS_T, 0, stream>>>(args, Awork); break;
    default: rb_lu_stream_kernel<6><<<grid, LUS_T, 0, stream>>>(args, Awork); break;
  }
  return (int)cudaGetLastError();
}

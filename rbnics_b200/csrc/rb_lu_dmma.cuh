// rb_lu_dmma.cuh -- blocked batched LU (partial pivoting) on the FP64 tensor pipe for 32 < N <= 104.
//
// Replaces numpy.linalg.solve (LAPACK dgesv; rbnics/backends/online/numpy/linear_solver.py:29-31) and the lifting rows
// (rbnics/backends/online/basic/wrapping/DirichletBC.py:37-56) for a chunk of parameters.
//
// One system = W warps.  The matrix (with the rhs as column N) is kept in REGISTERS in DMMA accumulator-fragment layout:
// warp w owns the 8-row tile rows 2w and 2w+1, all NC 8-column tiles of each (2 doubles per tile per lane).
// Pivoting is implicit (rows never move): per 8-column panel
//   A  every warp drops its panel tile(s) into shared memory            (fully parallel, STS.128)
//   B  warp 0 factors the N x 8 panel, lane = rows lane+32j: pivot = largest |a| among live rows, lowest row on ties
//      (idamax), REDUX.MAX/MIN on the IEEE bit pattern; produces the (negated) multipliers of every row and the pivot rows
//   C  the 8 pivot rows' trailing columns are gathered from registers into the U store (pivot order)
//   C2 8x8 unit-lower triangular solve on those rows, one thread per column  -> final U rows
//   D  rank-8 update of every register tile right of the panel: C -= L U as two DMMA.8x8x4 per tile; rows already
//      used as pivots have zero multipliers, rows pivoted in this panel carry only their within-panel multipliers
// then one warp back-substitutes with shuffles.  Only phase B is sequential (N pivot steps); two systems per SM overlap it.
#pragma once
#include "rb_lu_kernel.cuh"

#ifdef LUD_TIMING
__device__ unsigned long long lud_timing[32];
#define LUD_TDECL long long lud_t_last = clock64();
#define LUD_T(i)                                                             \
  if (lane == 0) {                                                           \
    const long long now__ = clock64();                                       \
    atomicAdd(&lud_timing[i], (unsigned long long)(now__ - lud_t_last));     \
    lud_t_last = now__;                                                      \
  }
#else
#define LUD_TDECL
#define LUD_T(i)
#endif

constexpr int LUD_PP = 12;  // pitch of the panel / multiplier buffers (doubles): conflict-free A fragments, 16 B rows

template <int NT, int NC>
struct LudCfg {
  static constexpr int W = (NT + 1) / 2;   // tile warps (two tile rows each); one more warp factors the panels
  static constexpr int WS = W + 1;          // warps per system
  static constexpr int NP8 = 8 * NT;
  static constexpr int UP = 8 * NC + 4;  // == 4 or 12 (mod 16): conflict-free B fragments
  static constexpr int SPC = WS >= 4 ? 1 : 4 / WS;
  static constexpr int THREADS = WS * 32 * SPC;
  static constexpr int P_OFF = 0;                                // [NP8][PP]
  static constexpr int L_OFF = P_OFF;                            // [NP8][PP]  negated multipliers; ALIASES the panel buffer:
                                                                 // B reads the whole panel into registers before it writes nL, and the tile warps
                                                                 // only store the next panel after their last read of nL (barrier LUD_BAR_C2)
  static constexpr int U_OFF = P_OFF + NP8 * LUD_PP;             // [NP8][UP]  U rows in pivot order (+ rhs column N)
  static constexpr int R_OFF = U_OFF + NP8 * UP;                 // [NP8] reciprocal pivots
  static constexpr int UB_OFF = R_OFF + NP8;                     // [2][32][10] candidate rows of the panel warp's lanes
  static constexpr int PIV_OFF = UB_OFF + 2 * 32 * 10;                     // [NP8] ints (NP8/2 doubles)
  static constexpr int DOUBLES = PIV_OFF + NP8 / 2 + 2;
  static constexpr int BYTES_PER_SYS = ((DOUBLES * 8 + 15) / 16) * 16;
  static constexpr int BYTES = BYTES_PER_SYS * SPC;
};

// branch-free select (the compiler turns "cond ? a : b" on a lane-varying condition into divergent branches here)
__device__ __forceinline__ double lud_selp(double a, double b, int x, int y) {  // x == y ? a : b
  double r;
  asm("{\n .reg .pred p;\n setp.eq.s32 p, %3, %4;\n selp.f64 %0, %1, %2, p;\n}" : "=d"(r) : "d"(a), "d"(b), "r"(x), "r"(y));
  return r;
}

// One pivot step T of phase B as a template: a "#pragma unroll" of the t loop is not honoured for a body of this size
// (the front end then indexes pv[j][t] dynamically through select ladders, ~10x more instructions per step).
template <int T>
__device__ __forceinline__ void lud_pivot_step(double (&pv)[4][8], int (&tst)[4], unsigned& dmask, int* pivrow,
                                               double* rinvs, double* ubuf, int p, int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int t = T;
  // local best among this lane's live rows (strict >: the lower row of the lane wins ties)
  unsigned long long key = 0ull;
  int jb = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const unsigned long long kj =
        ((dmask >> j) & 1u) ? 0ull : ((unsigned long long)__double_as_longlong(fabs(pv[j][t])) + 1ull);
    if (kj > key) {
      key = kj;
      jb = j;
    }
  }
  // arg-max of the 64-bit key over the warp, lowest row index on ties (idamax): REDUX.MAX on the high word, then
  // REDUX.MAX on {low word without its 7 lowest bits, 127 - row}.  Candidates that differ only below 2^-45 relative
  // are treated as equal (any of them is an equally good pivot).
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mhi = __reduce_max_sync(FULL, hi);
  // the lane's candidate row (columns >= t), selected branch-free; every lane stages it (and a speculative
  // reciprocal) in its shared-memory slot while the reductions are in flight; the winner's slot is then read by all
  // lanes (one __syncwarp per step)
  double sel[8];
#pragma unroll
  for (int c = t; c < 8; ++c) {
    double v = pv[0][c];
    v = lud_selp(pv[1][c], v, jb, 1);
    v = lud_selp(pv[2][c], v, jb, 2);
    v = lud_selp(pv[3][c], v, jb, 3);
    sel[c] = v;
  }
  const double myrinv = 1.0 / sel[t];
  double* slot = ubuf + ((t & 1) * 32 + lane) * 10;
#pragma unroll
  for (int c = t + 1; c < 8; ++c) slot[c] = sel[c];
  slot[8] = myrinv;
  const int myrow = jb * 32 + lane;
  const unsigned k2 = (hi == mhi && key != 0ull) ? ((lo & ~127u) | (unsigned)(127 - myrow)) : 0u;
  const unsigned m2 = __reduce_max_sync(FULL, k2);
  const int prow = 127 - (int)(m2 & 127u);
  const int ol = prow & 31;
  if ((lane == ol) && (k2 == m2)) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j == jb) tst[j] = t;
    dmask |= 1u << jb;
    pivrow[8 * p + t] = prow;
    rinvs[8 * p + t] = myrinv;
  }
  __syncwarp();
  const double* wslot = ubuf + ((t & 1) * 32 + ol) * 10;
  const double rinv = wslot[8];
  double u[8];
#pragma unroll
  for (int c = t + 1; c < 8; ++c) u[c] = wslot[c];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (!((dmask >> j) & 1u)) {
      const double l = pv[j][t] * rinv;
      pv[j][t] = l;
#pragma unroll
      for (int c = t + 1; c < 8; ++c) pv[j][c] = fma(-l, u[c], pv[j][c]);
    }
  }
}

// Phase B.  Returns the updated per-lane mask of used rows (bit j = row lane+32j).
__device__ __forceinline__ unsigned lud_panel_factor(double* P, double* nL,
                                                  double* __restrict__ Uf, int* __restrict__ pivrow,
                                                  double* __restrict__ rinvs, double* __restrict__ ubuf, int p, int N,
                                                  int NP8, int UP, unsigned dmask, int lane) {
  const unsigned dmask_in = dmask;
  double pv[4][8];
  int tst[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    tst[j] = -1;
    const int row = lane + 32 * j;
    const bool live = (row < NP8) && !((dmask >> j) & 1u);
    if (live) {
      const double2* src = reinterpret_cast<const double2*>(P + row * LUD_PP);
#pragma unroll
      for (int c2 = 0; c2 < 4; ++c2) {
        const double2 v = src[c2];
        pv[j][2 * c2] = v.x;
        pv[j][2 * c2 + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) pv[j][c] = 0.0;
    }
  }
  const int nst = min(8, N - 8 * p);
  lud_pivot_step<0>(pv, tst, dmask, pivrow, rinvs, ubuf, p, lane);
  if (nst > 1) lud_pivot_step<1>(pv, tst, dmask, pivrow, rinvs, ubuf, p, lane);
  if (nst > 2) lud_pivot_step<2>(pv, tst, dmask, pivrow, rinvs, ubuf, p, lane);
  if (nst > 3) lud_pivot_step<3>(pv, tst, dmask, pivrow, rinvs, ubuf, p, lane);
  if (nst > 4) lud_pivot_step<4>(pv, tst, dmask, pivrow, rinvs, ubuf, p, lane);
  if (nst > 5) lud_pivot_step<5>(pv, tst, dmask, pivrow, rinvs, ubuf, p, lane);
  if (nst > 6) lud_pivot_step<6>(pv, tst, dmask, pivrow, rinvs, ubuf, p, lane);
  if (nst > 7) lud_pivot_step<7>(pv, tst, dmask, pivrow, rinvs, ubuf, p, lane);
  // negated multipliers of every row for phases C2 / D, and the diagonal block of U
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = lane + 32 * j;
    if (row < NP8) {
      const bool dead = (dmask_in >> j) & 1u;
      double nl[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const bool is_mult = !dead && (tst[j] < 0 || s < tst[j]);
        nl[s] = is_mult ? -pv[j][s] : 0.0;
      }
      double2* dst = reinterpret_cast<double2*>(nL + row * LUD_PP);
#pragma unroll
      for (int c2 = 0; c2 < 4; ++c2) dst[c2] = make_double2(nl[2 * c2], nl[2 * c2 + 1]);
      if (tst[j] >= 0) {
        double* urow = Uf + (8 * p + tst[j]) * UP + 8 * p;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c >= tst[j]) urow[c] = pv[j][c];
      }
    }
  }
  return dmask;
}

// Named-barrier hand-offs between the tile warps and the panel warp (one system per CTA for NT >= 5).
__device__ __forceinline__ void lud_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void lud_bar_arrive(int id, int n) {
  __threadfence_block();  // publish this thread's shared-memory writes before signalling
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}
constexpr int LUD_BAR_P = 1;   // panel tiles are in P            (tile warps arrive, panel warp waits)
constexpr int LUD_BAR_B = 2;   // panel factored                  (panel warp arrives, tile warps wait)
constexpr int LUD_BAR_C = 3;   // pivot rows gathered             (tile warps arrive, panel warp waits)
constexpr int LUD_BAR_C2 = 4;  // final U rows of the panel ready (panel warp arrives, tile warps wait)
constexpr int LUD_BAR_END = 5;

template <int NT, int NC>
__global__ void __launch_bounds__(LudCfg<NT, NC>::THREADS, 2) rb_lu_dmma_kernel(LuArgs args) {
  using Cfg = LudCfg<NT, NC>;
  static_assert(Cfg::SPC == 1, "the DMMA LU runs one system per CTA");
  constexpr int W = Cfg::W, WS = Cfg::WS, NP8 = Cfg::NP8, UP = Cfg::UP, NTH = Cfg::WS * 32;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char lud_smem_raw[];

  const int tid = threadIdx.x;
  const int w = tid >> 5, lane = tid & 31;
  const long long sys = blockIdx.x;

  double* sm = reinterpret_cast<double*>(lud_smem_raw);
  double* P = sm + Cfg::P_OFF;
  double* nL = sm + Cfg::L_OFF;
  double* Uf = sm + Cfg::U_OFF;
  double* rinvs = sm + Cfg::R_OFF;
  double* ubuf = sm + Cfg::UB_OFF;
  int* pivrow = reinterpret_cast<int*>(sm + Cfg::PIV_OFF);

  const int N = args.N;
  const int npan = (N + 7) / 8;

  if (w == W) {
    // =================== panel warp: phases B and C2, then the back-substitution ===================
    unsigned dmask = 0u;  // rows >= N are dead from the start
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (lane + 32 * j >= N) dmask |= 1u << j;
    LUD_TDECL
    for (int p = 0; p < npan; ++p) {
      lud_bar_sync(LUD_BAR_P, NTH);
      LUD_T(0)
      dmask = lud_panel_factor(P, nL, Uf, pivrow, rinvs, ubuf, p, N, NP8, UP, dmask, lane);
      LUD_T(1)
      lud_bar_arrive(LUD_BAR_B, NTH);
    }
    lud_bar_sync(LUD_BAR_END, NTH);
    LUD_T(2)
    // back-substitution U x = y; lane owns rows lane, lane+32, ...; y is column N of the U store
    constexpr int R = (NP8 + 31) / 32;
    double acc[R], ri[R], xo[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = r * 32 + lane;
      acc[r] = row < N ? Uf[row * UP + N] : 0.0;
      ri[r] = row < N ? rinvs[row] : 0.0;
      xo[r] = 0.0;
    }
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
      const int kk_hi = min(31, N - 1 - r * 32);
      for (int kk = kk_hi; kk >= 0; --kk) {
        const int k = r * 32 + kk;
        const double xk = __shfl_sync(FULL, acc[r] * ri[r], kk);
        if (lane == kk) xo[r] = xk;
#pragma unroll
        for (int r2 = 0; r2 <= r; ++r2) {
          const int row = r2 * 32 + lane;
          if (row < k) acc[r2] = fma(-Uf[row * UP + k], xk, acc[r2]);
        }
      }
    }
    double* u = args.U + sys * (long long)args.ldu;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = r * 32 + lane;
      if (row < N) u[row] = xo[r];
    }
    LUD_T(3)
    return;
  }

  // =================== tile warps: phases A, C and D ===================
  LUD_TDECL
  const int g = lane >> 2, q4 = lane & 3;
  const int NPc = (N + 3) / 4 * 4;  // leading dimension of the assembled matrices (column-major)
  const int tr0 = 2 * w, tr1 = 2 * w + 1;
  const bool has1 = tr1 < NT;

  // load: c[h][tc][e] = A[row = 8*tr_h + g][col = 8*tc + 2*q4 + e]; column N is the rhs.  All loads are issued
  // unconditionally from clamped addresses first (independent, all in flight), the fix-ups come afterwards.
  double c[2][NC][2];
  {
    const double* Ab = args.Ab + sys * (long long)args.EP;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int rowc = min(8 * (2 * w + h) + g, N - 1);
#pragma unroll
      for (int tc = 0; tc < NC; ++tc) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int colc = min(8 * tc + 2 * q4 + e, N - 1);
          c[h][tc][e] = __ldcs(Ab + (long long)colc * NPc + rowc);
        }
      }
    }
    const double* th = args.theta + sys * (long long)args.QT + args.Ql;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = 8 * (2 * w + h) + g;
      const bool rv = (row < N) && (h == 0 || has1);
      int bc_t = -1;
      if (rv)
        for (int t = 0; t < args.n_bc; ++t)
          if (__ldg(args.bc_rows + t) == row) bc_t = t;
      double rhs = 0.0;
      if (rv) {
        if (bc_t >= 0)
          rhs = __ldg(args.thetaBC + sys * (long long)args.n_bc + bc_t);
        else
        {
          if (args.rhs_in)
            rhs = args.rhs_in[sys * (long long)N + row];
          else
            for (int q = 0; q < args.Qr; ++q) rhs = fma(__ldg(th + q), __ldg(args.F + q * NPc + row), rhs);
        }
      }
#pragma unroll
      for (int tc = 0; tc < NC; ++tc) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = 8 * tc + 2 * q4 + e;
          double v = c[h][tc][e];
          if (bc_t >= 0) v = (col == row) ? 1.0 : 0.0;
          if (col == N) v = rhs;
          if (!rv || col > N) v = 0.0;
          c[h][tc][e] = v;
        }
      }
    }
  }

  auto store_panel = [&](int p) {  // phase A
#pragma unroll
    for (int tc = 0; tc < NC; ++tc) {
      if (tc == p) {
        *reinterpret_cast<double2*>(P + (8 * tr0 + g) * LUD_PP + 2 * q4) = make_double2(c[0][tc][0], c[0][tc][1]);
        if (has1)
          *reinterpret_cast<double2*>(P + (8 * tr1 + g) * LUD_PP + 2 * q4) = make_double2(c[1][tc][0], c[1][tc][1]);
      }
    }
  };

  store_panel(0);
  if (w == 0) { LUD_T(10) }
  lud_bar_arrive(LUD_BAR_P, NTH);
  for (int p = 0; p < npan; ++p) {
    lud_bar_sync(LUD_BAR_B, NTH);  // panel p factored: pivot rows, multipliers, diagonal block of U
    if (w == 0) { LUD_T(11) }
    if (p + 1 < NC) {
      // ---- C: gather the trailing part of the 8 pivot rows (pivot order) from registers
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int prow = pivrow[8 * p + t];
        const int ptr_ = prow >> 3;
        if ((prow & 7) == g && (ptr_ == tr0 || ptr_ == tr1)) {
          double* urow = Uf + (8 * p + t) * UP + 2 * q4;
          if (ptr_ == tr0) {
#pragma unroll
            for (int tc = 0; tc < NC; ++tc)
              if (tc > p) *reinterpret_cast<double2*>(urow + 8 * tc) = make_double2(c[0][tc][0], c[0][tc][1]);
          } else {
#pragma unroll
            for (int tc = 0; tc < NC; ++tc)
              if (tc > p) *reinterpret_cast<double2*>(urow + 8 * tc) = make_double2(c[1][tc][0], c[1][tc][1]);
          }
        }
      }
      // multipliers of this warp's rows (A fragments of phase D): read now, nL is rewritten by the next panel
      const double a00 = nL[(8 * tr0 + g) * LUD_PP + q4], a01 = nL[(8 * tr0 + g) * LUD_PP + 4 + q4];
      double a10 = 0.0, a11 = 0.0;
      if (has1) {
        a10 = nL[(8 * tr1 + g) * LUD_PP + q4];
        a11 = nL[(8 * tr1 + g) * LUD_PP + 4 + q4];
      }
      lud_bar_sync(LUD_BAR_C, W * 32);  // tile warps only: the gathered rows are complete
      if (w == 0) { LUD_T(12) }
      // ---- C2: unit-lower 8x8 solve on the gathered rows, one thread per trailing column
      for (int col = 8 * (p + 1) + tid; col < 8 * NC; col += W * 32) {
        double x[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) x[t] = Uf[(8 * p + t) * UP + col];
#pragma unroll
        for (int t = 1; t < 8; ++t) {
          const double* nlrow = nL + pivrow[8 * p + t] * LUD_PP;
#pragma unroll
          for (int s = 0; s < t; ++s) x[t] = fma(nlrow[s], x[s], x[t]);
        }
#pragma unroll
        for (int t = 1; t < 8; ++t) Uf[(8 * p + t) * UP + col] = x[t];
      }
      if (w == 0) { LUD_T(13) }
      lud_bar_sync(LUD_BAR_C2, W * 32);  // tile warps only: final U rows of panel p
      if (w == 0) { LUD_T(14) }
      // ---- D: rank-8 update of the register tiles right of the panel (2 DMMA per tile); the next panel's tile
      // column goes first and is handed to the panel warp at once (look-ahead: phase B(p+1) overlaps the rest of D(p))
      const double* ub0 = Uf + (8 * p + q4) * UP + g;
      const double* ub1 = Uf + (8 * p + 4 + q4) * UP + g;
#pragma unroll
      for (int tc = 0; tc < NC; ++tc) {
        if (tc == p + 1) {
          const double b0 = ub0[8 * tc], b1 = ub1[8 * tc];
          dmma884(c[0][tc][0], c[0][tc][1], a00, b0);
          dmma884(c[0][tc][0], c[0][tc][1], a01, b1);
          if (has1) {
            dmma884(c[1][tc][0], c[1][tc][1], a10, b0);
            dmma884(c[1][tc][0], c[1][tc][1], a11, b1);
          }
        }
      }
      if (p + 1 < npan) {
        store_panel(p + 1);
        if (w == 0) { LUD_T(15) }
        lud_bar_arrive(LUD_BAR_P, NTH);
      }
#pragma unroll
      for (int tc = 0; tc < NC; ++tc) {
        if (tc > p + 1) {
          const double b0 = ub0[8 * tc], b1 = ub1[8 * tc];
          dmma884(c[0][tc][0], c[0][tc][1], a00, b0);
          dmma884(c[0][tc][0], c[0][tc][1], a01, b1);
          if (has1) {
            dmma884(c[1][tc][0], c[1][tc][1], a10, b0);
            dmma884(c[1][tc][0], c[1][tc][1], a11, b1);
          }
        }
      }
    }
  }
  lud_bar_arrive(LUD_BAR_END, NTH);
  if (w == 0) { LUD_T(16) }
}

template <int NT, int NC>
int rb_lu_dmma_launch(const LuArgs& args, cudaStream_t stream) {
  using Cfg = LudCfg<NT, NC>;
  auto kern = rb_lu_dmma_kernel<NT, NC>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::BYTES);
  if (e != cudaSuccess) return (int)e;
  kern<<<(unsigned)args.count, Cfg::THREADS, Cfg::BYTES, stream>>>(args);
  return (int)cudaGetLastError();
}

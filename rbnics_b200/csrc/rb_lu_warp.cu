// rb_lu_warp.cu -- batched LU (partial pivoting) + solves for 32 < N <= 128: ONE WARP PER SYSTEM, matrix streamed
// through L2, ~15 systems in flight per SM.
//
// Replaces numpy.linalg.solve (LAPACK dgesv; rbnics/backends/online/numpy/linear_solver.py:29-31) and the lifting rows
// (rbnics/backends/online/basic/wrapping/DirichletBC.py:37-56) for a chunk of parameters.
//
// The pivot chain of one system is sequential (N dependent steps of ~150 instructions), so throughput comes from the
// number of independent chains an SM can interleave.  The register-resident kernel of rb_lu_dmma.cuh holds 2 systems
// per SM (registers AND shared memory full); here the matrix stays in the assembly scratch in global memory (only the
// systems in flight need to be L2-resident) and a warp keeps just an 8-column panel in registers:
//   F  lane owns rows lane+32j (j < 4) of the N x 8 panel (+ the rhs entry of each row, carried through the
//      elimination); 8 pivot steps: arg-max of |a| with lowest row on ties (idamax) by two REDUX.MAX on the IEEE bit
//      pattern, pivot row broadcast through a shared-memory slot, rank-1 update of the panel.  Rows never move.
//   G  lane = trailing column: gathers the 8 pivot rows' entries from global memory, solves the 8x8 unit-lower
//      system -> final U rows (shared memory for the update, global memory for the back-substitution)
//   U  rank-8 update of the trailing matrix on the FP64 tensor pipe: per 8x8 tile 2 LDG + 2 DMMA.8x8x4 + 2 STG, the
//      multipliers staged in shared memory in A-fragment order; tile rows whose 8 rows are all used are skipped
//   back-substitution by panels in reverse (8x8 triangular solve with shuffles, then one gather-update per panel).
// No block-level barrier anywhere: CTA = one warp.
#include <stdlib.h>

#include "rb_lu_kernel.cuh"


template <int NT>
struct LwCfg {
  static constexpr int NP8 = 8 * NT;
  static constexpr int UP = ((NP8 - 8 + 15) / 16) * 16 + 4;  // == 4 (mod 16): conflict-free B fragments
  static constexpr int R = (NP8 + 31) / 32;                   // rows per lane
};

template <int NT>
struct LwSmem {
  using C = LwCfg<NT>;
  double nLf[NT][2][32];      // negated multipliers of the panel in DMMA A-fragment order [tile row][k-step][lane]
  union {
    double Ubar[8][C::UP];    // final U rows of the panel, trailing columns (B fragments)
    double cand[2][32][10];   // pivot candidate slots of phase F: panel tail [t+1..7], 1/pivot [8], rhs [9]
  };
  double Lpp[8][8];           // negated multipliers of the panel's pivot rows w.r.t. earlier pivots of the panel
  double ys[C::NP8];          // rhs in pivot order -> solution
  double rinvs[C::NP8];       // reciprocal pivots
  int pivrow[C::NP8];         // pivot order -> row
};

__device__ __forceinline__ double lw_selp(double a, double b, int x, int y) {  // x == y ? a : b, branch-free
  double r;
  asm("{\n .reg .pred p;\n setp.eq.s32 p, %3, %4;\n selp.f64 %0, %1, %2, p;\n}" : "=d"(r) : "d"(a), "d"(b), "r"(x), "r"(y));
  return r;
}

// One pivot step T of phase F (a template: the register indices must be static).
template <int T>
__device__ __forceinline__ void lw_pivot_step(double (&pv)[4][8], double (&b)[4], int (&tst)[4], unsigned& dmask,
                                              int* pivrow, double* rinvs, double (*cand)[32][10], int c0, int lane) {
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
  // arg-max of the 64-bit key over the warp, lowest row on ties: REDUX.MAX on the high word, then on
  // {low word without its 7 lowest bits, 127 - row}; candidates that differ only below 2^-45 relative count as equal.
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mhi = __reduce_max_sync(FULL, hi);
  double sel[8];
#pragma unroll
  for (int c = t; c < 8; ++c) {
    double v = pv[0][c];
    v = lw_selp(pv[1][c], v, jb, 1);
    v = lw_selp(pv[2][c], v, jb, 2);
    v = lw_selp(pv[3][c], v, jb, 3);
    sel[c] = v;
  }
  double selb = b[0];
  selb = lw_selp(b[1], selb, jb, 1);
  selb = lw_selp(b[2], selb, jb, 2);
  selb = lw_selp(b[3], selb, jb, 3);
  const double myrinv = 1.0 / sel[t];  // speculative: only the winner's is used
  double* slot = cand[t & 1][lane];
#pragma unroll
  for (int c = t + 1; c < 8; ++c) slot[c] = sel[c];
  slot[8] = myrinv;
  slot[9] = selb;
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
    pivrow[c0 + t] = prow;
    rinvs[c0 + t] = myrinv;
  }
  __syncwarp();
  const double* wslot = cand[t & 1][ol];
  const double rinv = wslot[8];
  const double ub = wslot[9];
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
      b[j] = fma(-l, ub, b[j]);
    }
  }
}

template <int NT>
__global__ void __launch_bounds__(32, 15) rb_lu_warp_kernel(LuArgs args, double* __restrict__ Awork) {
  using C = LwCfg<NT>;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ LwSmem<NT> sm;
  const int lane = threadIdx.x;
  const long long sys = blockIdx.x;
  const int N = args.N;
  const int NP = (N + 3) / 4 * 4;  // leading dimension of the assembled matrices (column-major)
  double* A = Awork + sys * (long long)args.EP;
  const int g = lane >> 2, q4 = lane & 3;

  // ---- rhs of this lane's rows; lifting rows become identity rows in the scratch
  double b[4];
  unsigned dmask = 0u;  // bit j: row lane+32j is used up (or does not exist)
  {
    const double* th = args.theta + sys * (long long)args.QT + args.Ql;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = lane + 32 * j;
      double v = 0.0;
      if (row < N) {
        for (int q = 0; q < args.Qr; ++q) v = fma(__ldg(th + q), __ldg(args.F + q * NP + row), v);
        for (int t = 0; t < args.n_bc; ++t)
          if (__ldg(args.bc_rows + t) == row) {
            v = __ldg(args.thetaBC + sys * (long long)args.n_bc + t);
            for (int c = 0; c < N; ++c) __stcg(A + (long long)c * NP + row, (c == row) ? 1.0 : 0.0);
          }
      } else {
        dmask |= 1u << j;
      }
      b[j] = v;
    }
  }
  __syncwarp();

  const int npan = (N + 7) / 8;
  for (int p = 0; p < npan; ++p) {
    const int c0 = 8 * p;
    const int nst = min(8, N - c0);
    const unsigned dmask_in = dmask;
    // ---- F: panel into registers, 8 pivot steps
    double pv[4][8];
    int tst[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      tst[j] = -1;
      const int row = lane + 32 * j;
      const bool live = !((dmask >> j) & 1u);
#pragma unroll
      for (int c = 0; c < 8; ++c) pv[j][c] = (live && c < nst) ? __ldcg(A + (long long)(c0 + c) * NP + row) : 0.0;
    }
    lw_pivot_step<0>(pv, b, tst, dmask, sm.pivrow, sm.rinvs, sm.cand, c0, lane);
    if (nst > 1) lw_pivot_step<1>(pv, b, tst, dmask, sm.pivrow, sm.rinvs, sm.cand, c0, lane);
    if (nst > 2) lw_pivot_step<2>(pv, b, tst, dmask, sm.pivrow, sm.rinvs, sm.cand, c0, lane);
    if (nst > 3) lw_pivot_step<3>(pv, b, tst, dmask, sm.pivrow, sm.rinvs, sm.cand, c0, lane);
    if (nst > 4) lw_pivot_step<4>(pv, b, tst, dmask, sm.pivrow, sm.rinvs, sm.cand, c0, lane);
    if (nst > 5) lw_pivot_step<5>(pv, b, tst, dmask, sm.pivrow, sm.rinvs, sm.cand, c0, lane);
    if (nst > 6) lw_pivot_step<6>(pv, b, tst, dmask, sm.pivrow, sm.rinvs, sm.cand, c0, lane);
    if (nst > 7) lw_pivot_step<7>(pv, b, tst, dmask, sm.pivrow, sm.rinvs, sm.cand, c0, lane);
    __syncwarp();  // every lane is done with the candidate slots (they alias Ubar)

    // ---- publish: multipliers (A-fragment order), rhs of the new pivots, their U entries inside the panel
    const int ntrail = N - (c0 + nst);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = lane + 32 * j;
      if (row < C::NP8) {
        const bool dead = (dmask_in >> j) & 1u;
        if (ntrail > 0) {
          const int tr = row >> 3, gg = row & 7;
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            const bool is_mult = !dead && (tst[j] < 0 || s < tst[j]);
            sm.nLf[tr][s >> 2][gg * 4 + (s & 3)] = is_mult ? -pv[j][s] : 0.0;
          }
        }
        if (tst[j] >= 0) {
          const int tt = tst[j];
          sm.ys[c0 + tt] = b[j];
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            if (s < tt) sm.Lpp[tt][s] = -pv[j][s];
            if (s > tt && s < nst) __stcg(A + (long long)(c0 + s) * NP + row, pv[j][s]);
          }
        }
      }
    }
    __syncwarp();
    if (ntrail > 0) {
      // ---- G: final U rows of the panel over the trailing columns, lane = column (+32 per pass).  The loads of
      // every pass are issued before the first solve: one L2 round trip per panel.
      constexpr int NPASS = (C::NP8 - 8 + 31) / 32;
      int pr[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) pr[t] = sm.pivrow[c0 + t];
      double v[NPASS][8];
#pragma unroll
      for (int ps = 0; ps < NPASS; ++ps) {
        const int cj = 32 * ps + lane;
        const double* col = A + (long long)(c0 + 8 + cj) * NP;
#pragma unroll
        for (int t = 0; t < 8; ++t) v[ps][t] = (cj < ntrail) ? __ldcg(col + pr[t]) : 0.0;
      }
#pragma unroll
      for (int ps = 0; ps < NPASS; ++ps) {
        const int cj = 32 * ps + lane;
        if (cj < ntrail) {
          double* col = A + (long long)(c0 + 8 + cj) * NP;
#pragma unroll
          for (int t = 1; t < 8; ++t) {
            double x = v[ps][t];
#pragma unroll
            for (int s = 0; s < t; ++s) x = fma(sm.Lpp[t][s], v[ps][s], x);
            v[ps][t] = x;
            __stcg(col + pr[t], x);
          }
#pragma unroll
          for (int t = 0; t < 8; ++t) sm.Ubar[t][cj] = v[ps][t];
        }
      }
      __syncwarp();
      // ---- U: C -= L U on the tensor pipe; tile rows with no live row are skipped.  All tiles of a tile row are
      // loaded before the first DMMA: one L2 round trip per live tile row.
      unsigned live_tr = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned bal = __ballot_sync(FULL, !((dmask >> j) & 1u));
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if ((bal >> (8 * q)) & 0xffu) live_tr |= 1u << (4 * j + q);
      }
      constexpr int NTC = NT - 1;
      const int ntc = (ntrail + 7) / 8;
#pragma unroll 1
      for (int tr = 0; tr < NT; ++tr) {
        if (!((live_tr >> tr) & 1u)) continue;
        const double a0 = sm.nLf[tr][0][lane], a1 = sm.nLf[tr][1][lane];
        const int row = 8 * tr + g;
        const bool rok = row < N;
        double* rbase = A + (long long)(c0 + 8 + 2 * q4) * NP + row;
        double c[NTC][2];
#pragma unroll
        for (int k = 0; k < NTC; ++k) {
#pragma unroll
          for (int e = 0; e < 2; ++e)
            c[k][e] = (rok && 8 * k + 2 * q4 + e < ntrail) ? __ldcg(rbase + (long long)(8 * k + e) * NP) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < NTC; ++k) {
          if (k < ntc) {  // uniform
            const double b0 = sm.Ubar[q4][8 * k + g], b1 = sm.Ubar[4 + q4][8 * k + g];
            dmma884(c[k][0], c[k][1], a0, b0);
            dmma884(c[k][0], c[k][1], a1, b1);
          }
        }
#pragma unroll
        for (int k = 0; k < NTC; ++k) {
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (rok && 8 * k + 2 * q4 + e < ntrail) __stcg(rbase + (long long)(8 * k + e) * NP, c[k][e]);
        }
      }
      __syncwarp();
    }
  }

  // ---- back-substitution U x = y by panels in reverse; ys is overwritten with the solution.
  // U[k][j] (j > k) = A[pivrow[k]][j]
  for (int p = npan - 1; p >= 0; --p) {
    const int c0 = 8 * p;
    const int nst = min(8, N - c0);
    // prefetch the entries of the earlier pivot rows in this panel's columns (independent of the solve below)
    double uv[C::R][8];
#pragma unroll
    for (int r = 0; r < C::R; ++r) {
      const int k = lane + 32 * r;
      const bool act = k < c0;
      const int pr = act ? sm.pivrow[k] : 0;
#pragma unroll
      for (int s = 0; s < 8; ++s) uv[r][s] = (act && s < nst) ? __ldcg(A + (long long)(c0 + s) * NP + pr) : 0.0;
    }
    {  // 8x8 upper-triangular solve: lane t < nst owns pivot c0 + t
      const bool act = lane < nst;
      const int pr = act ? sm.pivrow[c0 + lane] : 0;
      double y = act ? sm.ys[c0 + lane] : 0.0;
      const double ri = act ? sm.rinvs[c0 + lane] : 0.0;
      double ud[8];
#pragma unroll
      for (int s = 1; s < 8; ++s) ud[s] = (act && s > lane && s < nst) ? __ldcg(A + (long long)(c0 + s) * NP + pr) : 0.0;
      double xo = 0.0;
#pragma unroll
      for (int t = 7; t >= 0; --t) {
        if (t < nst) {  // uniform
          const double xt = __shfl_sync(FULL, y * ri, t);
          if (lane == t) xo = xt;
          if (t > 0) y = fma(-ud[t], xt, y);
        }
      }
      if (act) sm.ys[c0 + lane] = xo;
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < C::R; ++r) {
      const int k = lane + 32 * r;
      if (k < c0) {
        double acc = sm.ys[k];
#pragma unroll
        for (int s = 0; s < 8; ++s)
          if (s < nst) acc = fma(-uv[r][s], sm.ys[c0 + s], acc);
        sm.ys[k] = acc;
      }
    }
    __syncwarp();
  }
  double* u = args.U + sys * (long long)args.ldu;
#pragma unroll
  for (int r = 0; r < C::R; ++r) {
    const int k = lane + 32 * r;
    if (k < N) u[k] = sm.ys[k];
  }
}

// Systems in flight per SM.  The in-flight matrices must fit the 126 MB L2 (80 KB each at N = 100 -> at most ~10 per SM);
// occupancy is capped by padding the CTA's shared-memory footprint with unused dynamic shared memory.
static int lw_target_occupancy() {
  static const int occ = [] {
    const char* e = getenv("RB_LW_OCC");
    const int v = e ? atoi(e) : 8;
    return v < 1 ? 1 : (v > 16 ? 16 : v);
  }();
  return occ;
}

template <int NT>
static int lw_launch(const LuArgs& args, double* Awork, cudaStream_t stream) {
  const int occ = lw_target_occupancy();
  const long long per_cta = (227 * 1024) / occ - 1024;  // 1 KB per CTA is reserved by the driver
  long long pad = per_cta - (long long)sizeof(LwSmem<NT>);
  if (pad < 0) pad = 0;
  pad = pad / 128 * 128;
  cudaError_t e = cudaFuncSetAttribute(rb_lu_warp_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
  if (e != cudaSuccess) return (int)e;
  rb_lu_warp_kernel<NT><<<(unsigned)args.count, 32, (size_t)pad, stream>>>(args, Awork);
  return (int)cudaGetLastError();
}

int rb_lu_warp_launch(const LuArgs& args, double* Awork, cudaStream_t stream) {
  switch ((args.N + 7) / 8) {
    case 5: return lw_launch<5>(args, Awork, stream);
    case 6: return lw_launch<6>(args, Awork, stream);
    case 7: return lw_launch<7>(args, Awork, stream);
    case 8: return lw_launch<8>(args, Awork, stream);
    case 9: return lw_launch<9>(args, Awork, stream);
    case 10: return lw_launch<10>(args, Awork, stream);
    case 11: return lw_launch<11>(args, Awork, stream);
    case 12: return lw_launch<12>(args, Awork, stream);
    case 13: return lw_launch<13>(args, Awork, stream);
    case 14: return lw_launch<14>(args, Awork, stream);
    case 15: return lw_launch<15>(args, Awork, stream);
    case 16: return lw_launch<16>(args, Awork, stream);
    default: return -1000;
  }
}

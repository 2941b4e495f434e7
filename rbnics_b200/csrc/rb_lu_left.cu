// rb_lu_left.cu -- batched LU (partial pivoting) + solves for 32 < N <= 128: one warp per system, LEFT-LOOKING, with
// the finished L panels streamed back from L2 by asynchronous copies.
//
// Replaces numpy.linalg.solve (LAPACK dgesv; rbnics/backends/online/numpy/linear_solver.py:29-31) and the lifting rows
// (rbnics/backends/online/basic/wrapping/DirichletBC.py:37-56) for a chunk of parameters.
//
// Why left-looking: the pivot chain of a system is sequential, so an SM must interleave many systems, and then the
// matrices cannot live on chip.  A right-looking update reads AND writes the trailing matrix in global memory after
// every panel (latency-bound read-modify-write; measured, then dropped).  Left-looking, panel p (8 columns) is read ONCE from
// the assembled matrix into DMMA accumulator tiles that stay in registers while the records of the finished panels
// q < p stream by:
//     record q = { negated multipliers of panel q in DMMA A-fragment order, inverse of its 8x8 unit-lower block,
//                  its pivot rows, mask of tile rows with non-zero multipliers }
// written once, read-only afterwards, contiguous (7.2 KB at N = 100) -> prefetched with cp.async into a 2-stage
// shared-memory ring.  Per record:  X = the panel rows pivoted in q (final since the records >= q have zero
// multipliers for them), U_qp = L_qq^-1 X (2 DMMA.8x8x4), C -= L_q U_qp (2 DMMA.8x8x4 per live tile row).
// Then the panel is transposed through shared memory to one row set per lane and factored exactly like
// rb_lu_dmma.cuh (REDUX.MAX arg-max, idamax tie rule, implicit pivoting, rhs carried along).
// U is stored in 8x8 blocks, block column p contiguous, so the back-substitution reads coalesced rows.
// CTA = one warp; no block-level barrier; ~24 KB of shared memory -> 8 systems per SM.
// Measured alternative: filling the ring with one cp.async.bulk (TMA) per record instead of 14 cp.async per lane is
// 2 % SLOWER (higher copy latency at a prefetch distance of one record outweighs the saved issue slots).
#include <stdlib.h>

#include "rb_lu_kernel.cuh"

template <int NT>
struct LlCfg {
  static constexpr int NP8 = 8 * NT;
  static constexpr int R = (NP8 + 31) / 32;        // pivot indices / rows per lane
  static constexpr int REC = 64 + 8 + NT * 64;     // doubles per record: L_qq^-1 | pivrow[8], live mask | packed nLf
  static constexpr int SLOT = REC < 640 ? 640 : REC;  // ring slot (doubles): slot 0 also hosts the 640 candidate doubles
  static constexpr int PP = 8;                     // pitch of the transpose buffer (dump: conflict-free STS.128)
  static constexpr int UBLK = NT * (NT + 1) / 2;   // 8x8 blocks of the packed upper-triangular U store
  // scratch doubles per system beyond the assembled matrix: records of all panels + packed U
  static constexpr long long SCRATCH = (long long)NT * REC + (long long)UBLK * 64;
};

template <int NT>
constexpr long long ll_scratch_doubles() {
  return LlCfg<NT>::SCRATCH;
}

template <int NT>
struct LlSmem {
  using C = LlCfg<NT>;
  union {
    double ring[2][C::SLOT];    // prefetched records / staged panel of A / staged block column of U
    double cand[2][32][10];     // pivot candidate slots of phase F (the ring is idle then)
  };
  union {
    double P[C::NP8][C::PP];    // panel, row-major, for the gather of pivot rows and the transpose into phase F
    double rec[C::REC];         // record of the panel just factored, staged for the coalesced copy-out
  };
  double Ub[8][8];              // U_qp of the record being applied (B fragments)
  double ys[C::NP8];            // rhs in pivot order -> solution
  double rinvs[C::NP8];         // reciprocal pivots
  int pivrow[C::NP8];           // pivot order -> row
  int recsz[NT];                // 16-byte chunks of record q (header + live tile rows)
};

__device__ __forceinline__ double ll_selp(double a, double b, int x, int y) {  // x == y ? a : b, branch-free
  double r;
  asm("{\n .reg .pred p;\n setp.eq.s32 p, %3, %4;\n selp.f64 %0, %1, %2, p;\n}" : "=d"(r) : "d"(a), "d"(b), "r"(x), "r"(y));
  return r;
}
__device__ __forceinline__ void ll_cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void ll_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ll_cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// One pivot step T of phase F (a template: the register indices must be static).  Same arithmetic as lud_pivot_step of rb_lu_dmma.cuh, plus the rhs entry.
template <int T>
__device__ __forceinline__ void ll_pivot_step(double (&pv)[4][8], double (&b)[4], int (&tst)[4], unsigned& dmask,
                                              int* pivrow, double* rinvs, double (*cand)[32][10], int c0, int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int t = T;
  unsigned long long key = 0ull;
  int jb = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const unsigned long long kj =
        ((dmask >> j) & 1u) ? 0ull : ((unsigned long long)__double_as_longlong(fabs(pv[j][t])) + 1ull);
    if (kj > key) {  // strict: the lower row of the lane wins ties
      key = kj;
      jb = j;
    }
  }
  // arg-max of the 64-bit key over the warp, lowest row on ties (idamax): REDUX.MAX on the high word, then on
  // {low word without its 7 lowest bits, 127 - row}; candidates that differ only below 2^-45 relative count as equal.
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mhi = __reduce_max_sync(FULL, hi);
  double sel[8];
#pragma unroll
  for (int c = t; c < 8; ++c) {
    double v = pv[0][c];
    v = ll_selp(pv[1][c], v, jb, 1);
    v = ll_selp(pv[2][c], v, jb, 2);
    v = ll_selp(pv[3][c], v, jb, 3);
    sel[c] = v;
  }
  double selb = b[0];
  selb = ll_selp(b[1], selb, jb, 1);
  selb = ll_selp(b[2], selb, jb, 2);
  selb = ll_selp(b[3], selb, jb, 3);
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
__global__ void __launch_bounds__(32, 8) rb_lu_left_kernel(LuArgs args, double* __restrict__ scratch) {
  using C = LlCfg<NT>;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ __align__(16) LlSmem<NT> sm;
  const int lane = threadIdx.x;
  const long long sys = blockIdx.x;
  const int N = args.N;
  const int NP = (N + 3) / 4 * 4;  // leading dimension of the assembled matrices (column-major)
  const double* A = args.Ab + sys * (long long)args.EP;
  double* Lrec = scratch + sys * C::SCRATCH;   // [NT][REC]
  double* Ust = Lrec + (long long)NT * C::REC;               // block (q, p), q <= p, at (p(p+1)/2 + q) * 64, [t][s]
  const int g = lane >> 2, q4 = lane & 3;

  // ---- rhs of this lane's rows (rows lane+32j), lifting rows
  double b[4];
  unsigned dmask = 0u;       // bit j: row lane+32j is used up (or does not exist)
  unsigned bc_frag = 0u;     // bit tr: row 8tr+g is a lifting row (fragment layout of the panel tiles)
  {
    const double* th = args.theta + sys * (long long)args.QT + args.Ql;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = lane + 32 * j;
      double v = 0.0;
      if (row < N) {
        if (args.rhs_in)
          v = args.rhs_in[sys * (long long)N + row];
        else
          for (int q = 0; q < args.Qr; ++q) v = fma(__ldg(th + q), __ldg(args.F + q * NP + row), v);
        for (int t = 0; t < args.n_bc; ++t)
          if (__ldg(args.bc_rows + t) == row) v = __ldg(args.thetaBC + sys * (long long)args.n_bc + t);
      } else {
        dmask |= 1u << j;
      }
      b[j] = v;
    }
    for (int t = 0; t < args.n_bc; ++t) {
      const int r = __ldg(args.bc_rows + t);
      if ((r & 7) == g) bc_frag |= 1u << (r >> 3);
    }
  }

  const int npan = (N + 7) / 8;
  // 16-byte chunk copiers (cp.async, one commit group per call site)
  auto issue_chunks = [&](double* dst, const double* src, int n16, int maxtrips) {
    for (int k = 0; k < maxtrips; ++k) {
      const int i = lane + 32 * k;
      if (i < n16) ll_cp_async16(dst + 2 * i, src + 2 * i);
    }
  };
  auto issue_record = [&](int q, int slot) {
    issue_chunks(sm.ring[slot], Lrec + (long long)q * C::REC, sm.recsz[q], (C::REC / 2 + 31) / 32);
  };
  auto issue_apanel = [&](int pn) {  // 8 columns of the assembled matrix are contiguous: -> ring slot 1
    issue_chunks(sm.ring[1], A + (long long)(8 * pn) * NP, min(8, N - 8 * pn) * NP / 2, NT + 1);
  };
  issue_apanel(0);
  ll_cp_commit();

  for (int p = 0; p < npan; ++p) {
    const int c0 = 8 * p;
    const int nst = min(8, N - c0);
    // ---- first record on its way, then the staged panel of the assembled matrix into accumulator tiles
    if (p > 0) issue_record(0, 0);
    ll_cp_commit();
    ll_cp_wait<1>();  // the panel (older group) has landed
    __syncwarp();
    double c[NT][2];
    {
      const double* As = sm.ring[1];
#pragma unroll
      for (int tr = 0; tr < NT; ++tr) {
        const int row = 8 * tr + g;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int cl = 2 * q4 + e;
          double v = 0.0;
          if (row < N && cl < nst) {
            v = As[cl * NP + row];
            if ((bc_frag >> tr) & 1u) v = (c0 + cl == row) ? 1.0 : 0.0;
          }
          c[tr][e] = v;
        }
      }
    }
    __syncwarp();  // ring slot 1 is free for record 1
    // ---- apply the records q < p
    for (int q = 0; q < p; ++q) {
      if (q + 1 < p) issue_record(q + 1, (q + 1) & 1);
      ll_cp_commit();
      // dump the panel: the rows pivoted in panel q are final here
#pragma unroll
      for (int tr = 0; tr < NT; ++tr)
        *reinterpret_cast<double2*>(&sm.P[8 * tr + g][2 * q4]) = make_double2(c[tr][0], c[tr][1]);
      ll_cp_wait<1>();
      __syncwarp();
      const double* rq = sm.ring[q & 1];
      const int* prq = reinterpret_cast<const int*>(rq + 64);
      // U_qp = L_qq^-1 X on the tensor pipe (the inverse of the 8x8 unit-lower block travels in the record)
      double u0 = 0.0, u1 = 0.0;
      {
        const double la0 = rq[lane], la1 = rq[32 + lane];
        const double xb0 = sm.P[prq[q4]][g], xb1 = sm.P[prq[4 + q4]][g];
        dmma884(u0, u1, la0, xb0);
        dmma884(u0, u1, la1, xb1);
      }
      // U block (q, p) -> packed store straight from the accumulator layout (row g, columns 2q4, 2q4+1: 16 B per
      // lane, 512 B contiguous per warp; read once, at the very end)
      __stcs(reinterpret_cast<double2*>(Ust + ((long long)(p * (p + 1) / 2 + q)) * 64 + 8 * g + 2 * q4),
             make_double2(u0, u1));
      // accumulator layout -> B-fragment layout by shuffles: b0 = U[q4][g], b1 = U[4 + q4][g]
      double b0, b1;
      {
        const int src0 = q4 * 4 + (g >> 1), src1 = (4 + q4) * 4 + (g >> 1);
        const double e0 = __shfl_sync(FULL, u0, src0), o0 = __shfl_sync(FULL, u1, src0);
        const double e1 = __shfl_sync(FULL, u0, src1), o1 = __shfl_sync(FULL, u1, src1);
        b0 = (g & 1) ? o0 : e0;
        b1 = (g & 1) ? o1 : e1;
      }
      const unsigned live_q = (unsigned)prq[8];  // tile rows with non-zero multipliers (packed in this order)
      const double* af = rq + 72 + lane;
#pragma unroll
      for (int tr = 0; tr < NT; ++tr) {
        if ((live_q >> tr) & 1u) {  // uniform
          const double a0 = af[0], a1 = af[32];
          dmma884(c[tr][0], c[tr][1], a0, b0);
          dmma884(c[tr][0], c[tr][1], a1, b1);
          af += 64;
        }
      }
      __syncwarp();  // all lanes are done with this ring slot and P
    }
    ll_cp_wait<0>();
    // ---- transpose into phase F: lane owns rows lane+32j
#pragma unroll
    for (int tr = 0; tr < NT; ++tr)
      *reinterpret_cast<double2*>(&sm.P[8 * tr + g][2 * q4]) = make_double2(c[tr][0], c[tr][1]);
    __syncwarp();
    const unsigned dmask_in = dmask;
    double pv[4][8];
    int tst[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      tst[j] = -1;
      const int row = lane + 32 * j;
      const bool live = !((dmask >> j) & 1u);
#pragma unroll
      for (int cc = 0; cc < 8; cc += 2) {
        double2 v = make_double2(0.0, 0.0);
        if (live) v = *reinterpret_cast<const double2*>(&sm.P[row][cc]);
        pv[j][cc] = v.x;
        pv[j][cc + 1] = v.y;
      }
    }
    __syncwarp();
    // the next panel of the assembled matrix streams into ring slot 1 while this one is factored (slot 0 = cand)
    if (p + 1 < npan) issue_apanel(p + 1);
    ll_cp_commit();
    ll_pivot_step<0>(pv, b, tst, dmask, sm.pivrow, sm.rinvs, sm.cand, c0, lane);
    if (nst > 1) ll_pivot_step<1>(pv, b, tst, dmask, sm.pivrow, sm.rinvs, sm.cand, c0, lane);
    if (nst > 2) ll_pivot_step<2>(pv, b, tst, dmask, sm.pivrow, sm.rinvs, sm.cand, c0, lane);
    if (nst > 3) ll_pivot_step<3>(pv, b, tst, dmask, sm.pivrow, sm.rinvs, sm.cand, c0, lane);
    if (nst > 4) ll_pivot_step<4>(pv, b, tst, dmask, sm.pivrow, sm.rinvs, sm.cand, c0, lane);
    if (nst > 5) ll_pivot_step<5>(pv, b, tst, dmask, sm.pivrow, sm.rinvs, sm.cand, c0, lane);
    if (nst > 6) ll_pivot_step<6>(pv, b, tst, dmask, sm.pivrow, sm.rinvs, sm.cand, c0, lane);
    if (nst > 7) ll_pivot_step<7>(pv, b, tst, dmask, sm.pivrow, sm.rinvs, sm.cand, c0, lane);
    __syncwarp();

    // ---- publish: rhs and diagonal U block of the new pivots; the record of this panel
    const bool last = (p + 1 == npan);
    double* rec = sm.rec;  // aliases P (dead now)
    // tile rows that held a live row BEFORE this panel have (possibly) non-zero multipliers: only those are stored
    unsigned live_tr = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned bal = __ballot_sync(FULL, !((dmask_in >> j) & 1u));
#pragma unroll
      for (int qq = 0; qq < 4; ++qq)
        if ((bal >> (8 * qq)) & 0xffu) live_tr |= 1u << (4 * j + qq);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = lane + 32 * j;
      if (row < C::NP8) {
        const bool dead = (dmask_in >> j) & 1u;
        const int tr = row >> 3, gg = row & 7;
        if (!last && ((live_tr >> tr) & 1u)) {
          double* dstf = rec + 72 + 64 * __popc(live_tr & ((1u << tr) - 1u)) + gg * 4;
          double nl[8];
#pragma unroll
          for (int s = 0; s < 8; ++s) nl[s] = (!dead && (tst[j] < 0 || s < tst[j])) ? -pv[j][s] : 0.0;
          *reinterpret_cast<double2*>(dstf) = make_double2(nl[0], nl[1]);
          *reinterpret_cast<double2*>(dstf + 2) = make_double2(nl[2], nl[3]);
          *reinterpret_cast<double2*>(dstf + 32) = make_double2(nl[4], nl[5]);
          *reinterpret_cast<double2*>(dstf + 34) = make_double2(nl[6], nl[7]);
        }
        if (tst[j] >= 0) {
          const int tt = tst[j];
          sm.ys[c0 + tt] = b[j];
          double* ud = Ust + ((long long)(p * (p + 1) / 2 + p)) * 64 + tt * 8;
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            if (!last && s < tt) sm.Ub[tt][s] = -pv[j][s];
            if (s > tt && s < nst) __stcs(ud + s, pv[j][s]);
          }
        }
      }
    }
    if (!last) {
      const int nlive = __popc(live_tr);
      int* pri = reinterpret_cast<int*>(rec + 64);
      if (lane < 8) pri[lane] = sm.pivrow[c0 + lane];
      if (lane == 8) pri[8] = (int)live_tr;
      if (lane == 9) sm.recsz[p] = (72 + 64 * nlive) / 2;
      __syncwarp();
      if (lane < 8) {  // column `lane` of L_qq^-1 (unit lower; Ub holds the negated multipliers), A-fragment order
        double x[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          double acc = (t == lane) ? 1.0 : 0.0;
#pragma unroll
          for (int s2 = 0; s2 < t; ++s2)
            if (s2 >= lane) acc = fma(sm.Ub[t][s2], x[s2], acc);
          x[t] = (t >= lane) ? acc : 0.0;
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) rec[(lane >> 2) * 32 + t * 4 + (lane & 3)] = x[t];
      }
      __syncwarp();
      double* dst = Lrec + (long long)p * C::REC;
      const int n16 = (72 + 64 * nlive) / 2;
      for (int i = lane; i < n16; i += 32)
        __stcg(reinterpret_cast<double2*>(dst) + i, *reinterpret_cast<const double2*>(rec + 2 * i));
    }
    __syncwarp();
  }

  // ---- back-substitution U x = y by panels in reverse; ys is overwritten with the solution.
  // Block column p of the packed U store is contiguous ((p + 1) blocks: rows of the pivots k < 8p, then the diagonal
  // block) and is staged one panel ahead in the (now idle) ring.
  auto issue_ucol = [&](int pc) {
    issue_chunks(sm.ring[pc & 1], Ust + ((long long)(pc * (pc + 1) / 2)) * 64, (pc + 1) * 32, NT);
  };
  ll_cp_wait<0>();
  __syncwarp();
  issue_ucol(npan - 1);
  ll_cp_commit();
  for (int p = npan - 1; p >= 0; --p) {
    const int c0 = 8 * p;
    const int nst = min(8, N - c0);
    if (p > 0) issue_ucol(p - 1);
    ll_cp_commit();
    ll_cp_wait<1>();
    __syncwarp();
    const double* colp = sm.ring[p & 1];
    {  // 8x8 upper-triangular solve: lane t < nst owns pivot c0 + t
      const bool act = lane < nst;
      double y = act ? sm.ys[c0 + lane] : 0.0;
      const double ri = act ? sm.rinvs[c0 + lane] : 0.0;
      double ud[8];
#pragma unroll
      for (int s = 1; s < 8; ++s) ud[s] = (act && s > lane && s < nst) ? colp[(c0 + lane) * 8 + s] : 0.0;
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
    double xs8[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) xs8[s] = (s < nst) ? sm.ys[c0 + s] : 0.0;
#pragma unroll
    for (int r = 0; r < C::R; ++r) {
      const int k = lane + 32 * r;
      if (k < c0) {
        double acc = sm.ys[k];
#pragma unroll
        for (int s = 0; s < 8; s += 2) {
          const double2 v = *reinterpret_cast<const double2*>(colp + k * 8 + s);
          acc = fma(-v.x, xs8[s], acc);
          acc = fma(-v.y, xs8[s + 1], acc);
        }
        sm.ys[k] = acc;
      }
    }
    __syncwarp();
  }
  ll_cp_wait<0>();
  double* u = args.U + sys * (long long)args.ldu;
#pragma unroll
  for (int r = 0; r < C::R; ++r) {
    const int k = lane + 32 * r;
    if (k < N) u[k] = sm.ys[k];
  }
}

// Systems in flight per SM (A/B knob): occupancy is capped by padding with unused dynamic shared memory.
static int ll_target_occupancy() {
  static const int occ = [] {
    const char* e = getenv("RB_LL_OCC");
    const int v = e ? atoi(e) : 9;
    return v < 1 ? 1 : (v > 9 ? 9 : v);
  }();
  return occ;
}

template <int NT>
static int ll_launch(const LuArgs& args, double* scratch, cudaStream_t stream) {
  const int occ = ll_target_occupancy();
  long long pad = 0;
  if (occ < 9) {
    pad = (227 * 1024) / occ - 1024 - (long long)sizeof(LlSmem<NT>);
    pad = pad < 0 ? 0 : pad / 128 * 128;
    cudaError_t e = cudaFuncSetAttribute(rb_lu_left_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
    if (e != cudaSuccess) return (int)e;
  }
  rb_lu_left_kernel<NT><<<(unsigned)args.count, 32, (size_t)pad, stream>>>(args, scratch);
  return (int)cudaGetLastError();
}

// doubles of scratch per system (records + packed U) for a given N, or -1 if the kernel does not serve this N
long long rb_lu_left_scratch_doubles(int N) {
  switch ((N + 7) / 8) {
    case 5: return ll_scratch_doubles<5>();
    case 6: return ll_scratch_doubles<6>();
    case 7: return ll_scratch_doubles<7>();
    case 8: return ll_scratch_doubles<8>();
    case 9: return ll_scratch_doubles<9>();
    case 10: return ll_scratch_doubles<10>();
    case 11: return ll_scratch_doubles<11>();
    case 12: return ll_scratch_doubles<12>();
    case 13: return ll_scratch_doubles<13>();
    case 14: return ll_scratch_doubles<14>();
    case 15: return ll_scratch_doubles<15>();
    case 16: return ll_scratch_doubles<16>();
    default: return -1;
  }
}

int rb_lu_left_launch(const LuArgs& args, double* scratch, cudaStream_t stream) {
  switch ((args.N + 7) / 8) {
    case 5: return ll_launch<5>(args, scratch, stream);
    case 6: return ll_launch<6>(args, scratch, stream);
    case 7: return ll_launch<7>(args, scratch, stream);
    case 8: return ll_launch<8>(args, scratch, stream);
    case 9: return ll_launch<9>(args, scratch, stream);
    case 10: return ll_launch<10>(args, scratch, stream);
    case 11: return ll_launch<11>(args, scratch, stream);
    case 12: return ll_launch<12>(args, scratch, stream);
    case 13: return ll_launch<13>(args, scratch, stream);
    case 14: return ll_launch<14>(args, scratch, stream);
    case 15: return ll_launch<15>(args, scratch, stream);
    case 16: return ll_launch<16>(args, scratch, stream);
    default: return -1000;
  }
}

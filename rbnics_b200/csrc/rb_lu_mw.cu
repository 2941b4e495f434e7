// rb_lu_mw.cu -- batched LU (partial pivoting) + solves for 32 < N <= 128: ONE CTA OF UP TO FOUR WARPS PER SYSTEM,
// left-looking, persistent (every CTA walks over systems and re-uses ONE scratch slot, so the L records and the packed
// U stay in L2 instead of round-tripping a multi-GB HBM buffer).
//
// Replaces numpy.linalg.solve (LAPACK dgesv; rbnics/backends/online/numpy/linear_solver.py:29-31) and the lifting rows
// (rbnics/backends/online/basic/wrapping/DirichletBC.py:37-56) for a chunk of parameters.
//
// Same algorithm and record format idea as the one-warp kernel of round 1 (rb_lu_left.cu, kept as an A/B variant):
// panel p (8 columns) is read once into DMMA accumulator tiles that stay in registers while the records of the
// finished panels q < p stream by through a 2-slot cp.async ring; per record U_qp = L_qq^-1 X (2 DMMA) and
// C -= L_q U_qp (2 DMMA per live tile row); then the panel is transposed through shared memory and factored with
// implicit pivoting (two REDUX.MAX per warp on the IEEE bit pattern, LAPACK idamax tie rule), the rhs carried along.
//
// What changed, and why (ncu of the one-warp kernel: 49 k warp instructions per system, 2 warps per scheduler, issue
// slots 35 % busy, 226 registers per thread -> the SM was starved of warps, not of bandwidth):
//   * the 8*NT rows of a system are spread over NW = ceil(NT/4) warps, ONE row per lane in the pivot phase (no 4-way
//     row-select chains, 16 registers of panel per lane instead of 64) and tile rows tr = w, w+NW, .. per warp in the
//     update phase (8 accumulator doubles per lane instead of 26) -> <= 72 registers, 7-8 systems = 28-32 warps per SM;
//   * the U row, the rhs entry and the multipliers of a pivot row are published by its owner at pivot time, so rows that
//     are used up need no predicates afterwards;
//   * U_qp is computed by one warp per record (round robin) and handed to the others through shared memory.
#include <stdlib.h>

#include "rb_lu_kernel.cuh"

template <int NT>
struct MwCfg {
  static constexpr int NW = (NT + 3) / 4;          // warps per system
  static constexpr int NTH = 32 * NW;
  static constexpr int NP8 = 8 * NT;
  static constexpr int TPW = (NT + NW - 1) / NW;   // tile rows per warp
  static constexpr int REC = 72 + 64 * NT;         // doubles per record: L_qq^-1 | pivrow[8], live mask | L tiles
  static constexpr int SLOT = REC > 8 * NP8 ? REC : 8 * NP8;  // ring slot: a record, a staged panel of A, a block column of U
  static constexpr int UBLK = NT * (NT + 1) / 2;   // 8x8 blocks of the packed upper-triangular U store
  static constexpr long long SCRATCH = (long long)NT * REC + (long long)UBLK * 64;  // doubles per scratch slot
};

template <int NT>
struct MwSmem {
  using C = MwCfg<NT>;
  double ring[2][C::SLOT];
  double P[C::NP8][8];                  // the panel, row-major: gather of pivot rows, transpose into the pivot phase
  double Ub[8][8];                      // U_qp of the record being applied / the pivot rows of the panel being factored
  double slot[2][C::NW][10];            // pivot candidates of the warps: row entries t+1..7, 1/pivot, rhs entry
  unsigned long long keys[2][C::NW];
  double ys[C::NP8];                    // rhs in pivot order -> solution
  double rinvs[C::NP8];                 // reciprocal pivots
  int pivrow[C::NP8];                   // pivot order -> row
  int recsz[NT];                        // 16-byte chunks of record q
  unsigned livew[2][C::NW];             // tile rows (4 bits per warp) that still held a live row before this panel
};

__device__ __forceinline__ void mw_cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void mw_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void mw_cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
template <int NTH>
__device__ __forceinline__ void mw_sync() {
  if (NTH == 32)
    __syncwarp();
  else
    __syncthreads();
}

// One pivot step T of the panel (template: register indices must be static).  Row `tid` lives in pv[0..8).
template <int T, int NT>
__device__ __forceinline__ void mw_pivot_step(double (&pv)[8], double& b, bool& dead, int& tst, MwSmem<NT>& sm,
                                              double* __restrict__ udiag, int c0, int tid, int lane, int w) {
  using C = MwCfg<NT>;
  constexpr unsigned FULL = 0xffffffffu;
  const double v = pv[T];
  const unsigned long long key = dead ? 0ull : ((unsigned long long)__double_as_longlong(fabs(v)) + 1ull);
  // arg-max of the 64-bit key over the warp, lowest row on ties (idamax): REDUX.MAX on the high word, then on
  // {low word without its 8 lowest bits, 128 - row}; candidates that differ only below 2^-44 relative count as equal.
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mhi = __reduce_max_sync(FULL, hi);
  const unsigned k2 = (hi == mhi && key != 0ull) ? ((lo & ~255u) | (unsigned)(128 - tid)) : 0u;
  const unsigned m2 = __reduce_max_sync(FULL, k2);
  const double myrinv = 1.0 / v;  // speculative: only the winner's is used
  if (m2 != 0u ? (k2 == m2) : (lane == 0)) {  // the warp's candidate (or "none")
    double* s = sm.slot[T & 1][w];
#pragma unroll
    for (int c = T + 1; c < 8; ++c) s[c - 1] = pv[c];
    s[8] = myrinv;
    s[9] = b;
    sm.keys[T & 1][w] = ((unsigned long long)mhi << 32) | m2;
  }
  mw_sync<C::NTH>();
  unsigned long long best = sm.keys[T & 1][0];
  int bw = 0;
#pragma unroll
  for (int w2 = 1; w2 < C::NW; ++w2) {
    const unsigned long long k = sm.keys[T & 1][w2];
    if (k > best) {
      best = k;
      bw = w2;
    }
  }
  const int prow = 128 - (int)((unsigned)best & 255u);
  const double* ws = sm.slot[T & 1][bw];
  if (tid == prow) {
    // this row becomes pivot c0 + T: publish everything that is final now
    dead = true;
    tst = T;
    sm.pivrow[c0 + T] = prow;
    sm.rinvs[c0 + T] = myrinv;
    sm.ys[c0 + T] = b;
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
      *reinterpret_cast<double2*>(&sm.Ub[T][c]) = make_double2(pv[c], pv[c + 1]);      // multipliers s < T (for L_qq^-1)
      __stcg(reinterpret_cast<double2*>(udiag + T * 8 + c), make_double2(pv[c], pv[c + 1]));  // U entries s > T
    }
  } else {
    // rows that are used up compute garbage in columns > T, which nobody reads any more
    const double rinv = ws[8];
    const double l = v * rinv;
    pv[T] = l;
#pragma unroll
    for (int c = T + 1; c < 8; ++c) pv[c] = fma(-l, ws[c - 1], pv[c]);
    b = fma(-l, ws[9], b);
  }
}

template <int NT, int OCC>
__global__ void __launch_bounds__(MwCfg<NT>::NTH, OCC) rb_lu_mw_kernel(LuArgs args, double* __restrict__ scratch) {
  using C = MwCfg<NT>;
  constexpr int NW = C::NW, NTH = C::NTH, TPW = C::TPW;
  __shared__ __align__(16) MwSmem<NT> sm;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, q4 = lane & 3;
  const int N = args.N;
  const int NP = (N + 3) / 4 * 4;  // leading dimension of the assembled matrices (column-major)
  const int npan = (N + 7) / 8;
  double* Lrec = scratch + (long long)blockIdx.x * C::SCRATCH;  // [NT][REC]
  double* Ust = Lrec + (long long)NT * C::REC;                  // block (q, p), q <= p, at (p(p+1)/2 + q) * 64, [t][s]

  auto issue_chunks = [&](double* dst, const double* src, int n16) {
    for (int i = tid; i < n16; i += NTH) mw_cp_async16(dst + 2 * i, src + 2 * i);
  };

  for (long long sys = blockIdx.x; sys < args.count; sys += gridDim.x) {
    const double* A = args.Ab + sys * (long long)args.EP;
    auto issue_apanel = [&](int pn) {  // 8 columns of the assembled matrix are contiguous: -> ring slot 1
      issue_chunks(sm.ring[1], A + (long long)(8 * pn) * NP, min(8, N - 8 * pn) * NP / 2);
    };
    issue_apanel(0);
    mw_cp_commit();

    // ---- rhs entry of this lane's row, lifting rows
    double b = 0.0;
    bool dead = tid >= N;
    unsigned bc_frag = 0u;  // bit i: row 8 (w + NW i) + g is a lifting row (fragment layout of the panel tiles)
    {
      const double* th = args.theta + sys * (long long)args.QT + args.Ql;
      if (tid < N) {
        if (args.rhs_in)
          b = args.rhs_in[sys * (long long)N + tid];
        else
          for (int q = 0; q < args.Qr; ++q) b = fma(__ldg(th + q), __ldg(args.F + q * NP + tid), b);
      }
      for (int t = 0; t < args.n_bc; ++t) {
        const int r = __ldg(args.bc_rows + t);
        if (r == tid) b = __ldg(args.thetaBC + sys * (long long)args.n_bc + t);
        if ((r & 7) == g && ((r >> 3) % NW) == w) bc_frag |= 1u << ((r >> 3) / NW);
      }
    }

    for (int p = 0; p < npan; ++p) {
      const int c0 = 8 * p;
      const int nst = min(8, N - c0);
      // tile rows that hold a live row before this panel: only those get multipliers in the record of this panel
      {
        const unsigned bal = __ballot_sync(0xffffffffu, !dead);
        if (lane == 0)
          sm.livew[p & 1][w] = ((bal & 0xffu) ? 1u : 0u) | ((bal & 0xff00u) ? 2u : 0u) | ((bal & 0xff0000u) ? 4u : 0u) |
                        ((bal & 0xff000000u) ? 8u : 0u);
      }
      // ---- the staged panel of the assembled matrix into accumulator tiles, the first record on its way
      mw_cp_wait<0>();
      mw_sync<NTH>();  // the panel has landed; the record of the previous panel is complete
      if (p > 0) {
        issue_chunks(sm.ring[0], Lrec, sm.recsz[0]);
        mw_cp_commit();
      }
      double c[TPW][2];
      {
        const double* As = sm.ring[1];
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
          const int tr = w + NW * i;
          const int row = 8 * tr + g;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int cl = 2 * q4 + e;
            double v = 0.0;
            if (tr < NT && row < N && cl < nst) {
              v = As[cl * NP + row];
              if ((bc_frag >> i) & 1u) v = (c0 + cl == row) ? 1.0 : 0.0;
            }
            c[i][e] = v;
          }
        }
      }
      // ---- apply the records q < p
      for (int q = 0; q < p; ++q) {
        // dump the panel: the rows pivoted in panel q are final here
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
          const int tr = w + NW * i;
          if (tr < NT) *reinterpret_cast<double2*>(&sm.P[8 * tr + g][2 * q4]) = make_double2(c[i][0], c[i][1]);
        }
        mw_cp_wait<0>();
        mw_sync<NTH>();  // (A) record q and the dump are visible; everybody is done with record q - 1
        if (q + 1 < p) {
          issue_chunks(sm.ring[(q + 1) & 1], Lrec + (long long)(q + 1) * C::REC, sm.recsz[q + 1]);
          mw_cp_commit();
        }
        const double* rq = sm.ring[q & 1];
        const int* prq = reinterpret_cast<const int*>(rq + 64);
        if (w == q % NW) {
          // U_qp = L_qq^-1 X on the tensor pipe (the inverse of the 8x8 unit-lower block travels in the record)
          const double2 la = *reinterpret_cast<const double2*>(rq + 2 * lane);
          const double xb0 = sm.P[prq[q4]][g], xb1 = sm.P[prq[4 + q4]][g];
          double u0 = 0.0, u1 = 0.0;
          dmma884(u0, u1, la.x, xb0);
          dmma884(u0, u1, la.y, xb1);
          *reinterpret_cast<double2*>(&sm.Ub[g][2 * q4]) = make_double2(u0, u1);
          // U block (q, p) -> packed store straight from the accumulator layout (read once, at the very end)
          __stcg(reinterpret_cast<double2*>(Ust + ((long long)(p * (p + 1) / 2 + q)) * 64 + 8 * g + 2 * q4),
                 make_double2(u0, u1));
        }
        mw_sync<NTH>();  // (B) U_qp is visible; P may be overwritten
        const double b0 = sm.Ub[q4][g], b1 = sm.Ub[4 + q4][g];
        const unsigned live_q = (unsigned)prq[8];  // tile rows with non-zero multipliers (packed in this order)
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
          const int tr = w + NW * i;
          if (tr < NT && ((live_q >> tr) & 1u)) {  // uniform
            const int ci = __popc(live_q & ((1u << tr) - 1u));
            const double2 a = *reinterpret_cast<const double2*>(rq + 72 + 64 * ci + 2 * lane);
            dmma884(c[i][0], c[i][1], a.x, b0);
            dmma884(c[i][0], c[i][1], a.y, b1);
          }
        }
      }
      // ---- transpose into the pivot phase: thread tid owns row tid
#pragma unroll
      for (int i = 0; i < TPW; ++i) {
        const int tr = w + NW * i;
        if (tr < NT) *reinterpret_cast<double2*>(&sm.P[8 * tr + g][2 * q4]) = make_double2(c[i][0], c[i][1]);
      }
      mw_sync<NTH>();  // everybody is done with the last record: both ring slots are free
      const bool dead_in = dead;
      double pv[8];
#pragma unroll
      for (int cc = 0; cc < 8; cc += 2) {
        double2 v = make_double2(0.0, 0.0);
        if (!dead) v = *reinterpret_cast<const double2*>(&sm.P[tid][cc]);
        pv[cc] = v.x;
        pv[cc + 1] = v.y;
      }
      int tst = -1;
      // the next panel of the assembled matrix streams into ring slot 1 while this one is factored
      if (p + 1 < npan) issue_apanel(p + 1);
      mw_cp_commit();
      double* udiag = Ust + ((long long)(p * (p + 1) / 2 + p)) * 64;
      mw_pivot_step<0>(pv, b, dead, tst, sm, udiag, c0, tid, lane, w);
      if (nst > 1) mw_pivot_step<1>(pv, b, dead, tst, sm, udiag, c0, tid, lane, w);
      if (nst > 2) mw_pivot_step<2>(pv, b, dead, tst, sm, udiag, c0, tid, lane, w);
      if (nst > 3) mw_pivot_step<3>(pv, b, dead, tst, sm, udiag, c0, tid, lane, w);
      if (nst > 4) mw_pivot_step<4>(pv, b, dead, tst, sm, udiag, c0, tid, lane, w);
      if (nst > 5) mw_pivot_step<5>(pv, b, dead, tst, sm, udiag, c0, tid, lane, w);
      if (nst > 6) mw_pivot_step<6>(pv, b, dead, tst, sm, udiag, c0, tid, lane, w);
      if (nst > 7) mw_pivot_step<7>(pv, b, dead, tst, sm, udiag, c0, tid, lane, w);
      mw_sync<NTH>();  // Ub (multipliers of the pivot rows), pivrow are complete

      // ---- the record of this panel (not needed after the last one)
      if (p + 1 < npan) {
        unsigned live_tr = 0u;
#pragma unroll
        for (int w2 = 0; w2 < NW; ++w2) live_tr |= sm.livew[p & 1][w2] << (4 * w2);
        double* dst = Lrec + (long long)p * C::REC;
        const int tr = tid >> 3, gg = tid & 7;
        if (tid < C::NP8 && ((live_tr >> tr) & 1u)) {
          // negated multipliers of row tid in DMMA A-fragment order, both k-steps of a lane adjacent (one LDS.128)
          double nl[8];
#pragma unroll
          for (int s = 0; s < 8; ++s) nl[s] = (!dead_in && (tst < 0 || s < tst)) ? -pv[s] : 0.0;
          double2* d2 = reinterpret_cast<double2*>(dst + 72 + 64 * __popc(live_tr & ((1u << tr) - 1u)) + gg * 8);
          __stcg(d2 + 0, make_double2(nl[0], nl[4]));
          __stcg(d2 + 1, make_double2(nl[1], nl[5]));
          __stcg(d2 + 2, make_double2(nl[2], nl[6]));
          __stcg(d2 + 3, make_double2(nl[3], nl[7]));
        }
        if (tid < 8) {  // column `tid` of L_qq^-1 (unit lower; Ub holds the multipliers), A-fragment order
          double x[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            double acc = (t == tid) ? 1.0 : 0.0;
#pragma unroll
            for (int s2 = 0; s2 < t; ++s2)
              if (s2 >= tid) acc = fma(-sm.Ub[t][s2], x[s2], acc);
            x[t] = (t >= tid && t < nst) ? acc : ((t == tid) ? 1.0 : 0.0);
          }
          // entry (t, tid): lane (g = t, q4 = tid & 3), k-step tid >> 2
#pragma unroll
          for (int t = 0; t < 8; ++t) __stcg(dst + (t * 4 + (tid & 3)) * 2 + (tid >> 2), x[t]);
          reinterpret_cast<int*>(dst + 64)[tid] = sm.pivrow[c0 + tid];
        }
        if (tid == 8) {
          reinterpret_cast<int*>(dst + 64)[8] = (int)live_tr;
          sm.recsz[p] = (72 + 64 * __popc(live_tr)) / 2;
        }
      }
      // the barrier at the top of the next panel orders these writes before the record prefetches
    }

    // ---- back-substitution U x = y by panels in reverse; ys is overwritten with the solution.
    // Block column p of the packed U store is contiguous ((p + 1) blocks: rows of the pivots k < 8p, then the diagonal
    // block) and is staged one panel ahead in the (now idle) ring.
    mw_cp_wait<0>();
    mw_sync<NTH>();
    issue_chunks(sm.ring[(npan - 1) & 1], Ust + ((long long)((npan - 1) * npan / 2)) * 64, npan * 32);
    mw_cp_commit();
    for (int p = npan - 1; p >= 0; --p) {
      const int c0 = 8 * p;
      const int nst = min(8, N - c0);
      if (p > 0) {
        issue_chunks(sm.ring[(p - 1) & 1], Ust + ((long long)((p - 1) * p / 2)) * 64, p * 32);
        mw_cp_commit();
        mw_cp_wait<1>();
      } else {
        mw_cp_wait<0>();
      }
      mw_sync<NTH>();
      const double* colp = sm.ring[p & 1];
      if (w == 0) {  // 8x8 upper-triangular solve: lane t < nst owns pivot c0 + t
        constexpr unsigned FULL = 0xffffffffu;
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
      mw_sync<NTH>();
      if (tid < c0) {
        double acc = sm.ys[tid];
#pragma unroll
        for (int s = 0; s < 8; s += 2) {
          const double2 v = *reinterpret_cast<const double2*>(colp + tid * 8 + s);
          const double2 x = *reinterpret_cast<const double2*>(&sm.ys[c0 + s]);
          acc = fma(-v.x, (s < nst) ? x.x : 0.0, acc);
          acc = fma(-v.y, (s + 1 < nst) ? x.y : 0.0, acc);
        }
        sm.ys[tid] = acc;
      }
      mw_sync<NTH>();
    }
    if (tid < N) args.U[sys * (long long)args.ldu + tid] = sm.ys[tid];
    mw_sync<NTH>();  // ys, ring are re-used by the next system
  }
}

// Systems in flight per SM (A/B knob RB_MW_OCC, default: as many as registers and shared memory allow).
template <int NT, int OCC>
static int mw_launch_occ(const LuArgs& args, double* scratch, int slots, cudaStream_t stream) {
  rb_lu_mw_kernel<NT, OCC><<<(unsigned)slots, MwCfg<NT>::NTH, 0, stream>>>(args, scratch);
  return (int)cudaGetLastError();
}

template <int NT>
struct MwOcc {
  // shared memory per CTA (+1 KB reserved by the driver) against 227 KB, capped by 2048 threads per SM
  static constexpr int BY_SMEM = (227 * 1024) / ((int)sizeof(MwSmem<NT>) + 1024);
  static constexpr int BY_THREADS = 2048 / MwCfg<NT>::NTH;
  static constexpr int MAXOCC = BY_SMEM < BY_THREADS ? BY_SMEM : BY_THREADS;
  static constexpr int OCC = MAXOCC > 8 ? 8 : MAXOCC;  // 8 x 128 threads x 64 registers = the whole register file
};

static int mw_sm_count() {
  static const int n = [] {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
  }();
  return n;
}

template <int NT>
static int mw_slots(long long count) {
  const long long full = (long long)mw_sm_count() * MwOcc<NT>::OCC;
  return (int)(count < full ? count : full);
}

template <int NT>
static int mw_launch(const LuArgs& args, double* scratch, cudaStream_t stream) {
  return mw_launch_occ<NT, MwOcc<NT>::OCC>(args, scratch, mw_slots<NT>(args.count), stream);
}

// doubles of scratch (records + packed U of every resident CTA) for `count` systems of size N, or -1
long long rb_lu_mw_scratch_doubles(int N, long long count) {
  switch ((N + 7) / 8) {
#define RB_MW_CASE(NT) \
  case NT: return MwCfg<NT>::SCRATCH * mw_slots<NT>(count);
    RB_MW_CASE(5) RB_MW_CASE(6) RB_MW_CASE(7) RB_MW_CASE(8) RB_MW_CASE(9) RB_MW_CASE(10) RB_MW_CASE(11) RB_MW_CASE(12)
    RB_MW_CASE(13) RB_MW_CASE(14) RB_MW_CASE(15) RB_MW_CASE(16)
#undef RB_MW_CASE
    default: return -1;
  }
}

int rb_lu_mw_launch(const LuArgs& args, double* scratch, cudaStream_t stream) {
  switch ((args.N + 7) / 8) {
#define RB_MW_CASE(NT) \
  case NT: return mw_launch<NT>(args, scratch, stream);
    RB_MW_CASE(5) RB_MW_CASE(6) RB_MW_CASE(7) RB_MW_CASE(8) RB_MW_CASE(9) RB_MW_CASE(10) RB_MW_CASE(11) RB_MW_CASE(12)
    RB_MW_CASE(13) RB_MW_CASE(14) RB_MW_CASE(15) RB_MW_CASE(16)
#undef RB_MW_CASE
    default: return -1000;
  }
}

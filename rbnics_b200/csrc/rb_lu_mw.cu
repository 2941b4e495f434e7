// rb_lu_mw.cu -- batched LU (partial pivoting) + solves for 32 < N <= 128: one CTA of NW warps per system,
// left-looking, persistent (every CTA walks over systems and re-uses ONE scratch slot for its L records and packed U).
//
// Replaces numpy.linalg.solve (LAPACK dgesv; rbnics/backends/online/numpy/linear_solver.py:29-31) and the lifting rows
// (rbnics/backends/online/basic/wrapping/DirichletBC.py:37-56) for a chunk of parameters.
//
// Algorithm (same as the one-warp kernel of round 1, rb_lu_left.cu, which stays available as an A/B variant):
// panel p (8 columns) is read once into DMMA accumulator tiles that stay in registers while the records of the
// finished panels q < p stream by through a 2-slot ring filled by 1-D bulk copies (TMA engine, one instruction per
// record); per record U_qp = L_qq^-1 X (2 DMMA) and C -= L_q U_qp (2 DMMA per live tile row); then the panel is
// transposed through shared memory and factored with implicit pivoting (two REDUX.MAX per warp on the IEEE bit
// pattern, LAPACK idamax tie rule), the rhs carried along.
//
// Why more than one warp per system: ncu of the one-warp kernel showed 49 k warp instructions per system issued by
// 2 warps per scheduler (226 registers per thread) with the issue slots 35 % busy -- the SM was starved of warps, not
// of bandwidth.  With the 8 NT rows of a system spread over NW warps (RPL = rows per lane in the pivot phase, tile rows
// tr = w, w + NW, .. per warp in the update phase) the register need per thread falls roughly with NW and the number
// of resident warps rises; the price is work that every warp repeats (pivot decision, record header).  NW = 4 / RPL = 1
// measured 120 k instructions per system (slower than one warp); NW = 2 / RPL = 2 is the compromise shipped.
// Further differences from the one-warp kernel: the U row, the rhs entry and the multipliers of a pivot row are
// published by its owner at pivot time, so used-up rows need no predicates afterwards; U_qp is computed by one warp
// per record (round robin) and handed over through shared memory; reciprocal pivots by MUFU.RCP64H + two Newton steps.
#include <stdlib.h>

#include "rb_lu_kernel.cuh"

template <int NT, int NW_>
struct MwCfg {
  static constexpr int NW = NW_;                   // warps per system
  static constexpr int NTH = 32 * NW;
  static constexpr int NP8 = 8 * NT;
  static constexpr int RPL = (NP8 + NTH - 1) / NTH;  // rows per lane in the pivot phase: row = tid + NTH * j
  static constexpr int TPW = (NT + NW - 1) / NW;   // tile rows per warp: tr = w + NW * i
  static constexpr int REC = 72 + 64 * NT;         // doubles per record: L_qq^-1 | pivrow[8], live mask | L tiles
  static constexpr int SLOT = REC > 8 * NP8 ? REC : 8 * NP8;  // ring slot: a record, a staged panel of A, a block column of U
  static constexpr int UBLK = NT * (NT + 1) / 2;   // 8x8 blocks of the packed upper-triangular U store
  static constexpr long long SCRATCH = (long long)NT * REC + (long long)UBLK * 64;  // doubles per scratch slot
  static_assert(NW * RPL <= 4 && NP8 <= 128, "rows are encoded in 8 bits of the pivot key; live mask has 4 nibbles");
};

template <int NT, int NW>
struct MwSmem {
  using C = MwCfg<NT, NW>;
  double ring[2][C::SLOT];
  double P[C::NP8 * 8];                 // the panel, row-major, 16-byte chunks XOR-swizzled with bits 1-2 of the row
                                        // (mw_pidx: conflict-free LDS.128 of one row per lane without padding):
                                        // gather of pivot rows, transpose into the pivot phase
  double Ub[8][8];                      // U_qp of the record being applied / the pivot rows of the panel being factored
  double slot[2][NW][10];               // pivot candidates of the warps: row entries t+1..7, 1/pivot, rhs entry
  unsigned long long keys[2][NW];
  double ys[C::NP8];                    // rhs in pivot order -> solution
  double rinvs[C::NP8];                 // reciprocal pivots
  uint64_t full[2];                     // mbarriers of the ring slots
  int pivrow[C::NP8];                   // pivot order -> row
  int recsz[NT];                        // bytes of record q
  long long nextsys;                    // the system this CTA works on after the current one (dynamic schedule)
  unsigned livew[2][4];                 // nibble n: tile rows 4n..4n+3 that still held a live row before this panel
};

// index of entry (row, col) of the swizzled panel buffer MwSmem::P
__device__ __forceinline__ int mw_pidx(int row, int col) {
  return row * 8 + ((((col >> 1) ^ (row >> 1)) & 3) << 1) + (col & 1);
}

template <int NTH>
__device__ __forceinline__ void mw_sync() {
  if (NTH == 32)
    __syncwarp();
  else
    __syncthreads();
}
// generic-proxy writes (st.global of records / U blocks) before async-proxy reads (bulk copies) of the same addresses
__device__ __forceinline__ void mw_fence_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// L2 eviction-priority hints: the L records / U blocks of a scratch slot are re-read many times and re-written by the
// next system of the same CTA (keep: evict_last); the assembled matrices are read exactly once (evict_first).
__device__ __forceinline__ uint64_t mw_policy_keep() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t mw_policy_stream() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void mw_bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void mw_st2_keep(double* p, double a, double b, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(policy) : "memory");
}
__device__ __forceinline__ void mw_st1_keep(double* p, double a, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(a), "l"(policy) : "memory");
}

// 1 / a to double precision: MUFU.RCP64H seed + two Newton steps (a == 0 gives NaN: singular matrices stay visible)
__device__ __forceinline__ double mw_rcp(double a) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  double e = fma(-a, x, 1.0);
  x = fma(x, e, x);
  e = fma(-a, x, 1.0);
  return fma(x, e, x);
}

// One pivot step T of the panel (template: register indices must be static).  Row tid + NTH j lives in pv[j][0..8).
template <int T, int NT, int NW>
__device__ __forceinline__ void mw_pivot_step(double (&pv)[MwCfg<NT, NW>::RPL][8], double (&b)[MwCfg<NT, NW>::RPL],
                                              unsigned& deadm, int (&tst)[MwCfg<NT, NW>::RPL], MwSmem<NT, NW>& sm,
                                              double* __restrict__ udiag, uint64_t keep, int c0, int tid, int lane, int w) {
  using C = MwCfg<NT, NW>;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int RPL = C::RPL;
  // the lane's own candidate among its RPL rows (strict compare: the lower row wins ties)
  unsigned long long key = 0ull;
  int jb = 0;
  double cand = 1.0;
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
    const unsigned long long kj =
        ((deadm >> j) & 1u) ? 0ull : ((unsigned long long)__double_as_longlong(fabs(pv[j][T])) + 1ull);
    if (kj > key) {
      key = kj;
      jb = j;
      cand = pv[j][T];
    }
  }
  // every lane inverts its own candidate while the two reductions are in flight (same issue slots as one lane doing it
  // afterwards, but off the dependent chain of the step)
  const double cand_rinv = mw_rcp(cand);
  // arg-max of the 64-bit key over the warp, lowest row on ties (idamax): REDUX.MAX on the high word, then on
  // {low word without its 8 lowest bits, 128 - row}; candidates that differ only below 2^-44 relative count as equal.
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mhi = __reduce_max_sync(FULL, hi);
  const int myrow = tid + C::NTH * jb;
  const unsigned k2 = (hi == mhi && key != 0ull) ? ((lo & ~255u) | (unsigned)(128 - myrow)) : 0u;
  const unsigned m2 = __reduce_max_sync(FULL, k2);
  if (m2 != 0u ? (k2 == m2) : (lane == 0)) {  // the warp's candidate (or "none"): one lane
    double* s = sm.slot[T & 1][w];
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
      if (j == jb) {
#pragma unroll
        for (int c = T + 1; c < 8; ++c) s[c - 1] = pv[j][c];
        s[8] = cand_rinv;
        s[9] = b[j];
      }
    }
    sm.keys[T & 1][w] = ((unsigned long long)mhi << 32) | m2;
  }
  mw_sync<C::NTH>();
  unsigned long long best = sm.keys[T & 1][0];
  int bw = 0;
#pragma unroll
  for (int w2 = 1; w2 < NW; ++w2) {
    const unsigned long long k = sm.keys[T & 1][w2];
    if (k > best) {
      best = k;
      bw = w2;
    }
  }
  const int prow = 128 - (int)((unsigned)best & 255u);
  const double* ws = sm.slot[T & 1][bw];
  const double rinv = ws[8];
#pragma unroll
  for (int j = 0; j < RPL; ++j) {
    if (tid + C::NTH * j == prow) {
      // this row becomes pivot c0 + T: publish everything that is final now
      deadm |= 1u << j;
      tst[j] = T;
      sm.pivrow[c0 + T] = prow;
      sm.rinvs[c0 + T] = rinv;
      sm.ys[c0 + T] = b[j];
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        *reinterpret_cast<double2*>(&sm.Ub[T][c]) = make_double2(pv[j][c], pv[j][c + 1]);  // multipliers s < T
        mw_st2_keep(udiag + T * 8 + c, pv[j][c], pv[j][c + 1], keep);  // U entries s > T
      }
    } else {
      // rows that are used up compute garbage in columns > T, which nobody reads any more
      const double l = pv[j][T] * rinv;
      pv[j][T] = l;
#pragma unroll
      for (int c = T + 1; c < 8; ++c) pv[j][c] = fma(-l, ws[c - 1], pv[j][c]);
      b[j] = fma(-l, ws[9], b[j]);
    }
  }
}

#ifdef RB_MW_PROFILE
// per-phase cycle counters of thread 0 of every CTA, summed into rb_mw_prof[0..16) (instrumented build only:
// tools/lu_phase_probe.py; the table is in DESIGN.md 4.2)
__device__ unsigned long long rb_mw_prof[16];
#define MW_ACC(slot, since) do { const long long n__ = clock64(); if (tid == 0) pacc[slot] += n__ - (since); (since) = n__; } while (0)
#else
#define MW_ACC(slot, since) do { } while (0)
#endif

template <int NT, int NW, int OCC>
__global__ void __launch_bounds__(32 * NW, OCC) rb_lu_mw_kernel(LuArgs args, double* __restrict__ scratch,
                                                                 unsigned long long* __restrict__ next_sys) {
  using C = MwCfg<NT, NW>;
  constexpr int NTH = C::NTH, TPW = C::TPW, RPL = C::RPL;
  __shared__ __align__(16) MwSmem<NT, NW> sm;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, q4 = lane & 3;
  const int N = args.N;
  const int NP = (N + 3) / 4 * 4;  // leading dimension of the assembled matrices (column-major)
  const int npan = (N + 7) / 8;
  double* Lrec = scratch + (long long)blockIdx.x * C::SCRATCH;  // [NT][REC]
  double* Ust = Lrec + (long long)NT * C::REC;                  // block (q, p), q <= p, at (p(p+1)/2 + q) * 64, [t][s]

  if (tid == 0) {
    mbar_init(&sm.full[0], 1);
    mbar_init(&sm.full[1], 1);
    mbar_fence_init();
  }
  mw_sync<NTH>();
#ifdef RB_MW_PROFILE
  long long pacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tp = clock64();
#endif
  unsigned phase = 0u;  // bit s: parity of the next completion of ring slot s (uniform)
  const uint64_t keep = mw_policy_keep(), stream = mw_policy_stream();
  auto fill = [&](int slot, const double* src, int bytes, uint64_t policy) {  // one thread, one instruction per transfer
    if (tid == 0) {
      mbar_arrive_expect_tx(&sm.full[slot], (uint32_t)bytes);
      mw_bulk_g2s_hint(sm.ring[slot], src, (uint32_t)bytes, &sm.full[slot], policy);
    }
  };
  auto wait_slot = [&](int slot) {
    mbar_wait(&sm.full[slot], (phase >> slot) & 1u);
    phase ^= 1u << slot;
  };

  // dynamic schedule: systems differ in their pivot patterns (live tile rows), so CTAs pull the next index from a global
  // counter; it is fetched one system ahead, so its latency is hidden
  for (long long sys = blockIdx.x; sys < args.count; sys = sm.nextsys) {
    mw_sync<NTH>();  // everybody has read nextsys (and is done with ys of the previous system)
    if (tid == 0) sm.nextsys = (long long)gridDim.x + (long long)atomicAdd(next_sys, 1ull);
    const double* A = args.Ab + sys * (long long)args.EP;
    // 8 columns of the assembled matrix are contiguous: -> ring slot 1
    fill(1, A, min(8, N) * NP * 8, stream);

    // ---- rhs entries of this lane's rows, lifting rows
    double b[RPL];
    unsigned deadm = 0u;    // bit j: row tid + NTH j is used up (or does not exist)
    unsigned bc_frag = 0u;  // bit i: row 8 (w + NW i) + g is a lifting row (fragment layout of the panel tiles)
    {
      const double* th = args.theta + sys * (long long)args.QT + args.Ql;
#pragma unroll
      for (int j = 0; j < RPL; ++j) {
        const int row = tid + NTH * j;
        double v = 0.0;
        if (row < N) {
          if (args.rhs_in)
            v = args.rhs_in[sys * (long long)N + row];
          else
            for (int q = 0; q < args.Qr; ++q) v = fma(__ldg(th + q), __ldg(args.F + q * NP + row), v);
        } else {
          deadm |= 1u << j;
        }
        b[j] = v;
      }
      for (int t = 0; t < args.n_bc; ++t) {
        const int r = __ldg(args.bc_rows + t);
#pragma unroll
        for (int j = 0; j < RPL; ++j)
          if (r == tid + NTH * j) b[j] = __ldg(args.thetaBC + sys * (long long)args.n_bc + t);
        if ((r & 7) == g && ((r >> 3) % NW) == w) bc_frag |= 1u << ((r >> 3) / NW);
      }
    }

    for (int p = 0; p < npan; ++p) {
      const int c0 = 8 * p;
      const int nst = min(8, N - c0);
      // tile rows that hold a live row before this panel: only those get multipliers in the record of this panel
#pragma unroll
      for (int j = 0; j < RPL; ++j) {
        const unsigned bal = __ballot_sync(0xffffffffu, !((deadm >> j) & 1u));
        if (lane == 0)
          sm.livew[p & 1][w + NW * j] = ((bal & 0xffu) ? 1u : 0u) | ((bal & 0xff00u) ? 2u : 0u) |
                                        ((bal & 0xff0000u) ? 4u : 0u) | ((bal & 0xff000000u) ? 8u : 0u);
      }
      // ---- the staged panel of the assembled matrix into accumulator tiles, the first record on its way
      MW_ACC(0, tp);  // rhs / live mask / loop head
      mw_sync<NTH>();  // the record of the previous panel is complete (and fenced)
      if (p > 0) fill(0, Lrec, sm.recsz[0], keep);
      MW_ACC(1, tp);  // barrier at the panel top
      wait_slot(1);
      MW_ACC(2, tp);  // wait for the staged panel of A
      double c[TPW][2];
      {
        const double* As = sm.ring[1];
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
          const int tr = w + NW * i;
          const int row = 8 * tr + g;
          double2 v = make_double2(0.0, 0.0);
          if (tr < NT && row < N) {
            if (2 * q4 < nst) v.x = As[(2 * q4) * NP + row];
            if (2 * q4 + 1 < nst) v.y = As[(2 * q4 + 1) * NP + row];
            if ((bc_frag >> i) & 1u) v = make_double2((c0 + 2 * q4 == row) ? 1.0 : 0.0, (c0 + 2 * q4 + 1 == row) ? 1.0 : 0.0);
          }
          c[i][0] = v.x;
          c[i][1] = v.y;
        }
      }
      MW_ACC(3, tp);  // panel tiles into registers
      // ---- apply the records q < p
      for (int q = 0; q < p; ++q) {
        // dump the tile rows that hold the pivot rows of panel q (they are final here); the mask travels in the record
        wait_slot(q & 1);
        MW_ACC(4, tp);  // wait for record q
        const double* rq = sm.ring[q & 1];
        const int* prq = reinterpret_cast<const int*>(rq + 64);
        {
          const unsigned dm = (unsigned)prq[9] >> w;
#pragma unroll
          for (int i = 0; i < TPW; ++i)
            if ((dm >> (NW * i)) & 1u)  // uniform
              *reinterpret_cast<double2*>(&sm.P[mw_pidx(8 * (w + NW * i) + g, 2 * q4)]) = make_double2(c[i][0], c[i][1]);
        }
        MW_ACC(5, tp);  // dump
        mw_sync<NTH>();  // (A) the dump is visible; everybody is done with record q - 1 (and, for q = 0, the panel of A)
        if (q + 1 < p) fill((q + 1) & 1, Lrec + (long long)(q + 1) * C::REC, sm.recsz[q + 1], keep);
        MW_ACC(6, tp);  // barrier A + issue of the next fill
        if (w == q % NW) {
          // U_qp = L_qq^-1 X on the tensor pipe (the inverse of the 8x8 unit-lower block travels in the record)
          const double2 la = *reinterpret_cast<const double2*>(rq + 2 * lane);
          const double xb0 = sm.P[mw_pidx(prq[q4], g)], xb1 = sm.P[mw_pidx(prq[4 + q4], g)];
          double u0 = 0.0, u1 = 0.0;
          dmma884(u0, u1, la.x, xb0);
          dmma884(u0, u1, la.y, xb1);
          *reinterpret_cast<double2*>(&sm.Ub[g][2 * q4]) = make_double2(u0, u1);
          // U block (q, p) -> packed store straight from the accumulator layout (read once, at the very end)
          mw_st2_keep(Ust + ((long long)(p * (p + 1) / 2 + q)) * 64 + 8 * g + 2 * q4, u0, u1, keep);
        }
        MW_ACC(7, tp);  // U_qp (computing warp) / nothing (the other)
        mw_sync<NTH>();  // (B) U_qp is visible; P may be overwritten
        MW_ACC(8, tp);  // barrier B
        const double b0 = sm.Ub[q4][g], b1 = sm.Ub[4 + q4][g];
        unsigned lv = (unsigned)prq[8];  // tile rows with non-zero multipliers (packed in this order)
        const double* af = rq + 72 + 2 * lane + 64 * __popc(lv & ((1u << w) - 1u));
        lv >>= w;
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
          if (lv & 1u) {  // uniform
            const double2 a = *reinterpret_cast<const double2*>(af);
            dmma884(c[i][0], c[i][1], a.x, b0);
            dmma884(c[i][0], c[i][1], a.y, b1);
          }
          af += 64 * __popc(lv & ((1u << NW) - 1u));
          lv >>= NW;
        }
        MW_ACC(9, tp);  // tile updates with record q
      }
      // ---- transpose into the pivot phase
#pragma unroll
      for (int i = 0; i < TPW; ++i) {
        const int tr = w + NW * i;
        if (tr < NT) *reinterpret_cast<double2*>(&sm.P[mw_pidx(8 * tr + g, 2 * q4)]) = make_double2(c[i][0], c[i][1]);
      }
      mw_sync<NTH>();  // everybody is done with the last record: both ring slots are free
      // the next panel of the assembled matrix streams into ring slot 1 while this one is factored
      if (p + 1 < npan) fill(1, A + (long long)(c0 + 8) * NP, min(8, N - c0 - 8) * NP * 8, stream);
      const unsigned dead_in = deadm;
      double pv[RPL][8];
      int tst[RPL];
#pragma unroll
      for (int j = 0; j < RPL; ++j) {
        tst[j] = -1;
#pragma unroll
        for (int cc = 0; cc < 8; cc += 2) {
          double2 v = make_double2(0.0, 0.0);
          if (!((deadm >> j) & 1u)) v = *reinterpret_cast<const double2*>(&sm.P[mw_pidx(tid + NTH * j, cc)]);
          pv[j][cc] = v.x;
          pv[j][cc + 1] = v.y;
        }
      }
      double* udiag = Ust + ((long long)(p * (p + 1) / 2 + p)) * 64;
      MW_ACC(10, tp);  // transpose into rows
      mw_pivot_step<0, NT, NW>(pv, b, deadm, tst, sm, udiag, keep, c0, tid, lane, w);
      if (nst > 1) mw_pivot_step<1, NT, NW>(pv, b, deadm, tst, sm, udiag, keep, c0, tid, lane, w);
      if (nst > 2) mw_pivot_step<2, NT, NW>(pv, b, deadm, tst, sm, udiag, keep, c0, tid, lane, w);
      if (nst > 3) mw_pivot_step<3, NT, NW>(pv, b, deadm, tst, sm, udiag, keep, c0, tid, lane, w);
      if (nst > 4) mw_pivot_step<4, NT, NW>(pv, b, deadm, tst, sm, udiag, keep, c0, tid, lane, w);
      if (nst > 5) mw_pivot_step<5, NT, NW>(pv, b, deadm, tst, sm, udiag, keep, c0, tid, lane, w);
      if (nst > 6) mw_pivot_step<6, NT, NW>(pv, b, deadm, tst, sm, udiag, keep, c0, tid, lane, w);
      if (nst > 7) mw_pivot_step<7, NT, NW>(pv, b, deadm, tst, sm, udiag, keep, c0, tid, lane, w);
      MW_ACC(11, tp);  // the 8 pivot steps
      mw_sync<NTH>();  // Ub (multipliers of the pivot rows), pivrow are complete

      // ---- the record of this panel (not needed after the last one)
      if (p + 1 < npan) {
        unsigned live_tr = 0u;
#pragma unroll
        for (int n = 0; n < NW * RPL; ++n) live_tr |= sm.livew[p & 1][n] << (4 * n);
        double* dst = Lrec + (long long)p * C::REC;
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
          const int row = tid + NTH * j;
          const int tr = row >> 3, gg = row & 7;
          if (row < C::NP8 && ((live_tr >> tr) & 1u)) {
            // negated multipliers of the row in DMMA A-fragment order, both k-steps of a lane adjacent (one LDS.128)
            const bool was_dead = (dead_in >> j) & 1u;
            double nl[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) nl[s] = (!was_dead && (tst[j] < 0 || s < tst[j])) ? -pv[j][s] : 0.0;
            double* d2 = dst + 72 + 64 * __popc(live_tr & ((1u << tr) - 1u)) + gg * 8;
            mw_st2_keep(d2 + 0, nl[0], nl[4], keep);
            mw_st2_keep(d2 + 2, nl[1], nl[5], keep);
            mw_st2_keep(d2 + 4, nl[2], nl[6], keep);
            mw_st2_keep(d2 + 6, nl[3], nl[7], keep);
          }
        }
        if (tid < 8) {  // column `tid` of L_qq^-1 (unit lower; Ub holds the multipliers), A-fragment order
          double x[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            double acc = (t == tid) ? 1.0 : 0.0;
#pragma unroll
            for (int s2 = 0; s2 < t; ++s2)
              if (s2 >= tid) acc = fma(-sm.Ub[t][s2], x[s2], acc);
            x[t] = (t >= tid) ? acc : 0.0;
          }
          // entry (t, tid): lane (g = t, q4 = tid & 3), k-step tid >> 2
#pragma unroll
          for (int t = 0; t < 8; ++t) mw_st1_keep(dst + (t * 4 + (tid & 3)) * 2 + (tid >> 2), x[t], keep);
          reinterpret_cast<int*>(dst + 64)[tid] = sm.pivrow[c0 + tid];
        }
        if (tid == 9 % NTH) {  // tile rows that hold the pivot rows of this panel: the only ones later panels must dump
          unsigned dm = 0u;
#pragma unroll
          for (int t = 0; t < 8; ++t)
            if (t < nst) dm |= 1u << (sm.pivrow[c0 + t] >> 3);
          reinterpret_cast<int*>(dst + 64)[9] = (int)dm;
        }
        if (tid == 8 % NTH) {
          reinterpret_cast<int*>(dst + 64)[8] = (int)live_tr;
          sm.recsz[p] = (72 + 64 * __popc(live_tr)) * 8;
        }
      }
      MW_ACC(12, tp);  // publish (record, L_qq^-1)
      mw_fence_async();  // this thread's record / U stores before the bulk copies issued after the next barrier
      MW_ACC(13, tp);  // proxy fence
    }

    // ---- back-substitution U x = y by panels in reverse; ys is overwritten with the solution.
    // Block column p of the packed U store is contiguous ((p + 1) blocks: rows of the pivots k < 8p, then the diagonal
    // block) and is staged one panel ahead in the (now idle) ring.
    mw_sync<NTH>();
    fill((npan - 1) & 1, Ust + ((long long)((npan - 1) * npan / 2)) * 64, npan * 512, keep);
    for (int p = npan - 1; p >= 0; --p) {
      const int c0 = 8 * p;
      const int nst = min(8, N - c0);
      if (p > 0) fill((p - 1) & 1, Ust + ((long long)((p - 1) * p / 2)) * 64, p * 512, keep);
      wait_slot(p & 1);
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
#pragma unroll
      for (int j = 0; j < RPL; ++j) {
        const int k = tid + NTH * j;
        if (k < c0) {
          double acc = sm.ys[k];
#pragma unroll
          for (int s = 0; s < 8; s += 2) {
            const double2 v = *reinterpret_cast<const double2*>(colp + k * 8 + s);
            const double2 x = *reinterpret_cast<const double2*>(&sm.ys[c0 + s]);
            acc = fma(-v.x, (s < nst) ? x.x : 0.0, acc);
            acc = fma(-v.y, (s + 1 < nst) ? x.y : 0.0, acc);
          }
          sm.ys[k] = acc;
        }
      }
      mw_sync<NTH>();  // ys is updated; everybody is done with this ring slot
    }
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
      const int k = tid + NTH * j;
      if (k < N) args.U[sys * (long long)args.ldu + k] = sm.ys[k];
    }
    MW_ACC(14, tp);  // back-substitution + store
  }
#ifdef RB_MW_PROFILE
  if (tid == 0)
    for (int k = 0; k < 16; ++k) atomicAdd(&rb_mw_prof[k], (unsigned long long)pacc[k]);
#endif
}

#ifdef RB_MW_PROFILE
extern "C" int rb_mw_profile_read(unsigned long long* out16, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out16, rb_mw_prof, sizeof(unsigned long long) * 16);
  if (e == cudaSuccess && reset) {
    unsigned long long z[16] = {0};
    e = cudaMemcpyToSymbol(rb_mw_prof, z, sizeof(z));
  }
  return (int)e;
}
#endif

// Warps per system and systems in flight per SM (env RB_MW_NW = 1 | 2 | 4 selects the A/B variants in builds with
// -DRB_AB_VARIANTS; the default is 2 warps).
template <int NT, int NW>
struct MwOcc {
  // shared memory per CTA (+1 KB reserved by the driver) against 227 KB, capped by the 64 K registers of the SM
  static constexpr int BY_SMEM = (227 * 1024) / ((int)sizeof(MwSmem<NT, NW>) + 1024);
  static constexpr int REGS = NW == 4 ? 64 : (NW == 2 ? 112 : 224);
  static constexpr int BY_REGS = 65536 / (32 * NW * REGS);
  static constexpr int OCC = BY_SMEM < BY_REGS ? BY_SMEM : BY_REGS;
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

#ifdef RB_AB_VARIANTS
static int mw_warps() {
  static const int nw = [] {
    const char* e = getenv("RB_MW_NW");
    const int v = e ? atoi(e) : 2;
    return (v == 1 || v == 4) ? v : 2;
  }();
  return nw;
}
#endif

template <int NT, int NW>
static int mw_slots(long long count) {
  int occ = MwOcc<NT, NW>::OCC;
#ifdef RB_AB_VARIANTS
  static const int lim = getenv("RB_MW_OCC") ? atoi(getenv("RB_MW_OCC")) : 0;  // fewer systems in flight per SM (A/B)
  if (lim > 0 && lim < occ) occ = lim;
#endif
  const long long full = (long long)mw_sm_count() * occ;
  return (int)(count < full ? count : full);
}

template <int NT, int NW>
static int mw_launch(const LuArgs& args, double* scratch, cudaStream_t stream) {
  const int slots = mw_slots<NT, NW>(args.count);
  // the work counter lives behind the scratch slots
  unsigned long long* counter = reinterpret_cast<unsigned long long*>(scratch + MwCfg<NT, NW>::SCRATCH * slots);
  cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return (int)e;
  rb_lu_mw_kernel<NT, NW, MwOcc<NT, NW>::OCC><<<(unsigned)slots, 32 * NW, 0, stream>>>(args, scratch, counter);
  return (int)cudaGetLastError();
}

template <int NT>
static long long mw_scratch_nt(long long count) {
#ifdef RB_AB_VARIANTS
  if (mw_warps() == 4) return MwCfg<NT, 4>::SCRATCH * mw_slots<NT, 4>(count) + 2;
  if (mw_warps() == 1) return MwCfg<NT, 1>::SCRATCH * mw_slots<NT, 1>(count) + 2;
#endif
  return MwCfg<NT, 2>::SCRATCH * mw_slots<NT, 2>(count) + 2;  // + the work counter
}

template <int NT>
static int mw_launch_nt(const LuArgs& args, double* scratch, cudaStream_t stream) {
#ifdef RB_AB_VARIANTS
  if (mw_warps() == 4) return mw_launch<NT, 4>(args, scratch, stream);
  if (mw_warps() == 1) return mw_launch<NT, 1>(args, scratch, stream);
#endif
  return mw_launch<NT, 2>(args, scratch, stream);
}

// number of systems the persistent grid works on at a time (chunks that are a multiple of it leave no partial wave)
int rb_lu_mw_quantum(int N) {
  switch ((N + 7) / 8) {
#define RB_MW_CASE(NT) \
  case NT: return (int)(mw_scratch_nt<NT>(1LL << 40) / MwCfg<NT, 2>::SCRATCH);
    RB_MW_CASE(5) RB_MW_CASE(6) RB_MW_CASE(7) RB_MW_CASE(8) RB_MW_CASE(9) RB_MW_CASE(10) RB_MW_CASE(11) RB_MW_CASE(12)
    RB_MW_CASE(13) RB_MW_CASE(14) RB_MW_CASE(15) RB_MW_CASE(16)
#undef RB_MW_CASE
    default: return 0;
  }
}

// doubles of scratch (records + packed U of every resident CTA) for `count` systems of size N, or -1
long long rb_lu_mw_scratch_doubles(int N, long long count) {
  switch ((N + 7) / 8) {
#define RB_MW_CASE(NT) \
  case NT: return mw_scratch_nt<NT>(count);
    RB_MW_CASE(5) RB_MW_CASE(6) RB_MW_CASE(7) RB_MW_CASE(8) RB_MW_CASE(9) RB_MW_CASE(10) RB_MW_CASE(11) RB_MW_CASE(12)
    RB_MW_CASE(13) RB_MW_CASE(14) RB_MW_CASE(15) RB_MW_CASE(16)
#undef RB_MW_CASE
    default: return -1;
  }
}

int rb_lu_mw_launch(const LuArgs& args, double* scratch, cudaStream_t stream) {
  switch ((args.N + 7) / 8) {
#define RB_MW_CASE(NT) \
  case NT: return mw_launch_nt<NT>(args, scratch, stream);
    RB_MW_CASE(5) RB_MW_CASE(6) RB_MW_CASE(7) RB_MW_CASE(8) RB_MW_CASE(9) RB_MW_CASE(10) RB_MW_CASE(11) RB_MW_CASE(12)
    RB_MW_CASE(13) RB_MW_CASE(14) RB_MW_CASE(15) RB_MW_CASE(16)
#undef RB_MW_CASE
    default: return -1000;
  }
}

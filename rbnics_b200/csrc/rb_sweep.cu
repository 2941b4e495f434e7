// rb_sweep.cu -- the greedy sweep on one B200: affine assembly GEMM, estimator GEMM (z^T G z), finalize + arg-max,
// operator packing and the C ABI around them.  Reference call stack being replaced: SURVEY.md 3.1 / 8a rows A1-A11.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "rb_common.cuh"
#include "rb_parabolic.cuh"

// =====================================================================================================================
// Packing kernels (run once per greedy iteration, when the reduced operators change)
// =====================================================================================================================

// A (Ql x N x N row-major)  ->  Aq [QlP][EP] with entry (i,j) at j*NP+i, zero padded.
__global__ void rb_pack_system_kernel(const double* __restrict__ A, double* __restrict__ Aq, int N, int NP, int Ql,
                                      int QlP, int EP) {
  const long long total = (long long)QlP * EP;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(t / EP), e = (int)(t % EP);
    const int j = e / NP, i = e % NP;
    double v = 0.0;
    if (q < Ql && e < NP * NP && i < N && j < N) v = A[((long long)q * N + i) * N + j];
    Aq[t] = v;
  }
}

// Dense G (Kz x Kz) -> packed G'.  z^T G z only sees the symmetric part of G, and blocks of z that scale the SAME
// slice of V (z_b = theta_b v, z_b' = theta_b' v: the Q_a blocks theta^a_q u of the elliptic estimator) make the
// coefficient of G[(b,n),(b',m)] symmetric in (n, m) as well.  Both symmetries are folded into the lower triangles:
//   b == b'              : n > m : G[k,c] + G[c,k]          n == m : G[k,c]
//   b >  b', same slice  : n > m : S(n,m) + S(m,n)          n == m : S(n,n)      S(n,m) = G[(b,n),(b',m)] + G[(b',m),(b,n)]
//   b >  b', other slice : G[k,c] + G[c,k]
//   everything else 0.
// A column tile (64 columns = two 32-column halves) takes the same 32-column range [32r, 32r+32) of TWO blocks of one
// slice class, so every row block of that class contributes only its rows n >= 32r: the k range of the tile skips 8r
// k-steps per block.  The stream of a tile is the concatenation of its k-step records; each record is stored in DMMA
// B-fragment order  [wn (2)][pair (2)][lane (32)][e (2)]  <-  G'[row 4*ks + lane%4][column wn*32 + (2*pair+e)*8 + lane/4].
struct PackArgs {
  const double* G;
  double* Gp;
  long long total_rec;
  const int* rec_tile;    // [total_rec] column tile of the record
  const int* rec_ks;      // [total_rec] k-step (padded row / 4) of the record
  const int* row_unpad;   // [Kp] padded row -> dense index or -1
  const int* row_blk;     // [Kp]
  const int* row_idx;     // [Kp] index within the block
  const int* col_unpad;   // [n_ctile * 64]
  const int* col_blk;
  const int* col_idx;
  const int* blk_cls;     // [nblk] slice class
  const int* blk_ubegin;  // [nblk] dense index of the block's first entry
  int Kz;
};

__global__ void rb_pack_estimator_kernel(PackArgs a) {
  const long long total = a.total_rec * EST_KSTEP_DOUBLES;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long rec = t / EST_KSTEP_DOUBLES;
    const int r = (int)(t % EST_KSTEP_DOUBLES);
    const int ct = a.rec_tile[rec], ks = a.rec_ks[rec];
    const int wn = r >> 7, pair = (r >> 6) & 1, lane = (r >> 1) & 31, e = r & 1;
    const int kp = 4 * ks + (lane & 3);
    const int cpos = ct * EST_BN + wn * 32 + (2 * pair + e) * 8 + (lane >> 2);
    const long long k = a.row_unpad[kp], c = a.col_unpad[cpos];
    double v = 0.0;
    if (k >= 0 && c >= 0) {
      const int b = a.row_blk[kp], n = a.row_idx[kp], bp = a.col_blk[cpos], m = a.col_idx[cpos];
      const long long Kz = a.Kz;
      if (b == bp) {
        if (n > m)
          v = a.G[k * Kz + c] + a.G[c * Kz + k];
        else if (n == m)
          v = a.G[k * Kz + c];
      } else if (b > bp) {
        if (a.blk_cls[b] == a.blk_cls[bp]) {
          if (n >= m) {
            v = a.G[k * Kz + c] + a.G[c * Kz + k];
            if (n > m) {
              const long long k2 = a.blk_ubegin[b] + m, c2 = a.blk_ubegin[bp] + n;
              v += a.G[k2 * Kz + c2] + a.G[c2 * Kz + k2];
            }
          }
        } else {
          v = a.G[k * Kz + c] + a.G[c * Kz + k];
        }
      }
    }
    a.Gp[t] = v;
  }
}

// =====================================================================================================================
// K_asm: Ab[p][e] = sum_q Theta[p][q] * Aq[q][e]   (DMMA GEMM, M = parameters, N = matrix entries, K = Ql)
// product order-1 of rbnics/backends/online/basic/product.py:22-33 for a whole chunk of parameters.
// =====================================================================================================================
struct AsmArgs {
  const double* theta;  // [count][QT] (lhs thetas are the first Ql columns)
  const double* Aq;     // [QlP][EP]
  double* Ab;           // [count][EP]
  long long count;
  int QT, Ql, QlP, EP;
};

__device__ __forceinline__ void asm_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// One CTA = one block of ASM_BM parameters x a strided set of 64-entry tiles: the theta tile is loaded once, the A_q
// tiles stream through a 2-stage cp.async ring, so global-load latency is off the critical path and the kernel runs
// at the rate of its 80 KB-per-parameter output.
__global__ void __launch_bounds__(256, 2) rb_assemble_kernel(AsmArgs a) {
  extern __shared__ __align__(16) double asm_smem[];
  const int TS = a.QlP + ((a.QlP % 8 == 4) ? 0 : 4);  // theta tile pitch == 4 (mod 8): conflict-free A fragments
  constexpr int BS = ASM_BE + 4;                       // Aq tile pitch == 4 (mod 16): conflict-free B fragments
  double* ths = asm_smem;                              // [ASM_BM][TS]
  double* bsm0 = asm_smem + ASM_BM * TS;               // [2][QlP][BS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const long long p0 = (long long)blockIdx.y * ASM_BM;
  const int ntile = a.EP / ASM_BE;

  auto issue_tile = [&](int t, int slot) {
    double* bsm = bsm0 + (size_t)slot * a.QlP * BS;
    const double* src = a.Aq + (long long)t * ASM_BE;
    for (int i = tid; i < a.QlP * (ASM_BE / 2); i += 256) {
      const int q = i / (ASM_BE / 2), c2 = i % (ASM_BE / 2);
      asm_cp_async16(bsm + q * BS + 2 * c2, src + (long long)q * a.EP + 2 * c2);
    }
  };
  int t = blockIdx.x;
  if (t < ntile) issue_tile(t, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = tid; i < ASM_BM * a.QlP; i += 256) {
    const int r = i / a.QlP, q = i % a.QlP;
    const long long p = p0 + r;
    ths[r * TS + q] = (p < a.count && q < a.Ql) ? __ldg(a.theta + p * a.QT + q) : 0.0;
  }
  const int ar = wm * 32 + (lane >> 2), ak = lane & 3;
  const int bc = wn * 32 + (lane >> 2);
  int slot = 0;
  for (; t < ntile; t += gridDim.x, slot ^= 1) {
    if (t + (int)gridDim.x < ntile) issue_tile(t + gridDim.x, slot ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const double* bsm = bsm0 + (size_t)slot * a.QlP * BS;
    double acc[4][4][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    for (int ks = 0; ks < a.QlP / 4; ++ks) {
      double af[4], bf[4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) af[mt] = ths[(ar + mt * 8) * TS + 4 * ks + ak];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bf[nt] = bsm[(4 * ks + ak) * BS + bc + nt * 8];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
    }
    const int e0 = t * ASM_BE;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const long long p = p0 + wm * 32 + mt * 8 + (lane >> 2);
      if (p < a.count) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int e = e0 + wn * 32 + nt * 8 + (lane & 3) * 2;
          __stcs(reinterpret_cast<double2*>(a.Ab + p * a.EP + e), make_double2(acc[mt][nt][0], acc[mt][nt][1]));
        }
      }
    }
    __syncthreads();  // every warp is done with this slot before it is refilled
  }
}

// =====================================================================================================================
// K_est: eps2 partial[split][p] = sum over the split's column tiles of (Z G')[p, c] * z[p, c]
// order-2 products of rbnics/backends/online/basic/product.py:34-46 + the quadratic forms of
// rbnics/problems/elliptic/elliptic_rb_reduced_problem.py:55-64, for EST_BM parameters per CTA.
// =====================================================================================================================
struct EstArgs {
  const double* Gp;
  const long long* tile_off;
  const int* tile_skip;    // k-steps skipped at the start of every row block of the tile's slice class
  const int* tile_cls;
  const int* tile_b0;      // first row block of the tile
  const int* tile_cb;      // [4 n_ctile] per 16-column range of the tile: first row block that has data for it
  const int* blk_ks_begin;
  const int* blk_scale;
  const int* blk_src;
  const int* blk_cls;
  const int* col_scale;
  const int* col_src;
  const int* split_begin;  // [nsplit+1] column-tile boundaries
  const double* U;         // [B][ldu]
  const double* theta;     // [B][QT]
  double* partial;         // [nsplit][B]
  long long B;
  int nblk, KS, N, ldu, QT, v_off, v_len, LVS, nsplit, n_ctile, stage_ksteps;
  int nstage;  // stages of the G' ring; a stage never crosses a z-block boundary
  int tri16;   // 1: tiles are made of 16-column ranges (enables the half-width head of folded blocks)
};

// shared-memory footprint of rb_estimator_kernel (host + device agree through this one function)
__host__ __device__ inline size_t est_smem_bytes(int nstage, int stage_ksteps, int LVS, int nblk, int n_ctile) {
  size_t b = (size_t)nstage * stage_ksteps * EST_KSTEP_DOUBLES * 8;       // G' stages
  b += (size_t)EST_NSTAGE * 8;                                            // full barriers (EST_NSTAGE = maximum)
  b += (size_t)3 * EST_BM * 8;                                            // cross-warp reduction
  b += (size_t)EST_BM * LVS * 8;                                          // V tile
  b += (size_t)n_ctile * 8;                                               // tile offsets
  b += (size_t)(7 * n_ctile + 4 * nblk + 1 + EST_NSTAGE) * 4;             // int tables + release counters
  return b + 16;
}

// Fragments of one k-step: B from the staged G' (fragment order, conflict-free LDS.128), A from the V tile.
template <int NTN>
struct EstFrag {
  double2 b[NTN / 2];
  double a[4];
};

template <int NTN>
__device__ __forceinline__ void est_load_frag(EstFrag<NTN>& f, const double* gs, const double* vp, int LVS) {
#pragma unroll
  for (int h = 0; h < NTN / 2; ++h) f.b[h] = *reinterpret_cast<const double2*>(gs + 64 * h);
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) f.a[mt] = vp[mt * 8 * LVS];
}

// HEAD: only the first 8-column tile of every 16-column range (the second one is structurally zero in the first two
// k-steps of a folded block: columns m >= 16r+8 against rows n < 16r+8).
// FIRST_ONLY: only the warp's first 16-column range (the block of its second range lies below this row block).
// INIT: 1 = every accumulator this call touches starts from zero (D = A*B), 2 = only the odd 8-column tiles do
// (the even ones were started by the HEAD steps), 0 = accumulate.
template <int NTN, bool HEAD = false, bool FIRST_ONLY = false, int INIT = 0>
__device__ __forceinline__ void est_mma(double (&T)[4][NTN][2], const EstFrag<NTN>& f) {
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
    for (int h = 0; h < (FIRST_ONLY ? 1 : NTN / 2); ++h) {
      if (INIT == 1)
        dmma884_init(T[mt][2 * h][0], T[mt][2 * h][1], f.a[mt], f.b[h].x);
      else
        dmma884(T[mt][2 * h][0], T[mt][2 * h][1], f.a[mt], f.b[h].x);
      if (!HEAD) {
        if (INIT != 0)
          dmma884_init(T[mt][2 * h + 1][0], T[mt][2 * h + 1][1], f.a[mt], f.b[h].y);
        else
          dmma884(T[mt][2 * h + 1][0], T[mt][2 * h + 1][1], f.a[mt], f.b[h].y);
      }
    }
  }
}

// NTN = 8-column tiles per warp: 4 -> 8 warps (4 m x 2 n, warp tile 32 x 32), 2 -> 16 warps (4 m x 4 n, 32 x 16).
template <int NTN>
__global__ void __launch_bounds__(32 * 4 * (8 / NTN), 1) rb_estimator_kernel(EstArgs a) {
  constexpr int WN = 8 / NTN;              // warps along n
  constexpr int NWARPS = 4 * WN;
  constexpr int NTHREADS = 32 * NWARPS;
  extern __shared__ __align__(128) unsigned char est_smem_raw[];
  const int SK = a.stage_ksteps;
  const int STAGE_DOUBLES = SK * EST_KSTEP_DOUBLES;
  double* Gs = reinterpret_cast<double*>(est_smem_raw);                     // [NSTAGE][STAGE_DOUBLES]
  const int NSTAGE = a.nstage;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(Gs + NSTAGE * STAGE_DOUBLES);
  double* red = reinterpret_cast<double*>(full_bar + EST_NSTAGE);          // [3][EST_BM]
  double* Vs = red + 3 * EST_BM;                                            // [EST_BM][LVS]
  long long* s_tile_off = reinterpret_cast<long long*>(Vs + EST_BM * a.LVS);  // [n_ctile]
  int* s_tile_skip = reinterpret_cast<int*>(s_tile_off + a.n_ctile);        // [n_ctile]
  int* s_tile_cls = s_tile_skip + a.n_ctile;                                // [n_ctile]
  int* s_tile_b0 = s_tile_cls + a.n_ctile;                                  // [n_ctile]
  int* s_ks_begin = s_tile_b0 + a.n_ctile;                                  // [nblk+1]
  int* s_scale = s_ks_begin + (a.nblk + 1);                                 // [nblk]
  int* s_src = s_scale + a.nblk;                                            // [nblk]
  int* s_cls = s_src + a.nblk;                                              // [nblk]
  int* s_tile_cb = s_cls + a.nblk;                                          // [4 n_ctile] first row block with data per 16-column range
  int* s_cnt = s_tile_cb + 4 * a.n_ctile;                                   // [NSTAGE] release counters

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int split = blockIdx.x % a.nsplit;
  const long long rb = blockIdx.x / a.nsplit;
  const long long p0 = rb * EST_BM;
  const int ct_begin = a.split_begin[split], ct_end = a.split_begin[split + 1];
  const int LVS = a.LVS;

  // V tile: [u | theta[v_off : v_off+v_len] | 1 | 0 ...]
  for (int t = tid; t < EST_BM * LVS; t += NTHREADS) {
    const int r = t / LVS, c = t % LVS;
    const long long p = p0 + r;
    double v = 0.0;
    if (p < a.B) {
      if (c < a.N)
        v = a.U[p * a.ldu + c];
      else if (c < a.N + a.v_len)
        v = a.theta[p * a.QT + a.v_off + (c - a.N)];
      else if (c == a.N + a.v_len)
        v = 1.0;
    }
    Vs[t] = v;
  }
  for (int t = tid; t < a.n_ctile; t += NTHREADS) {
    s_tile_off[t] = a.tile_off[t];
    s_tile_skip[t] = a.tile_skip[t];
    s_tile_cls[t] = a.tile_cls[t];
    s_tile_b0[t] = a.tile_b0[t];
  }
  for (int t = tid; t < 4 * a.n_ctile; t += NTHREADS) s_tile_cb[t] = a.tile_cb[t];
  for (int t = tid; t <= a.nblk; t += NTHREADS) {
    s_ks_begin[t] = a.blk_ks_begin[t];
    if (t < a.nblk) {
      s_scale[t] = a.blk_scale[t];
      s_src[t] = a.blk_src[t];
      s_cls[t] = a.blk_cls[t];
    }
  }
  if (tid == 0) {
    for (int s = 0; s < EST_NSTAGE; ++s) {
      mbar_init(full_bar + s, 1);
      s_cnt[s] = 0;
    }
    mbar_fence_init();
  }
  __syncthreads();

  // ---- G' stream: 1-D bulk copies (TMA engine).  Cursor = next stage to issue; every warp advances it identically,
  // the warp that is LAST to release a stage slot issues the refill of that slot (no dedicated producer warp).
  // cursor: tile pct, row block pb, k-step pks (absolute), element offset poff inside the tile's stream
  auto blk_first_ks = [&](int ct, int b) { return s_ks_begin[b] + ((s_cls[b] == s_tile_cls[ct]) ? s_tile_skip[ct] : 0); };
  int pct = ct_begin, pb = 0, pks = 0;
  long long poff = 0;
  if (pct < ct_end) {
    pb = s_tile_b0[pct];
    pks = blk_first_ks(pct, pb);
  }
  auto stage_len = [&]() { return min(SK, s_ks_begin[pb + 1] - pks); };  // k-steps of the stage at the cursor
  auto issue_stage = [&](int slot) {  // called by exactly one thread per stage
    const uint32_t bytes = (uint32_t)stage_len() * EST_KSTEP_DOUBLES * 8u;
    mbar_arrive_expect_tx(full_bar + slot, bytes);
    bulk_g2s(Gs + slot * STAGE_DOUBLES, a.Gp + s_tile_off[pct] + poff, bytes, full_bar + slot);
  };
  auto advance_cursor = [&]() {
    const int n = stage_len();
    pks += n;
    poff += (long long)n * EST_KSTEP_DOUBLES;
    if (pks == s_ks_begin[pb + 1]) {
      ++pb;
      if (pb == a.nblk) {
        ++pct;
        poff = 0;
        if (pct < ct_end) pb = s_tile_b0[pct];
      }
      if (pct < ct_end) pks = blk_first_ks(pct, pb);
    }
  };
  for (int s = 0; s < NSTAGE; ++s) {
    if (pct < ct_end) {
      if (tid == 0) issue_stage(s);
      advance_cursor();
    }
  }

  const int wm = warp & 3, wn = warp >> 2;  // wn < WN
  const int g = lane >> 2, q4 = lane & 3;
  const int r0 = wm * 32 + g;  // + mt*8
  const double* vrow = Vs + r0 * LVS + q4;
  const double* gsw = Gs + wn * (32 * NTN) + lane * 2;
  const double* thp[4];  // theta rows of this thread's four parameters
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) thp[mt] = a.theta + min(p0 + r0 + mt * 8, a.B - 1) * a.QT;
  double rowacc[4] = {0.0, 0.0, 0.0, 0.0};
  int stage = 0;
  uint32_t phase = 0;

  double T[4][NTN][2];  // accumulator of the current block
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < NTN; ++nt) T[mt][nt][0] = T[mt][nt][1] = 0.0;

  for (int ct = ct_begin; ct < ct_end; ++ct) {
    double Y[4][NTN][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTN; ++nt) Y[mt][nt][0] = Y[mt][nt][1] = 0.0;

    const int tcls = s_tile_cls[ct], tskip = s_tile_skip[ct];
    const int cb0 = s_tile_cb[4 * ct + 2 * wn], cb1 = s_tile_cb[4 * ct + 2 * wn + 1];
    auto first_ks = [&](int bb) { return s_ks_begin[bb] + ((s_cls[bb] == tcls) ? tskip : 0); };
    int b = s_tile_b0[ct];
    int ks = first_ks(b);
    int kend = s_ks_begin[b + 1];
    const int src = s_src[b] + 4 * (ks - s_ks_begin[b]);
    int kin = 0;                   // k-step within the current stage
    bool tri_head = a.tri16 && (s_cls[b] == tcls);  // the next segment starts a folded block
    int nvalid = (b >= cb0) + (b >= cb1);           // ranges with data
    // fresh: T holds the (already folded) sums of the previous block; the first k-step of the block overwrites it
    // with D = A*B instead of a clearing pass.  Partially empty ranges (nvalid < 2) clear T explicitly instead.
    int fresh = 1;
    if (nvalid < 2) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTN; ++nt) T[mt][nt][0] = T[mt][nt][1] = 0.0;
      fresh = 0;
    }
    int nin = min(SK, kend - ks);  // k-steps held by the current stage (a stage never crosses a block boundary)
    double th[4], thn[4];
    {
      const int sc = s_scale[b];
      const int scn = (b + 1 < a.nblk) ? s_scale[b + 1] : -1;
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        th[mt] = (sc < 0) ? 1.0 : __ldg(thp[mt] + sc);
        thn[mt] = (scn < 0) ? 1.0 : __ldg(thp[mt] + scn);
      }
    }
    mbar_wait(full_bar + stage, phase);
    const double* gp = gsw + stage * STAGE_DOUBLES;   // this warp's B fragments of the current k-step
    const double* vp = vrow + src;                    // this thread's A fragments of the current k-step
    EstFrag<NTN> f0, f1;
    est_load_frag(f0, gp, vp, LVS);

    // The k loop is cut into segments that end at the next stage or block boundary, so that the steady-state body is
    // 16 DMMAs + the fragment loads of the next k-step and nothing else (two warps per SM sub-partition then run in
    // anti-phase: one issues DMMAs while the other does its loads / boundary bookkeeping).
    while (true) {
      const int nseg = min(kend - ks, nin - kin);
      int i = nseg;
      if (nvalid < 2) {
        // one of the first row blocks of a tile group: the column blocks of one or both of this warp's 16-column
        // ranges lie below it (structural zeros) -> half or none of the DMMAs; fragments are still stepped through
        for (; i > 1; --i) {
          if (nvalid == 1) est_mma<NTN, false, true>(T, f0);
          gp += EST_KSTEP_DOUBLES;
          vp += 4;
          est_load_frag(f0, gp, vp, LVS);
        }
        if (nvalid == 1) est_mma<NTN, false, true>(T, f0);
      } else {
      int init_last = fresh;    // how the last k-step of the segment starts T if no earlier k-step did
      if (tri_head && i > 2) {  // first two k-steps of a folded block (fresh): half of the column tiles are zero
        est_mma<NTN, true, false, 1>(T, f0);
        gp += EST_KSTEP_DOUBLES;
        vp += 4;
        est_load_frag(f0, gp, vp, LVS);
        est_mma<NTN, true>(T, f0);
        gp += EST_KSTEP_DOUBLES;
        vp += 4;
        est_load_frag(f0, gp, vp, LVS);
        i -= 2;
        init_last = 2;
        if (i > 1) {
          est_mma<NTN, false, false, 2>(T, f0);
          gp += EST_KSTEP_DOUBLES;
          vp += 4;
          est_load_frag(f0, gp, vp, LVS);
          --i;
          init_last = 0;
        }
      } else if (fresh && i > 1) {
        est_mma<NTN, false, false, 1>(T, f0);
        gp += EST_KSTEP_DOUBLES;
        vp += 4;
        est_load_frag(f0, gp, vp, LVS);
        --i;
        init_last = 0;
      }
      while (i > 2) {  // steady state, two k-steps per trip (ping-pong fragment buffers)
        est_mma(T, f0);
        est_load_frag(f1, gp + EST_KSTEP_DOUBLES, vp + 4, LVS);
        est_mma(T, f1);
        gp += 2 * EST_KSTEP_DOUBLES;
        vp += 8;
        est_load_frag(f0, gp, vp, LVS);
        i -= 2;
      }
      if (i == 2) {
        est_mma(T, f0);
        gp += EST_KSTEP_DOUBLES;
        vp += 4;
        est_load_frag(f0, gp, vp, LVS);
      }
      // last k-step of the segment; its successor is loaded after the boundary bookkeeping
      if (init_last == 0)
        est_mma(T, f0);
      else if (init_last == 1)
        est_mma<NTN, false, false, 1>(T, f0);
      else
        est_mma<NTN, false, false, 2>(T, f0);
      fresh = 0;
      }
      ks += nseg;
      kin += nseg;
      gp += EST_KSTEP_DOUBLES;
      vp += 4;
      const bool blk_done = (ks == kend);
      const bool tile_done = blk_done && (b + 1 == a.nblk);
      tri_head = false;
      if (blk_done && !tile_done) {
        ++b;
        nvalid = (b >= cb0) + (b >= cb1);
        tri_head = a.tri16 && (s_cls[b] == tcls);
        ks = first_ks(b);
        kend = s_ks_begin[b + 1];
        vp = vrow + s_src[b] + 4 * (ks - s_ks_begin[b]);
      }
      if (kin == nin) {  // stage consumed: release it; the last warp to do so refills the slot
        __syncwarp();
        if (lane == 0) {
          const int old = atomicAdd(&s_cnt[stage], 1);
          if (old == NWARPS - 1) {
            s_cnt[stage] = 0;
            if (pct < ct_end) issue_stage(stage);
          }
        }
        if (pct < ct_end) advance_cursor();
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1u;
        }
        kin = 0;
        nin = min(SK, kend - ks);
        gp = gsw + stage * STAGE_DOUBLES;
        if (!tile_done) mbar_wait(full_bar + stage, phase);
      }
      if (!tile_done) est_load_frag(f0, gp, vp, LVS);
      if (blk_done) {  // fold the finished block: Y += theta_b * T
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
          for (int nt = 0; nt < NTN; ++nt) {
            Y[mt][nt][0] = fma(th[mt], T[mt][nt][0], Y[mt][nt][0]);
            Y[mt][nt][1] = fma(th[mt], T[mt][nt][1], Y[mt][nt][1]);
          }
          th[mt] = thn[mt];
        }
        if (!tile_done) {
          const int scn = (b + 1 < a.nblk) ? s_scale[b + 1] : -1;
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) thn[mt] = (scn < 0) ? 1.0 : __ldg(thp[mt] + scn);
          fresh = 1;
          if (nvalid < 2) {  // (nvalid already describes the next block)
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
              for (int nt = 0; nt < NTN; ++nt) T[mt][nt][0] = T[mt][nt][1] = 0.0;
            fresh = 0;
          }
        }
      }
      if (tile_done) break;
    }

    // epilogue of the column tile: rowacc[p] += sum_c Y[p,c] * z[p,c]
    const int cbase = ct * EST_BN + wn * (8 * NTN) + q4 * 2;
#pragma unroll
    for (int nt = 0; nt < NTN; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = cbase + nt * 8 + e;
        const int csc = __ldg(a.col_scale + c);
        const int csr = __ldg(a.col_src + c);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          const double thc = (csc < 0) ? 1.0 : __ldg(thp[mt] + csc);
          const double zc = thc * Vs[(r0 + mt * 8) * LVS + csr];
          rowacc[mt] = fma(Y[mt][nt][e], zc, rowacc[mt]);
        }
      }
    }
  }

  // reduce over the quad (columns) and over the two n-warps
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    rowacc[mt] += __shfl_xor_sync(0xffffffffu, rowacc[mt], 1);
    rowacc[mt] += __shfl_xor_sync(0xffffffffu, rowacc[mt], 2);
  }
  if (wn > 0 && q4 == 0) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) red[(wn - 1) * EST_BM + r0 + mt * 8] = rowacc[mt];
  }
  __syncthreads();
  if (wn == 0 && q4 == 0) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const long long p = p0 + r0 + mt * 8;
      double e = rowacc[mt];
#pragma unroll
      for (int w2 = 0; w2 < WN - 1; ++w2) e += red[w2 * EST_BM + r0 + mt * 8];  // fixed order: deterministic
      if (p < a.B) a.partial[(long long)split * a.B + p] = e;
    }
  }
}

// =====================================================================================================================
// K_fin: eps2 = sum of split partials (fixed order), Delta, per-block arg-max; then one block reduces the blocks.
// rbnics/problems/elliptic/elliptic_rb_reduced_problem.py:35-40 and parameter_space_subset.py:67,75.
// =====================================================================================================================
struct MaxLoc {
  double val;
  long long idx;
  long long n_nonfinite;
};

__device__ __forceinline__ void block_argmax(double& v, long long& ix, long long& nf) {
  __shared__ double sv[32];
  __shared__ long long si[32];
  __shared__ long long sn[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int off = 16; off > 0; off >>= 1) {
    const double v2 = __shfl_xor_sync(0xffffffffu, v, off);
    const long long i2 = __shfl_xor_sync(0xffffffffu, ix, off);
    nf += __shfl_xor_sync(0xffffffffu, nf, off);
    if (argmax_better(v2, i2, v, ix)) {
      v = v2;
      ix = i2;
    }
  }
  if (lane == 0) {
    sv[warp] = v;
    si[warp] = ix;
    sn[warp] = nf;
  }
  __syncthreads();
  if (warp == 0) {
    v = lane < nw ? sv[lane] : -INFINITY;
    ix = lane < nw ? si[lane] : 0x7fffffffffffffffLL;
    nf = lane < nw ? sn[lane] : 0;
    for (int off = 16; off > 0; off >>= 1) {
      const double v2 = __shfl_xor_sync(0xffffffffu, v, off);
      const long long i2 = __shfl_xor_sync(0xffffffffu, ix, off);
      nf += __shfl_xor_sync(0xffffffffu, nf, off);
      if (argmax_better(v2, i2, v, ix)) {
        v = v2;
        ix = i2;
      }
    }
  }
}

__global__ void __launch_bounds__(256) rb_finalize_kernel(const double* __restrict__ partial, int nsplit,
                                                          const double* __restrict__ beta, long long B, int kind,
                                                          long long idx_off, long long idx_stride,
                                                          double* __restrict__ delta, double* __restrict__ eps2_out,
                                                          double* __restrict__ blk_val, long long* __restrict__ blk_idx,
                                                          long long* __restrict__ blk_nf) {
  double best = -INFINITY;
  long long besti = 0x7fffffffffffffffLL, nf = 0;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < B; p += (long long)gridDim.x * blockDim.x) {
    double e = 0.0;
    for (int s = 0; s < nsplit; ++s) e += partial[(long long)s * B + p];
    const double bt = beta[p];
    const double d = (kind == RB_EST_SQRT_EPS2_OVER_BETA) ? sqrt(fabs(e)) / bt : sqrt(fabs(e) / bt);
    delta[p] = d;
    eps2_out[p] = e;
    if (!isfinite(d)) ++nf;
    const long long gi = idx_off + p * idx_stride;
    if (argmax_better(d, gi, best, besti)) {
      best = d;
      besti = gi;
    }
  }
  block_argmax(best, besti, nf);
  if (threadIdx.x == 0) {
    blk_val[blockIdx.x] = best;
    blk_idx[blockIdx.x] = besti;
    blk_nf[blockIdx.x] = nf;
  }
}

__global__ void __launch_bounds__(256) rb_argmax_final_kernel(const double* __restrict__ blk_val,
                                                              const long long* __restrict__ blk_idx,
                                                              const long long* __restrict__ blk_nf, int nblocks,
                                                              MaxLoc* __restrict__ out) {
  double best = -INFINITY;
  long long besti = 0x7fffffffffffffffLL, nf = 0;
  for (int t = threadIdx.x; t < nblocks; t += blockDim.x) {
    nf += blk_nf[t];
    if (argmax_better(blk_val[t], blk_idx[t], best, besti)) {
      best = blk_val[t];
      besti = blk_idx[t];
    }
  }
  block_argmax(best, besti, nf);
  if (threadIdx.x == 0) {
    out->val = best;
    out->idx = besti;
    out->n_nonfinite = nf;
  }
}

// =====================================================================================================================
// Host side
// =====================================================================================================================
static int roundup(int x, int m) { return (x + m - 1) / m * m; }

template <typename T>
static int dev_realloc(rb_ctx* ctx, T** p, size_t count) {
  if (*p) {
    RB_CUDA(cudaFree(*p));
    *p = nullptr;
  }
  if (count) RB_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
  return RB_OK;
}

extern "C" int rb_ctx_create(int device, rb_ctx** out) {
  if (!out) return RB_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return RB_ERR_CUDA;
  rb_ctx* ctx = new rb_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete ctx;
    return RB_ERR_CUDA;
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cc_major = prop.major;
  ctx->cc_minor = prop.minor;
  if (prop.major != 10) {  // sm_100a only: no fallback path exists
    delete ctx;
    return RB_ERR_UNSUPPORTED;
  }
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return RB_ERR_CUDA;
  }
  bool ok = true;
  for (int i = 0; i < 6; ++i) ok = ok && cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->dResult, sizeof(MaxLoc)) == cudaSuccess;
  if (!ok) {  // release what was created so far
    for (int i = 0; i < 6; ++i)
      if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return RB_ERR_CUDA;
  }
  *out = ctx;
  return RB_OK;
}

extern "C" void rb_ctx_destroy(rb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  void* ptrs[] = {ctx->dAq,       ctx->dF,       ctx->dBcRows,  ctx->dGp,      ctx->dTileOff, ctx->dTileKs0,
                  ctx->dTileB0,   ctx->dBlkKsBegin, ctx->dBlkScale, ctx->dBlkSrc, ctx->dColScale, ctx->dColSrc,
                  ctx->dTileCls,  ctx->dBlkCls,  ctx->dTileCb,
                  ctx->dTheta,    ctx->dThetaBC, ctx->dBeta,    ctx->dAb,      ctx->dU,       ctx->dPart,
                  ctx->dDelta,    ctx->dEps2,    ctx->dBlkVal,  ctx->dBlkIdx,  ctx->dResult,  ctx->dSplitBegin,
                  ctx->dLuScratch, ctx->dGdense,  ctx->dZsrc,    ctx->dZscale,  ctx->dWeights};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  rb_pair_free(ctx);
  rb_asmrec_free(ctx);
  for (int t = 0; t < RB_OFF_SLOTS; ++t) {
    for (void* p : {(void*)ctx->off_csr[t].ip, (void*)ctx->off_csr[t].ix, (void*)ctx->off_csr[t].dt, (void*)ctx->off_blk[t].d,
                    (void*)ctx->off_blk[t].zt, (void*)ctx->off_blk[t].xzt})
      if (p) cudaFree(p);
  }
  for (void* p : ctx->off_scratch)
    if (p) cudaFree(p);
  for (int i = 0; i < 6; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (cudaEvent_t e : ctx->up_ev) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* rb_last_error(const rb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null handle"; }

extern "C" int rb_device_info(const rb_ctx* ctx, int* cc_major, int* cc_minor, int* sm_count) {
  if (!ctx) return RB_ERR_ARG;
  if (cc_major) *cc_major = ctx->cc_major;
  if (cc_minor) *cc_minor = ctx->cc_minor;
  if (sm_count) *sm_count = ctx->sm_count;
  return RB_OK;
}

extern "C" void* rb_host_alloc(int64_t bytes) {
  void* p = nullptr;
  if (bytes <= 0 || cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) return nullptr;
  return p;
}
extern "C" void rb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

extern "C" int rb_set_system(rb_ctx* ctx, int N, int Ql, int Qr, const double* A, const double* F, int n_bc,
                             const int* bc_rows) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (N < 0 || Ql < 0 || Qr < 0 || n_bc < 0 || n_bc > N) RB_FAIL(RB_ERR_ARG, "rb_set_system: negative size or n_bc > N");
  if (N > 0 && (Ql == 0 || !A || (Qr > 0 && !F))) RB_FAIL(RB_ERR_ARG, "rb_set_system: N > 0 needs A (and F when Qr > 0)");
  if (n_bc > 0 && !bc_rows) RB_FAIL(RB_ERR_ARG, "rb_set_system: bc_rows is NULL");
  for (int t = 0; t < n_bc; ++t)
    if (bc_rows[t] < 0 || bc_rows[t] >= N) RB_FAIL(RB_ERR_ARG, "rb_set_system: bc row out of range");
  const int NP = roundup(N, 4);
  if (NP > LU_MAX_NP) RB_FAIL(RB_ERR_UNSUPPORTED, "rb_set_system: N exceeds the supported maximum (160)");
  ctx->N = N;
  ctx->NP = NP;
  ctx->Ql = Ql;
  ctx->Qr = Qr;
  ctx->QlP = std::max(4, roundup(Ql, 4));
  ctx->n_bc = n_bc;
  ctx->E = NP * NP;
  ctx->EP = roundup(std::max(ctx->E, 1), ASM_BE);
  ctx->ldu = std::max(N, 1);
  ctx->have_system = false;
  if (N > 0) {
    if (int rc = dev_realloc(ctx, &ctx->dAq, (size_t)ctx->QlP * ctx->EP)) return rc;
    if (int rc = dev_realloc(ctx, &ctx->dF, (size_t)std::max(Qr, 1) * NP)) return rc;
    if (int rc = dev_realloc(ctx, &ctx->dBcRows, (size_t)std::max(n_bc, 1))) return rc;
    double* dA = nullptr;  // (freed on every path below)
    RB_CUDA(cudaMalloc(&dA, (size_t)Ql * N * N * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(dA, A, (size_t)Ql * N * N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
      rb_pack_system_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(dA, ctx->dAq, N, NP, Ql, ctx->QlP, ctx->EP);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dA);
    if (e != cudaSuccess) RB_FAIL(RB_ERR_CUDA, std::string("rb_set_system: ") + cudaGetErrorString(e));
    std::vector<double> Fp((size_t)std::max(Qr, 1) * NP, 0.0);
    for (int q = 0; q < Qr; ++q)
      for (int i = 0; i < N; ++i) Fp[(size_t)q * NP + i] = F[(size_t)q * N + i];
    RB_CUDA(cudaMemcpy(ctx->dF, Fp.data(), Fp.size() * sizeof(double), cudaMemcpyHostToDevice));
    if (n_bc) RB_CUDA(cudaMemcpy(ctx->dBcRows, bc_rows, n_bc * sizeof(int), cudaMemcpyHostToDevice));
  }
  ctx->have_system = true;
  ctx->have_est = false;  // the estimator layout depends on N
  return rb_asmrec_setup(ctx);
}

extern "C" int rb_set_estimator(rb_ctx* ctx, int nblk, const int* blk_scale_col, const int* blk_src_off,
                                const int* blk_len, int v_theta_off, int v_theta_len, const double* G) {
  if (!ctx) return RB_ERR_ARG;
  return rb_set_estimator_ex(ctx, nblk, blk_scale_col, blk_src_off, blk_len, v_theta_off, v_theta_len, G, ctx->N);
}

extern "C" int rb_set_estimator_ex(rb_ctx* ctx, int nblk, const int* blk_scale_col, const int* blk_src_off,
                                   const int* blk_len, int v_theta_off, int v_theta_len, const double* G,
                                   int solution_len) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->have_system) RB_FAIL(RB_ERR_STATE, "rb_set_estimator: call rb_set_system first");
  if (nblk <= 0 || !blk_scale_col || !blk_src_off || !blk_len || !G) RB_FAIL(RB_ERR_ARG, "rb_set_estimator: bad arguments");
  const int QT = ctx->Ql + ctx->Qr;
  if (v_theta_off < 0 || v_theta_len < 0 || v_theta_off + v_theta_len > QT)
    RB_FAIL(RB_ERR_ARG, "rb_set_estimator: v_theta range outside [0, Ql+Qr)");
  if (solution_len < 0) RB_FAIL(RB_ERR_ARG, "rb_set_estimator: negative solution length");
  const int LV = solution_len + v_theta_len + 1;
  int Kz = 0, Kp = 0;
  std::vector<int> ks_begin(nblk + 1, 0);
  for (int b = 0; b < nblk; ++b) {
    if (blk_len[b] <= 0) RB_FAIL(RB_ERR_ARG, "rb_set_estimator: block length must be positive");
    if (blk_src_off[b] < 0 || blk_src_off[b] + blk_len[b] > LV) RB_FAIL(RB_ERR_ARG, "rb_set_estimator: block source outside V");
    if (blk_scale_col[b] >= QT) RB_FAIL(RB_ERR_ARG, "rb_set_estimator: scale column outside theta");
    Kz += blk_len[b];
    Kp += roundup(blk_len[b], 4);
    ks_begin[b + 1] = Kp / 4;
  }
  const int KS = Kp / 4;
  if (nblk > EST_MAX_BLOCKS) RB_FAIL(RB_ERR_UNSUPPORTED, "rb_set_estimator: too many z blocks");
  // ---- slice classes: blocks that scale the same slice of V are foldable against each other (see the pack kernel)
  std::vector<int> cls(nblk), ubegin(nblk), scale_norm(nblk);
  for (int b = 0, u = 0; b < nblk; ++b) {
    cls[b] = b;
    for (int b2 = 0; b2 < b; ++b2)
      if (blk_src_off[b2] == blk_src_off[b] && blk_len[b2] == blk_len[b]) {
        cls[b] = cls[b2];
        break;
      }
    ubegin[b] = u;
    u += blk_len[b];
    scale_norm[b] = blk_scale_col[b] < 0 ? -1 : blk_scale_col[b];
  }
  // ---- padded rows
  std::vector<int> row_unpad(Kp, -1), row_blk(Kp, 0), row_idx(Kp, 0);
  for (int b = 0; b < nblk; ++b)
    for (int t = 0; t < roundup(blk_len[b], 4); ++t) {
      const int kp = 4 * ks_begin[b] + t;
      row_blk[kp] = b;
      row_idx[kp] = t;
      if (t < blk_len[b]) row_unpad[kp] = ubegin[b] + t;
    }
  // ---- column tiles: 64 columns = two 32-column halves, each the range [32 r, 32 r + 32) of one block; both blocks
  // of a tile belong to one class, so the tile's k range skips 8 r k-steps in every row block of that class.  A last
  // range of <= 8 columns is packed 8 blocks to a tile instead.
  struct Tile {
    int b0, skip, cls;
  };
  static const int RW = [] {  // column-range width: 16 by default (4 blocks per tile); 8 / 32 for A/B measurements
    const int v = getenv("RB_EST_RW") ? atoi(getenv("RB_EST_RW")) : 16;
    return (v == 8 || v == 32) ? v : 16;
  }();
  const size_t GRP = EST_BN / RW;  // blocks per tile
  const int zero_src = LV;  // V[LV...] is zero filled
  std::vector<Tile> tiles;
  std::vector<int> col_unpad, col_blk, col_idx, col_scale, col_src;
  auto new_tile = [&](int b0, int skip, int c) {
    tiles.push_back({b0, skip, c});
    col_unpad.resize(tiles.size() * EST_BN, -1);
    col_blk.resize(tiles.size() * EST_BN, 0);
    col_idx.resize(tiles.size() * EST_BN, 0);
    col_scale.resize(tiles.size() * EST_BN, -1);
    col_src.resize(tiles.size() * EST_BN, zero_src);
    return (int)(tiles.size() - 1) * EST_BN;
  };
  auto put_cols = [&](int base, int pos, int b, int m0, int w) {
    for (int j = 0; j < w; ++j) {
      col_unpad[base + pos + j] = ubegin[b] + m0 + j;
      col_blk[base + pos + j] = b;
      col_idx[base + pos + j] = m0 + j;
      col_scale[base + pos + j] = scale_norm[b];
      col_src[base + pos + j] = blk_src_off[b] + m0 + j;
    }
  };
  for (int c = 0; c < nblk; ++c) {
    if (cls[c] != c) continue;
    std::vector<int> members;
    for (int b = c; b < nblk; ++b)
      if (cls[b] == c) members.push_back(b);
    const int L = blk_len[c];
    const int nr = (L + RW - 1) / RW;
    const int last_w = L - RW * (nr - 1);
    const bool leftover = (nr > 1 && last_w <= 8);
    const int nfull = leftover ? nr - 1 : nr;
    for (size_t i = 0; i < members.size(); i += GRP) {
      for (int r = 0; r < nfull; ++r) {
        const int w = std::min(RW, L - RW * r);
        const int base = new_tile(members[i], (RW * r) / 4, c);
        // A warp (wn) multiplies two of the four 16-column ranges of a tile, and row block b only meets column blocks
        // j <= b: with the blocks of a full group placed as [j0, j3 | j1, j2] both warps become busy one row block
        // earlier than with [j0, j1 | j2, j3] (the CTA advances at the pace of its busiest warp)
        static const bool est_perm = !(getenv("RB_EST_PERM") && atoi(getenv("RB_EST_PERM")) == 0);
        const bool full_group = (GRP == 4 && i + GRP <= members.size() && est_perm);
        static const int perm4[4] = {0, 2, 3, 1};
        for (size_t j = i; j < std::min(members.size(), i + GRP); ++j)
          put_cols(base, RW * (full_group ? perm4[j - i] : (int)(j - i)), members[j], RW * r, w);
      }
    }
    if (leftover)
      for (size_t i = 0; i < members.size(); i += 8) {
        const int base = new_tile(members[i], (RW * (nr - 1)) / 4, c);
        for (size_t j = i; j < std::min(members.size(), i + 8); ++j)
          put_cols(base, 8 * (int)(j - i), members[j], RW * (nr - 1), last_w);
      }
  }
  const int n_ctile = (int)tiles.size();
  const int Cp = n_ctile * EST_BN;
  // ---- records (k-steps) of every tile, in stream order
  std::vector<long long> tile_off(n_ctile);
  std::vector<int> tile_b0(n_ctile), tile_skip(n_ctile), tile_cls(n_ctile), tile_cost(n_ctile), rec_tile, rec_ks;
  // per 16-column range of a tile: the lowest column block in it (row blocks above it are structural zeros there);
  // -1 = always process (empty ranges cost nothing to skip but keep the fast path for mixed / narrow layouts)
  std::vector<int> tile_cb(4 * (size_t)n_ctile, -1);
  static const bool est_skip = !(getenv("RB_EST_SKIP") && atoi(getenv("RB_EST_SKIP")) == 0);
  for (int ct = 0; ct < n_ctile && est_skip; ++ct)
    for (int rg = 0; rg < 4; ++rg) {
      int lo = 0x7fffffff;
      for (int j = 0; j < 16; ++j)
        if (col_unpad[(size_t)ct * EST_BN + 16 * rg + j] >= 0) lo = std::min(lo, col_blk[(size_t)ct * EST_BN + 16 * rg + j]);
      tile_cb[4 * (size_t)ct + rg] = lo;  // an empty range never has data: 0x7fffffff skips it for every row block
    }
  for (int ct = 0; ct < n_ctile; ++ct) {
    tile_off[ct] = (long long)rec_tile.size() * EST_KSTEP_DOUBLES;
    tile_b0[ct] = tiles[ct].b0;
    tile_skip[ct] = tiles[ct].skip;
    tile_cls[ct] = tiles[ct].cls;
    for (int b = tiles[ct].b0; b < nblk; ++b)
      for (int ks = ks_begin[b] + (cls[b] == tiles[ct].cls ? tiles[ct].skip : 0); ks < ks_begin[b + 1]; ++ks) {
        rec_tile.push_back(ct);
        rec_ks.push_back(ks);
      }
    tile_cost[ct] = (int)(rec_tile.size() - tile_off[ct] / EST_KSTEP_DOUBLES);
  }
  // multiply-adds per parameter issued on the tensor pipe: every record is 4 rows x 64 columns, minus the half-width
  // heads (first two k-steps of every folded block visit longer than two k-steps, see est_mma<HEAD>)
  const bool tri16 = (RW == 16 && !(getenv("RB_EST_TRI") && atoi(getenv("RB_EST_TRI")) == 0));
  long long issued_macs = (long long)rec_tile.size() * EST_KSTEP_DOUBLES;
  if (tri16)
    for (int ct = 0; ct < n_ctile; ++ct)
      for (int b = tiles[ct].b0; b < nblk; ++b)
        if (cls[b] == tiles[ct].cls && ks_begin[b + 1] - ks_begin[b] - tiles[ct].skip > 2) issued_macs -= EST_KSTEP_DOUBLES;
  const long long total_rec = (long long)rec_tile.size();
  const long long off = total_rec * EST_KSTEP_DOUBLES;
  // V tile pitch: >= LV + 4 (A fragments of a padded block may read 3 entries past it; one zero slot) and == 4 mod 8
  int LVS = LV + 4;
  while (LVS % 8 != 4) ++LVS;
  // G' ring: 2 stages that end at z-block boundaries (stage = block for N <= 128: the stage hand-over and the block
  // fold then happen at the same k-step)
  int max_blk_ks = 1;
  for (int b = 0; b < nblk; ++b) max_blk_ks = std::max(max_blk_ks, ks_begin[b + 1] - ks_begin[b]);
  const int nstage = 2;
  int stage_ksteps = std::min(32, max_blk_ks);
  while (stage_ksteps > 2 && est_smem_bytes(nstage, stage_ksteps, LVS, nblk, n_ctile) > 227 * 1024) --stage_ksteps;
  const size_t smem = est_smem_bytes(nstage, stage_ksteps, LVS, nblk, n_ctile);
  if (smem > 227 * 1024) RB_FAIL(RB_ERR_UNSUPPORTED, "rb_set_estimator: N + v_theta_len too large for the estimator tile");

  ctx->nblk = nblk;
  ctx->Kz = Kz;
  ctx->Kp = Kp;
  ctx->KS = KS;
  ctx->n_ctile = n_ctile;
  ctx->v_off = v_theta_off;
  ctx->v_len = v_theta_len;
  ctx->LV = LV;
  ctx->LVS = LVS;
  ctx->stage_ksteps = stage_ksteps;
  ctx->est_nstage = nstage;
  ctx->est_vecN = solution_len;
  ctx->est_tri16 = tri16 ? 1 : 0;
  ctx->est_issued_macs = issued_macs;
  ctx->h_blk_ks_begin = ks_begin;
  ctx->h_tile_cost = tile_cost;
  ctx->gp_doubles = off;
  ctx->have_est = false;

  if (int rc = dev_realloc(ctx, &ctx->dGp, (size_t)std::max<long long>(off, 1))) return rc;
  if (int rc = dev_realloc(ctx, &ctx->dTileOff, (size_t)n_ctile)) return rc;
  if (int rc = dev_realloc(ctx, &ctx->dTileKs0, (size_t)n_ctile)) return rc;   // tile_skip
  if (int rc = dev_realloc(ctx, &ctx->dTileCls, (size_t)n_ctile)) return rc;
  if (int rc = dev_realloc(ctx, &ctx->dTileB0, (size_t)n_ctile)) return rc;
  if (int rc = dev_realloc(ctx, &ctx->dTileCb, (size_t)4 * n_ctile)) return rc;
  if (int rc = dev_realloc(ctx, &ctx->dBlkKsBegin, (size_t)nblk + 1)) return rc;
  if (int rc = dev_realloc(ctx, &ctx->dBlkScale, (size_t)nblk)) return rc;
  if (int rc = dev_realloc(ctx, &ctx->dBlkSrc, (size_t)nblk)) return rc;
  if (int rc = dev_realloc(ctx, &ctx->dBlkCls, (size_t)nblk)) return rc;
  if (int rc = dev_realloc(ctx, &ctx->dColScale, (size_t)Cp)) return rc;
  if (int rc = dev_realloc(ctx, &ctx->dColSrc, (size_t)Cp)) return rc;
  RB_CUDA(cudaMemcpy(ctx->dTileOff, tile_off.data(), n_ctile * sizeof(long long), cudaMemcpyHostToDevice));
  RB_CUDA(cudaMemcpy(ctx->dTileKs0, tile_skip.data(), n_ctile * sizeof(int), cudaMemcpyHostToDevice));
  RB_CUDA(cudaMemcpy(ctx->dTileCls, tile_cls.data(), n_ctile * sizeof(int), cudaMemcpyHostToDevice));
  RB_CUDA(cudaMemcpy(ctx->dTileB0, tile_b0.data(), n_ctile * sizeof(int), cudaMemcpyHostToDevice));
  RB_CUDA(cudaMemcpy(ctx->dTileCb, tile_cb.data(), 4 * n_ctile * sizeof(int), cudaMemcpyHostToDevice));
  RB_CUDA(cudaMemcpy(ctx->dBlkKsBegin, ks_begin.data(), (nblk + 1) * sizeof(int), cudaMemcpyHostToDevice));
  RB_CUDA(cudaMemcpy(ctx->dBlkScale, scale_norm.data(), nblk * sizeof(int), cudaMemcpyHostToDevice));
  RB_CUDA(cudaMemcpy(ctx->dBlkSrc, blk_src_off, nblk * sizeof(int), cudaMemcpyHostToDevice));
  RB_CUDA(cudaMemcpy(ctx->dBlkCls, cls.data(), nblk * sizeof(int), cudaMemcpyHostToDevice));
  RB_CUDA(cudaMemcpy(ctx->dColScale, col_scale.data(), Cp * sizeof(int), cudaMemcpyHostToDevice));
  RB_CUDA(cudaMemcpy(ctx->dColSrc, col_src.data(), Cp * sizeof(int), cudaMemcpyHostToDevice));

  // dense G + packing tables -> device -> packed G'
  double* dG = nullptr;
  int* dTab = nullptr;
  std::vector<int> tab;   // rec_tile | rec_ks | row_unpad | row_blk | row_idx | col_unpad | col_blk | col_idx | ubegin
  const size_t o_rec_tile = 0, o_rec_ks = o_rec_tile + rec_tile.size(), o_row_unpad = o_rec_ks + rec_ks.size(),
               o_row_blk = o_row_unpad + Kp, o_row_idx = o_row_blk + Kp, o_col_unpad = o_row_idx + Kp,
               o_col_blk = o_col_unpad + Cp, o_col_idx = o_col_blk + Cp, o_ubegin = o_col_idx + Cp;
  tab.reserve(o_ubegin + nblk);
  for (const std::vector<int>* v : {&rec_tile, &rec_ks, &row_unpad, &row_blk, &row_idx, &col_unpad, &col_blk, &col_idx, &ubegin})
    tab.insert(tab.end(), v->begin(), v->end());
  RB_CUDA(cudaMalloc(&dG, (size_t)Kz * Kz * sizeof(double)));
  cudaError_t e = cudaMalloc(&dTab, tab.size() * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpyAsync(dG, G, (size_t)Kz * Kz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dTab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && total_rec > 0) {
    PackArgs pa;
    pa.G = dG;
    pa.Gp = ctx->dGp;
    pa.total_rec = total_rec;
    pa.rec_tile = dTab + o_rec_tile;
    pa.rec_ks = dTab + o_rec_ks;
    pa.row_unpad = dTab + o_row_unpad;
    pa.row_blk = dTab + o_row_blk;
    pa.row_idx = dTab + o_row_idx;
    pa.col_unpad = dTab + o_col_unpad;
    pa.col_blk = dTab + o_col_blk;
    pa.col_idx = dTab + o_col_idx;
    pa.blk_cls = ctx->dBlkCls;
    pa.blk_ubegin = dTab + o_ubegin;
    pa.Kz = Kz;
    const long long nthr = total_rec * EST_KSTEP_DOUBLES;
    const int blocks = (int)std::min<long long>((nthr + 255) / 256, (long long)ctx->sm_count * 16);
    rb_pack_estimator_kernel<<<blocks, 256, 0, ctx->stream>>>(pa);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // tab is a local vector
  // the fused parabolic sweep multiplies with the dense G itself
  if (ctx->dGdense) cudaFree(ctx->dGdense);
  if (ctx->dZsrc) cudaFree(ctx->dZsrc);
  if (ctx->dZscale) cudaFree(ctx->dZscale);
  ctx->dGdense = nullptr;
  ctx->dZsrc = ctx->dZscale = nullptr;
  if (e == cudaSuccess && Kz <= 128) {
    std::vector<int> zsrc(Kz), zscale(Kz);
    for (int b = 0; b < nblk; ++b)
      for (int t = 0; t < blk_len[b]; ++t) {
        zsrc[ubegin[b] + t] = blk_src_off[b] + t;
        zscale[ubegin[b] + t] = scale_norm[b];
      }
    e = cudaMalloc(&ctx->dZsrc, Kz * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->dZscale, Kz * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(ctx->dZsrc, zsrc.data(), Kz * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(ctx->dZscale, zscale.data(), Kz * sizeof(int), cudaMemcpyHostToDevice);
    const size_t nfrag = rb_parabolic_pack_g_doubles(Kz);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->dGdense, nfrag * sizeof(double));
    if (e == cudaSuccess) {
      std::vector<double> frag(nfrag);
      rb_parabolic_pack_g(G, Kz, frag.data());
      e = cudaMemcpy(ctx->dGdense, frag.data(), nfrag * sizeof(double), cudaMemcpyHostToDevice);
    }
  }
  int pair_rc = RB_OK;
  if (e == cudaSuccess)
    pair_rc = rb_pair_setup(ctx, nblk, scale_norm.data(), blk_src_off, blk_len, cls.data(), ubegin.data(), v_theta_off,
                            v_theta_len, solution_len, dG, Kz);
  cudaFree(dG);
  if (dTab) cudaFree(dTab);
  if (e != cudaSuccess) RB_FAIL(RB_ERR_CUDA, std::string("rb_set_estimator: ") + cudaGetErrorString(e));
  if (pair_rc != RB_OK) return pair_rc;
  RB_CUDA(cudaFuncSetAttribute(rb_estimator_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#ifdef RB_AB_VARIANTS
  RB_CUDA(cudaFuncSetAttribute(rb_estimator_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
  ctx->have_est = true;
  return RB_OK;
}

// pieces == false: everything is resident when the call returns.  pieces == true (rb_sweep): beta / theta_bc first, then
// the Theta rows in pieces on the copy stream with one event per piece; the caller's buffers must stay valid until the
// sweep that follows has returned (rb_sweep synchronises both streams before it does).
static int set_training_set_impl(rb_ctx* ctx, int64_t B, int n_theta, const double* Theta, int n_bc,
                                 const double* ThetaBC, const double* beta, int64_t index_offset, int64_t index_stride,
                                 bool pieces) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (B <= 0 || n_theta <= 0 || !Theta || !beta || n_bc < 0 || (n_bc > 0 && !ThetaBC) || index_stride <= 0)
    RB_FAIL(RB_ERR_ARG, "rb_set_training_set: bad arguments");
  if (B > ctx->Bcap || n_theta != ctx->QT || n_bc != ctx->n_bc_train) {
    if (int rc = dev_realloc(ctx, &ctx->dTheta, (size_t)B * n_theta)) return rc;
    if (int rc = dev_realloc(ctx, &ctx->dThetaBC, (size_t)B * std::max(n_bc, 1))) return rc;
    if (int rc = dev_realloc(ctx, &ctx->dBeta, (size_t)B)) return rc;
    ctx->Bcap = B;
  }
  ctx->B = B;
  ctx->QT = n_theta;
  ctx->n_bc_train = n_bc;
  ctx->idx_off = index_offset;
  ctx->idx_stride = index_stride;
  ctx->up_pieces = 0;
  ctx->have_train = false;
  if (!pieces) {
    RB_CUDA(cudaMemcpyAsync(ctx->dTheta, Theta, (size_t)B * n_theta * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (n_bc)
      RB_CUDA(cudaMemcpyAsync(ctx->dThetaBC, ThetaBC, (size_t)B * n_bc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    RB_CUDA(cudaMemcpyAsync(ctx->dBeta, beta, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->have_train = true;
    return RB_OK;
  }
  if (!ctx->copy_stream) RB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  // ~16 MB per piece: long enough for full PCIe rate, short enough that the first chunk of the sweep starts at once
  const long long piece = std::max<long long>(1024, (16ll << 20) / ((long long)n_theta * 8));
  const int np = (int)((B + piece - 1) / piece);
  while ((int)ctx->up_ev.size() < np) {
    cudaEvent_t e;
    RB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->up_ev.push_back(e);
  }
  cudaStream_t cs = ctx->copy_stream;
  if (n_bc) RB_CUDA(cudaMemcpyAsync(ctx->dThetaBC, ThetaBC, (size_t)B * n_bc * sizeof(double), cudaMemcpyHostToDevice, cs));
  RB_CUDA(cudaMemcpyAsync(ctx->dBeta, beta, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, cs));
  for (int i = 0; i < np; ++i) {
    const long long r0 = i * piece, r1 = std::min<long long>(B, r0 + piece);
    RB_CUDA(cudaMemcpyAsync(ctx->dTheta + r0 * n_theta, Theta + r0 * n_theta, (size_t)(r1 - r0) * n_theta * sizeof(double),
                            cudaMemcpyHostToDevice, cs));
    RB_CUDA(cudaEventRecord(ctx->up_ev[i], cs));
  }
  ctx->up_piece = piece;
  ctx->up_pieces = np;
  ctx->have_train = true;
  return RB_OK;
}

// the compute stream waits until the Theta rows [0, row_end) (and beta / theta_bc) have arrived
static int wait_uploaded(rb_ctx* ctx, long long row_end) {
  if (ctx->up_pieces <= 0) return RB_OK;
  const int idx = (int)std::min<long long>(ctx->up_pieces - 1, (row_end - 1) / ctx->up_piece);
  RB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->up_ev[idx], 0));
  return RB_OK;
}

extern "C" int rb_set_training_set(rb_ctx* ctx, int64_t B, int n_theta, const double* Theta, int n_bc,
                                   const double* ThetaBC, const double* beta, int64_t index_offset,
                                   int64_t index_stride) {
  return set_training_set_impl(ctx, B, n_theta, Theta, n_bc, ThetaBC, beta, index_offset, index_stride, false);
}

// choose the number of column splits so that units = row-blocks x splits fill the SMs evenly
static int choose_nsplit(long long nrb, int sms, int n_ctile) {
  int best = 1;
  double best_cost = 1e300;
  for (int s = 1; s <= 16 && s <= n_ctile; ++s) {
    const long long units = nrb * s;
    const long long waves = (units + sms - 1) / sms;
    const double cost = (double)waves / s * (1.0 + 0.004 * (s - 1));
    if (cost < best_cost - 1e-12) {
      best_cost = cost;
      best = s;
    }
  }
  return best;
}

// ---- helpers shared by the elliptic and the parabolic sweep -------------------------------------------------------------
// column splits of the estimator for `rows` parameters: chooses nsplit, uploads the split table (balanced by the number
// of k-step records of the tiles) and sizes the partial-sum buffer
static int prepare_estimator_splits(rb_ctx* ctx, long long rows, int* nsplit_out) {
  cudaStream_t st = ctx->stream;
  const long long nrb = (rows + EST_BM - 1) / EST_BM;
  const int n_ctile = ctx->pair.ok ? ctx->pair.n_tile : ctx->n_ctile;
  const std::vector<int>& tile_cost = ctx->pair.ok ? ctx->pair.h_tile_cost : ctx->h_tile_cost;
  const int nsplit = choose_nsplit(nrb, ctx->sm_count, n_ctile);
  // Pair-form kernel, large sweeps: the row blocks that fill whole waves of the SMs run at the coarse split; the
  // remaining ones go into a second launch with a finer split, so that the last, partial wave consists of short CTAs
  // (10^6 parameters: 7813 row blocks x 3 splits = 158.4 waves cost 159; 158 waves + 19 row blocks x 7 splits cost 158.5).
  // Time unit below: one CTA that processes all column tiles of a row block.
  rb_pair_t& pr = ctx->pair;
  pr.bulk_nsplit = nsplit;
  pr.bulk_nrb = nrb;
  pr.tail_nsplit = 0;
  pr.tail_nrb = 0;
  static const bool two_part = !(getenv("RB_EST_TAIL") && atoi(getenv("RB_EST_TAIL")) == 0);
  if (pr.ok && two_part && rows >= 20000) {
    const long long sms = ctx->sm_count, units = nrb * nsplit, waves_full = units / sms;
    if (units % sms != 0 && waves_full >= 40) {  // (measured: no gain at 19 waves, -0.34 % at 158)
      const long long nrb_bulk = waves_full * sms / nsplit, nrb_tail = nrb - nrb_bulk;
      int best_s = 1;
      double best_cost = 1e300;
      for (int s = 1; s <= 16 && s <= n_ctile; ++s) {
        const double cost = (double)((nrb_tail * s + sms - 1) / sms) / s * (1.0 + 0.01 * (s - 1));
        if (cost < best_cost - 1e-12) {
          best_cost = cost;
          best_s = s;
        }
      }
      const double t_single = (double)((units + sms - 1) / sms) / nsplit;
      const double t_plan = (double)waves_full / nsplit + best_cost;
      if (nrb_tail > 0 && t_plan < 0.997 * t_single) {
        pr.bulk_nrb = nrb_bulk;
        pr.tail_nrb = nrb_tail;
        pr.tail_nsplit = best_s;
      }
    }
  }
  const int nsum = std::max(nsplit, pr.tail_nsplit);  // partial sums per parameter that the finalize kernel adds
  if ((long long)nsum * rows > ctx->part_cap) {
    if (int rc = dev_realloc(ctx, &ctx->dPart, (size_t)nsum * rows)) return rc;
    ctx->part_cap = (long long)nsum * rows;
  }
  if (nsplit + 1 > ctx->split_cap) {
    if (int rc = dev_realloc(ctx, &ctx->dSplitBegin, (size_t)nsplit + 1)) return rc;
    ctx->split_cap = nsplit + 1;
  }
  if (pr.tail_nsplit + 1 > pr.tail_cap) {
    if (int rc = dev_realloc(ctx, &pr.dSplitBeginTail, (size_t)pr.tail_nsplit + 1)) return rc;
    pr.tail_cap = pr.tail_nsplit + 1;
  }
  long long total = 0;
  for (int ct = 0; ct < n_ctile; ++ct) total += tile_cost[ct];
  auto split_table = [&](int ns) {  // contiguous tile ranges balanced by the number of records
    std::vector<int> sb(ns + 1, n_ctile);
    long long accum = 0;
    int s = 0;
    sb[0] = 0;
    for (int ct = 0; ct < n_ctile && s + 1 < ns; ++ct) {
      accum += tile_cost[ct];
      if (accum * ns >= total * (s + 1)) sb[++s] = ct + 1;
    }
    for (int t = s + 1; t <= ns; ++t) sb[t] = n_ctile;
    return sb;
  };
  const std::vector<int> sb = split_table(nsplit), sbt = split_table(std::max(pr.tail_nsplit, 1));
  RB_CUDA(cudaMemcpyAsync(ctx->dSplitBegin, sb.data(), (nsplit + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  if (pr.tail_nsplit > 0)
    RB_CUDA(cudaMemcpyAsync(pr.dSplitBeginTail, sbt.data(), (pr.tail_nsplit + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  RB_CUDA(cudaStreamSynchronize(st));  // the tables are stack vectors
  *nsplit_out = nsum;
  return RB_OK;
}

// one estimator launch over `rows` parameters: V = [Uvec (vecN entries, pitch ldu) | theta[v_off : v_off+v_len] | 1]
static int launch_estimator(rb_ctx* ctx, const double* Uvec, int ldu, int vecN, const double* theta, long long rows,
                            int nsplit) {
  if (ctx->pair.ok && vecN == ctx->pair.vecN) return rb_pair_launch(ctx, Uvec, ldu, theta, rows, nsplit);
  if (ctx->pair.tail_nrb > 0) RB_FAIL(RB_ERR_STATE, "estimator: two-part launch plan without the pair-form kernel");
  EstArgs ea;
  ea.Gp = ctx->dGp;
  ea.tile_off = ctx->dTileOff;
  ea.tile_skip = ctx->dTileKs0;
  ea.tile_cls = ctx->dTileCls;
  ea.tile_b0 = ctx->dTileB0;
  ea.tile_cb = ctx->dTileCb;
  ea.blk_ks_begin = ctx->dBlkKsBegin;
  ea.blk_scale = ctx->dBlkScale;
  ea.blk_src = ctx->dBlkSrc;
  ea.blk_cls = ctx->dBlkCls;
  ea.col_scale = ctx->dColScale;
  ea.col_src = ctx->dColSrc;
  ea.split_begin = ctx->dSplitBegin;
  ea.U = Uvec;
  ea.theta = theta;
  ea.partial = ctx->dPart;
  ea.B = rows;
  ea.nblk = ctx->nblk;
  ea.KS = ctx->KS;
  ea.N = vecN;
  ea.ldu = ldu;
  ea.QT = ctx->QT;
  ea.v_off = ctx->v_off;
  ea.v_len = ctx->v_len;
  ea.LVS = ctx->LVS;
  ea.nsplit = nsplit;
  ea.n_ctile = ctx->n_ctile;
  ea.stage_ksteps = ctx->stage_ksteps;
  ea.nstage = ctx->est_nstage;
  ea.tri16 = ctx->est_tri16;
  const long long nrb = (rows + EST_BM - 1) / EST_BM;
  const size_t smem = est_smem_bytes(ctx->est_nstage, ctx->stage_ksteps, ctx->LVS, ctx->nblk, ctx->n_ctile);
#ifdef RB_AB_VARIANTS
  static const bool est16 = getenv("RB_EST_WARPS") && atoi(getenv("RB_EST_WARPS")) == 16;  // A/B switch (default 8)
  if (est16)
    rb_estimator_kernel<2><<<(unsigned)(nrb * nsplit), 512, smem, ctx->stream>>>(ea);
  else
#endif
    rb_estimator_kernel<4><<<(unsigned)(nrb * nsplit), 256, smem, ctx->stream>>>(ea);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}

// chunk size of the assembly scratch and its (re)allocation
int rb_lu_mw_quantum(int N);
static int prepare_assembly(rb_ctx* ctx, long long B, long long* chunk_out, size_t* asm_smem_out) {
  // at least one wave of the assembly GEMM (148 x 128 parameters), and as many parameters as fit 6 GB of assembled
  // matrices (69,264 at N = 100): every LU launch ends with a tail of one system time (~0.15 ms) in which the SMs run
  // dry, so fewer, longer launches pay (1.5 GB chunks: 157.1 ms per 10^6 parameters, 6 GB: 153.8, 12 GB: 153.7)
  long long chunk = (long long)ctx->sm_count * ASM_BM;
  {
    const long long budget = 6144LL << 20;
    const long long fit = budget / ((long long)ctx->EP * (long long)sizeof(double));
    if (fit > chunk) chunk = fit / chunk * chunk;
  }
  if (const char* e = getenv("RB_LU_CHUNK")) {  // tuning knob; anything that is not a positive count is ignored
    const long long v = atoll(e);
    if (v > 0) chunk = v;
  }
  // the persistent LU grid deals systems round robin to its CTAs: a chunk that is a multiple of the grid leaves no
  // partial wave (18,944 systems over 1,332 CTAs: 14.2 waves -> one extra system time per launch, ~7 %)
  if (ctx->NP > 32 && ctx->N <= 128) {
    const long long quantum = rb_lu_mw_quantum(ctx->N);
    if (quantum > 0 && chunk > quantum) chunk = chunk / quantum * quantum;
  }
  chunk = std::max<long long>(1, std::min<long long>(B, chunk));
  if (chunk * ctx->EP > ctx->ab_cap) {
    if (int rc = dev_realloc(ctx, &ctx->dAb, (size_t)chunk * ctx->EP)) return rc;
    ctx->ab_cap = chunk * ctx->EP;
  }
  const int TS = ctx->QlP + ((ctx->QlP % 8 == 4) ? 0 : 4);
  const size_t asm_smem = ((size_t)ASM_BM * TS + 2 * (size_t)ctx->QlP * (ASM_BE + 4)) * sizeof(double);
  if (asm_smem > 227 * 1024) RB_FAIL(RB_ERR_UNSUPPORTED, "rb_sweep: too many lhs terms for the assembly tile");
  RB_CUDA(cudaFuncSetAttribute(rb_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)asm_smem));
  *chunk_out = chunk;
  *asm_smem_out = asm_smem;
  return RB_OK;
}

// Ab[0..n) = sum_q Theta[c0 + p][q] A_q for the chunk starting at parameter c0
static int launch_assembly(rb_ctx* ctx, long long c0, long long n, size_t asm_smem) {
  if (ctx->asmrec_ok) return rb_asmrec_launch(ctx, ctx->dTheta + c0 * ctx->QT, n);
  AsmArgs aa;
  aa.theta = ctx->dTheta + c0 * ctx->QT;
  aa.Aq = ctx->dAq;
  aa.Ab = ctx->dAb;
  aa.count = n;
  aa.QT = ctx->QT;
  aa.Ql = ctx->Ql;
  aa.QlP = ctx->QlP;
  aa.EP = ctx->EP;
  // row blocks x entry-tile splits: 2 resident CTAs per SM.  The split is the one in [2 waves .. 4 waves] that wastes the
  // least of its last wave (146 row blocks x 5 splits = 730 CTAs on 296 slots ran 3 waves for 2.47 waves of work)
  const unsigned nrb_asm = (unsigned)((n + ASM_BM - 1) / ASM_BM);
  const unsigned slots = 2u * (unsigned)ctx->sm_count;
  const unsigned max_split = (unsigned)(ctx->EP / ASM_BE);
  unsigned esplit = 1;
  double best_eff = 0.0;
  for (unsigned e = 1; e <= max_split && (unsigned long long)e * nrb_asm <= 4ull * slots + nrb_asm; ++e) {
    const unsigned long long total = (unsigned long long)e * nrb_asm;
    const unsigned long long waves = (total + slots - 1) / slots;
    double eff = (double)total / (double)(waves * slots);
    if (total < slots) eff *= 0.5;  // less than one wave: not enough CTAs to hide the tile loads
    if (eff > best_eff - 1e-9) {  // ties: the finer split (shorter CTAs, smoother tail)
      best_eff = std::max(eff, best_eff);
      esplit = e;
    }
  }
  dim3 grid(esplit, nrb_asm);
  rb_assemble_kernel<<<grid, 256, asm_smem, ctx->stream>>>(aa);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}

extern "C" int rb_sweep_resident(rb_ctx* ctx, int estimator_kind, double* max_val, int64_t* max_idx, double* delta_out,
                                 double* u_out, double* eps2_out) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->have_system || !ctx->have_est || !ctx->have_train)
    RB_FAIL(RB_ERR_STATE, "rb_sweep: system, estimator and training set must all be set");
  if (ctx->QT != ctx->Ql + ctx->Qr) RB_FAIL(RB_ERR_ARG, "rb_sweep: training set theta width != Ql + Qr");
  if (ctx->n_bc_train != ctx->n_bc) RB_FAIL(RB_ERR_ARG, "rb_sweep: training set n_bc != system n_bc");
  if (estimator_kind != RB_EST_SQRT_EPS2_OVER_BETA && estimator_kind != RB_EST_SQRT_OF_EPS2_OVER_BETA)
    RB_FAIL(RB_ERR_ARG, "rb_sweep: unknown estimator kind");
  const long long B = ctx->B;
  const int N = ctx->N;
  cudaStream_t st = ctx->stream;
  int launches = 0;

  if (ctx->est_vecN != N) RB_FAIL(RB_ERR_STATE, "rb_sweep: the estimator was set for a different solution length");
  // ---- work buffers
  int nsplit = 1;
  if (int rc = prepare_estimator_splits(ctx, B, &nsplit)) return rc;
  if ((long long)B * ctx->ldu > ctx->u_cap) {
    if (int rc = dev_realloc(ctx, &ctx->dU, (size_t)B * ctx->ldu)) return rc;
    ctx->u_cap = (long long)B * ctx->ldu;
  }
  if (B > ctx->delta_cap) {
    if (int rc = dev_realloc(ctx, &ctx->dDelta, (size_t)B)) return rc;
    if (int rc = dev_realloc(ctx, &ctx->dEps2, (size_t)B)) return rc;
    ctx->delta_cap = B;
  }
  const int fin_blocks = (int)std::min<long long>((B + 255) / 256, ctx->sm_count * 8);
  if (fin_blocks > ctx->blk_cap) {
    if (int rc = dev_realloc(ctx, &ctx->dBlkVal, (size_t)fin_blocks)) return rc;
    if (int rc = dev_realloc(ctx, &ctx->dBlkIdx, (size_t)fin_blocks * 2)) return rc;
    ctx->blk_cap = fin_blocks;
  }

  RB_CUDA(cudaEventRecord(ctx->ev[0], st));
  if (N > 0) {
    // chunked assembly + LU; scratch = chunk x EP doubles
    long long chunk = 0;
    size_t asm_smem = 0;
    if (int rc = prepare_assembly(ctx, B, &chunk, &asm_smem)) return rc;
    // while the Theta rows are still arriving (rb_sweep from host buffers) the first chunk is a short one, so that the
    // solve starts after ~8 MB of the upload instead of after a whole 6 GB chunk's worth of rows
    const long long first = ctx->up_pieces > 0 ? std::min(chunk, std::max<long long>(1, chunk / 4)) : chunk;
    for (long long c0 = 0, n = 0; c0 < B; c0 += n) {
      n = std::min(c0 == 0 ? first : chunk, B - c0);
      if (int rc = wait_uploaded(ctx, c0 + n)) return rc;
      if (int rc = launch_assembly(ctx, c0, n, asm_smem)) return rc;
      ++launches;
      if (int rc = rb_launch_lu(ctx, n, ctx->dAb, ctx->dTheta + c0 * ctx->QT,
                                ctx->n_bc ? ctx->dThetaBC + c0 * ctx->n_bc : nullptr, ctx->dU + c0 * ctx->ldu))
        return rc;
      ++launches;
    }
  }
  RB_CUDA(cudaEventRecord(ctx->ev[1], st));
  if (int rc = wait_uploaded(ctx, B)) return rc;
  if (int rc = launch_estimator(ctx, ctx->dU, ctx->ldu, N, ctx->dTheta, B, nsplit)) return rc;
  launches += ctx->pair.tail_nrb > 0 ? 2 : 1;
  RB_CUDA(cudaEventRecord(ctx->ev[2], st));
  {
    long long* blk_nf = ctx->dBlkIdx + fin_blocks;
    rb_finalize_kernel<<<fin_blocks, 256, 0, st>>>(ctx->dPart, nsplit, ctx->dBeta, B, estimator_kind, ctx->idx_off,
                                                   ctx->idx_stride, ctx->dDelta, ctx->dEps2, ctx->dBlkVal,
                                                   ctx->dBlkIdx, blk_nf);
    RB_CUDA(cudaGetLastError());
    rb_argmax_final_kernel<<<1, 256, 0, st>>>(ctx->dBlkVal, ctx->dBlkIdx, blk_nf, fin_blocks, (MaxLoc*)ctx->dResult);
    RB_CUDA(cudaGetLastError());
    launches += 2;
  }
  RB_CUDA(cudaEventRecord(ctx->ev[3], st));
  MaxLoc res;
  RB_CUDA(cudaMemcpyAsync(&res, ctx->dResult, sizeof(MaxLoc), cudaMemcpyDeviceToHost, st));
  if (delta_out) RB_CUDA(cudaMemcpyAsync(delta_out, ctx->dDelta, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (eps2_out) RB_CUDA(cudaMemcpyAsync(eps2_out, ctx->dEps2, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (u_out && N > 0) RB_CUDA(cudaMemcpyAsync(u_out, ctx->dU, (size_t)B * N * sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaStreamSynchronize(st));
  float t01 = 0, t12 = 0, t23 = 0, t03 = 0;
  cudaEventElapsedTime(&t01, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&t12, ctx->ev[1], ctx->ev[2]);
  cudaEventElapsedTime(&t23, ctx->ev[2], ctx->ev[3]);
  cudaEventElapsedTime(&t03, ctx->ev[0], ctx->ev[3]);
  ctx->last_ms[0] = t01;  // assembly + LU (interleaved per chunk)
  ctx->last_ms[1] = 0.f;
  ctx->last_ms[2] = t12;
  ctx->last_ms[3] = t23;
  ctx->last_ms[4] = t03;
  ctx->last_launches = launches;
  if (max_val) *max_val = res.val;
  if (max_idx) *max_idx = res.idx;
  if (res.n_nonfinite > 0) {
    ctx->err = "rb_sweep: " + std::to_string(res.n_nonfinite) +
               " non-finite error estimators (singular reduced matrix or beta == 0); numpy.linalg.solve would raise "
               "LinAlgError";
    return RB_ERR_SINGULAR;
  }
  return RB_OK;
}

// =====================================================================================================================
// Parabolic (POD-Greedy) sweep: K implicit-Euler steps per parameter on top of the same assembly / LU / estimator
// kernels.  rbnics/reduction_methods/base/time_dependent_rb_reduction.py:262-295.
// =====================================================================================================================
// r[p][i] = sum_q theta_f[p][q] F[q][i] + sum_q theta_m[p][q] sum_j (M_q/dt)[i][j] u_prev[p][j]
// (Aq holds the lhs terms column-major: entry (i, j) of term q at q*EP + j*NP + i; the first Qm terms are M_q/dt)
__global__ void __launch_bounds__(128) rb_parabolic_rhs_kernel(const double* __restrict__ theta, int QT, int Ql, int Qr,
                                                               int Qm, const double* __restrict__ Aq, int EP, int NP,
                                                               const double* __restrict__ F, const double* __restrict__ UV,
                                                               int ldu, int N, long long n, double* __restrict__ rhs) {
  extern __shared__ double prs[];  // u_prev of the 4 parameters of this CTA
  const int lp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * 4 + lp;
  if (p < n)
    for (int j = lane; j < N; j += 32) prs[lp * N + j] = UV[p * ldu + j];
  __syncthreads();
  if (p >= n) return;
  const double* th = theta + p * QT;
  const double* up = prs + lp * N;
  for (int i = lane; i < N; i += 32) {
    double acc = 0.0;
    for (int q = 0; q < Qr; ++q) acc = fma(th[Ql + q], __ldg(F + q * NP + i), acc);
    for (int q = 0; q < Qm; ++q) {
      const double* Mq = Aq + (long long)q * EP + i;
      double s = 0.0;
      for (int j = 0; j < N; ++j) s = fma(__ldg(Mq + (long long)j * NP), up[j], s);
      acc = fma(th[q], s, acc);
    }
    rhs[p * N + i] = acc;
  }
}

// UV[p] = [u_new | (u_new - u_prev)/dt]; optional history u_hist[(c0 + p)][k][:]
__global__ void rb_parabolic_advance_kernel(const double* __restrict__ unew, double* __restrict__ UV, int ldu, int N,
                                            long long n, double inv_dt, double* __restrict__ u_hist, long long hist_pitch) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n * N) return;
  const long long p = t / N;
  const int i = (int)(t % N);
  const double un = unew[p * N + i], up = UV[p * ldu + i];
  UV[p * ldu + i] = un;
  UV[p * ldu + N + i] = (un - up) * inv_dt;
  if (u_hist) u_hist[p * hist_pitch + i] = un;
}

// eps2_k = sum of split partials; acc += w_k * |eps2_k| / beta
__global__ void rb_parabolic_accumulate_kernel(const double* __restrict__ partial, int nsplit, long long n,
                                               const double* __restrict__ beta, double w, double* __restrict__ acc,
                                               double* __restrict__ eps2_hist, long long hist_pitch) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  double e = 0.0;
  for (int s = 0; s < nsplit; ++s) e += partial[(long long)s * n + p];
  acc[p] = fma(w, fabs(e) / beta[p], acc[p]);
  if (eps2_hist) eps2_hist[p * hist_pitch] = e;
}

// Delta = sqrt(acc), block arg-max
__global__ void __launch_bounds__(256) rb_parabolic_finalize_kernel(const double* __restrict__ acc, long long B,
                                                                    long long idx_off, long long idx_stride,
                                                                    double* __restrict__ delta,
                                                                    double* __restrict__ blk_val,
                                                                    long long* __restrict__ blk_idx,
                                                                    long long* __restrict__ blk_nf) {
  double best = -INFINITY;
  long long besti = 0x7fffffffffffffffLL, nf = 0;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < B; p += (long long)gridDim.x * blockDim.x) {
    const double d = sqrt(acc[p]);
    delta[p] = d;
    if (!isfinite(d)) ++nf;
    const long long gi = idx_off + p * idx_stride;
    if (argmax_better(d, gi, best, besti)) {
      best = d;
      besti = gi;
    }
  }
  block_argmax(best, besti, nf);
  if (threadIdx.x == 0) {
    blk_val[blockIdx.x] = best;
    blk_idx[blockIdx.x] = besti;
    blk_nf[blockIdx.x] = nf;
  }
}

int rb_launch_lu_rhs(rb_ctx* ctx, int64_t count, const double* Ab, const double* theta, const double* thetaBC,
                     const double* rhs_in, double* u, int ldu);

// UV[p] = [u0[p] | 0]; acc[p] = w0 * |e0[p]|  (non-homogeneous initial condition)
__global__ void rb_parabolic_init_kernel(const double* __restrict__ u0, const double* __restrict__ e0, double w0,
                                         double* __restrict__ UV, int ldu, int N, long long n, double* __restrict__ acc,
                                         double* __restrict__ u_hist, long long hist_pitch) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n * N) return;
  const long long p = t / N;
  const int i = (int)(t % N);
  const double v = u0 ? u0[p * N + i] : 0.0;
  UV[p * ldu + i] = v;
  UV[p * ldu + N + i] = 0.0;
  if (u_hist) u_hist[p * hist_pitch + i] = v;
  if (i == 0) acc[p] = e0 ? w0 * fabs(e0[p]) : 0.0;
}

extern "C" int rb_sweep_parabolic(rb_ctx* ctx, int Qm, int K, double dt, const double* weights, double* max_val,
                                  int64_t* max_idx, double* delta_out, double* eps2_out, double* u_out) {
  return rb_sweep_parabolic_ic(ctx, Qm, K, dt, weights, nullptr, nullptr, max_val, max_idx, delta_out, eps2_out, u_out);
}

extern "C" int rb_sweep_parabolic_ic(rb_ctx* ctx, int Qm, int K, double dt, const double* weights, const double* U0,
                                     const double* init_err2, double* max_val, int64_t* max_idx, double* delta_out,
                                     double* eps2_out, double* u_out) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->have_system || !ctx->have_est || !ctx->have_train)
    RB_FAIL(RB_ERR_STATE, "rb_sweep_parabolic: system, estimator and training set must all be set");
  const int N = ctx->N;
  if (N <= 0) RB_FAIL(RB_ERR_ARG, "rb_sweep_parabolic: empty reduced space");
  if (Qm <= 0 || Qm >= ctx->Ql || K <= 0 || !(dt > 0.0) || !weights) RB_FAIL(RB_ERR_ARG, "rb_sweep_parabolic: bad arguments");
  if (ctx->QT != ctx->Ql + ctx->Qr) RB_FAIL(RB_ERR_ARG, "rb_sweep_parabolic: training set theta width != Ql + Qr");
  if (ctx->n_bc_train != ctx->n_bc) RB_FAIL(RB_ERR_ARG, "rb_sweep_parabolic: training set n_bc != system n_bc");
  if (ctx->est_vecN != 2 * N) RB_FAIL(RB_ERR_STATE, "rb_sweep_parabolic: set the estimator with solution_len = 2 N");
  const long long B = ctx->B;
  cudaStream_t st = ctx->stream;
  // small systems (NP <= 32, Kz <= 128: tutorial 06) run the whole time loop of a parameter in one kernel; everything
  // else goes through the per-step launches below (RB_PARABOLIC_FUSED=0 forces them: tests compare the two)
  const int VP = roundup(2 * ctx->NP + ctx->v_len + 2, 2);  // [u (NP) | udot (NP) | theta slice | 1 | 0]
  const bool allow_fused = !(getenv("RB_PARABOLIC_FUSED") && atoi(getenv("RB_PARABOLIC_FUSED")) == 0);
  const bool fused = allow_fused && ctx->dGdense && rb_parabolic_fused_supported(ctx->NP, ctx->Kz, Qm, VP);
  long long chunk = B;
  size_t asm_smem = 0;
  int nsplit = 1;
  if (!fused) {
    if (int rc = prepare_assembly(ctx, B, &chunk, &asm_smem)) return rc;
    if (int rc = prepare_estimator_splits(ctx, chunk, &nsplit)) return rc;
  }
  if (B > ctx->delta_cap) {
    if (int rc = dev_realloc(ctx, &ctx->dDelta, (size_t)B)) return rc;
    if (int rc = dev_realloc(ctx, &ctx->dEps2, (size_t)B)) return rc;
    ctx->delta_cap = B;
  }
  const int fin_blocks = (int)std::min<long long>((B + 255) / 256, ctx->sm_count * 8);
  if (fin_blocks > ctx->blk_cap) {
    if (int rc = dev_realloc(ctx, &ctx->dBlkVal, (size_t)fin_blocks)) return rc;
    if (int rc = dev_realloc(ctx, &ctx->dBlkIdx, (size_t)fin_blocks * 2)) return rc;
    ctx->blk_cap = fin_blocks;
  }
  // chunk-sized work arrays: [u | udot], rhs, new solution; optional histories (device side, whole training set)
  double *dUV = nullptr, *dRhs = nullptr, *dUnew = nullptr, *dHistE = nullptr, *dHistU = nullptr, *dU0 = nullptr,
         *dE0 = nullptr;
  auto free_all = [&]() {
    for (double* q : {dUV, dRhs, dUnew, dHistE, dHistU, dU0, dE0})
      if (q) cudaFree(q);
  };
#define RB_PB(call)                                                                                  \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      free_all();                                                                                    \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                                \
      return RB_ERR_CUDA;                                                                            \
    }                                                                                                \
  } while (0)
  if (!fused) {
    RB_PB(cudaMalloc(&dUV, (size_t)chunk * 2 * N * sizeof(double)));
    RB_PB(cudaMalloc(&dRhs, (size_t)chunk * N * sizeof(double)));
    RB_PB(cudaMalloc(&dUnew, (size_t)chunk * N * sizeof(double)));
  }
  if (eps2_out) RB_PB(cudaMalloc(&dHistE, (size_t)B * (K + 1) * sizeof(double)));
  if (u_out) RB_PB(cudaMalloc(&dHistU, (size_t)B * (K + 1) * N * sizeof(double)));
  if (dHistE) RB_PB(cudaMemsetAsync(dHistE, 0, (size_t)B * (K + 1) * sizeof(double), st));
  if (dHistU) RB_PB(cudaMemsetAsync(dHistU, 0, (size_t)B * (K + 1) * N * sizeof(double), st));
  if (U0) {  // reduced initial condition of every parameter (time_dependent_reduced_problem.py:277-302 + projection)
    RB_PB(cudaMalloc(&dU0, (size_t)B * N * sizeof(double)));
    RB_PB(cudaMemcpyAsync(dU0, U0, (size_t)B * N * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  if (init_err2) {  // get_initial_error_estimate_squared(): the bound at t0 (abstract_parabolic_rb_reduced_problem.py:41-44)
    RB_PB(cudaMalloc(&dE0, (size_t)B * sizeof(double)));
    RB_PB(cudaMemcpyAsync(dE0, init_err2, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  int launches = 0;
  if (fused) {
    if (K + 1 > ctx->weights_cap) {
      if (ctx->dWeights) cudaFree(ctx->dWeights);
      ctx->dWeights = nullptr;
      ctx->weights_cap = 0;
      RB_PB(cudaMalloc(&ctx->dWeights, (size_t)(K + 1) * sizeof(double)));
      ctx->weights_cap = K + 1;
    }
    RB_PB(cudaMemcpyAsync(ctx->dWeights, weights, (size_t)(K + 1) * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  RB_PB(cudaEventRecord(ctx->ev[0], st));
  if (fused) {
    PbArgs pa;
    pa.theta = ctx->dTheta;
    pa.thetaBC = ctx->n_bc ? ctx->dThetaBC : nullptr;
    pa.beta = ctx->dBeta;
    pa.Aq = ctx->dAq;
    pa.F = ctx->dF;
    pa.bc_rows = ctx->dBcRows;
    pa.Gfrag = ctx->dGdense;
    pa.zsrc = ctx->dZsrc;
    pa.zscale = ctx->dZscale;
    pa.weights = ctx->dWeights;
    pa.U0 = dU0;
    pa.E0 = dE0;
    pa.acc_out = ctx->dEps2;
    pa.eps2_hist = dHistE;
    pa.u_hist = dHistU;
    pa.n = B;
    pa.N = N;
    pa.EP = ctx->EP;
    pa.QT = ctx->QT;
    pa.Ql = ctx->Ql;
    pa.Qr = ctx->Qr;
    pa.Qm = Qm;
    pa.n_bc = ctx->n_bc;
    pa.K = K;
    pa.Kz = ctx->Kz;
    pa.v_off = ctx->v_off;
    pa.v_len = ctx->v_len;
    pa.VP = VP;
    pa.inv_dt = 1.0 / dt;
    if (int e = rb_parabolic_fused_launch(pa, ctx->NP, st)) {
      free_all();
      ctx->err = std::string("rb_parabolic_fused_kernel: ") + cudaGetErrorString((cudaError_t)e);
      return RB_ERR_CUDA;
    }
    ++launches;
  }
  for (long long c0 = 0; c0 < B && !fused; c0 += chunk) {
    const long long n = std::min(chunk, B - c0);
    if (int rc = launch_assembly(ctx, c0, n, asm_smem)) {
      free_all();
      return rc;
    }
    ++launches;
    rb_parabolic_init_kernel<<<(unsigned)((n * N + 255) / 256), 256, 0, st>>>(
        dU0 ? dU0 + c0 * N : nullptr, dE0 ? dE0 + c0 : nullptr, weights[0], dUV, 2 * N, N, n, ctx->dEps2 + c0,
        dHistU ? dHistU + c0 * (K + 1) * N : nullptr, (long long)(K + 1) * N);
    RB_PB(cudaGetLastError());
    ++launches;
    const double* th = ctx->dTheta + c0 * ctx->QT;
    for (int k = 1; k <= K; ++k) {
      rb_parabolic_rhs_kernel<<<(unsigned)((n + 3) / 4), 128, 4 * N * sizeof(double), st>>>(
          th, ctx->QT, ctx->Ql, ctx->Qr, Qm, ctx->dAq, ctx->EP, ctx->NP, ctx->dF, dUV, 2 * N, N, n, dRhs);
      RB_PB(cudaGetLastError());
      if (int rc = rb_launch_lu_rhs(ctx, n, ctx->dAb, th, ctx->n_bc ? ctx->dThetaBC + c0 * ctx->n_bc : nullptr, dRhs,
                                    dUnew, N)) {
        free_all();
        return rc;
      }
      rb_parabolic_advance_kernel<<<(unsigned)((n * N + 255) / 256), 256, 0, st>>>(
          dUnew, dUV, 2 * N, N, n, 1.0 / dt, dHistU ? dHistU + (c0 * (K + 1) + k) * N : nullptr, (long long)(K + 1) * N);
      RB_PB(cudaGetLastError());
      if (int rc = launch_estimator(ctx, dUV, 2 * N, 2 * N, th, n, nsplit)) {
        free_all();
        return rc;
      }
      rb_parabolic_accumulate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
          ctx->dPart, nsplit, n, ctx->dBeta + c0, weights[k], ctx->dEps2 + c0,
          dHistE ? dHistE + c0 * (K + 1) + k : nullptr, (long long)(K + 1));
      RB_PB(cudaGetLastError());
      launches += 5;
    }
  }
  long long* blk_nf = ctx->dBlkIdx + fin_blocks;
  rb_parabolic_finalize_kernel<<<fin_blocks, 256, 0, st>>>(ctx->dEps2, B, ctx->idx_off, ctx->idx_stride, ctx->dDelta,
                                                           ctx->dBlkVal, ctx->dBlkIdx, blk_nf);
  RB_PB(cudaGetLastError());
  rb_argmax_final_kernel<<<1, 256, 0, st>>>(ctx->dBlkVal, ctx->dBlkIdx, blk_nf, fin_blocks, (MaxLoc*)ctx->dResult);
  RB_PB(cudaGetLastError());
  launches += 2;
  RB_PB(cudaEventRecord(ctx->ev[3], st));
  MaxLoc res;
  RB_PB(cudaMemcpyAsync(&res, ctx->dResult, sizeof(MaxLoc), cudaMemcpyDeviceToHost, st));
  if (delta_out) RB_PB(cudaMemcpyAsync(delta_out, ctx->dDelta, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (eps2_out) RB_PB(cudaMemcpyAsync(eps2_out, dHistE, (size_t)B * (K + 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (u_out) RB_PB(cudaMemcpyAsync(u_out, dHistU, (size_t)B * (K + 1) * N * sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_PB(cudaStreamSynchronize(st));
#undef RB_PB
  free_all();
  float t03 = 0;
  cudaEventElapsedTime(&t03, ctx->ev[0], ctx->ev[3]);
  ctx->last_ms[0] = ctx->last_ms[1] = ctx->last_ms[2] = ctx->last_ms[3] = 0.f;
  ctx->last_ms[4] = t03;
  ctx->last_launches = launches;
  if (max_val) *max_val = res.val;
  if (max_idx) *max_idx = res.idx;
  if (res.n_nonfinite > 0) {
    ctx->err = "rb_sweep_parabolic: " + std::to_string(res.n_nonfinite) + " non-finite error estimators";
    return RB_ERR_SINGULAR;
  }
  return RB_OK;
}

extern "C" int rb_sweep(rb_ctx* ctx, int64_t B, int n_theta, const double* Theta, int n_bc, const double* ThetaBC,
                        const double* beta, int64_t index_offset, int64_t index_stride, int estimator_kind,
                        double* max_val, int64_t* max_idx, double* delta_out, double* u_out, double* eps2_out) {
  // host buffers in: the upload of the Theta rows overlaps the assembly + LU of the chunks that are already there
  int rc = set_training_set_impl(ctx, B, n_theta, Theta, n_bc, ThetaBC, beta, index_offset, index_stride, true);
  if (rc == RB_OK) rc = rb_sweep_resident(ctx, estimator_kind, max_val, max_idx, delta_out, u_out, eps2_out);
  if (ctx && ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);  // also on error paths: the caller's buffers are free again
    ctx->up_pieces = 0;
  }
  return rc;
}

// out[p] = sum_q ThetaS[p][q] * (S[q] . u_p): one warp per parameter
__global__ void __launch_bounds__(256) rb_outputs_kernel(const double* __restrict__ U, int ldu, int N, long long B, int Qs,
                                                         const double* __restrict__ S, const double* __restrict__ ThetaS,
                                                         double* __restrict__ out) {
  const long long p = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= B) return;
  double acc = 0.0;
  for (int q = 0; q < Qs; ++q) {
    double d = 0.0;
    for (int i = lane; i < N; i += 32) d = fma(__ldg(S + q * N + i), U[p * ldu + i], d);
    for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
    acc = fma(ThetaS[p * Qs + q], d, acc);
  }
  if (lane == 0) out[p] = acc;
}

extern "C" int rb_compute_outputs(rb_ctx* ctx, int Qs, const double* S, const double* ThetaS, double* out) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->have_system || !ctx->have_train || !ctx->dU || ctx->N <= 0)
    RB_FAIL(RB_ERR_STATE, "rb_compute_outputs: run a sweep first");
  if (Qs <= 0 || !S || !ThetaS || !out) RB_FAIL(RB_ERR_ARG, "rb_compute_outputs: bad arguments");
  const long long B = ctx->B;
  const int N = ctx->N;
  cudaStream_t st = ctx->stream;
  double *dS = nullptr, *dT = nullptr, *dO = nullptr;
  cudaError_t e = cudaMalloc(&dS, (size_t)Qs * N * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dT, (size_t)B * Qs * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dO, (size_t)B * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpyAsync(dS, S, (size_t)Qs * N * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dT, ThetaS, (size_t)B * Qs * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    rb_outputs_kernel<<<(unsigned)((B * 32 + 255) / 256), 256, 0, st>>>(ctx->dU, ctx->ldu, N, B, Qs, dS, dT, dO);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dO, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  for (double* q : {dS, dT, dO})
    if (q) cudaFree(q);
  if (e != cudaSuccess) RB_FAIL(RB_ERR_CUDA, std::string("rb_compute_outputs: ") + cudaGetErrorString(e));
  ctx->last_launches = 1;
  return RB_OK;
}

// Reduced supremizers of a batch of parameters (rbnics/problems/stokes/stokes_reduced_problem.py:80-103): the left-hand
// side X_s is the same for every parameter, so with W_q = X_s^{-1} Bt_q (host, once) the per-parameter solve becomes
//   s_p = sum_q theta_pq W_q u_p,   Wt[q][j][i] = W_q[i][j].
// One CTA of 128 threads per 8 parameters: theta and u of the 8 in shared memory, thread i owns entry i of the 8
// supremizers and walks Wt (contiguous in i, L2 resident) once for all 8.
constexpr int SUP_PB = 8;
__global__ void __launch_bounds__(128) rb_supremizer_kernel(const double* __restrict__ Wt, int Ns, int Nu, int Q,
                                                            const double* __restrict__ Theta,
                                                            const double* __restrict__ U, int ldu, long long B,
                                                            double* __restrict__ S) {
  extern __shared__ double sup_smem[];
  double* th = sup_smem;               // [PB][Q]
  double* us = sup_smem + SUP_PB * Q;  // [PB][Nu]
  const long long p0 = (long long)blockIdx.x * SUP_PB;
  for (int t = threadIdx.x; t < SUP_PB * Q; t += blockDim.x) {
    const long long p = min(p0 + t / Q, B - 1);
    th[t] = Theta[p * Q + t % Q];
  }
  for (int t = threadIdx.x; t < SUP_PB * Nu; t += blockDim.x) {
    const long long p = min(p0 + t / Nu, B - 1);
    us[t] = U[p * ldu + t % Nu];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Ns; i += blockDim.x) {
    double acc[SUP_PB];
#pragma unroll
    for (int p = 0; p < SUP_PB; ++p) acc[p] = 0.0;
    for (int q = 0; q < Q; ++q) {
      for (int j = 0; j < Nu; ++j) {
        const double w = __ldg(Wt + ((long long)q * Nu + j) * Ns + i);
#pragma unroll
        for (int p = 0; p < SUP_PB; ++p) acc[p] = fma(w, th[p * Q + q] * us[p * Nu + j], acc[p]);
      }
    }
#pragma unroll
    for (int p = 0; p < SUP_PB; ++p)
      if (p0 + p < B) S[(p0 + p) * Ns + i] = acc[p];
  }
}

extern "C" int rb_supremizers(rb_ctx* ctx, int Ns, int Nu, int Q, const double* W, int64_t B, const double* Theta,
                              const double* U, double* S) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (Ns <= 0 || Nu <= 0 || Q <= 0 || B <= 0 || !W || !Theta || !S) RB_FAIL(RB_ERR_ARG, "rb_supremizers: bad arguments");
  if (!U && (!ctx->dU || ctx->B != B || ctx->N != Nu))
    RB_FAIL(RB_ERR_STATE, "rb_supremizers: no resident solutions of a sweep with B parameters and N == Nu");
  if ((size_t)SUP_PB * (Q + Nu) * sizeof(double) > 200 * 1024) RB_FAIL(RB_ERR_ARG, "rb_supremizers: Q + Nu too large");
  cudaStream_t st = ctx->stream;
  std::vector<double> Wt((size_t)Q * Nu * Ns);
  for (int q = 0; q < Q; ++q)
    for (int i = 0; i < Ns; ++i)
      for (int j = 0; j < Nu; ++j) Wt[((size_t)q * Nu + j) * Ns + i] = W[((size_t)q * Ns + i) * Nu + j];
  double *dW = nullptr, *dT = nullptr, *dUu = nullptr, *dS = nullptr;
  cudaError_t e = cudaMalloc(&dW, Wt.size() * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dT, (size_t)B * Q * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dS, (size_t)B * Ns * sizeof(double));
  if (e == cudaSuccess && U) e = cudaMalloc(&dUu, (size_t)B * Nu * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpyAsync(dW, Wt.data(), Wt.size() * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dT, Theta, (size_t)B * Q * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && U) e = cudaMemcpyAsync(dUu, U, (size_t)B * Nu * sizeof(double), cudaMemcpyHostToDevice, st);
  const size_t smem = (size_t)SUP_PB * (Q + Nu) * sizeof(double);
  if (e == cudaSuccess && smem > 48 * 1024)
    e = cudaFuncSetAttribute(rb_supremizer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    cudaEventRecord(ctx->ev[4], st);
    rb_supremizer_kernel<<<(unsigned)((B + SUP_PB - 1) / SUP_PB), 128, smem, st>>>(
        dW, Ns, Nu, Q, dT, U ? dUu : ctx->dU, U ? Nu : ctx->ldu, B, dS);
    e = cudaGetLastError();
    cudaEventRecord(ctx->ev[5], st);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(S, dS, (size_t)B * Ns * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  for (double* q : {dW, dT, dUu, dS})
    if (q) cudaFree(q);
  if (e != cudaSuccess) RB_FAIL(RB_ERR_CUDA, std::string("rb_supremizers: ") + cudaGetErrorString(e));
  float a = 0;
  cudaEventElapsedTime(&a, ctx->ev[4], ctx->ev[5]);
  ctx->last_ms[0] = a;
  ctx->last_ms[4] = a;
  ctx->last_launches = 1;
  return RB_OK;
}

extern "C" int rb_result_device_ptrs(rb_ctx* ctx, void** maxloc16, double** delta) {
  if (!ctx) return RB_ERR_ARG;
  if (maxloc16) *maxloc16 = ctx->dResult;
  if (delta) *delta = ctx->dDelta;
  return RB_OK;
}

extern "C" int rb_estimator_work(rb_ctx* ctx, int64_t* issued_macs_per_param, int* n_column_tiles) {
  if (!ctx) return RB_ERR_ARG;
  if (!ctx->have_est) RB_FAIL(RB_ERR_STATE, "rb_estimator_work: no estimator set");
  if (issued_macs_per_param) *issued_macs_per_param = ctx->pair.ok ? ctx->pair.issued_macs : ctx->est_issued_macs;
  if (n_column_tiles) *n_column_tiles = ctx->pair.ok ? ctx->pair.n_tile : ctx->n_ctile;
  return RB_OK;
}

extern "C" int rb_last_timings(rb_ctx* ctx, float* ms5, int* n_launches) {
  if (!ctx) return RB_ERR_ARG;
  if (ms5) memcpy(ms5, ctx->last_ms, sizeof(ctx->last_ms));
  if (n_launches) *n_launches = ctx->last_launches;
  return RB_OK;
}

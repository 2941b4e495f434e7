// rb_offline.cu -- offline basis linear algebra on one B200: C = S^T (X S2) for the POD correlation matrix
// (rbnics/backends/basic/transpose.py:304-313 called from proper_orthogonal_decomposition_base.py:44-47), the
// Galerkin projections Z^T A_q Z / Z^T f_q (transpose.py:441-468, 365-400; parametrized_reduced_differential_
// problem.py:727-790) and the Riesz Gram blocks R_q^T X R_q' (rb_reduced_problem.py:238-258).
//
// The reference does Ns sparse mat-vecs and Ns^2 PETSc VecDot calls (each an MPI all-reduce).  Here:
//   K_fused : C += S^T (X S2) for <= 64 columns, Y = X S2 only ever exists as 16-row blocks in shared memory
//             (warp-specialised gather + DMMA CTA)                                                -- one pass over S
//   K_spmm  : Y = X S2        CSR x dense (dof-major, so every non-zero reads one contiguous row of S2)   -- HBM / L2
//   K_tsmm2 : C += S^T Y      tall-skinny FP64 GEMM on DMMA.8x8x4, 128x128 / 64x64 tiles fed by 1-D bulk copies,
//             split over dof rows, symmetric tiles mirrored when X == X^T         (K_tsmm: cp.async fallback for
//             operands whose rows are not 16-byte aligned)
//   K_red   : fixed-order sum of the split partials (deterministic)
// Rows [row_begin, row_end) select this rank's dof partition; the host layer all-reduces the Ns x Ns2 partials.
// Note on bulk copies: cp.async.bulk takes its operands from uniform registers, so a copy per LANE is issued one
// lane at a time (~40 clk each) -- fine for 16-64 row copies per stage next to thousands of DMMA cycles, but a
// Z = S V kernel built on 160 small row copies per stage measured 2x SLOWER than the cp.async one below (kept).
#include <algorithm>

#include <cuda.h>  // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no link against libcuda)

#include "rb_common.cuh"

constexpr int TS_TM = 64, TS_TN = 64, TS_KC = 32, TS_THREADS = 128, TS_STAGES = 3;
constexpr int TS_PITCH = TS_TM + 4;  // == 4 (mod 16): conflict-free DMMA fragment loads

// ---------------------------------------------------------------------------------------------------------------------
// Y[r, :] = sum_k data[k] * S2[indices[k], :]  for r in [row_begin, row_end); one warp per (row, 128-column chunk)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rb_spmm_kernel(const long long* __restrict__ indptr,
                                                      const int* __restrict__ indices,
                                                      const double* __restrict__ data, const double* __restrict__ S2,
                                                      int ld2, int ns2, double* __restrict__ Y, int ldy,
                                                      long long row_begin, long long row_end) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int nchunk = (ns2 + 127) / 128;
  const long long r = row_begin + warp / nchunk;
  const int c0 = (int)(warp % nchunk) * 128;
  if (r >= row_end) return;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const long long k0 = indptr[r], k1 = indptr[r + 1];
  for (long long k = k0; k < k1; ++k) {
    const double v = __ldg(data + k);
    const double* row = S2 + (long long)__ldg(indices + k) * ld2 + c0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int c = lane + 32 * t;
      if (c0 + c < ns2) acc[t] = fma(v, __ldg(row + c), acc[t]);
    }
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int c = lane + 32 * t;
    if (c0 + c < ns2) Y[r * ldy + c0 + c] = acc[t];
  }
}

// Y[r, c] = sum_k data[k] * S[indices[k], c] for a FEW columns (n <= 32): one thread per (dof row, column), rows
// [row_begin, row_end) -- the one-warp-per-row kernel above idles most of its lanes and pays its latency chain per row.
__global__ void __launch_bounds__(256) rb_spmm_narrow_kernel(const long long* __restrict__ indptr,
                                                             const int* __restrict__ indices,
                                                             const double* __restrict__ data,
                                                             const double* __restrict__ S, int lds, int n,
                                                             double* __restrict__ Y, int ldy, long long row_begin,
                                                             long long row_end) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long r = row_begin + t / n;
  const int c = (int)(t % n);
  if (r >= row_end) return;
  double acc = 0.0;
  const long long k1 = __ldg(indptr + r + 1);
  for (long long k = __ldg(indptr + r); k < k1; ++k)
    acc = fma(__ldg(data + k), __ldg(S + (long long)__ldg(indices + k) * lds + c), acc);
  Y[r * ldy + c] = acc;
}

// Same with one thread per (dof row, PAIR of columns) for 33..128 columns of 16-byte aligned operands (even pitches).
__global__ void __launch_bounds__(256) rb_spmm_pair_kernel(const long long* __restrict__ indptr,
                                                           const int* __restrict__ indices,
                                                           const double* __restrict__ data,
                                                           const double* __restrict__ S, int lds, int npair,
                                                           double* __restrict__ Y, int ldy, long long row_begin,
                                                           long long row_end) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long r = row_begin + t / npair;
  const int c = 2 * (int)(t % npair);
  if (r >= row_end) return;
  double2 acc = make_double2(0.0, 0.0);
  const long long k1 = __ldg(indptr + r + 1);
  for (long long k = __ldg(indptr + r); k < k1; ++k) {
    const double v = __ldg(data + k);
    const double2 x = __ldg(reinterpret_cast<const double2*>(S + (long long)__ldg(indices + k) * lds + c));
    acc.x = fma(v, x.x, acc.x);
    acc.y = fma(v, x.y, acc.y);
  }
  *reinterpret_cast<double2*>(Y + r * ldy + c) = acc;
}

// ---------------------------------------------------------------------------------------------------------------------
// Partial C tile (64 x 64) over a contiguous range of dof rows: ws[split][i][j] = sum_r S[r][i] * Y[r][j]
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(void* dst, const void* src, bool valid) {
  const uint32_t d = smem_u32(dst);
  const int sz = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(TS_THREADS) rb_tsmm_kernel(const double* __restrict__ S, int lds, int ns,
                                                             const double* __restrict__ Y, int ldy, int ns2,
                                                             long long row_begin, long long row_end,
                                                             long long rows_per_split, double* __restrict__ ws,
                                                             int ldw) {
  extern __shared__ __align__(16) double ts_smem[];
  double* As = ts_smem;                                   // [STAGES][KC][PITCH]  (rows of S, columns i0..i0+63)
  double* Bs = ts_smem + TS_STAGES * TS_KC * TS_PITCH;    // [STAGES][KC][PITCH]  (rows of Y, columns j0..j0+63)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;
  const int i0 = blockIdx.x * TS_TM, j0 = blockIdx.y * TS_TN;
  const int split = blockIdx.z;
  const long long rs = row_begin + split * rows_per_split;
  const long long re = min(row_end, rs + rows_per_split);
  const int nchunks = (int)((re - rs + TS_KC - 1) / TS_KC);

  auto load_stage = [&](int stage, int chunk) {
    const long long r0 = rs + (long long)chunk * TS_KC;
    double* as = As + stage * TS_KC * TS_PITCH;
    double* bs = Bs + stage * TS_KC * TS_PITCH;
    for (int t = tid; t < TS_KC * TS_TM; t += TS_THREADS) {
      const int rr = t / TS_TM, cc = t % TS_TM;
      const long long r = r0 + rr;
      const bool rv = r < re;
      const bool va = rv && (i0 + cc < ns);
      const bool vb = rv && (j0 + cc < ns2);
      cp_async8(as + rr * TS_PITCH + cc, va ? (S + r * lds + i0 + cc) : S, va);
      cp_async8(bs + rr * TS_PITCH + cc, vb ? (Y + r * ldy + j0 + cc) : Y, vb);
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

  for (int s = 0; s < TS_STAGES - 1; ++s) {
    if (s < nchunks) load_stage(s, s);
    cp_async_commit();
  }
  const int g = lane >> 2, q4 = lane & 3;
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<TS_STAGES - 2>();
    __syncthreads();
    {
      const int nxt = c + TS_STAGES - 1;
      if (nxt < nchunks) load_stage(nxt % TS_STAGES, nxt);
      cp_async_commit();
    }
    const double* as = As + (c % TS_STAGES) * TS_KC * TS_PITCH;
    const double* bs = Bs + (c % TS_STAGES) * TS_KC * TS_PITCH;
#pragma unroll
    for (int ks = 0; ks < TS_KC / 4; ++ks) {
      double af[4], bf[4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) af[mt] = as[(4 * ks + q4) * TS_PITCH + wm * 32 + mt * 8 + g];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bf[nt] = bs[(4 * ks + q4) * TS_PITCH + wn * 32 + nt * 8 + g];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
    }
  }
  cp_async_wait<0>();
  double* w = ws + (long long)split * ldw * ldw;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const int i = i0 + wm * 32 + mt * 8 + g;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int j = j0 + wn * 32 + nt * 8 + q4 * 2;
      *reinterpret_cast<double2*>(w + (long long)i * ldw + j) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Second generation of the same product for 16-byte aligned operands (even pitches): TM x TM tile per CTA
// ((TM/32)^2 warps of 32 x 32), rows of S and Y brought in by 1-D bulk copies (TMA engine) into a 3-stage ring of
// 32-row chunks with padded rows (pitch == 4 mod 16: conflict-free fragment loads); the warp that is last to release
// a stage refills it, so there is no producer warp.  Warps whose 32 x 32 tile lies outside ns x ns2 retire at once.
// Per DMMA k-step a 128 x 128 tile moves 8 KB from L2 for 256 DMMAs (a 64 x 64 tile: 4 KB for 64).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int T2_KC = 32, T2_NST = 3;

__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// One 32-row chunk of a warp's MT x NT sub-tiles (8 x 8 each); every (MT, NT) is its own fully unrolled instantiation,
// chosen once per warp, so there are no predicates inside the DMMA stream.
template <int MT, int NT>
__device__ __forceinline__ void t2_chunk(double (&acc)[4][4][2], const double* __restrict__ as,
                                         const double* __restrict__ bs, int PA, int PB) {
#pragma unroll
  for (int ks = 0; ks < T2_KC / 4; ++ks) {
    double af[MT], bf[NT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) af[mt] = as[4 * ks * PA + mt * 8];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) bf[nt] = bs[4 * ks * PB + nt * 8];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
  }
}

template <int MT>
__device__ __forceinline__ void t2_chunk_nt(int nt_w, double (&acc)[4][4][2], const double* as, const double* bs, int PA,
                                            int PB) {
  switch (nt_w) {
    case 1: t2_chunk<MT, 1>(acc, as, bs, PA, PB); break;
    case 2: t2_chunk<MT, 2>(acc, as, bs, PA, PB); break;
    case 3: t2_chunk<MT, 3>(acc, as, bs, PA, PB); break;
    default: t2_chunk<MT, 4>(acc, as, bs, PA, PB); break;
  }
}

__device__ __forceinline__ void t2_chunk_dispatch(int mt_w, int nt_w, double (&acc)[4][4][2], const double* as,
                                                  const double* bs, int PA, int PB) {
  switch (mt_w) {
    case 1: t2_chunk_nt<1>(nt_w, acc, as, bs, PA, PB); break;
    case 2: t2_chunk_nt<2>(nt_w, acc, as, bs, PA, PB); break;
    case 3: t2_chunk_nt<3>(nt_w, acc, as, bs, PA, PB); break;
    default: t2_chunk_nt<4>(nt_w, acc, as, bs, PA, PB); break;
  }
}

template <int TM>
__global__ void __launch_bounds__(32 * (TM / 32) * (TM / 32), TM == 64 ? 2 : 1)
    rb_tsmm2_kernel(const double* __restrict__ S, int lds, int ns, const double* __restrict__ Y, int ldy, int ns2,
                    long long row_begin, long long row_end, long long rows_per_split, double* __restrict__ ws, int ldw,
                    int sym, int use_tmap, const __grid_constant__ CUtensorMap tmS,
                    const __grid_constant__ CUtensorMap tmY) {
  constexpr int WM = TM / 32, P = TM + 4;
  if (sym && blockIdx.y < blockIdx.x) return;  // C == C^T: tiles below the diagonal are mirrored by the reduction
  extern __shared__ __align__(128) unsigned char t2_smem[];
  double* As = reinterpret_cast<double*>(t2_smem);               // [NST][KC][P]
  double* Bs = As + T2_NST * T2_KC * P;                          // [NST][KC][P]
  uint64_t* full = reinterpret_cast<uint64_t*>(Bs + T2_NST * T2_KC * P);
  int* cnt = reinterpret_cast<int*>(full + T2_NST);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int wm = warp / WM, wn = (warp + wm) % WM;  // rotated so that a single active row or column of warps still
                                              // spreads over the four SM sub-partitions
  // rows of the tile grid are TM wide (the last one narrower); its columns deal the ceil(ns2/32) warp columns out
  // evenly (4,3,3,3 rather than 4,4,4,1), which keeps am*an a multiple of 4 in most tiles -- a CTA takes as long as
  // its busiest SM sub-partition.  The symmetric variant needs the same cut in both directions.
  const int i0 = blockIdx.x * TM;
  const int am = min(WM, (ns - i0 + 31) / 32);
  int j0, an;
  if (sym) {
    j0 = blockIdx.y * TM;
    an = min(WM, (ns2 - j0 + 31) / 32);
  } else {
    const int Wb = (ns2 + 31) / 32, Tb = gridDim.y, bb = Wb / Tb, rbm = Wb % Tb, tb = blockIdx.y;
    an = bb + (tb < rbm);
    j0 = 32 * (tb * bb + min(tb, rbm));
  }
  const int split = blockIdx.z;
  const long long rs = row_begin + split * rows_per_split;
  const long long re = min(row_end, rs + rows_per_split);
  const int nchunks = (int)((re - rs + T2_KC - 1) / T2_KC);
  // a diagonal tile of a symmetric product needs its warp tiles on and above the diagonal only (the reduction mirrors at
  // warp-tile granularity): they are dealt to warps 0, 1, 2, ... row by row, so that the 10 tiles of a full 4 x 4 tile
  // load the four SM sub-partitions 3, 3, 2, 2 instead of 4, 3, 2, 1
  const bool diag = sym && blockIdx.x == blockIdx.y;
  if (diag) {
    int k = warp;
    wm = 0;
    while (wm < am && k >= an - wm) {
      k -= an - wm;
      ++wm;
    }
    wn = wm + k;
  }
  const int nactive = diag ? an * (an + 1) / 2 : am * an;
  const int wa = min(am * 32, ((ns + 1) & ~1) - i0), wb = min(an * 32, ((ns2 + 1) & ~1) - j0);  // columns copied per row
  if (tid == 0) {
    for (int s = 0; s < T2_NST; ++s) {
      mbar_init(full + s, 1);
      cnt[s] = 0;
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (wm >= am || wn >= an) return;

  auto issue = [&](int st, int c) {  // executed by one whole warp: lane l brings row l of the chunk
    const long long r0 = rs + (long long)c * T2_KC;
    if (use_tmap) {
      // 2-D tensor maps (box = 32 rows x (TM + 4) columns): ONE TMA instruction per operand and stage.  The box is 4
      // columns wider than the tile on purpose -- its dense rows in shared memory are then exactly the padded pitch the
      // fragment loads want; columns past the operand and rows past row_end are filled with zeros by the TMA unit.
      if (lane == 0) {
        mbar_arrive_expect_tx(full + st, 2u * (uint32_t)(T2_KC * P * 8));
        tma_load_2d(As + st * T2_KC * P, &tmS, i0, (int)r0, full + st);
        tma_load_2d(Bs + st * T2_KC * P, &tmY, j0, (int)r0, full + st);
      }
      return;
    }
    const int nrows = (int)min((long long)T2_KC, re - r0);
    double* as = As + (st * T2_KC + lane) * P;
    double* bs = Bs + (st * T2_KC + lane) * P;
    if (lane >= nrows) {  // rows past the end of the split count as zeros
      for (int c2 = 0; c2 < TM; ++c2) as[c2] = bs[c2] = 0.0;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_expect_tx(full + st, (uint32_t)nrows * (uint32_t)(wa + wb) * 8u);
    __syncwarp();
    if (lane < nrows) {
      bulk_g2s(as, S + (r0 + lane) * lds + i0, (uint32_t)wa * 8u, full + st);
      bulk_g2s(bs, Y + (r0 + lane) * ldy + j0, (uint32_t)wb * 8u, full + st);
    }
  };
  if (warp == 0) {
    for (int s = 0; s < T2_NST && s < nchunks; ++s) issue(s, s);
  }

  double acc[4][4][2];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  const int g = lane >> 2, q4 = lane & 3;
  const int aoff = q4 * P + wm * 32 + g, boff = q4 * P + wn * 32 + g;
  // 8-column sub-tiles of this warp that hold columns of the product (the last warp row / column of the grid may be
  // partly outside: 400 columns = 12 full warp columns + one of two sub-tiles)
  const int mt_w = min(4, (ns - (i0 + wm * 32) + 7) / 8), nt_w = min(4, (ns2 - (j0 + wn * 32) + 7) / 8);
  int st = 0;
  uint32_t phase = 0;
  for (int c = 0; c < nchunks; ++c) {
    mbar_wait(full + st, phase);
    t2_chunk_dispatch(mt_w, nt_w, acc, As + st * T2_KC * P + aoff, Bs + st * T2_KC * P + boff, P, P);
    __syncwarp();
    int old = 0;
    if (lane == 0) old = atomicAdd(&cnt[st], 1);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old == nactive - 1) {  // last warp out refills the slot
      if (lane == 0) cnt[st] = 0;
      if (c + T2_NST < nchunks) issue(st, c + T2_NST);
    }
    if (++st == T2_NST) {
      st = 0;
      phase ^= 1u;
    }
  }
  double* w = ws + (long long)split * ldw * ldw;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const int i = i0 + wm * 32 + mt * 8 + g;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int j = j0 + wn * 32 + nt * 8 + q4 * 2;
      if (mt < mt_w && nt < nt_w)
        *reinterpret_cast<double2*>(w + (long long)i * ldw + j) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Single-tile variant of rb_tsmm2_kernel (ns, ns2 <= TM) with the 8-column sub-tiles dealt evenly to the warps: a
// product with 100 columns has 13 x 13 sub-tiles, which 4 x 4 warps of fixed 32 x 32 tiles cover with 16 x 16 (61 %
// useful DMMAs, and the busiest SM sub-partition decides).  Here warp row r owns M8/WM (+1) consecutive sub-tile rows,
// warp column likewise; the k loop of every (rows, columns) count is its own fully unrolled instantiation, chosen
// once per warp, so there are no predicates inside the DMMA stream.
// ---------------------------------------------------------------------------------------------------------------------
template <int TM>
__global__ void __launch_bounds__(32 * (TM / 32) * (TM / 32), TM == 64 ? 2 : 1)
    rb_tsmm2_bal_kernel(const double* __restrict__ S, int lds, int ns, const double* __restrict__ Y, int ldy, int ns2,
                        long long row_begin, long long row_end, long long rows_per_split, double* __restrict__ ws,
                        int ldw, int contig, int use_tmap, const __grid_constant__ CUtensorMap tmS,
                        const __grid_constant__ CUtensorMap tmY) {
  constexpr int WM = TM / 32, P = TM + 4;
  // contig: both operands start at column 0 of rows no longer than a stage row, so the 32 rows of a chunk are ONE piece
  // of memory per operand -- one bulk copy each instead of 64 per-lane copies (~40 clk apiece on the refilling warp,
  // more than its DMMA time per chunk once the sub-tiles are balanced); the stage then keeps the global pitch.
  const int PA = contig ? lds : P, PB = contig ? ldy : P;
  extern __shared__ __align__(128) unsigned char t2_smem[];
  double* As = reinterpret_cast<double*>(t2_smem);               // [NST][KC][P]
  double* Bs = As + T2_NST * T2_KC * P;                          // [NST][KC][P]
  uint64_t* full = reinterpret_cast<uint64_t*>(Bs + T2_NST * T2_KC * P);
  int* cnt = reinterpret_cast<int*>(full + T2_NST);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WM, wn = (warp + wm) % WM;
  const int split = blockIdx.x;
  const long long rs = row_begin + split * rows_per_split;
  const long long re = min(row_end, rs + rows_per_split);
  const int nchunks = (int)((re - rs + T2_KC - 1) / T2_KC);
  const int M8 = (ns + 7) / 8, N8 = (ns2 + 7) / 8;  // 8-column sub-tiles (<= 4 WM each way)
  const int mt_w = M8 / WM + (wm < M8 % WM), m8 = wm * (M8 / WM) + min(wm, M8 % WM);
  const int nt_w = N8 / WM + (wn < N8 % WM), n8 = wn * (N8 / WM) + min(wn, N8 % WM);
  const int nactive = min(WM, M8) * min(WM, N8);
  const int wa = (ns + 1) & ~1, wb = (ns2 + 1) & ~1;  // columns copied per row
  if (tid == 0) {
    for (int s = 0; s < T2_NST; ++s) {
      mbar_init(full + s, 1);
      cnt[s] = 0;
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (mt_w == 0 || nt_w == 0) return;

  auto issue = [&](int st, int c) {  // executed by one whole warp: lane l brings row l of the chunk
    const long long r0 = rs + (long long)c * T2_KC;
    if (!contig && use_tmap) {  // column sub-range of a wider block: one 2-D TMA box per operand (see rb_tsmm2_kernel)
      if (lane == 0) {
        mbar_arrive_expect_tx(full + st, 2u * (uint32_t)(T2_KC * P * 8));
        tma_load_2d(As + st * T2_KC * P, &tmS, 0, (int)r0, full + st);
        tma_load_2d(Bs + st * T2_KC * P, &tmY, 0, (int)r0, full + st);
      }
      return;
    }
    const int nrows = (int)min((long long)T2_KC, re - r0);
    double* as = As + st * T2_KC * P + lane * PA;
    double* bs = Bs + st * T2_KC * P + lane * PB;
    if (lane >= nrows) {  // rows past the end of the split count as zeros
      for (int c2 = 0; c2 < PA; ++c2) as[c2] = 0.0;
      for (int c2 = 0; c2 < PB; ++c2) bs[c2] = 0.0;
    }
    __syncwarp();
    if (contig) {
      if (lane == 0) {
        const uint32_t ba = (uint32_t)nrows * (uint32_t)lds * 8u, bb = (uint32_t)nrows * (uint32_t)ldy * 8u;
        mbar_arrive_expect_tx(full + st, ba + bb);
        bulk_g2s(As + st * T2_KC * P, S + r0 * lds, ba, full + st);
        bulk_g2s(Bs + st * T2_KC * P, Y + r0 * ldy, bb, full + st);
      }
      return;
    }
    if (lane == 0) mbar_arrive_expect_tx(full + st, (uint32_t)nrows * (uint32_t)(wa + wb) * 8u);
    __syncwarp();
    if (lane < nrows) {
      bulk_g2s(as, S + (r0 + lane) * lds, (uint32_t)wa * 8u, full + st);
      bulk_g2s(bs, Y + (r0 + lane) * ldy, (uint32_t)wb * 8u, full + st);
    }
  };
  if (warp == 0) {  // (warp 0 owns sub-tile row 0 and column 0: always active)
    for (int s = 0; s < T2_NST && s < nchunks; ++s) issue(s, s);
  }

  double acc[4][4][2];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  const int g = lane >> 2, q4 = lane & 3;
  const int aoff = q4 * PA + m8 * 8 + g, boff = q4 * PB + n8 * 8 + g;
  int st = 0;
  uint32_t phase = 0;
  for (int c = 0; c < nchunks; ++c) {
    mbar_wait(full + st, phase);
    const double* as = As + st * T2_KC * P + aoff;
    const double* bs = Bs + st * T2_KC * P + boff;
    switch (mt_w) {
      case 1: t2_chunk_nt<1>(nt_w, acc, as, bs, PA, PB); break;
      case 2: t2_chunk_nt<2>(nt_w, acc, as, bs, PA, PB); break;
      case 3: t2_chunk_nt<3>(nt_w, acc, as, bs, PA, PB); break;
      default: t2_chunk_nt<4>(nt_w, acc, as, bs, PA, PB); break;
    }
    __syncwarp();
    int old = 0;
    if (lane == 0) old = atomicAdd(&cnt[st], 1);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old == nactive - 1) {  // last warp out refills the slot
      if (lane == 0) cnt[st] = 0;
      if (c + T2_NST < nchunks) issue(st, c + T2_NST);
    }
    if (++st == T2_NST) {
      st = 0;
      phase ^= 1u;
    }
  }
  double* w = ws + (long long)split * ldw * ldw;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const int i = (m8 + mt) * 8 + g;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int j = (n8 + nt) * 8 + q4 * 2;
      if (mt < mt_w && nt < nt_w)
        *reinterpret_cast<double2*>(w + (long long)i * ldw + j) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Fused Gram for ns, ns2 <= 64: C += S^T (X S2) over a range of dof rows WITHOUT writing Y = X S2 to HBM
// (algorithmic bytes 12 nnz + 8 Nh Ns).  One warp-specialised CTA of 28 warps per SM; registers move
// from the gather warps (64) to the DMMA warps (120) with setmaxnreg -- the sum has to fit the CTA's launch
// allocation of 28 x 72, an inc beyond what the decs released blocks for ever:
//   24 gather warps, 6 teams of 4: team t owns the 16-row chunks c == t (mod 6).  Its first warp starts the bulk
//                    copies of the chunk's S rows (A operand); every warp builds 4 rows of the 16 x ns2 block of Y in
//                    shared memory with one pair of columns per lane.  The CSR entries of its four rows sit one per
//                    lane (prefetched one chunk ahead, the row pointers two chunks ahead); a ballot splits them into
//                    entries whose column lies inside the chunk -- those read the S rows the bulk copy is bringing
//                    in anyway (when S2 is S) -- and the others, whose rows come from L2 four loads at a time.
//                    Two FP64 FMAs per entry and lane: the FP64 pipe is shared with the DMMA warps.
//    4 DMMA warps  : 32 x 32 tile each of C += A^T B per stage, release the stage.
// Eight stages: six in production, one being consumed, one slack.  Measured at Nh = 10^6, Ns = 61, 5 nnz/row:
// 0.49 ms (SpMM 0.49 ms + tall-skinny product 0.36 ms unfused); 0.34 ms with the gather switched off, i.e. the
// per-chunk barrier hand-over of the DMMA warps, not the gather, is what is left above the 0.22 ms DMMA time.
// 16-row chunks are dealt round-robin to the CTAs: at any moment the grid works on one moving window of
// gridDim.x * 16 consecutive dof rows, so the matrix columns outside a chunk (neighbouring mesh rows) are rows that
// some other CTA has just brought through L2 -- with one contiguous row range per CTA they came from HBM (measured:
// 1.59 GB DRAM reads against 0.55 GB algorithmic; 0.56 GB with the round-robin deal).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FG_KC = 16, FG_NST = 8, FG_NWD = 4, FG_NKG = FG_NWD / 4, FG_TEAMS = 6, FG_TW = FG_KC / 4,
              FG_NWG = FG_TW * FG_TEAMS, FG_P = 68;  // FG_TW warps per team, 4 rows each
constexpr int FG_THREADS = 32 * (FG_NWD + FG_NWG);

__global__ void __launch_bounds__(FG_THREADS, 1)
    rb_gram_fused_kernel(const double* __restrict__ S, int lds, int ns, const double* __restrict__ S2, int ld2, int ns2,
                         int reuse, const long long* __restrict__ ip, const int* __restrict__ ix,
                         const double* __restrict__ dt, long long row_begin, long long row_end,
                         double* __restrict__ ws, int ldw, int use_tmap, const __grid_constant__ CUtensorMap tmS) {
  extern __shared__ __align__(128) unsigned char fg_smem[];
  double* As = reinterpret_cast<double*>(fg_smem);               // [NST][KC][P]
  double* Bs = As + FG_NST * FG_KC * FG_P;                       // [NST][KC][P]
  uint64_t* fullA = reinterpret_cast<uint64_t*>(Bs + FG_NST * FG_KC * FG_P);
  uint64_t* fullB = fullA + FG_NST;
  uint64_t* empty = fullB + FG_NST;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long CS = (long long)gridDim.x * FG_KC;               // row stride between consecutive chunks of a CTA
  const long long rs = row_begin + (long long)blockIdx.x * FG_KC;  // first row of chunk 0 of this CTA
  const long long re = row_end;
  const long long tot_chunks = (row_end - row_begin + FG_KC - 1) / FG_KC;
  const int nchunks = tot_chunks > blockIdx.x ? (int)((tot_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  const int wa = (ns + 1) & ~1, wb = (ns2 + 1) & ~1;
  if (tid == 0) {
    for (int s = 0; s < FG_NST; ++s) {
      mbar_init(fullA + s, 1);
      mbar_init(fullB + s, FG_TW);
      mbar_init(empty + s, FG_NWD);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp < FG_NWD) {
    // ---------------------------------------------------------------- DMMA warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;\n");
    const int kgrp = warp >> 2, tw = warp & 3;
    const int wm = tw & 1, wn = tw >> 1;
    double acc[4][4][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    const int g = lane >> 2, q4 = lane & 3;
    const int aoff = q4 * FG_P + wm * 32 + g, boff = q4 * FG_P + wn * 32 + g;
    int st = 0;
    uint32_t phase = 0;
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait(fullA + st, phase);
      mbar_wait(fullB + st, phase);
      const double* as = As + st * FG_KC * FG_P + aoff;
      const double* bs = Bs + st * FG_KC * FG_P + boff;
#pragma unroll
      for (int kk = 0; kk < FG_KC / 4 / FG_NKG; ++kk) {
        const int ks = FG_NKG * kk + kgrp;
        double af[4], bf[4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) af[mt] = as[4 * ks * FG_P + mt * 8];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) bf[nt] = bs[4 * ks * FG_P + nt * 8];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + st);
      if (++st == FG_NST) {
        st = 0;
        phase ^= 1u;
      }
    }
    double* w = ws + ((long long)blockIdx.x * FG_NKG + kgrp) * ldw * ldw;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int i = wm * 32 + mt * 8 + g;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int j = wn * 32 + nt * 8 + q4 * 2;
        *reinterpret_cast<double2*>(w + (long long)i * ldw + j) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
      }
    }
    return;
  }

  // ------------------------------------------------------------------ gather warps
  asm volatile("setmaxnreg.dec.sync.aligned.u32 64;\n");
  const int team = (warp - FG_NWD) / FG_TW, u = (warp - FG_NWD) % FG_TW;  // this warp builds rows 4u .. 4u+3 of a chunk
  const int cp = min(2 * lane, wb - 2);                            // this lane's pair of columns (clamped)
  // rows of the warp in chunk c: [r0 + 4u, r0 + 4u + 4) cut at row_end; lane l <= 4 holds indptr of row 4u + l
  auto first_row = [&](int c) { return rs + (long long)c * CS + 4 * u; };
  auto load_ptr = [&](int c) -> long long {
    if (c >= nchunks) return 0;
    const long long r = min(first_row(c) + min(lane, 4), re);
    return __ldg(ip + r);
  };
  auto load_entries = [&](long long p, int& col, double& val) {  // lane l: entry l of the warp's four rows
    const long long base = __shfl_sync(0xffffffffu, p, 0), end = __shfl_sync(0xffffffffu, p, 4);
    col = 0;
    val = 0.0;
    if (base + lane < end) {
      col = __ldg(ix + base + lane);
      val = __ldg(dt + base + lane);
    }
  };
  int c = team;
  long long p_cur = load_ptr(c), p_nxt = load_ptr(c + FG_TEAMS);
  int col_cur = 0, col_nxt = 0;
  double val_cur = 0.0, val_nxt = 0.0;
  load_entries(p_cur, col_cur, val_cur);
  for (; c < nchunks; c += FG_TEAMS) {
    const long long p_nn = load_ptr(c + 2 * FG_TEAMS);
    load_entries(p_nxt, col_nxt, val_nxt);
    const int st = c % FG_NST, n = c / FG_NST;
    if (n > 0) mbar_wait(empty + st, (uint32_t)(n - 1) & 1u);
    const long long r0 = rs + (long long)c * CS;
    const int nrows = (int)min((long long)FG_KC, re - r0);
    if (u == 0 && use_tmap) {  // the chunk's S rows as ONE 2-D TMA box (16 rows x FG_P columns, zero filled outside)
      if (lane == 0) {
        mbar_arrive_expect_tx(fullA + st, (uint32_t)(FG_KC * FG_P * 8));
        tma_load_2d(As + st * FG_KC * FG_P, &tmS, 0, (int)r0, fullA + st);
      }
    } else if (u == 0) {
      double* as = As + (st * FG_KC + lane) * FG_P;
      if (lane < FG_KC && lane >= nrows) {  // rows past the end count as zeros
        for (int c2 = 0; c2 < 64; ++c2) as[c2] = 0.0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(fullA + st, (uint32_t)nrows * (uint32_t)wa * 8u);
      __syncwarp();
      if (lane < nrows) bulk_g2s(as, S + (r0 + lane) * lds, (uint32_t)wa * 8u, fullA + st);
    }
    const long long base = __shfl_sync(0xffffffffu, p_cur, 0);
    const int r0i = (int)r0;  // CSR column indices are 32-bit, so dof rows are too
    const double* a_stage = As + st * FG_KC * FG_P + cp;
    double2 acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = make_double2(0.0, 0.0);
    int kb1, kb2, kb3, cnt;  // entry offsets (relative to base) at which rows 1, 2, 3 of the warp start; total count
    kb1 = (int)(__shfl_sync(0xffffffffu, p_cur, 1) - base);
    kb2 = (int)(__shfl_sync(0xffffffffu, p_cur, 2) - base);
    kb3 = (int)(__shfl_sync(0xffffffffu, p_cur, 3) - base);
    cnt = (int)(__shfl_sync(0xffffffffu, p_cur, 4) - base);
    const bool mine = lane < cnt;                                             // this lane holds an entry
    const bool in_chunk = reuse && (unsigned)(col_cur - r0i) < (unsigned)nrows;
    auto add = [&](int e, double v, const double2& x) {  // entry e belongs to row (e>=kb1)+(e>=kb2)+(e>=kb3)
      // e is warp-uniform: plain branches, two FMAs per entry (the FP64 pipe is shared with the DMMA warps)
      if (e < kb2) {
        if (e < kb1) {
          acc[0].x = fma(v, x.x, acc[0].x);
          acc[0].y = fma(v, x.y, acc[0].y);
        } else {
          acc[1].x = fma(v, x.x, acc[1].x);
          acc[1].y = fma(v, x.y, acc[1].y);
        }
      } else {
        if (e < kb3) {
          acc[2].x = fma(v, x.x, acc[2].x);
          acc[2].y = fma(v, x.y, acc[2].y);
        } else {
          acc[3].x = fma(v, x.x, acc[3].x);
          acc[3].y = fma(v, x.y, acc[3].y);
        }
      }
    };
    // entries whose column lies outside the chunk (or all of them if S2 is not S): rows of S2 from L2 / HBM.  The
    // entries sit one per lane; a ballot marks the far ones and the warp walks the set bits four at a time, so the
    // loads of four entries (of any of the four rows) are in flight together.
    unsigned farmask = __ballot_sync(0xffffffffu, mine && !in_chunk);
    while (farmask) {
      double2 x[4];
      double v[4];
      int eb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        x[j] = make_double2(0.0, 0.0);
        v[j] = 0.0;
        eb[j] = 0;
        if (farmask) {
          eb[j] = __ffs(farmask) - 1;
          farmask &= farmask - 1;
          const int col = __shfl_sync(0xffffffffu, col_cur, eb[j]);
          v[j] = __shfl_sync(0xffffffffu, val_cur, eb[j]);
          x[j] = __ldg(reinterpret_cast<const double2*>(S2 + (long long)col * ld2 + cp));
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) add(eb[j], v[j], x[j]);
    }
    for (int e = 32; e < cnt; ++e) {  // (rows with so many entries that four of them exceed one entry per lane)
      const int col = __ldg(ix + base + e);
      if (!(reuse && (unsigned)(col - r0i) < (unsigned)nrows))
        add(e, __ldg(dt + base + e), __ldg(reinterpret_cast<const double2*>(S2 + (long long)col * ld2 + cp)));
    }
    if (reuse) {
      // entries inside the chunk: the S rows are (being) staged in shared memory as the A operand of this stage
      mbar_wait(fullA + st, (uint32_t)n & 1u);
      unsigned nearmask = __ballot_sync(0xffffffffu, mine && in_chunk);
      while (nearmask) {
        const int e = __ffs(nearmask) - 1;
        nearmask &= nearmask - 1;
        const int d = __shfl_sync(0xffffffffu, col_cur, e) - r0i;
        const double v = __shfl_sync(0xffffffffu, val_cur, e);
        add(e, v, *reinterpret_cast<const double2*>(a_stage + d * FG_P));
      }
      for (int e = 32; e < cnt; ++e) {
        const int d = __ldg(ix + base + e) - r0i;
        if ((unsigned)d < (unsigned)nrows) add(e, __ldg(dt + base + e), *reinterpret_cast<const double2*>(a_stage + d * FG_P));
      }
    }
    double* dst = Bs + (st * FG_KC + 4 * u) * FG_P + 2 * lane;
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<double2*>(dst + q * FG_P) = acc[q];
    __syncwarp();
    if (lane == 0) mbar_arrive(fullB + st);
    p_cur = p_nxt;
    p_nxt = p_nn;
    col_cur = col_nxt;
    val_cur = val_nxt;
  }
}

// C[i][j] = sum over splits, in a fixed order that depends on nsplit only (deterministic): four quarter sums of
// consecutive splits, then their sum.  256 threads = 64 entries x 4 quarters.
__global__ void __launch_bounds__(256) rb_split_reduce_kernel(const double* __restrict__ ws, int ldw, int nsplit,
                                                              int ns, int ns2, double* __restrict__ C,
                                                              int sym_tm) {
  __shared__ double part[4][64];
  const int e = blockIdx.x * 64 + (threadIdx.x & 63), quarter = threadIdx.x >> 6;
  const int per = (nsplit + 3) / 4;
  const int k0 = quarter * per, k1 = min(nsplit, k0 + per);
  double s = 0.0;
  if (e < ns * ns2) {
    int i = e / ns2, j = e % ns2;
    if (sym_tm && j / sym_tm < i / sym_tm) {  // (warp) tile below the diagonal: take the mirrored entry
      const int t = i;
      i = j;
      j = t;
    }
    const double* w = ws + (long long)i * ldw + j;
    const long long stride = (long long)ldw * ldw;
    int k = k0;
    for (; k + 8 <= k1; k += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(w + (k + u) * stride);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; k < k1; ++k) s += __ldg(w + k * stride);
  }
  part[quarter][threadIdx.x & 63] = s;
  __syncthreads();
  if (quarter == 0 && e < ns * ns2) C[e] = ((part[0][threadIdx.x] + part[1][threadIdx.x]) + part[2][threadIdx.x]) + part[3][threadIdx.x];
}

// ---------------------------------------------------------------------------------------------------------------------
// Host
// ---------------------------------------------------------------------------------------------------------------------
__global__ void rb_csr_symmetric_kernel(const long long* __restrict__ ip, const int* __restrict__ ix,
                                        const double* __restrict__ dt, long long Nh, int* __restrict__ flag) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= Nh) return;
  for (long long k = ip[r]; k < ip[r + 1]; ++k) {
    const int c = ix[k];
    if (c == r) continue;
    bool found = false;
    if (c >= 0 && c < Nh) {
      for (long long k2 = ip[c]; k2 < ip[c + 1]; ++k2)
        if (ix[k2] == r) {
          found = dt[k2] == dt[k];
          break;
        }
    }
    if (!found) {
      *flag = 0;
      return;
    }
  }
}

namespace {
struct DevBuf;
}
__global__ void rb_off_scatter_kernel(const double* __restrict__ src, int n, double* __restrict__ dst, int cap, int c0,
                                      long long Nh);
namespace {
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
  template <typename T>
  T* as() {
    return reinterpret_cast<T*>(p);
  }
};

// Scratch of the Gram path kept in the context between calls (cudaMalloc / cudaFree cost more than the kernels of a
// small product); grow-only, freed with the context.  Same interface as DevBuf.
struct PoolBuf {
  rb_ctx* ctx;
  int slot;
  void* p = nullptr;
  PoolBuf(rb_ctx* c, int s) : ctx(c), slot(s) {}
  cudaError_t alloc(size_t bytes) {
    bytes = std::max<size_t>(bytes, 16);
    if (ctx->off_scratch_cap[slot] < bytes) {
      if (ctx->off_scratch[slot]) cudaFree(ctx->off_scratch[slot]);
      ctx->off_scratch[slot] = nullptr;
      ctx->off_scratch_cap[slot] = 0;
      cudaError_t e = cudaMalloc(&ctx->off_scratch[slot], bytes);
      if (e != cudaSuccess) return e;
      ctx->off_scratch_cap[slot] = bytes;
    }
    p = ctx->off_scratch[slot];
    return cudaSuccess;
  }
  template <typename T>
  T* as() {
    return reinterpret_cast<T*>(p);
  }
};

// Row pitch of a dense dof-major operand with n columns: even (16-byte aligned rows for the bulk copies); 65..128
// columns go through the single-tile product with one bulk copy per chunk and the global pitch kept in shared memory,
// where a pitch == 4 (mod 16) keeps the fragment loads conflict-free.
inline int dense_pitch(int n) { return (n > 64 && n <= 128) ? (n + 11) / 16 * 16 + 4 : ((n + 1) & ~1); }

// Host matrix (Nh x n, row-major, dense) -> device with an even row pitch *ld (16-byte aligned rows).
int upload_cols(rb_ctx* ctx, const double* host, long long Nh, int n, DevBuf& out, int* ld) {
  cudaStream_t st = ctx->stream;
  *ld = dense_pitch(n);
  RB_CUDA(out.alloc((size_t)Nh * *ld * sizeof(double)));
  if (*ld == n) {
    RB_CUDA(cudaMemcpyAsync(out.p, host, (size_t)Nh * n * sizeof(double), cudaMemcpyHostToDevice, st));
    return RB_OK;
  }
  DevBuf tmp;
  RB_CUDA(tmp.alloc((size_t)Nh * n * sizeof(double)));
  RB_CUDA(cudaMemsetAsync(out.p, 0, (size_t)Nh * *ld * sizeof(double), st));
  RB_CUDA(cudaMemcpyAsync(tmp.p, host, (size_t)Nh * n * sizeof(double), cudaMemcpyHostToDevice, st));
  rb_off_scatter_kernel<<<(unsigned)((Nh * n + 255) / 256), 256, 0, st>>>(tmp.as<double>(), n, out.as<double>(), *ld, 0, Nh);
  RB_CUDA(cudaGetLastError());
  RB_CUDA(cudaStreamSynchronize(st));  // tmp is freed on return
  return RB_OK;
}

// *out = 1 iff the CSR matrix equals its transpose entry by entry (one thread per row, rows are short).
int csr_is_symmetric(rb_ctx* ctx, long long Nh, const long long* dIp, const int* dIx, const double* dDt, int* out) {
  DevBuf flag;
  RB_CUDA(flag.alloc(sizeof(int)));
  const int one = 1;
  RB_CUDA(cudaMemcpyAsync(flag.p, &one, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  rb_csr_symmetric_kernel<<<(unsigned)((Nh + 255) / 256), 256, 0, ctx->stream>>>(dIp, dIx, dDt, Nh, flag.as<int>());
  RB_CUDA(cudaGetLastError());
  RB_CUDA(cudaMemcpyAsync(out, flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  RB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RB_OK;
}

// row-major [rows][ld] doubles seen from `ptr` as a 2-D tensor of `cols` x `rows`; box = box_cols x box_rows
typedef CUresult (*rb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
bool make_tensor_map(CUtensorMap* map, const double* ptr, long long cols, long long rows, long long ld, int box_cols,
                     int box_rows = T2_KC) {
  static const rb_encode_tiled_fn encode = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (rb_encode_tiled_fn)f;
  }();
  if (!encode || cols <= 0 || rows <= 0) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  static const int promo = getenv("RB_TMAP_L2PROMO") ? atoi(getenv("RB_TMAP_L2PROMO")) : 3;
  const CUtensorMapL2promotion l2 = promo == 0   ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                    : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                    : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                 : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}


// Fused path of gram_device for ns, ns2 <= 64 with an operator: one kernel + the split reduction.
int gram_fused_device(rb_ctx* ctx, long long Nh, int ns, int ns2, const double* dS, int lds, const double* dS2, int ld2,
                      const long long* dIndptr, const int* dIndices, const double* dData, long long rb, long long re,
                      double* C_host, float* ms_spmm, float* ms_tsmm) {
  cudaStream_t st = ctx->stream;
  PoolBuf ws(ctx, 1), dC(ctx, 2);
  const int ldw = 64;
  const long long rows = std::max<long long>(re - rb, 0);
  long long nsplit = (long long)ctx->sm_count;  // one wave, one CTA per SM
  nsplit = std::max<long long>(1, std::min(nsplit, (rows + FG_KC - 1) / FG_KC));
  RB_CUDA(ws.alloc((size_t)FG_NKG * nsplit * ldw * ldw * sizeof(double)));  // one partial per k-step group
  RB_CUDA(dC.alloc((size_t)ns * ns2 * sizeof(double)));
  const size_t smem = (size_t)2 * FG_NST * FG_KC * FG_P * sizeof(double) + 3 * FG_NST * sizeof(uint64_t);
  RB_CUDA(cudaFuncSetAttribute(rb_gram_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // a partially filled tile leaves accumulator rows / columns >= ns, ns2 undefined; they are never read back
  RB_CUDA(cudaEventRecord(ctx->ev[4], st));
  // in-chunk rows of S2 are read from the staged A operand only if that holds all the columns S2 needs
  const int reuse = (dS == dS2 && lds == ld2 && ((ns2 + 1) & ~1) <= ((ns + 1) & ~1)) ? 1 : 0;
  static const int no_tmap = getenv("RB_TSMM_TMAP") && atoi(getenv("RB_TSMM_TMAP")) == 0;
  CUtensorMap tmS;
  memset(&tmS, 0, sizeof(tmS));
  const int use_tmap = !no_tmap && re > 0 && make_tensor_map(&tmS, dS, (ns + 1) & ~1, re, lds, FG_P, FG_KC);
  rb_gram_fused_kernel<<<(unsigned)nsplit, FG_THREADS, smem, st>>>(dS, lds, ns, dS2, ld2, ns2, reuse, dIndptr,
                                                                              dIndices, dData, rb, re,
                                                                              ws.as<double>(), ldw, use_tmap, tmS);
  RB_CUDA(cudaGetLastError());
  rb_split_reduce_kernel<<<(ns * ns2 + 63) / 64, 256, 0, st>>>(ws.as<double>(), ldw, (int)(FG_NKG * nsplit), ns, ns2,
                                                               dC.as<double>(), 0);
  RB_CUDA(cudaGetLastError());
  RB_CUDA(cudaEventRecord(ctx->ev[0], st));
  if (ctx->allreduce) {  // sum of the ranks' partials, in place in HBM
    RB_CUDA(cudaStreamSynchronize(st));
    if (ctx->allreduce(ctx->allreduce_user, dC.as<double>(), (int64_t)ns * ns2))
      RB_FAIL(RB_ERR_CUDA, "rb_gram: the all-reduce callback failed");
  }
  RB_CUDA(cudaMemcpyAsync(C_host, dC.p, (size_t)ns * ns2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaStreamSynchronize(st));
  float a = 0;
  cudaEventElapsedTime(&a, ctx->ev[4], ctx->ev[0]);
  if (ms_spmm) *ms_spmm = 0.f;
  if (ms_tsmm) *ms_tsmm = a;
  ctx->last_ms[0] = 0.f;
  ctx->last_ms[1] = a;
  ctx->last_ms[4] = a;
  ctx->last_launches = 2;
  return RB_OK;
}

// dst[r * ldd + c] = src[r * lds + c], c < n: an operand whose rows are not 16-byte aligned (odd column offset inside a
// resident block) is copied once into aligned scratch instead of taking the slow 8-byte cp.async product.
__global__ void rb_copy_cols_kernel(const double* __restrict__ src, int lds, int n, double* __restrict__ dst, int ldd,
                                    long long r0, long long r1) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long r = r0 + t / n;
  if (r >= r1) return;
  dst[r * ldd + t % n] = src[r * lds + t % n];
}

// C (host, ns x ns2) = S^T (X S2) over rows [rb, re); all operands already on the device.
int gram_device(rb_ctx* ctx, long long Nh, int ns, int ns2, const double* dS, int lds, const double* dS2, int ld2,
                const long long* dIndptr, const int* dIndices, const double* dData, long long rb, long long re,
                double* C_host, float* ms_spmm, float* ms_tsmm, int x_sym = -1, bool at_col0 = true) {
  cudaStream_t st = ctx->stream;
  PoolBuf Ybuf(ctx, 0), ws(ctx, 1), dC(ctx, 2), alS(ctx, 3), alS2(ctx, 4);
  bool s_at0 = at_col0, y_at0 = at_col0 || dIndptr != nullptr;  // operand starts at column 0 of its rows (Y always does)
  static const int force_v1 = getenv("RB_TSMM_V1") ? atoi(getenv("RB_TSMM_V1")) : 0;
  if (!force_v1) {
    // realign: S always feeds the product; S2 does only without an operator (with one, the SpMM reads it at any
    // alignment and writes an aligned Y)
    const bool same = dS == dS2 && lds == ld2 && ns == ns2;
    auto misaligned = [](const double* p, int ld) { return ((uintptr_t)p & 15) != 0 || (ld & 1) != 0; };
    if (misaligned(dS, lds)) {
      const int l = dense_pitch(ns);
      RB_CUDA(alS.alloc((size_t)Nh * l * sizeof(double)));
      const long long rows = std::max<long long>(re - rb, 0);
      if (rows > 0)
        rb_copy_cols_kernel<<<(unsigned)((rows * ns + 255) / 256), 256, 0, st>>>(dS, lds, ns, alS.as<double>(), l, rb, re);
      RB_CUDA(cudaGetLastError());
      dS = alS.as<double>();
      lds = l;
      s_at0 = true;
      if (same && !dIndptr) {
        dS2 = dS;
        ld2 = l;
        y_at0 = true;
      }
    }
    if (!dIndptr && misaligned(dS2, ld2)) {
      const int l = dense_pitch(ns2);
      RB_CUDA(alS2.alloc((size_t)Nh * l * sizeof(double)));
      const long long rows = std::max<long long>(re - rb, 0);
      if (rows > 0)
        rb_copy_cols_kernel<<<(unsigned)((rows * ns2 + 255) / 256), 256, 0, st>>>(dS2, ld2, ns2, alS2.as<double>(), l, rb, re);
      RB_CUDA(cudaGetLastError());
      dS2 = alS2.as<double>();
      ld2 = l;
      y_at0 = true;
    }
  }
  const double* dY = dS2;
  // every allocation happens BEFORE the timed region (cudaMalloc synchronises and costs milliseconds)
  const int ldy = dIndptr ? dense_pitch(ns2) : ld2;
  static const int no_fuse = getenv("RB_GRAM_NOFUSE") ? atoi(getenv("RB_GRAM_NOFUSE")) : 0;
  // (up to 48 columns of S2 the thread-per-entry / per-pair SpMM + the single-tile product are faster than the fused
  //  kernel, whose time is set by its per-chunk hand-over, not by the width: 0.28 ms against 0.48 ms at 20 columns,
  //  0.38 against 0.49 ms at 40, 0.60 against 0.50 ms at 61; N_h = 10^6)
  const bool fused = dIndptr && !no_fuse && ns <= 64 && ns2 > 48 && ns2 <= 64 &&
                     (((uintptr_t)dS | (uintptr_t)dS2) & 15) == 0 && (lds % 2 == 0) && (ld2 % 2 == 0);
  if (fused) return gram_fused_device(ctx, Nh, ns, ns2, dS, lds, dS2, ld2, dIndptr, dIndices, dData, rb, re, C_host, ms_spmm, ms_tsmm);
  if (dIndptr) {
    RB_CUDA(Ybuf.alloc((size_t)Nh * ldy * sizeof(double)));
    dY = Ybuf.as<double>();
  }
  // bulk-copy kernels need 16-byte aligned rows (guaranteed by the realignment above; RB_TSMM_V1=1 selects the
  // first-generation cp.async product for A/B runs)
  const bool v2 = !force_v1 && (((uintptr_t)dS | (uintptr_t)dY) & 15) == 0 && (lds % 2 == 0) && (ldy % 2 == 0);
  const int TM = !v2 ? TS_TM : ((ns <= 64 && ns2 <= 64) ? 64 : 128);
  const int ti = (ns + TM - 1) / TM, tj = (ns2 + TM - 1) / TM;
  // S^T X S with X == X^T is symmetric: only the tiles on and above the diagonal are computed and the reduction mirrors
  // them (S^T A S for a general operator A takes the full grid).  The partial of a dof-row partition is not symmetric
  // by itself, but mirroring is linear: the SUM over the ranks of the mirrored partials is the mirrored sum, which is
  // what the host layer all-reduces -- so with a symmetric X a rank's block below the diagonal tiles holds the
  // transposed upper tiles of ITS partial, and only the sum over ranks is meaningful there.
  static const int no_sym = getenv("RB_GRAM_NOSYM") ? atoi(getenv("RB_GRAM_NOSYM")) : 0;
  bool sym = v2 && !no_sym && ti > 1 && dS == dS2 && lds == ld2 && ns == ns2;
  if (sym && dIndptr) {
    if (x_sym < 0) {
      if (int rc = csr_is_symmetric(ctx, Nh, dIndptr, dIndices, dData, &x_sym)) return rc;
    }
    sym = x_sym == 1;
  }
  const int ntiles = sym ? ti * (ti + 1) / 2 : ti * tj;
  static const int no_bal = getenv("RB_TSMM_NOBAL") ? atoi(getenv("RB_TSMM_NOBAL")) : 0;
  const bool bal = v2 && ti == 1 && tj == 1 && !no_bal;  // one tile: sub-tiles dealt evenly to the warps
  const int ldw = v2 ? 32 * std::max((ns + 31) / 32, (ns2 + 31) / 32) : std::max(ti, tj) * TM;
  const long long rows = std::max<long long>(re - rb, 0);
  const int KC = v2 ? T2_KC : TS_KC;
  long long nsplit;
  if (v2) {
    // one wave of equal CTAs for a single tile; several tiles have unequal numbers of active warps, so ~6 waves
    // let the block scheduler balance them
    const long long slots = (long long)ctx->sm_count * (TM == 64 ? 2 : 1);
    nsplit = std::max<long long>(1, slots * (ntiles > 1 ? 6 : 1) / ntiles);
  } else {
    // enough splits for ~3 waves of 148 SMs x 3 resident CTAs
    nsplit = std::max<long long>(1, (long long)ctx->sm_count * 9 / (ti * tj));
  }
  nsplit = std::min(nsplit, std::max<long long>(1, rows / (KC * 8)));
  nsplit = std::min<long long>(nsplit, 65535);
  long long rps = (rows + nsplit - 1) / nsplit;
  rps = (rps + KC - 1) / KC * KC;
  if (rps == 0) rps = KC;
  nsplit = std::max<long long>(1, (rows + rps - 1) / rps);
  RB_CUDA(ws.alloc((size_t)nsplit * ldw * ldw * sizeof(double)));
  RB_CUDA(dC.alloc((size_t)ns * ns2 * sizeof(double)));
  const size_t smem = v2 ? (size_t)2 * T2_NST * T2_KC * (TM + 4) * sizeof(double) + T2_NST * 16
                         : (size_t)2 * TS_STAGES * TS_KC * TS_PITCH * sizeof(double);
  if (!v2)
    RB_CUDA(cudaFuncSetAttribute(rb_tsmm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (bal && TM == 64)
    RB_CUDA(cudaFuncSetAttribute(rb_tsmm2_bal_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (bal)
    RB_CUDA(cudaFuncSetAttribute(rb_tsmm2_bal_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (TM == 64)
    RB_CUDA(cudaFuncSetAttribute(rb_tsmm2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else
    RB_CUDA(cudaFuncSetAttribute(rb_tsmm2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RB_CUDA(cudaEventRecord(ctx->ev[4], st));
  if (dIndptr) {
    const int nchunk = (ns2 + 127) / 128;
    const long long warps = (re - rb) * nchunk;
    const long long blocks = (warps * 32 + 255) / 256;
    static const int pair_max = getenv("RB_SPMM_PAIR_MAX") ? atoi(getenv("RB_SPMM_PAIR_MAX")) : 96;
    const bool s2_aligned = ((uintptr_t)dS2 & 15) == 0 && (ld2 % 2 == 0);
    // measured at N_h = 10^6, 5 nnz/row: per entry 0.095 ms at 8 columns (per pair: 0.158), per pair 0.131 ms at 20
    // (per entry: 0.184) and 0.181 ms at 32 (0.293)
    if (blocks > 0 && ns2 <= 32 && (ns2 <= 12 || !s2_aligned)) {
      const long long threads = (re - rb) * ns2;
      rb_spmm_narrow_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(dIndptr, dIndices, dData, dS2, ld2, ns2,
                                                                              Ybuf.as<double>(), ldy, rb, re);
      RB_CUDA(cudaGetLastError());
    } else if (blocks > 0 && ns2 <= pair_max && s2_aligned) {
      // (a pair that straddles ns2 reads / writes the padding column of the even pitch)
      const int npair = (ns2 + 1) / 2;
      const long long threads = (re - rb) * npair;
      rb_spmm_pair_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(dIndptr, dIndices, dData, dS2, ld2, npair,
                                                                            Ybuf.as<double>(), ldy, rb, re);
      RB_CUDA(cudaGetLastError());
    } else if (blocks > 0) {
      rb_spmm_kernel<<<(unsigned)blocks, 256, 0, st>>>(dIndptr, dIndices, dData, dS2, ld2, ns2, Ybuf.as<double>(), ldy,
                                                      rb, re);
      RB_CUDA(cudaGetLastError());
    }
  }
  RB_CUDA(cudaEventRecord(ctx->ev[5], st));
  dim3 grid(ti, tj, (unsigned)nsplit);
  if (!v2)
    rb_tsmm_kernel<<<grid, TS_THREADS, smem, st>>>(dS, lds, ns, dY, ldy, ns2, rb, re, rps, ws.as<double>(), ldw);
  else {
    // operands as 2-D tensor maps (RB_TSMM_TMAP=0: one 1-D bulk copy per row instead, for A/B runs); a single tile
    // whose operands are whole rows takes one contiguous bulk copy per operand and stage instead
    static const int no_tmap = getenv("RB_TSMM_TMAP") && atoi(getenv("RB_TSMM_TMAP")) == 0;
    const int contig = (bal && s_at0 && y_at0 && lds <= TM + 4 && ldy <= TM + 4) ? 1 : 0;
    CUtensorMap tmS, tmY;
    memset(&tmS, 0, sizeof(tmS));
    memset(&tmY, 0, sizeof(tmY));
    const int use_tmap = !no_tmap && !contig && re > 0 && make_tensor_map(&tmS, dS, (ns + 1) & ~1, re, lds, TM + 4) &&
                         make_tensor_map(&tmY, dY, (ns2 + 1) & ~1, re, ldy, TM + 4);
    if (bal && TM == 64)
      rb_tsmm2_bal_kernel<64><<<dim3((unsigned)nsplit), 128, smem, st>>>(dS, lds, ns, dY, ldy, ns2, rb, re, rps,
                                                                        ws.as<double>(), ldw, contig, use_tmap, tmS, tmY);
    else if (bal)
      rb_tsmm2_bal_kernel<128><<<dim3((unsigned)nsplit), 512, smem, st>>>(dS, lds, ns, dY, ldy, ns2, rb, re, rps,
                                                                         ws.as<double>(), ldw, contig, use_tmap, tmS, tmY);
    else if (TM == 64)
      rb_tsmm2_kernel<64><<<grid, 128, smem, st>>>(dS, lds, ns, dY, ldy, ns2, rb, re, rps, ws.as<double>(), ldw, 0,
                                                   use_tmap, tmS, tmY);
    else
      rb_tsmm2_kernel<128><<<grid, 512, smem, st>>>(dS, lds, ns, dY, ldy, ns2, rb, re, rps, ws.as<double>(), ldw,
                                                    sym ? 1 : 0, use_tmap, tmS, tmY);
  }
  RB_CUDA(cudaGetLastError());
  rb_split_reduce_kernel<<<(ns * ns2 + 63) / 64, 256, 0, st>>>(ws.as<double>(), ldw, (int)nsplit, ns, ns2,
                                                               dC.as<double>(), sym ? 32 : 0);
  RB_CUDA(cudaGetLastError());
  RB_CUDA(cudaEventRecord(ctx->ev[0], st));
  if (ctx->allreduce) {  // sum of the ranks' partials, in place in HBM
    RB_CUDA(cudaStreamSynchronize(st));
    if (ctx->allreduce(ctx->allreduce_user, dC.as<double>(), (int64_t)ns * ns2))
      RB_FAIL(RB_ERR_CUDA, "rb_gram: the all-reduce callback failed");
  }
  RB_CUDA(cudaMemcpyAsync(C_host, dC.p, (size_t)ns * ns2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaStreamSynchronize(st));
  float a = 0, b = 0;
  cudaEventElapsedTime(&a, ctx->ev[4], ctx->ev[5]);
  cudaEventElapsedTime(&b, ctx->ev[5], ctx->ev[0]);
  if (ms_spmm) *ms_spmm = a;
  if (ms_tsmm) *ms_tsmm = b;
  ctx->last_ms[0] = a;
  ctx->last_ms[1] = b;
  ctx->last_ms[4] = a + b;
  ctx->last_launches = (dIndptr ? 1 : 0) + 2;
  return RB_OK;
}
}  // namespace

extern "C" int rb_gram(rb_ctx* ctx, int64_t Nh, int Ns, int Ns2, const double* S, const double* S2,
                       const int64_t* indptr, const int32_t* indices, const double* data, int64_t row_begin,
                       int64_t row_end, double* C) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (Nh <= 0 || Ns <= 0 || !S || !C || row_begin < 0 || row_end > Nh || row_begin > row_end)
    RB_FAIL(RB_ERR_ARG, "rb_gram: bad arguments");
  if (indptr && (!indices || !data)) RB_FAIL(RB_ERR_ARG, "rb_gram: CSR arrays incomplete");
  if (!S2) Ns2 = Ns;
  if (Ns2 <= 0) RB_FAIL(RB_ERR_ARG, "rb_gram: Ns2 must be positive");
  cudaStream_t st = ctx->stream;
  DevBuf dS, dS2, dIp, dIx, dDt;
  int lds = Ns, ld2 = Ns;
  if (int rc = upload_cols(ctx, S, Nh, Ns, dS, &lds)) return rc;
  const double* pS2 = dS.as<double>();
  ld2 = lds;
  if (S2) {
    if (int rc = upload_cols(ctx, S2, Nh, Ns2, dS2, &ld2)) return rc;
    pS2 = dS2.as<double>();
  }
  if (indptr) {
    const long long nnz = indptr[Nh];
    RB_CUDA(dIp.alloc((size_t)(Nh + 1) * sizeof(long long)));
    RB_CUDA(dIx.alloc((size_t)nnz * sizeof(int)));
    RB_CUDA(dDt.alloc((size_t)nnz * sizeof(double)));
    RB_CUDA(cudaMemcpyAsync(dIp.p, indptr, (size_t)(Nh + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    RB_CUDA(cudaMemcpyAsync(dIx.p, indices, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, st));
    RB_CUDA(cudaMemcpyAsync(dDt.p, data, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  return gram_device(ctx, Nh, Ns, Ns2, dS.as<double>(), lds, pS2, ld2, indptr ? dIp.as<long long>() : nullptr,
                     indptr ? dIx.as<int>() : nullptr, indptr ? dDt.as<double>() : nullptr, row_begin, row_end, C,
                     nullptr, nullptr);
}

extern "C" int rb_project(rb_ctx* ctx, int64_t Nh, int N, int Q, const double* Z, const int64_t* indptr_all,
                          const int32_t* indices_all, const double* data_all, const int64_t* nnz_off, double* out) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (Nh <= 0 || N <= 0 || Q <= 0 || !Z || !indptr_all || !indices_all || !data_all || !nnz_off || !out)
    RB_FAIL(RB_ERR_ARG, "rb_project: bad arguments");
  cudaStream_t st = ctx->stream;
  DevBuf dZ, dIp, dIx, dDt;
  int ldz = N;
  if (int rc = upload_cols(ctx, Z, Nh, N, dZ, &ldz)) return rc;
  const long long nnz_total = nnz_off[Q];
  RB_CUDA(dIp.alloc((size_t)Q * (Nh + 1) * sizeof(long long)));
  RB_CUDA(dIx.alloc((size_t)nnz_total * sizeof(int)));
  RB_CUDA(dDt.alloc((size_t)nnz_total * sizeof(double)));
  RB_CUDA(cudaMemcpyAsync(dIp.p, indptr_all, (size_t)Q * (Nh + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
  RB_CUDA(cudaMemcpyAsync(dIx.p, indices_all, (size_t)nnz_total * sizeof(int), cudaMemcpyHostToDevice, st));
  RB_CUDA(cudaMemcpyAsync(dDt.p, data_all, (size_t)nnz_total * sizeof(double), cudaMemcpyHostToDevice, st));
  float tot_a = 0, tot_b = 0;
  for (int q = 0; q < Q; ++q) {
    float a = 0, b = 0;
    if (int rc = gram_device(ctx, Nh, N, N, dZ.as<double>(), ldz, dZ.as<double>(), ldz, dIp.as<long long>() + (size_t)q * (Nh + 1),
                             dIx.as<int>() + nnz_off[q], dDt.as<double>() + nnz_off[q], 0, Nh,
                             out + (size_t)q * N * N, &a, &b))
      return rc;
    tot_a += a;
    tot_b += b;
  }
  ctx->last_ms[0] = tot_a;
  ctx->last_ms[1] = tot_b;
  ctx->last_ms[4] = tot_a + tot_b;
  ctx->last_launches = 3 * Q;
  return RB_OK;
}

// =====================================================================================================================
// Z = S V  (tall GEMM, M = Nh; rbnics/backends/dolfin/wrapping/functions_list_mul.py:10-36).  S is read once and Z
// written once, but n is small (Nmax modes), so the DMMA work on a padded tile is what decides: the first version
// (128 x 64 tile) issued 2 Nh ns 64 flops for n = 20 and sat at 22 % of HBM.  Now: 128 rows x 32 columns per CTA (4
// warps of 32 x 32, grid.y walks wider n), only the NTC = ceil(columns / 8) 8-column tiles that exist are multiplied
// (compile-time count), K chunks of 16 through a 4-stage ring of 16-byte cp.async copies (operands are uploaded with
// even pitches), two CTAs per SM.  A ring of per-row bulk copies measured slower: cp.async.bulk is issued one lane
// at a time.
// =====================================================================================================================
constexpr int LC_BM = 128, LC_BN = 32, LC_KC = 16, LC_THREADS = 128, LC_NST = 4;
constexpr int LC_APITCH = LC_KC + 4;  // == 4 (mod 8)
constexpr int LC_BPITCH = LC_BN + 4;  // == 4 (mod 16)

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = smem_u32(dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

template <int NTC>
__global__ void __launch_bounds__(LC_THREADS, 2) rb_lincomb_kernel(const double* __restrict__ S, int lds, int ns,
                                                                   const double* __restrict__ V, int ldv, int n,
                                                                   double* __restrict__ Z, int ldz, long long Nh,
                                                                   int col0, int accumulate) {
  extern __shared__ __align__(16) double lc_smem[];
  double(*As)[LC_BM * LC_APITCH] = reinterpret_cast<double(*)[LC_BM * LC_APITCH]>(lc_smem);
  double(*Bs)[LC_KC * LC_BPITCH] = reinterpret_cast<double(*)[LC_KC * LC_BPITCH]>(lc_smem + LC_NST * LC_BM * LC_APITCH);
  const int tid = threadIdx.x, lane = tid & 31, wm = tid >> 5;
  const int g = lane >> 2, q4 = lane & 3;
  const long long r0 = (long long)blockIdx.x * LC_BM;
  const int n0 = col0 + blockIdx.y * LC_BN;
  const int nchunks = (ns + LC_KC - 1) / LC_KC;
  // pairs of doubles; lds and ldv are even, a pair that straddles ns / n brings one padding element that meets a zero
  // (rows of V past ns are zero-filled) or lands in a column that is never stored
  auto load_stage = [&](int stage, int chunk) {
    const int k0 = chunk * LC_KC;
    for (int t = tid; t < LC_BM * LC_KC / 2; t += LC_THREADS) {
      const int rr = t / (LC_KC / 2), kk = 2 * (t % (LC_KC / 2));
      const bool v = (r0 + rr < Nh) && (k0 + kk < ns);
      cp_async16(&As[stage][rr * LC_APITCH + kk], v ? (S + (r0 + rr) * lds + k0 + kk) : S, v);
    }
    for (int t = tid; t < LC_KC * LC_BN / 2; t += LC_THREADS) {
      const int kk = t / (LC_BN / 2), cc = 2 * (t % (LC_BN / 2));
      const bool v = (k0 + kk < ns) && (n0 + cc < n);
      cp_async16(&Bs[stage][kk * LC_BPITCH + cc], v ? (V + (long long)(k0 + kk) * ldv + n0 + cc) : V, v);
    }
  };
  double acc[4][NTC][2];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  for (int s = 0; s < LC_NST - 1; ++s) {
    if (s < nchunks) load_stage(s, s);
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<LC_NST - 2>();
    __syncthreads();
    {
      const int nxt = c + LC_NST - 1;
      if (nxt < nchunks) load_stage(nxt % LC_NST, nxt);
      cp_async_commit();
    }
    const double* as = As[c % LC_NST];
    const double* bs = Bs[c % LC_NST];
#pragma unroll
    for (int ks = 0; ks < LC_KC / 4; ++ks) {
      double af[4], bf[NTC];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) af[mt] = as[(wm * 32 + mt * 8 + g) * LC_APITCH + 4 * ks + q4];
#pragma unroll
      for (int nt = 0; nt < NTC; ++nt) bf[nt] = bs[(4 * ks + q4) * LC_BPITCH + nt * 8 + g];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTC; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const long long r = r0 + wm * 32 + mt * 8 + g;
    if (r >= Nh) continue;
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt) {
      const int c = n0 + nt * 8 + q4 * 2;
      if (c < n) Z[r * ldz + c] = accumulate ? Z[r * ldz + c] + acc[mt][nt][0] : acc[mt][nt][0];
      if (c + 1 < n) Z[r * ldz + c + 1] = accumulate ? Z[r * ldz + c + 1] + acc[mt][nt][1] : acc[mt][nt][1];
    }
  }
}

// ny column tiles of 32 starting at column col0 of V / Z
template <int NTC>
static cudaError_t lincomb_launch(cudaStream_t st, size_t smem, const double* dS, int lds, int ns, const double* dV,
                                  int ldv, int n, double* dZ, int ldz, long long Nh, int ny, int col0,
                                  int accumulate) {
  cudaError_t e = cudaFuncSetAttribute(rb_lincomb_kernel<NTC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)((Nh + LC_BM - 1) / LC_BM), ny);
  rb_lincomb_kernel<NTC><<<grid, LC_THREADS, smem, st>>>(dS, lds, ns, dV, ldv, n, dZ, ldz, Nh, col0, accumulate);
  return cudaGetLastError();
}

// Z (pitch ldz) = S (Nh x Ns, pitch lds, 16-byte aligned rows) * V (Ns x n, even pitch ldv); all on the device
static int lincomb_device(rb_ctx* ctx, long long Nh, int Ns, int n, const double* pS, int lds, const double* pV, int ldv,
                          double* pZ, int ldz, int accumulate = 0) {
  cudaStream_t st = ctx->stream;
  const size_t smem = (size_t)LC_NST * (LC_BM * LC_APITCH + LC_KC * LC_BPITCH) * sizeof(double);
  // full 32-column tiles in one launch, the narrower last tile with the 8-column tile count it needs
  const int nfull = n / LC_BN, rem = n % LC_BN;
  if (nfull > 0) RB_CUDA(lincomb_launch<4>(st, smem, pS, lds, Ns, pV, ldv, n, pZ, ldz, Nh, nfull, 0, accumulate));
  if (rem > 0) {
    const int c0 = nfull * LC_BN, ntc = (rem + 7) / 8;
    RB_CUDA(ntc == 1   ? lincomb_launch<1>(st, smem, pS, lds, Ns, pV, ldv, n, pZ, ldz, Nh, 1, c0, accumulate)
            : ntc == 2 ? lincomb_launch<2>(st, smem, pS, lds, Ns, pV, ldv, n, pZ, ldz, Nh, 1, c0, accumulate)
            : ntc == 3 ? lincomb_launch<3>(st, smem, pS, lds, Ns, pV, ldv, n, pZ, ldz, Nh, 1, c0, accumulate)
                       : lincomb_launch<4>(st, smem, pS, lds, Ns, pV, ldv, n, pZ, ldz, Nh, 1, c0, accumulate));
  }
  return RB_OK;
}

extern "C" int rb_lincomb(rb_ctx* ctx, int64_t Nh, int Ns, int n, const double* S, const double* V, double* Z) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (Nh <= 0 || Ns <= 0 || n <= 0 || !S || !V || !Z) RB_FAIL(RB_ERR_ARG, "rb_lincomb: bad arguments");
  cudaStream_t st = ctx->stream;
  DevBuf dS, dV, dZ;
  int lds = Ns, ldv = n;
  if (int rc = upload_cols(ctx, S, Nh, Ns, dS, &lds)) return rc;
  if (int rc = upload_cols(ctx, V, Ns, n, dV, &ldv)) return rc;
  RB_CUDA(dZ.alloc((size_t)Nh * n * sizeof(double)));
  RB_CUDA(cudaEventRecord(ctx->ev[4], st));
  if (int rc = lincomb_device(ctx, Nh, Ns, n, dS.as<double>(), lds, dV.as<double>(), ldv, dZ.as<double>(), n)) return rc;
  RB_CUDA(cudaEventRecord(ctx->ev[5], st));
  RB_CUDA(cudaMemcpyAsync(Z, dZ.p, (size_t)Nh * n * sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaStreamSynchronize(st));
  float a = 0;
  cudaEventElapsedTime(&a, ctx->ev[4], ctx->ev[5]);
  ctx->last_ms[0] = a;
  ctx->last_ms[4] = a;
  ctx->last_launches = (n / LC_BN > 0) + (n % LC_BN > 0);
  return RB_OK;
}

// =====================================================================================================================
// Gram-Schmidt step: sequential dot + axpy over the basis columns (deterministic two-stage dots)
// =====================================================================================================================
constexpr int GS_BLOCKS = 592, GS_THREADS = 256;

__global__ void __launch_bounds__(GS_THREADS) rb_dot_partial_kernel(const double* __restrict__ x, int incx,
                                                                    const double* __restrict__ y, int incy,
                                                                    long long n, double* __restrict__ partial) {
  __shared__ double sh[GS_THREADS / 32];
  double s = 0.0;
  for (long long i = blockIdx.x * (long long)GS_THREADS + threadIdx.x; i < n; i += (long long)GS_BLOCKS * GS_THREADS)
    s = fma(x[i * incx], y[i * incy], s);
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < GS_THREADS / 32; ++w) t += sh[w];
    partial[blockIdx.x] = t;
  }
}

__device__ __forceinline__ double gs_sum_partials(const double* partial) {
  // every block sums the same partials in the same order -> identical, deterministic scalar
  __shared__ double tot;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int b = 0; b < GS_BLOCKS; ++b) t += partial[b];
    tot = t;
  }
  __syncthreads();
  return tot;
}

// z -= d * b  with d = sum(partial)
__global__ void __launch_bounds__(GS_THREADS) rb_axpy_kernel(const double* __restrict__ partial,
                                                             const double* __restrict__ b, int incb,
                                                             double* __restrict__ z, long long n) {
  const double d = gs_sum_partials(partial);
  for (long long i = blockIdx.x * (long long)GS_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * GS_THREADS)
    z[i] = fma(-d, b[i * incb], z[i]);
}

// One MGS step on contiguous vectors: z -= d_k b_k with d_k = sum(partial_in), and in the same pass the partial dots of
// the NEXT step <z, xz_next> (same per-thread order as rb_dot_partial_kernel, so the results are bitwise those of the
// separate kernels); xz_next == nullptr for the last basis function.
__global__ void __launch_bounds__(GS_THREADS) rb_gs_step_kernel(const double* __restrict__ partial_in,
                                                                const double* __restrict__ b, double* __restrict__ z,
                                                                const double* __restrict__ xz_next, long long n,
                                                                double* __restrict__ partial_out) {
  __shared__ double sh[GS_THREADS / 32];
  const double d = gs_sum_partials(partial_in);
  double s = 0.0;
  for (long long i = blockIdx.x * (long long)GS_THREADS + threadIdx.x; i < n; i += (long long)GS_BLOCKS * GS_THREADS) {
    const double zi = fma(-d, b[i], z[i]);
    z[i] = zi;
    if (xz_next) s = fma(zi, xz_next[i], s);
  }
  if (!xz_next) return;
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < GS_THREADS / 32; ++w) t += sh[w];
    partial_out[blockIdx.x] = t;
  }
}

// z /= sqrt(sum(partial)) unless that is zero; writes the norm
__global__ void __launch_bounds__(GS_THREADS) rb_normalize_kernel(const double* __restrict__ partial,
                                                                  double* __restrict__ z, long long n,
                                                                  double* __restrict__ norm_out) {
  const double nrm = sqrt(gs_sum_partials(partial));
  if (blockIdx.x == 0 && threadIdx.x == 0) *norm_out = nrm;
  if (nrm != 0.0) {
    for (long long i = blockIdx.x * (long long)GS_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * GS_THREADS)
      z[i] = z[i] / nrm;
  }
}

namespace {
// z (device, Nh) is orthonormalised in place against the n columns of dZ (row-major, pitch ldz) in the X inner
// product (identity if dIp == nullptr): single-pass MGS, deterministic dots.  *launches is incremented.
int gs_device(rb_ctx* ctx, long long Nh, int n, const double* dZ, int ldz, const long long* dIp, const int* dIx,
              const double* dDt, double* dz, double* dNorm, int* launches) {
  cudaStream_t st = ctx->stream;
  DevBuf dXZ, dXz, dPart;
  RB_CUDA(dPart.alloc(GS_BLOCKS * sizeof(double)));
  const double* dXZp = dZ;
  int ldxz = ldz;
  if (dIp && n > 0) {
    RB_CUDA(dXZ.alloc((size_t)Nh * n * sizeof(double)));
    const int nchunk = (n + 127) / 128;
    const long long blocks = (Nh * nchunk * 32 + 255) / 256;
    rb_spmm_kernel<<<(unsigned)blocks, 256, 0, st>>>(dIp, dIx, dDt, dZ, ldz, n, dXZ.as<double>(), n, 0, Nh);
    RB_CUDA(cudaGetLastError());
    dXZp = dXZ.as<double>();
    ldxz = n;
    ++*launches;
  }
  for (int k = 0; k < n; ++k) {
    rb_dot_partial_kernel<<<GS_BLOCKS, GS_THREADS, 0, st>>>(dz, 1, dXZp + k, ldxz, Nh, dPart.as<double>());
    rb_axpy_kernel<<<GS_BLOCKS, GS_THREADS, 0, st>>>(dPart.as<double>(), dZ + k, ldz, dz, Nh);
    *launches += 2;
  }
  RB_CUDA(cudaGetLastError());
  const double* dXzp = dz;
  if (dIp) {
    RB_CUDA(dXz.alloc((size_t)Nh * sizeof(double)));
    const long long blocks = (Nh * 32 + 255) / 256;
    rb_spmm_kernel<<<(unsigned)blocks, 256, 0, st>>>(dIp, dIx, dDt, dz, 1, 1, dXz.as<double>(), 1, 0, Nh);
    dXzp = dXz.as<double>();
    ++*launches;
  }
  rb_dot_partial_kernel<<<GS_BLOCKS, GS_THREADS, 0, st>>>(dz, 1, dXzp, 1, Nh, dPart.as<double>());
  rb_normalize_kernel<<<GS_BLOCKS, GS_THREADS, 0, st>>>(dPart.as<double>(), dz, Nh, dNorm);
  *launches += 2;
  RB_CUDA(cudaGetLastError());
  RB_CUDA(cudaStreamSynchronize(st));  // the temporaries are freed on return
  return RB_OK;
}
}  // namespace

extern "C" int rb_gram_schmidt(rb_ctx* ctx, int64_t Nh, int n, const double* Zb, const int64_t* indptr,
                               const int32_t* indices, const double* data, double* z, double* norm_out) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (Nh <= 0 || n < 0 || (n > 0 && !Zb) || !z) RB_FAIL(RB_ERR_ARG, "rb_gram_schmidt: bad arguments");
  if (indptr && (!indices || !data)) RB_FAIL(RB_ERR_ARG, "rb_gram_schmidt: CSR arrays incomplete");
  cudaStream_t st = ctx->stream;
  DevBuf dZ, dz, dIp, dIx, dDt, dNorm;
  RB_CUDA(dz.alloc((size_t)Nh * sizeof(double)));
  RB_CUDA(dNorm.alloc(sizeof(double)));
  RB_CUDA(cudaMemcpyAsync(dz.p, z, (size_t)Nh * sizeof(double), cudaMemcpyHostToDevice, st));
  if (n > 0) {
    RB_CUDA(dZ.alloc((size_t)Nh * n * sizeof(double)));
    RB_CUDA(cudaMemcpyAsync(dZ.p, Zb, (size_t)Nh * n * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  if (indptr) {
    const long long nnz = indptr[Nh];
    RB_CUDA(dIp.alloc((size_t)(Nh + 1) * sizeof(long long)));
    RB_CUDA(dIx.alloc((size_t)nnz * sizeof(int)));
    RB_CUDA(dDt.alloc((size_t)nnz * sizeof(double)));
    RB_CUDA(cudaMemcpyAsync(dIp.p, indptr, (size_t)(Nh + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    RB_CUDA(cudaMemcpyAsync(dIx.p, indices, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, st));
    RB_CUDA(cudaMemcpyAsync(dDt.p, data, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  int launches = 0;
  if (int rc = gs_device(ctx, Nh, n, dZ.as<double>(), n, indptr ? dIp.as<long long>() : nullptr,
                         indptr ? dIx.as<int>() : nullptr, indptr ? dDt.as<double>() : nullptr, dz.as<double>(),
                         dNorm.as<double>(), &launches))
    return rc;
  double nrm = 0.0;
  RB_CUDA(cudaMemcpyAsync(z, dz.p, (size_t)Nh * sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaMemcpyAsync(&nrm, dNorm.p, sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaStreamSynchronize(st));
  if (norm_out) *norm_out = nrm;
  ctx->last_launches = launches;
  return RB_OK;
}

// =====================================================================================================================
// Device-resident offline workspace: sparse operators and dense column blocks stay in HBM across greedy iterations,
// only the new snapshot / basis function / Riesz representer crosses PCIe.
// =====================================================================================================================
static rb_off_csr_t* off_csr(rb_ctx* ctx, int slot) {
  return (slot >= 0 && slot < RB_OFF_SLOTS && ctx->off_csr[slot].ip) ? &ctx->off_csr[slot] : nullptr;
}
static rb_off_blk_t* off_blk(rb_ctx* ctx, int slot) {
  return (slot >= 0 && slot < RB_OFF_SLOTS && ctx->off_blk[slot].d) ? &ctx->off_blk[slot] : nullptr;
}

extern "C" int rb_off_csr(rb_ctx* ctx, int slot, int64_t Nh, const int64_t* indptr, const int32_t* indices,
                          const double* data) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (slot < 0 || slot >= RB_OFF_SLOTS || Nh <= 0 || !indptr || !indices || !data) RB_FAIL(RB_ERR_ARG, "rb_off_csr: bad arguments");
  rb_off_csr_t& c = ctx->off_csr[slot];
  for (void* q : {(void*)c.ip, (void*)c.ix, (void*)c.dt})
    if (q) cudaFree(q);
  c = rb_off_csr_t();
  c.generation = ++ctx->off_generation;  // mirrors of X z_k built with the previous matrix of this slot are stale
  const long long nnz = indptr[Nh];
  RB_CUDA(cudaMalloc((void**)&c.ip, (size_t)(Nh + 1) * sizeof(long long)));
  RB_CUDA(cudaMalloc((void**)&c.ix, (size_t)std::max<long long>(nnz, 1) * sizeof(int)));
  RB_CUDA(cudaMalloc((void**)&c.dt, (size_t)std::max<long long>(nnz, 1) * sizeof(double)));
  RB_CUDA(cudaMemcpy(c.ip, indptr, (size_t)(Nh + 1) * sizeof(long long), cudaMemcpyHostToDevice));
  RB_CUDA(cudaMemcpy(c.ix, indices, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice));
  RB_CUDA(cudaMemcpy(c.dt, data, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice));
  c.Nh = Nh;
  c.nnz = nnz;
  {
    // (DevBuf lives in the anonymous namespace above)
    if (int rc = csr_is_symmetric(ctx, Nh, c.ip, c.ix, c.dt, &c.sym)) return rc;
  }
  return RB_OK;
}

extern "C" int rb_off_block(rb_ctx* ctx, int slot, int64_t Nh, int capacity) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  if (slot < 0 || slot >= RB_OFF_SLOTS || Nh <= 0 || capacity <= 0) RB_FAIL(RB_ERR_ARG, "rb_off_block: bad arguments");
  rb_off_blk_t& b = ctx->off_blk[slot];
  for (void* q : {(void*)b.d, (void*)b.zt, (void*)b.xzt})
    if (q) cudaFree(q);
  b = rb_off_blk_t();
  const int ld = (capacity + 1) & ~1;
  RB_CUDA(cudaMalloc((void**)&b.d, (size_t)Nh * ld * sizeof(double)));
  RB_CUDA(cudaMemset(b.d, 0, (size_t)Nh * ld * sizeof(double)));
  b.Nh = Nh;
  b.cap = capacity;
  b.ld = ld;
  b.n = 0;
  return RB_OK;
}

// dst[r * cap + c0 + j] = src[r * n + j]
__global__ void rb_off_scatter_kernel(const double* __restrict__ src, int n, double* __restrict__ dst, int cap, int c0,
                                      long long Nh) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= Nh * n) return;
  dst[(t / n) * cap + c0 + (t % n)] = src[t];
}
__global__ void rb_off_gather_kernel(const double* __restrict__ src, int cap, int c0, int n, double* __restrict__ dst,
                                     long long Nh) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= Nh * n) return;
  dst[t] = src[(t / n) * cap + c0 + (t % n)];
}

// columns c0 .. c0+n of the rows outside [lo, hi) := 0
__global__ void rb_off_zero_rows_outside_kernel(double* __restrict__ d, int ld, int c0, int n, long long Nh, long long lo,
                                                long long hi) {
  const long long outside = Nh - (hi - lo);
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= outside * n) return;
  long long r = t / n;
  const int c = (int)(t % n);
  if (r >= lo) r += hi - lo;
  d[r * ld + c0 + c] = 0.0;
}

extern "C" int rb_off_append(rb_ctx* ctx, int slot, int n, const double* cols) {
  if (!ctx) return RB_ERR_ARG;
  rb_off_blk_t* b = off_blk(ctx, slot);
  if (!b) RB_FAIL(RB_ERR_ARG, "rb_off_append: empty slot");
  return rb_off_append_rows(ctx, slot, n, cols, 0, b->Nh);
}

// Only the dof rows [row_lo, row_hi) of the new columns cross PCIe (a rank's own rows plus the rows its operator rows
// reference); the others are zero-filled on the device and must not be read by this rank's products.
extern "C" int rb_off_append_rows(rb_ctx* ctx, int slot, int n, const double* cols, int64_t row_lo, int64_t row_hi) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  rb_off_blk_t* b = off_blk(ctx, slot);
  if (!b || n <= 0 || !cols) RB_FAIL(RB_ERR_ARG, "rb_off_append: bad arguments");
  if (b->n + n > b->cap) RB_FAIL(RB_ERR_ARG, "rb_off_append: block capacity exceeded");
  if (row_lo < 0 || row_hi > b->Nh || row_lo > row_hi) RB_FAIL(RB_ERR_ARG, "rb_off_append_rows: bad row range");
  const long long rows = row_hi - row_lo;
  if (rows > 0) {
    DevBuf tmp;
    RB_CUDA(tmp.alloc((size_t)rows * n * sizeof(double)));
    RB_CUDA(cudaMemcpyAsync(tmp.p, cols + row_lo * n, (size_t)rows * n * sizeof(double), cudaMemcpyHostToDevice,
                            ctx->stream));
    rb_off_scatter_kernel<<<(unsigned)((rows * n + 255) / 256), 256, 0, ctx->stream>>>(
        tmp.as<double>(), n, b->d + row_lo * b->ld, b->ld, b->n, rows);
    RB_CUDA(cudaGetLastError());
    RB_CUDA(cudaStreamSynchronize(ctx->stream));  // tmp is freed on return
  }
  if (rows < b->Nh) {
    const long long outside = (b->Nh - rows) * n;
    rb_off_zero_rows_outside_kernel<<<(unsigned)((outside + 255) / 256), 256, 0, ctx->stream>>>(b->d, b->ld, b->n, n, b->Nh,
                                                                                              row_lo, row_hi);
    RB_CUDA(cudaGetLastError());
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  b->n += n;
  return RB_OK;
}

extern "C" int rb_off_size(rb_ctx* ctx, int slot, int64_t* Nh, int* ncols) {
  if (!ctx) return RB_ERR_ARG;
  rb_off_blk_t* b = off_blk(ctx, slot);
  if (!b) RB_FAIL(RB_ERR_ARG, "rb_off_size: empty slot");
  if (Nh) *Nh = b->Nh;
  if (ncols) *ncols = b->n;
  return RB_OK;
}

extern "C" int rb_off_read(rb_ctx* ctx, int slot, int c0, int n, double* out) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  rb_off_blk_t* b = off_blk(ctx, slot);
  if (!b || c0 < 0 || n <= 0 || c0 + n > b->n || !out) RB_FAIL(RB_ERR_ARG, "rb_off_read: bad arguments");
  DevBuf tmp;
  RB_CUDA(tmp.alloc((size_t)b->Nh * n * sizeof(double)));
  rb_off_gather_kernel<<<(unsigned)((b->Nh * n + 255) / 256), 256, 0, ctx->stream>>>(b->d, b->ld, c0, n, tmp.as<double>(),
                                                                                    b->Nh);
  RB_CUDA(cudaGetLastError());
  RB_CUDA(cudaMemcpyAsync(out, tmp.p, (size_t)b->Nh * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  RB_CUDA(cudaStreamSynchronize(ctx->stream));
  return RB_OK;
}

extern "C" int rb_off_gram(rb_ctx* ctx, int blkS, int s0, int ns, int csrX, int blkS2, int t0, int ns2,
                           int64_t row_begin, int64_t row_end, double* C) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  rb_off_blk_t* S = off_blk(ctx, blkS);
  rb_off_blk_t* S2 = off_blk(ctx, blkS2);
  rb_off_csr_t* X = csrX >= 0 ? off_csr(ctx, csrX) : nullptr;
  if (!S || !S2 || !C || (csrX >= 0 && !X)) RB_FAIL(RB_ERR_ARG, "rb_off_gram: empty slot");
  if (S->Nh != S2->Nh || (X && X->Nh != S->Nh)) RB_FAIL(RB_ERR_ARG, "rb_off_gram: inconsistent number of dofs");
  if (s0 < 0 || ns <= 0 || s0 + ns > S->n || t0 < 0 || ns2 <= 0 || t0 + ns2 > S2->n)
    RB_FAIL(RB_ERR_ARG, "rb_off_gram: column range outside the block");
  if (row_begin < 0 || row_end > S->Nh || row_begin > row_end) RB_FAIL(RB_ERR_ARG, "rb_off_gram: bad row range");
  return gram_device(ctx, S->Nh, ns, ns2, S->d + s0, S->ld, S2->d + t0, S2->ld, X ? X->ip : nullptr,
                     X ? X->ix : nullptr, X ? X->dt : nullptr, row_begin, row_end, C, nullptr, nullptr, X ? X->sym : -1,
                     s0 == 0 && t0 == 0);
}

extern "C" int rb_off_spmm(rb_ctx* ctx, int csrA, int blkS, int s0, int n, int blkY, int y0) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  rb_off_csr_t* A = off_csr(ctx, csrA);
  rb_off_blk_t* S = off_blk(ctx, blkS);
  rb_off_blk_t* Y = off_blk(ctx, blkY);
  if (!A || !S || !Y) RB_FAIL(RB_ERR_ARG, "rb_off_spmm: empty slot");
  if (A->Nh != S->Nh || A->Nh != Y->Nh) RB_FAIL(RB_ERR_ARG, "rb_off_spmm: inconsistent number of dofs");
  if (s0 < 0 || n <= 0 || s0 + n > S->n || y0 < 0 || y0 + n > Y->cap)  // (columns never written are zero)
    RB_FAIL(RB_ERR_ARG, "rb_off_spmm: column range outside the block");
  if (S == Y && !(y0 >= s0 + n || y0 + n <= s0)) RB_FAIL(RB_ERR_ARG, "rb_off_spmm: source and target columns overlap");
  cudaStream_t st = ctx->stream;
  RB_CUDA(cudaEventRecord(ctx->ev[4], st));
  if (n <= 32) {
    const long long threads = A->Nh * n;
    rb_spmm_narrow_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(A->ip, A->ix, A->dt, S->d + s0, S->ld, n,
                                                                            Y->d + y0, Y->ld, 0, A->Nh);
  } else {
    const int nchunk = (n + 127) / 128;
    const long long blocks = (A->Nh * nchunk * 32 + 255) / 256;
    rb_spmm_kernel<<<(unsigned)blocks, 256, 0, st>>>(A->ip, A->ix, A->dt, S->d + s0, S->ld, n, Y->d + y0, Y->ld, 0, A->Nh);
  }
  RB_CUDA(cudaGetLastError());
  RB_CUDA(cudaEventRecord(ctx->ev[5], st));
  RB_CUDA(cudaStreamSynchronize(st));
  float a = 0;
  cudaEventElapsedTime(&a, ctx->ev[4], ctx->ev[5]);
  ctx->last_ms[0] = a;
  ctx->last_ms[1] = 0.f;
  ctx->last_ms[4] = a;
  ctx->last_launches = 1;
  if (Y->n < y0 + n) Y->n = y0 + n;
  if (y0 < Y->gs_n) Y->gs_n = y0;  // mirrored columns were overwritten: re-mirror them in the next Gram-Schmidt call
  return RB_OK;
}

// z[i] = 0 outside [rb, re): the rank's piece of a vector that is then summed over the ranks (= gathered)
__global__ void rb_zero_outside_kernel(double* __restrict__ z, long long n, long long rb, long long re) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (i < rb || i >= re) z[i] = 0.0;
}

extern "C" int rb_set_allreduce(rb_ctx* ctx, rb_allreduce_fn fn, void* user) {
  if (!ctx) return RB_ERR_ARG;
  ctx->allreduce = fn;
  ctx->allreduce_user = user;
  return RB_OK;
}

extern "C" int rb_off_gram_schmidt_rows(rb_ctx* ctx, int blkZ, int csrX, double* z, int append, int64_t row_begin,
                                        int64_t row_end, double* norm_out) {
  return rb_off_gram_schmidt_from(ctx, blkZ, csrX, z, append, 0, row_begin, row_end, norm_out);
}

// first_col > 0: the first columns of the block are not projected out -- lifting functions of non-homogeneous Dirichlet
// conditions, GS.apply(new, basis_functions[N_bc:]) in rbnics/reduction_methods/base/rb_reduction.py:125-139
extern "C" int rb_off_gram_schmidt_from(rb_ctx* ctx, int blkZ, int csrX, double* z, int append, int first_col,
                                        int64_t row_begin, int64_t row_end, double* norm_out) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  rb_off_blk_t* Z = off_blk(ctx, blkZ);
  rb_off_csr_t* X = csrX >= 0 ? off_csr(ctx, csrX) : nullptr;
  if (!Z || !z || (csrX >= 0 && !X)) RB_FAIL(RB_ERR_ARG, "rb_off_gram_schmidt: empty slot");
  if (first_col < 0 || first_col > Z->n) RB_FAIL(RB_ERR_ARG, "rb_off_gram_schmidt: first column outside the block");
  if (append && Z->n + 1 > Z->cap) RB_FAIL(RB_ERR_ARG, "rb_off_gram_schmidt: block capacity exceeded");
  cudaStream_t st = ctx->stream;
  const long long Nh = Z->Nh;
  if (row_begin < 0 || row_end > Nh || row_begin > row_end) RB_FAIL(RB_ERR_ARG, "rb_off_gram_schmidt: bad row range");
  const bool partitioned = !(row_begin == 0 && row_end == Nh);
  if (partitioned && !ctx->allreduce)
    RB_FAIL(RB_ERR_STATE, "rb_off_gram_schmidt_rows: a partial row range needs rb_set_allreduce");
  const long long rb = row_begin, nrows = row_end - row_begin;
  const unsigned nb1 = (unsigned)((Nh + 255) / 256);
  PoolBuf dz(ctx, 5), dNorm(ctx, 6), dXz(ctx, 7), dPart(ctx, 8);
  RB_CUDA(dz.alloc((size_t)Nh * sizeof(double)));
  RB_CUDA(dNorm.alloc(sizeof(double)));
  RB_CUDA(dPart.alloc(2 * GS_BLOCKS * sizeof(double)));
  if (X) RB_CUDA(dXz.alloc((size_t)Nh * sizeof(double)));
  RB_CUDA(cudaMemcpyAsync(dz.p, z, (size_t)Nh * sizeof(double), cudaMemcpyHostToDevice, st));
  int launches = 0;
  auto reduce = [&](double* buf, long long count) -> int {  // sum over the ranks (no-op on one rank)
    if (!ctx->allreduce) return RB_OK;
    RB_CUDA(cudaStreamSynchronize(st));
    if (ctx->allreduce(ctx->allreduce_user, buf, (int64_t)count))
      RB_FAIL(RB_ERR_CUDA, "rb_off_gram_schmidt: the all-reduce callback failed");
    return RB_OK;
  };
  // column mirrors (see rb_off_blk_t): bring them up to date with the block -- columns appended by rb_off_append or
  // overwritten by rb_off_spmm, a different operator slot, a NEW matrix in the same slot (generation), another row range
  if (!Z->zt) RB_CUDA(cudaMalloc((void**)&Z->zt, (size_t)Z->cap * Nh * sizeof(double)));
  if (X && !Z->xzt) RB_CUDA(cudaMalloc((void**)&Z->xzt, (size_t)Z->cap * Nh * sizeof(double)));
  if (Z->gs_csr != csrX || (X && Z->gs_csr_generation != X->generation) || Z->gs_row_begin != row_begin ||
      Z->gs_row_end != row_end)
    Z->gs_n = 0;
  auto mirror_column = [&](int k) -> int {  // zt[k] = column k of the block, xzt[k] = (X zt[k]) on the rank's rows
    rb_off_gather_kernel<<<nb1, 256, 0, st>>>(Z->d, Z->ld, k, 1, Z->zt + (size_t)k * Nh, Nh);
    if (X && nrows > 0)
      rb_spmm_narrow_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, st>>>(
          X->ip, X->ix, X->dt, Z->zt + (size_t)k * Nh, 1, 1, Z->xzt + (size_t)k * Nh, 1, row_begin, row_end);
    RB_CUDA(cudaGetLastError());
    launches += X ? 2 : 1;
    return RB_OK;
  };
  for (int k = Z->gs_n; k < Z->n; ++k)
    if (int rc = mirror_column(k)) return rc;
  Z->gs_n = Z->n;
  Z->gs_csr = csrX;
  Z->gs_csr_generation = X ? X->generation : -1;
  Z->gs_row_begin = row_begin;
  Z->gs_row_end = row_end;
  // single-pass MGS in the reference's order (rbnics/backends/basic/gram_schmidt.py:22-36), deterministic dots
  const double* xz = X ? Z->xzt : Z->zt;
  // dot of step 0, then one kernel per step: axpy of step k fused with the dots of step k+1 (N+1 launches, not 2N).
  // Partitioned: the kernels run on the rank's rows only and the per-CTA partials are summed over the ranks before the
  // next kernel reads them, so every rank subtracts the same d_k.
  double* part[2] = {dPart.as<double>(), dPart.as<double>() + GS_BLOCKS};
  double* zr = dz.as<double>() + rb;
  if (Z->n > first_col) {
    rb_dot_partial_kernel<<<GS_BLOCKS, GS_THREADS, 0, st>>>(zr, 1, xz + (size_t)first_col * Nh + rb, 1, nrows, part[0]);
    ++launches;
    if (int rc = reduce(part[0], GS_BLOCKS)) return rc;
  }
  for (int k = first_col; k < Z->n; ++k) {
    const int b = (k - first_col) & 1;
    rb_gs_step_kernel<<<GS_BLOCKS, GS_THREADS, 0, st>>>(part[b], Z->zt + (size_t)k * Nh + rb, zr,
                                                       k + 1 < Z->n ? xz + (size_t)(k + 1) * Nh + rb : nullptr, nrows,
                                                       part[b ^ 1]);
    ++launches;
    if (k + 1 < Z->n)
      if (int rc = reduce(part[b ^ 1], GS_BLOCKS)) return rc;
  }
  RB_CUDA(cudaGetLastError());
  if (partitioned) {  // gather the rows the other ranks projected: sum of the zero-padded pieces
    rb_zero_outside_kernel<<<nb1, 256, 0, st>>>(dz.as<double>(), Nh, row_begin, row_end);
    ++launches;
    if (int rc = reduce(dz.as<double>(), Nh)) return rc;
  }
  const double* dXzp = dz.as<double>();
  if (X) {
    if (nrows > 0)
      rb_spmm_narrow_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, st>>>(X->ip, X->ix, X->dt, dz.as<double>(), 1, 1,
                                                                            dXz.as<double>(), 1, row_begin, row_end);
    dXzp = dXz.as<double>();
    ++launches;
  }
  rb_dot_partial_kernel<<<GS_BLOCKS, GS_THREADS, 0, st>>>(zr, 1, dXzp + rb, 1, nrows, dPart.as<double>());
  ++launches;
  if (int rc = reduce(dPart.as<double>(), GS_BLOCKS)) return rc;
  rb_normalize_kernel<<<GS_BLOCKS, GS_THREADS, 0, st>>>(dPart.as<double>(), dz.as<double>(), Nh, dNorm.as<double>());
  ++launches;
  RB_CUDA(cudaGetLastError());
  if (append) {
    rb_off_scatter_kernel<<<nb1, 256, 0, st>>>(dz.as<double>(), 1, Z->d, Z->ld, Z->n, Nh);
    RB_CUDA(cudaGetLastError());
    ++launches;
  }
  double nrm = 0.0;
  RB_CUDA(cudaMemcpyAsync(z, dz.p, (size_t)Z->Nh * sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaMemcpyAsync(&nrm, dNorm.p, sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaStreamSynchronize(st));
  if (append) Z->n += 1;  // (its mirror is made by the next call, from the block itself)
  if (norm_out) *norm_out = nrm;
  ctx->last_launches = launches;
  return RB_OK;
}

extern "C" int rb_off_gram_schmidt(rb_ctx* ctx, int blkZ, int csrX, double* z, int append, double* norm_out) {
  if (!ctx) return RB_ERR_ARG;
  rb_off_blk_t* Z = off_blk(ctx, blkZ);
  if (!Z) RB_FAIL(RB_ERR_ARG, "rb_off_gram_schmidt: empty slot");
  return rb_off_gram_schmidt_rows(ctx, blkZ, csrX, z, append, 0, Z->Nh, norm_out);
}

// =====================================================================================================================
// Resident POD helpers: modes Z = S V, their X-norms from the DIAGONAL only, column scaling
// (rbnics/backends/basic/proper_orthogonal_decomposition_base.py:77-92, dolfin/wrapping/functions_list_mul.py:26-36)
// =====================================================================================================================
constexpr int CD_BLOCKS = 592;

// partial[block][c] = sum over the block's rows of Z[r, c] * Y[r, c]  (n <= 128 columns, lanes over columns)
__global__ void __launch_bounds__(256) rb_coldot_partial_kernel(const double* __restrict__ Z, int ldz,
                                                                const double* __restrict__ Y, int ldy, int n,
                                                                long long row_begin, long long row_end,
                                                                double* __restrict__ partial) {
  __shared__ double sh[8][128];
  const int lane = threadIdx.x & 31, rr = threadIdx.x >> 5;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (long long r = row_begin + blockIdx.x * 8LL + rr; r < row_end; r += 8LL * CD_BLOCKS) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int c = lane + 32 * t;
      if (c < n) acc[t] = fma(Z[r * ldz + c], Y[r * ldy + c], acc[t]);
    }
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) sh[rr][lane + 32 * t] = acc[t];
  __syncthreads();
  if (threadIdx.x < 128) {
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += sh[k][threadIdx.x];  // fixed order: deterministic
    if ((int)threadIdx.x < n) partial[(long long)blockIdx.x * 128 + threadIdx.x] = s;
  }
}

__global__ void rb_coldot_final_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
  const int c = threadIdx.x;
  if (c >= n) return;
  double s = 0.0;
  for (int b = 0; b < CD_BLOCKS; ++b) s += partial[(long long)b * 128 + c];
  out[c] = s;
}

__global__ void rb_scale_cols_kernel(double* __restrict__ Z, int ldz, int n, long long Nh, const double* __restrict__ sc) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= Nh * n) return;
  const long long r = t / n;
  const int c = (int)(t % n);
  Z[r * ldz + c] *= sc[c];
}

extern "C" int rb_off_lincomb(rb_ctx* ctx, int blkS, int s0, int ns, const double* V, int n, int blkZ, int z0) {
  return rb_off_lincomb_add(ctx, blkS, s0, ns, V, n, blkZ, z0, 0);
}

// accumulate != 0: Z[:, z0:z0+n] += S[:, s0:s0+ns] V on existing columns -- e.g. snapshot -= basis_functions * projection
// (rbnics/reduction_methods/base/time_dependent_rb_reduction.py:178-189) without the snapshots leaving the device
extern "C" int rb_off_lincomb_add(rb_ctx* ctx, int blkS, int s0, int ns, const double* V, int n, int blkZ, int z0,
                                  int accumulate) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  rb_off_blk_t* S = off_blk(ctx, blkS);
  rb_off_blk_t* Z = off_blk(ctx, blkZ);
  if (!S || !Z || !V) RB_FAIL(RB_ERR_ARG, "rb_off_lincomb: empty slot");
  if (S->Nh != Z->Nh) RB_FAIL(RB_ERR_ARG, "rb_off_lincomb: inconsistent number of dofs");
  if (s0 < 0 || ns <= 0 || s0 + ns > S->n || n <= 0 || z0 < 0 || z0 + n > Z->cap || S == Z)
    RB_FAIL(RB_ERR_ARG, "rb_off_lincomb: column range outside the block");
  if (accumulate && z0 + n > Z->n) RB_FAIL(RB_ERR_ARG, "rb_off_lincomb_add: the target columns do not exist yet");
  if (s0 & 1) RB_FAIL(RB_ERR_UNSUPPORTED, "rb_off_lincomb: the first source column must be even (16-byte aligned rows)");
  cudaStream_t st = ctx->stream;
  DevBuf dV;
  int ldv = n;
  if (int rc = upload_cols(ctx, V, ns, n, dV, &ldv)) return rc;
  RB_CUDA(cudaEventRecord(ctx->ev[4], st));
  if (int rc = lincomb_device(ctx, S->Nh, ns, n, S->d + s0, S->ld, dV.as<double>(), ldv, Z->d + z0, Z->ld, accumulate))
    return rc;
  RB_CUDA(cudaEventRecord(ctx->ev[5], st));
  RB_CUDA(cudaStreamSynchronize(st));
  float a = 0;
  cudaEventElapsedTime(&a, ctx->ev[4], ctx->ev[5]);
  ctx->last_ms[0] = a;
  ctx->last_ms[4] = a;
  ctx->last_launches = (n / LC_BN > 0) + (n % LC_BN > 0);
  if (Z->n < z0 + n) Z->n = z0 + n;
  if (z0 < Z->gs_n) Z->gs_n = z0;
  return RB_OK;
}

extern "C" int rb_off_column_norms2(rb_ctx* ctx, int blkZ, int z0, int n, int csrX, int64_t row_begin, int64_t row_end,
                                    double* out) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  rb_off_blk_t* Z = off_blk(ctx, blkZ);
  rb_off_csr_t* X = csrX >= 0 ? off_csr(ctx, csrX) : nullptr;
  if (!Z || !out || (csrX >= 0 && !X)) RB_FAIL(RB_ERR_ARG, "rb_off_column_norms2: empty slot");
  if (z0 < 0 || n <= 0 || n > 128 || z0 + n > Z->n) RB_FAIL(RB_ERR_ARG, "rb_off_column_norms2: bad column range");
  const long long Nh = Z->Nh;
  if (row_begin < 0 || row_end > Nh || row_begin > row_end) RB_FAIL(RB_ERR_ARG, "rb_off_column_norms2: bad row range");
  cudaStream_t st = ctx->stream;
  PoolBuf Ybuf(ctx, 0), part(ctx, 1), dOut(ctx, 2);
  RB_CUDA(part.alloc((size_t)CD_BLOCKS * 128 * sizeof(double)));
  RB_CUDA(dOut.alloc(128 * sizeof(double)));
  const double* Y = Z->d + z0;
  int ldy = Z->ld;
  RB_CUDA(cudaEventRecord(ctx->ev[4], st));
  if (X) {
    RB_CUDA(Ybuf.alloc((size_t)Nh * n * sizeof(double)));
    const long long nrows = row_end - row_begin;
    if (nrows > 0) {
      if (n <= 32) {
        rb_spmm_narrow_kernel<<<(unsigned)((nrows * n + 255) / 256), 256, 0, st>>>(X->ip, X->ix, X->dt, Z->d + z0, Z->ld, n,
                                                                                  Ybuf.as<double>(), n, row_begin, row_end);
      } else {
        const int nchunk = (n + 127) / 128;
        rb_spmm_kernel<<<(unsigned)((nrows * nchunk * 32 + 255) / 256), 256, 0, st>>>(
            X->ip, X->ix, X->dt, Z->d + z0, Z->ld, n, Ybuf.as<double>(), n, row_begin, row_end);
      }
      RB_CUDA(cudaGetLastError());
    }
    Y = Ybuf.as<double>();
    ldy = n;
  }
  rb_coldot_partial_kernel<<<CD_BLOCKS, 256, 0, st>>>(Z->d + z0, Z->ld, Y, ldy, n, row_begin, row_end, part.as<double>());
  rb_coldot_final_kernel<<<1, 128, 0, st>>>(part.as<double>(), n, dOut.as<double>());
  RB_CUDA(cudaGetLastError());
  RB_CUDA(cudaEventRecord(ctx->ev[5], st));
  if (ctx->allreduce) {
    RB_CUDA(cudaStreamSynchronize(st));
    if (ctx->allreduce(ctx->allreduce_user, dOut.as<double>(), n))
      RB_FAIL(RB_ERR_CUDA, "rb_off_column_norms2: the all-reduce callback failed");
  }
  RB_CUDA(cudaMemcpyAsync(out, dOut.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaStreamSynchronize(st));
  float a = 0;
  cudaEventElapsedTime(&a, ctx->ev[4], ctx->ev[5]);
  ctx->last_ms[0] = a;
  ctx->last_ms[4] = a;
  ctx->last_launches = X ? 3 : 2;
  return RB_OK;
}

extern "C" int rb_off_scale_columns(rb_ctx* ctx, int blkZ, int z0, int n, const double* scale) {
  if (!ctx) return RB_ERR_ARG;
  RB_CUDA(cudaSetDevice(ctx->device));
  rb_off_blk_t* Z = off_blk(ctx, blkZ);
  if (!Z || !scale) RB_FAIL(RB_ERR_ARG, "rb_off_scale_columns: empty slot");
  if (z0 < 0 || n <= 0 || z0 + n > Z->n) RB_FAIL(RB_ERR_ARG, "rb_off_scale_columns: bad column range");
  cudaStream_t st = ctx->stream;
  PoolBuf dSc(ctx, 2);
  RB_CUDA(dSc.alloc((size_t)n * sizeof(double)));
  RB_CUDA(cudaMemcpyAsync(dSc.p, scale, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  rb_scale_cols_kernel<<<(unsigned)((Z->Nh * n + 255) / 256), 256, 0, st>>>(Z->d + z0, Z->ld, n, Z->Nh, dSc.as<double>());
  RB_CUDA(cudaGetLastError());
  RB_CUDA(cudaStreamSynchronize(st));
  if (z0 < Z->gs_n) Z->gs_n = z0;
  ctx->last_launches = 1;
  return RB_OK;
}

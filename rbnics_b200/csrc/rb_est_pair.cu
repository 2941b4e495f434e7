// rb_est_pair.cu -- the error estimator as ONE dense GEMM over factor pairs (sm_100a, FP64 DMMA.8x8x4).
//
// rbnics/problems/elliptic/elliptic_rb_reduced_problem.py:35-64 (and the compliant / Stokes / parabolic siblings) evaluate
//   eps2(mu) = z^T G z,   z = concatenation of blocks  s_b(mu) * V[src_b : src_b + len_b],   V = [u | theta | 1].
// Blocks that scale the same slice of V form a class.  For a pair of classes (c, c') every term of the quadratic form is
//   (s_b s_b') * M[(b,b'), (n,m)] * (v_n v'_m):   a product of TWO entries of V' = [u | theta | 1 | 0] on the left, a
// constant, and a product of two entries of V' on the right.  With both symmetries folded into M (b >= b' and, inside a
// class pair c == c', n >= m) this is, per class pair, a plain dense GEMM
//   Y[p, col] = sum_k  A[p, k] M[k, col],   A[p, k] = V'[p, i1(k)] V'[p, i2(k)]        (k: block pairs)
//   eps2[p]  += sum_col Y[p, col] W[p, col], W[p, col] = V'[p, j1(col)] V'[p, j2(col)]  (col: entry pairs)
// with NO triangular tiles: the multiply-adds issued are the algorithmic ones (Q_a(Q_a+1)/2 * N(N+1)/2 + ...) plus the
// padding of K to 8 and of the columns to 64 per class pair (0.6 % at N = 100, Q_a = 50, Q_f = 10; the triangular
// 16-column tiles of rb_estimator_kernel issue 12 % more than the algorithmic count).  The A operand is never stored:
// every lane builds its fragment entries from the V' tile in shared memory (2 LDS.128 + 2 DMUL per k-step next to 16
// DMMAs), the two factor columns of the four k rows of a step travel with the B fragments in the same 1-D bulk copy.
// The orientation (which pair list is K, which one the columns) is chosen per class pair to minimise the padded work.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "rb_common.cuh"

constexpr int PR_REC_DOUBLES = 4 * EST_BN + 4;  // one k4-step of a 64-column tile: B fragments + 4 x (int, int) offsets
constexpr int PR_MAXSTAGE = 4;
constexpr int PR_THREADS = EST_THREADS + 32;  // 8 consumer warps + the producer warp of the record stream
constexpr int PR_MAX_TILES = 4096;

// ---------------------------------------------------------------------------------------------------------------------
// packing: dense G (device) -> record stream
// ---------------------------------------------------------------------------------------------------------------------
struct PrDesc {  // one padded row or column of a class-pair matrix
  int x, y;      // kind 0: blocks (b1, b2); kind 1: entries (n, m)
  int kind;      // -1: padding
  int voff1, voff2;  // byte offsets of the two factors' columns in the V' tile (column * 16)
};

struct PrPackArgs {
  const double* G;
  double* out;
  long long total_rec;
  const int* rec_tile;   // [total_rec]
  const int* rec_ks;     // [total_rec] k-step inside the tile's stream
  const int* tile_row0;  // [n_tile] first row descriptor of the tile's class pair
  const int* tile_same;  // [n_tile] 1: c == c' (folded in the entry pair as well)
  const PrDesc* row;     // row descriptors of all class pairs
  const PrDesc* col;     // [n_tile * 64]
  const int* ubegin;     // [nblk] dense index of a block's first entry
  int Kz;
};

__global__ void rb_pair_pack_kernel(PrPackArgs a) {
  const long long total = a.total_rec * PR_REC_DOUBLES;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long rec = t / PR_REC_DOUBLES;
    const int r = (int)(t % PR_REC_DOUBLES);
    const int tile = a.rec_tile[rec], ks = a.rec_ks[rec];
    if (r >= 4 * EST_BN) {  // factor offsets of k row 4 ks + (r - 256)
      const PrDesc d = a.row[a.tile_row0[tile] + 4 * ks + (r - 4 * EST_BN)];
      reinterpret_cast<int2*>(a.out)[t] = make_int2(d.voff1, d.voff2);
      continue;
    }
    // B-fragment order [wn (2)][pair (2)][lane (32)][e (2)] <- M[row 4 ks + lane % 4][column wn*32 + (2 pair + e)*8 + lane / 4]
    const int wn = r >> 7, pair = (r >> 6) & 1, lane = (r >> 1) & 31, e = r & 1;
    const PrDesc dr = a.row[a.tile_row0[tile] + 4 * ks + (lane & 3)];
    const PrDesc dc = a.col[(long long)tile * EST_BN + wn * 32 + (2 * pair + e) * 8 + (lane >> 2)];
    double v = 0.0;
    if (dr.kind >= 0 && dc.kind >= 0) {
      const PrDesc& db = dr.kind == 0 ? dr : dc;  // the block pair
      const PrDesc& de = dr.kind == 0 ? dc : dr;  // the entry pair
      const int b1 = db.x, b2 = db.y, n = de.x, m = de.y;
      const long long Kz = a.Kz;
      const long long i = a.ubegin[b1] + n, j = a.ubegin[b2] + m;
      if (!a.tile_same[tile]) {
        v = a.G[i * Kz + j] + a.G[j * Kz + i];
      } else if (b1 == b2) {
        v = a.G[i * Kz + j];
        if (n != m) v += a.G[j * Kz + i];
      } else {
        v = a.G[i * Kz + j] + a.G[j * Kz + i];
        if (n != m) {
          const long long i2 = a.ubegin[b1] + m, j2 = a.ubegin[b2] + n;
          v += a.G[i2 * Kz + j2] + a.G[j2 * Kz + i2];
        }
      }
    }
    a.out[t] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------------
struct PairArgs {
  const double* stream;
  const long long* tile_off;  // [n_tile] offset of the tile's first record, in doubles
  const int* tile_nrec;       // [n_tile]
  const int2* col;            // [n_tile * 64] byte offsets of the two factors of a column
  const int* split_begin;     // [nsplit + 1]
  const double* U;            // [B][ldu]
  const double* theta;        // [B][QT]
  double* partial;            // [nsplit][B]
  long long B;
  long long rb0;  // first row block of this launch
  int vecN, ldu, QT, LVP, nsplit, SK, nstage, n_tile;
};

__host__ __device__ inline size_t pair_smem_bytes(int nstage, int SK, int LVP, int n_tile) {
  size_t b = (size_t)nstage * SK * PR_REC_DOUBLES * 8;  // record ring
  b += 2 * PR_MAXSTAGE * 8;                             // full + empty barriers
  b += (size_t)EST_BM * LVP * 8;                        // V' tile
  b += (size_t)n_tile * 12;                             // tile offsets + record counts
  return b + 16;
}

struct PrFrag {
  double2 b[4];  // B fragments of the eight 8-column tiles
  double av[2];  // A entries of the lane's two parameter rows (products of the two factors)
};
struct PrRaw {
  double2 p, q;  // the two factors of the A entries
};

__device__ __forceinline__ void pr_load_b(PrFrag& f, const double* rec, int lane, int2& idx) {
  idx = *reinterpret_cast<const int2*>(rec + 4 * EST_BN + (lane & 3));
#pragma unroll
  for (int h = 0; h < 4; ++h) f.b[h] = *reinterpret_cast<const double2*>(rec + 64 * h + 2 * lane);
}
__device__ __forceinline__ void pr_load_a(PrRaw& r, const char* vlane, int2 idx) {
  r.p = *reinterpret_cast<const double2*>(vlane + idx.x);
  r.q = *reinterpret_cast<const double2*>(vlane + idx.y);
}
__device__ __forceinline__ void pr_mul_a(PrFrag& f, const PrRaw& r) {
  f.av[0] = r.p.x * r.q.x;
  f.av[1] = r.p.y * r.q.y;
}
// DMMAs [4 H, 4 H + 4) of the 16 of one record: 8-column tiles 2 H and 2 H + 1, both row tiles
template <int H>
__device__ __forceinline__ void pr_mma(double (&T)[2][8][2], const PrFrag& f) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    dmma884(T[mt][2 * H][0], T[mt][2 * H][1], f.av[mt], f.b[H].x);
    dmma884(T[mt][2 * H + 1][0], T[mt][2 * H + 1][1], f.av[mt], f.b[H].y);
  }
}

// CTA = 128 parameters x a contiguous range of column tiles; 8 warps, each one 16 parameters x all 64 columns of the
// tile (warp tile 16 x 64): one multiplication builds the A entry of 8 DMMAs -- a DMUL between DMMAs costs far more
// than its two cycles of the FP64 pipe (measured: 1.7 % of the kernel per DMUL and record), so the warp tile is as
// wide as the column tile.  V' tile layout: [row pair rq = warp*8 + g][column][slot s] holding parameter row
// warp*16 + g + 8 s; LVP = 4 mod 8, so the eight lanes of a quarter warp (two g x four consecutive columns) read eight
// different 16-byte bank groups.
__global__ void __launch_bounds__(PR_THREADS, 1) rb_estimator_pair_kernel(PairArgs a) {
  extern __shared__ __align__(128) unsigned char pr_smem_raw[];
  const int SK = a.SK, NSTAGE = a.nstage, LVP = a.LVP;
  const int STAGE_DOUBLES = SK * PR_REC_DOUBLES;
  double* Gs = reinterpret_cast<double*>(pr_smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(Gs + (size_t)NSTAGE * STAGE_DOUBLES);
  uint64_t* empty_bar = full_bar + PR_MAXSTAGE;
  double* Vs = reinterpret_cast<double*>(empty_bar + PR_MAXSTAGE);
  long long* s_tile_off = reinterpret_cast<long long*>(Vs + (size_t)EST_BM * LVP);
  int* s_tile_nrec = reinterpret_cast<int*>(s_tile_off + a.n_tile);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int split = blockIdx.x % a.nsplit;
  const long long rb = a.rb0 + blockIdx.x / a.nsplit;
  const long long p0 = rb * EST_BM;
  const int ct_begin = a.split_begin[split], ct_end = a.split_begin[split + 1];

  // ---- V' tile: [u | theta | 1 | 0 ...]
  {
    const int vecN = a.vecN, QT = a.QT;
#pragma unroll 4
    for (int it = tid; it < 64 * LVP; it += PR_THREADS) {
      const int rq = it / LVP, c = it - rq * LVP;
      const int rbase = (rq >> 3) * 16 + (rq & 7);
      double v[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const long long p = p0 + rbase + 8 * s;
        double x = 0.0;
        if (p < a.B) {
          if (c < vecN)
            x = a.U[p * a.ldu + c];
          else if (c < vecN + QT)
            x = a.theta[p * QT + (c - vecN)];
          else if (c == vecN + QT)
            x = 1.0;
        }
        v[s] = x;
      }
      *reinterpret_cast<double2*>(Vs + (size_t)it * 2) = make_double2(v[0], v[1]);
    }
  }
  for (int t = ct_begin + tid; t < ct_end; t += PR_THREADS) {
    s_tile_off[t] = a.tile_off[t];
    s_tile_nrec[t] = a.tile_nrec[t];
  }
  if (tid == 0) {
    for (int s = 0; s < PR_MAXSTAGE; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, EST_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  // ---- record stream: 1-D bulk copies, a stage (slot of the ring) holds up to SK records of ONE tile.  Warp 8 is the
  // producer: one lane walks the CTA's tiles, waits until the eight consumer warps have released a slot (empty
  // barrier) and refills it; a consumer's stage hand-over is then one arrive and one (normally satisfied) wait.
  if (warp == EST_WARPS) {
    if (lane == 0) {
      int slot = 0;
      uint32_t eph = 1;  // parity that completes immediately on a fresh barrier: the first NSTAGE fills do not wait
      for (int t = ct_begin; t < ct_end; ++t) {
        const int nrec = s_tile_nrec[t];
        const double* src = a.stream + s_tile_off[t];
        for (int r = 0; r < nrec; r += SK) {
          mbar_wait(empty_bar + slot, eph);
          const uint32_t bytes = (uint32_t)min(SK, nrec - r) * (PR_REC_DOUBLES * 8u);
          mbar_arrive_expect_tx(full_bar + slot, bytes);
          bulk_g2s(Gs + (size_t)slot * STAGE_DOUBLES, src + (long long)r * PR_REC_DOUBLES, bytes, full_bar + slot);
          if (++slot == NSTAGE) {
            slot = 0;
            eph ^= 1u;
          }
        }
      }
    }
    return;
  }

  const int g = lane >> 2, q4 = lane & 3;
  const char* vlane = reinterpret_cast<const char*>(Vs) + (size_t)((warp * 8 + g) * LVP) * 16;
  double rowacc[2] = {0.0, 0.0};

  if (ct_begin < ct_end) {
    double T[2][8][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) T[mt][nt][0] = T[mt][nt][1] = 0.0;
    int ct = ct_begin;
    int rem = s_tile_nrec[ct];  // records of the tile not yet consumed, counted in whole stages
    int stage = 0, nin = min(SK, rem);
    uint32_t phase = 0;
    const double* cur = Gs;
    PrFrag f0, f1;
    PrRaw raw;
    int2 idx;
    mbar_wait(full_bar, 0);
    pr_load_b(f0, cur, lane, idx);
    pr_load_a(raw, vlane, idx);
    pr_mul_a(f0, raw);

    // A record inside a stage: the fragments of the NEXT record are requested (and its A entries multiplied) while the
    // 16 DMMAs of this one issue; nothing but the loop counter sits between two bursts.
    auto body = [&](const PrFrag& fc, PrFrag& fn, const double* nxt) {
      pr_load_b(fn, nxt, lane, idx);
      pr_mma<0>(T, fc);
      pr_load_a(raw, vlane, idx);
      pr_mma<1>(T, fc);
      pr_mma<2>(T, fc);
      pr_mul_a(fn, raw);
      pr_mma<3>(T, fc);
    };
    // The last record of a stage: the next one (if any) is the first of the next slot.  Returns false at the end of the
    // CTA's work.
    auto stage_end = [&](const PrFrag& fc, PrFrag& fn) -> bool {
      const bool last_in_tile = (rem == nin);
      const bool have_next = !(last_in_tile && ct + 1 == ct_end);
      int ns = stage + 1;
      uint32_t nph = phase;
      if (ns == NSTAGE) {
        ns = 0;
        nph ^= 1u;
      }
      const double* nxt = Gs + (size_t)ns * STAGE_DOUBLES;
      if (have_next) {
        mbar_wait(full_bar + ns, nph);
        pr_load_b(fn, nxt, lane, idx);
      }
      pr_mma<0>(T, fc);
      if (have_next) pr_load_a(raw, vlane, idx);
      pr_mma<1>(T, fc);
      pr_mma<2>(T, fc);
      if (have_next) pr_mul_a(fn, raw);
      pr_mma<3>(T, fc);
      cur = nxt;
      // stage consumed: release the slot to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar + stage);
      rem -= nin;
      if (last_in_tile) {
        // epilogue of the column tile: rowacc[p] += sum_c Y[p, c] * W[p, c]
        const int cbase = ct * EST_BN + q4 * 2;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int2 jj = __ldg(a.col + cbase + nt * 8 + e);
            const double2 w = *reinterpret_cast<const double2*>(vlane + jj.x);
            const double2 x = *reinterpret_cast<const double2*>(vlane + jj.y);
            rowacc[0] = fma(T[0][nt][e], w.x * x.x, rowacc[0]);
            rowacc[1] = fma(T[1][nt][e], w.y * x.y, rowacc[1]);
            T[0][nt][e] = T[1][nt][e] = 0.0;
          }
        ++ct;
        if (ct < ct_end) rem = s_tile_nrec[ct];
      }
      stage = ns;
      phase = nph;
      nin = min(SK, rem);
      return have_next;
    };
    // per stage: f0 holds the first record; every stage has an even number of records (K is padded to 8 per class pair
    // and SK is even), so the fragment buffers play the same roles in every stage and no registers move at a hand-over
    while (true) {
      for (int i = nin - 2; i > 0; i -= 2) {
        body(f0, f1, cur + PR_REC_DOUBLES);
        body(f1, f0, cur + 2 * PR_REC_DOUBLES);
        cur += 2 * PR_REC_DOUBLES;
      }
      body(f0, f1, cur + PR_REC_DOUBLES);
      if (!stage_end(f1, f0)) break;
    }
  }

  // every warp owns its 16 parameters: reduce over the quad (columns) and store
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    rowacc[j] += __shfl_xor_sync(0xffffffffu, rowacc[j], 1);
    rowacc[j] += __shfl_xor_sync(0xffffffffu, rowacc[j], 2);
  }
  if (q4 == 0) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const long long p = p0 + warp * 16 + g + 8 * j;
      if (p < a.B) a.partial[(long long)split * a.B + p] = rowacc[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------------
void rb_pair_free(rb_ctx* ctx) {
  rb_pair_t& pr = ctx->pair;
  if (pr.dStream) cudaFree(pr.dStream);
  if (pr.dTileOff) cudaFree(pr.dTileOff);
  if (pr.dTileNrec) cudaFree(pr.dTileNrec);
  if (pr.dCol) cudaFree(pr.dCol);
  if (pr.dSplitBeginTail) cudaFree(pr.dSplitBeginTail);
  pr.dSplitBeginTail = nullptr;
  pr.tail_cap = 0;
  pr.dStream = nullptr;
  pr.dTileOff = nullptr;
  pr.dTileNrec = nullptr;
  pr.dCol = nullptr;
  pr.ok = false;
}

static int roundup_i(int x, int m) { return (x + m - 1) / m * m; }

// Builds the record stream from the dense G on the device (dG, Kz x Kz, row-major).  Returns RB_OK also when this form
// is not used (RB_EST_KERNEL=fold, or a V' tile that does not fit): ctx->pair.ok tells.
int rb_pair_setup(rb_ctx* ctx, int nblk, const int* scale_norm, const int* blk_src_off, const int* blk_len,
                  const int* cls, const int* ubegin, int v_theta_off, int v_theta_len, int solution_len,
                  const double* dG, int Kz) {
  rb_pair_free(ctx);
  rb_pair_t& pr = ctx->pair;
  if (const char* k = getenv("RB_EST_KERNEL"))
    if (!strcmp(k, "fold")) return RB_OK;
  const int QT = ctx->Ql + ctx->Qr;
  const int vecN = solution_len;
  const int col_one = vecN + QT, col_zero = vecN + QT + 1;
  int LVP = vecN + QT + 2;  // pitch of the V' tile in columns: 4 mod 8 (bank groups, see the kernel)
  while (LVP % 8 != 4) ++LVP;
  auto vmap = [&](int vidx) {  // index into V = [u | theta[v_off : v_off + v_len] | 1]  ->  column of V'
    if (vidx < vecN) return vidx;
    if (vidx < vecN + v_theta_len) return vecN + v_theta_off + (vidx - vecN);
    return col_one;
  };
  auto smap = [&](int sc) { return sc < 0 ? col_one : vecN + sc; };

  std::vector<int> classes;
  for (int b = 0; b < nblk; ++b)
    if (cls[b] == b) classes.push_back(b);
  std::vector<PrDesc> rows, cols;
  std::vector<int> tile_row0, tile_same, tile_nrec, rec_tile, rec_ks;
  std::vector<long long> tile_off;
  const PrDesc pad = {0, 0, -1, col_zero * 16, col_zero * 16};
  for (size_t ci = 0; ci < classes.size(); ++ci)
    for (size_t cj = 0; cj <= ci; ++cj) {
      const int c1 = classes[ci], c2 = classes[cj];
      const bool same = (ci == cj);
      std::vector<int> m1, m2;
      for (int b = 0; b < nblk; ++b) {
        if (cls[b] == c1) m1.push_back(b);
        if (cls[b] == c2) m2.push_back(b);
      }
      // the two pair lists: block pairs (scales) and entry pairs (slices of V)
      std::vector<PrDesc> bp, ep;
      for (size_t i = 0; i < m1.size(); ++i)
        for (size_t j = 0; j < (same ? i + 1 : m2.size()); ++j) {
          const int b1 = m1[i], b2 = m2[j];
          bp.push_back({b1, b2, 0, smap(scale_norm[b1]) * 16, smap(scale_norm[b2]) * 16});
        }
      const int L1 = blk_len[c1], L2 = blk_len[c2];
      for (int n = 0; n < L1; ++n)
        for (int m = 0; m < (same ? n + 1 : L2); ++m)
          ep.push_back({n, m, 1, vmap(blk_src_off[c1] + n) * 16, vmap(blk_src_off[c2] + m) * 16});
      // orientation: K = the list that gives less padded work
      const long long cost_bp_k = (long long)roundup_i((int)bp.size(), 8) * roundup_i((int)ep.size(), EST_BN);
      const long long cost_ep_k = (long long)roundup_i((int)ep.size(), 8) * roundup_i((int)bp.size(), EST_BN);
      const std::vector<PrDesc>& kl = cost_bp_k <= cost_ep_k ? bp : ep;
      const std::vector<PrDesc>& cl = cost_bp_k <= cost_ep_k ? ep : bp;
      const int row0 = (int)rows.size();
      rows.insert(rows.end(), kl.begin(), kl.end());
      rows.resize(row0 + roundup_i((int)kl.size(), 8), pad);  // an even number of records (see the kernel's stage loop)
      const int nrec = roundup_i((int)kl.size(), 8) / 4;
      for (size_t j0 = 0; j0 < cl.size(); j0 += EST_BN) {
        const int tile = (int)tile_row0.size();
        tile_row0.push_back(row0);
        tile_same.push_back(same ? 1 : 0);
        tile_nrec.push_back(nrec);
        tile_off.push_back((long long)rec_tile.size() * PR_REC_DOUBLES);
        for (int j = 0; j < EST_BN; ++j) cols.push_back(j0 + j < cl.size() ? cl[j0 + j] : pad);
        for (int ks = 0; ks < nrec; ++ks) {
          rec_tile.push_back(tile);
          rec_ks.push_back(ks);
        }
      }
    }
  const int n_tile = (int)tile_row0.size();
  if (n_tile == 0 || n_tile > PR_MAX_TILES) return RB_OK;
  int nstage = 2;
  if (const char* k = getenv("RB_PAIR_NSTAGE")) nstage = std::min(PR_MAXSTAGE, std::max(2, atoi(k)));  // tuning knob
  int max_nrec = 1;
  for (int t = 0; t < n_tile; ++t) max_nrec = std::max(max_nrec, tile_nrec[t]);
  int SK = std::min(16, max_nrec);  // even, like every tile's record count
  if (const char* k = getenv("RB_PAIR_SK")) SK = std::min(max_nrec, std::max(2, atoi(k) & ~1));
  while (SK > 2 && pair_smem_bytes(nstage, SK, LVP, n_tile) > 227 * 1024) SK -= 2;
  if (pair_smem_bytes(nstage, SK, LVP, n_tile) > 227 * 1024) return RB_OK;  // V' tile too wide: the folded kernel serves

  const long long total_rec = (long long)rec_tile.size();
  const size_t stream_doubles = (size_t)total_rec * PR_REC_DOUBLES;
  cudaError_t e = cudaMalloc(&pr.dStream, stream_doubles * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&pr.dTileOff, n_tile * sizeof(long long));
  if (e == cudaSuccess) e = cudaMalloc(&pr.dTileNrec, n_tile * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&pr.dCol, (size_t)n_tile * EST_BN * sizeof(int2));
  PrDesc *dRow = nullptr, *dColDesc = nullptr;
  int* dTab = nullptr;
  std::vector<int> tab;  // rec_tile | rec_ks | tile_row0 | tile_same | ubegin
  const size_t o_rec_ks = rec_tile.size(), o_row0 = o_rec_ks + rec_ks.size(), o_same = o_row0 + n_tile,
               o_ubegin = o_same + n_tile;
  tab.insert(tab.end(), rec_tile.begin(), rec_tile.end());
  tab.insert(tab.end(), rec_ks.begin(), rec_ks.end());
  tab.insert(tab.end(), tile_row0.begin(), tile_row0.end());
  tab.insert(tab.end(), tile_same.begin(), tile_same.end());
  tab.insert(tab.end(), ubegin, ubegin + nblk);
  std::vector<int2> colv(cols.size());
  for (size_t i = 0; i < cols.size(); ++i) colv[i] = make_int2(cols[i].voff1, cols[i].voff2);
  if (e == cudaSuccess) e = cudaMalloc(&dRow, rows.size() * sizeof(PrDesc));
  if (e == cudaSuccess) e = cudaMalloc(&dColDesc, cols.size() * sizeof(PrDesc));
  if (e == cudaSuccess) e = cudaMalloc(&dTab, tab.size() * sizeof(int));
  cudaStream_t st = ctx->stream;
  if (e == cudaSuccess) e = cudaMemcpyAsync(dRow, rows.data(), rows.size() * sizeof(PrDesc), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dColDesc, cols.data(), cols.size() * sizeof(PrDesc), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dTab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(pr.dTileOff, tile_off.data(), n_tile * sizeof(long long), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(pr.dTileNrec, tile_nrec.data(), n_tile * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(pr.dCol, colv.data(), colv.size() * sizeof(int2), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    PrPackArgs pa;
    pa.G = dG;
    pa.out = pr.dStream;
    pa.total_rec = total_rec;
    pa.rec_tile = dTab;
    pa.rec_ks = dTab + o_rec_ks;
    pa.tile_row0 = dTab + o_row0;
    pa.tile_same = dTab + o_same;
    pa.row = dRow;
    pa.col = dColDesc;
    pa.ubegin = dTab + o_ubegin;
    pa.Kz = Kz;
    const long long nthr = total_rec * PR_REC_DOUBLES;
    const int blocks = (int)std::min<long long>((nthr + 255) / 256, (long long)ctx->sm_count * 16);
    rb_pair_pack_kernel<<<blocks, 256, 0, st>>>(pa);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the tables are local vectors
  if (dRow) cudaFree(dRow);
  if (dColDesc) cudaFree(dColDesc);
  if (dTab) cudaFree(dTab);
  const size_t smem = pair_smem_bytes(nstage, SK, LVP, n_tile);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(rb_estimator_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    rb_pair_free(ctx);
    RB_FAIL(RB_ERR_CUDA, std::string("rb_set_estimator (pair form): ") + cudaGetErrorString(e));
  }
  pr.n_tile = n_tile;
  pr.LVP = LVP;
  pr.SK = SK;
  pr.nstage = nstage;
  pr.vecN = vecN;
  pr.issued_macs = total_rec * 4 * EST_BN;
  pr.h_tile_cost = tile_nrec;
  pr.ok = true;
  return RB_OK;
}

// `nsum` partial sums per parameter are added by the finalize kernel; the launch plan (one launch, or whole waves at
// the coarse split + the remaining row blocks at a finer one) was made by prepare_estimator_splits
int rb_pair_launch(rb_ctx* ctx, const double* Uvec, int ldu, const double* theta, long long rows, int nsum) {
  rb_pair_t& pr = ctx->pair;
  PairArgs a;
  a.stream = pr.dStream;
  a.tile_off = pr.dTileOff;
  a.tile_nrec = pr.dTileNrec;
  a.col = pr.dCol;
  a.U = Uvec;
  a.theta = theta;
  a.partial = ctx->dPart;
  a.B = rows;
  a.vecN = pr.vecN;
  a.ldu = ldu;
  a.QT = ctx->QT;
  a.LVP = pr.LVP;
  a.SK = pr.SK;
  a.nstage = pr.nstage;
  a.n_tile = pr.n_tile;
  const size_t smem = pair_smem_bytes(pr.nstage, pr.SK, pr.LVP, pr.n_tile);
  const long long nrb = (rows + EST_BM - 1) / EST_BM;
  const bool planned = pr.tail_nrb > 0 && nrb == pr.bulk_nrb + pr.tail_nrb;  // (a shorter last chunk: one launch)
  if (nsum != pr.bulk_nsplit)  // the parts write different numbers of partial sums per parameter: the others are zero
    RB_CUDA(cudaMemsetAsync(ctx->dPart, 0, (size_t)nsum * rows * sizeof(double), ctx->stream));
  a.split_begin = ctx->dSplitBegin;
  a.nsplit = pr.bulk_nsplit;
  a.rb0 = 0;
  rb_estimator_pair_kernel<<<(unsigned)((planned ? pr.bulk_nrb : nrb) * pr.bulk_nsplit), PR_THREADS, smem, ctx->stream>>>(a);
  RB_CUDA(cudaGetLastError());
  if (planned) {
    a.split_begin = pr.dSplitBeginTail;
    a.nsplit = pr.tail_nsplit;
    a.rb0 = pr.bulk_nrb;
    rb_estimator_pair_kernel<<<(unsigned)(pr.tail_nrb * pr.tail_nsplit), PR_THREADS, smem, ctx->stream>>>(a);
    RB_CUDA(cudaGetLastError());
  }
  return RB_OK;
}

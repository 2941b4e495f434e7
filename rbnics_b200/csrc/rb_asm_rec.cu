// rb_asm_rec.cu -- affine assembly  Ab[p][e] = sum_q Theta[p][q] A_q[e]  as a DMMA GEMM fed by a producer warp.
//
// product order 1 of rbnics/backends/online/basic/product.py:22-33 (+ sum.py:8-9) for a whole chunk of parameters.
// Same arithmetic as rb_assemble_kernel (rb_sweep.cu: k ascending per entry, FP64 DMMA accumulation) in the loop
// organisation that the pair-form estimator (rb_est_pair.cu) arrived at:
//   * A_q is packed ONCE per system (rb_asmrec_setup) into records = one k4-step of a 64-entry tile in DMMA B-fragment
//     order (2 KB); a tile is one contiguous run of QlP/4 records and travels with ONE 1-D bulk copy (TMA) into a ring
//     of tile slots; warp 8 is the producer (empty / full mbarriers), nobody else touches the copy engine;
//   * 8 consumer warps, warp tile 16 parameters x all 64 entries of the tile: the Theta fragment of a k-step is one
//     LDS.128 (tile layout [row pair][k][2], pitch = 4 mod 8: conflict-free), the B fragments four;
//   * the fragments of the next record are requested while the 16 DMMAs of the current one issue; a tile with an odd
//     number of records hands the fragment double-buffer over in the other role, so the tile loop exists in both
//     roles (no register moves at a tile boundary);
//   * the 16 x 64 result of a tile goes out with streaming 16-byte stores straight from the accumulators.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "rb_common.cuh"

constexpr int AR_REC_DOUBLES = 4 * ASM_BE;  // 256 doubles = one k4-step of a 64-entry tile
constexpr int AR_MAXSLOT = 6;
constexpr int AR_THREADS = 32 * 9;

// Aq [QlP][EP] (entry-major per term) -> records [tile][ks][h (4)][lane (32)][e (2)] <- Aq[4 ks + lane % 4][tile*64 + (2 h + e)*8 + lane / 4]
__global__ void rb_asmrec_pack_kernel(const double* __restrict__ Aq, double* __restrict__ rec, int QlP, int EP) {
  const int nrec = QlP / 4;
  const long long total = (long long)(EP / ASM_BE) * nrec * AR_REC_DOUBLES;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(t % AR_REC_DOUBLES);
    const long long rr = t / AR_REC_DOUBLES;
    const int ks = (int)(rr % nrec), tile = (int)(rr / nrec);
    const int h = r >> 6, lane = (r >> 1) & 31, e = r & 1;
    const int q = 4 * ks + (lane & 3);
    const int ent = tile * ASM_BE + (2 * h + e) * 8 + (lane >> 2);
    rec[t] = Aq[(long long)q * EP + ent];
  }
}

struct AsmRecArgs {
  const double* theta;  // [count][QT]
  const double* rec;    // packed A_q
  double* Ab;           // [count][EP]
  long long count;
  int QT, Ql, nrec, KP, EP, nslot;
};

__host__ __device__ inline size_t asmrec_smem_bytes(int nslot, int nrec, int KP) {
  return (size_t)nslot * nrec * AR_REC_DOUBLES * 8 + 2 * AR_MAXSLOT * 8 + (size_t)ASM_BM * KP * 8 + 16;
}

struct ArFrag {
  double2 b[4];
  double2 a;  // Theta entries of the lane's two parameter rows
};

__device__ __forceinline__ void ar_load(ArFrag& f, const double* rec, const char* tlane, int lane, int koff) {
#pragma unroll
  for (int h = 0; h < 4; ++h) f.b[h] = *reinterpret_cast<const double2*>(rec + 64 * h + 2 * lane);
  f.a = *reinterpret_cast<const double2*>(tlane + koff);
}
__device__ __forceinline__ void ar_mma(double (&T)[2][8][2], const ArFrag& f) {
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    dmma884(T[0][2 * h][0], T[0][2 * h][1], f.a.x, f.b[h].x);
    dmma884(T[0][2 * h + 1][0], T[0][2 * h + 1][1], f.a.x, f.b[h].y);
    dmma884(T[1][2 * h][0], T[1][2 * h][1], f.a.y, f.b[h].x);
    dmma884(T[1][2 * h + 1][0], T[1][2 * h + 1][1], f.a.y, f.b[h].y);
  }
}

__global__ void __launch_bounds__(AR_THREADS, 1) rb_assemble_rec_kernel(AsmRecArgs a) {
  extern __shared__ __align__(128) unsigned char ar_smem_raw[];
  const int nrec = a.nrec, KP = a.KP, NSLOT = a.nslot;
  const int SLOT_DOUBLES = nrec * AR_REC_DOUBLES;
  double* ring = reinterpret_cast<double*>(ar_smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + (size_t)NSLOT * SLOT_DOUBLES);
  uint64_t* empty_bar = full_bar + AR_MAXSLOT;
  double* Ts = reinterpret_cast<double*>(empty_bar + AR_MAXSLOT);  // Theta tile [row pair (64)][k (KP)][2]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long p0 = (long long)blockIdx.y * ASM_BM;
  const int ntile = a.EP / ASM_BE;

  for (int it = tid; it < 64 * KP; it += AR_THREADS) {
    const int rq = it / KP, k = it - rq * KP;
    const int rbase = (rq >> 3) * 16 + (rq & 7);
    double v[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const long long p = p0 + rbase + 8 * s;
      v[s] = (p < a.count && k < a.Ql) ? __ldg(a.theta + p * a.QT + k) : 0.0;
    }
    *reinterpret_cast<double2*>(Ts + (size_t)it * 2) = make_double2(v[0], v[1]);
  }
  if (tid == 0) {
    for (int s = 0; s < AR_MAXSLOT; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 8);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 8) {  // producer: one bulk copy per tile
    if (lane == 0) {
      int slot = 0;
      uint32_t eph = 1;
      const uint32_t bytes = (uint32_t)SLOT_DOUBLES * 8u;
      for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
        mbar_wait(empty_bar + slot, eph);
        mbar_arrive_expect_tx(full_bar + slot, bytes);
        bulk_g2s(ring + (size_t)slot * SLOT_DOUBLES, a.rec + (long long)t * SLOT_DOUBLES, bytes, full_bar + slot);
        if (++slot == NSLOT) {
          slot = 0;
          eph ^= 1u;
        }
      }
    }
    return;
  }
  if ((int)blockIdx.x >= ntile) return;

  const int g = lane >> 2, q4 = lane & 3;
  const char* tlane = reinterpret_cast<const char*>(Ts) + (size_t)((warp * 8 + g) * KP + q4) * 16;
  const long long prow0 = p0 + warp * 16 + g;  // + 8 mt
  double T[2][8][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) T[mt][nt][0] = T[mt][nt][1] = 0.0;

  int t = blockIdx.x, slot = 0;
  uint32_t phase = 0;
  const double* cur = ring;
  ArFrag f0, f1;
  mbar_wait(full_bar, 0);
  ar_load(f0, cur, tlane, lane, 0);

  // record ks of the current tile is in fc; the next one (same tile) goes to fn
  auto body = [&](const ArFrag& fc, ArFrag& fn, const double* nxt, int ks_next) {
    ar_load(fn, nxt, tlane, lane, ks_next * 64);
    ar_mma(T, fc);
  };
  // last record of a tile: the next record is the first one of the next tile (next slot); then the tile goes out
  auto last = [&](const ArFrag& fc, ArFrag& fn) -> bool {
    const int tn = t + gridDim.x;
    const bool have_next = tn < ntile;
    int ns = slot + 1;
    uint32_t nph = phase;
    if (ns == NSLOT) {
      ns = 0;
      nph ^= 1u;
    }
    const double* nxt = ring + (size_t)ns * SLOT_DOUBLES;
    if (have_next) {
      mbar_wait(full_bar + ns, nph);
      ar_load(fn, nxt, tlane, lane, 0);
    }
    ar_mma(T, fc);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar + slot);
    const int e0 = t * ASM_BE + q4 * 2;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const long long p = prow0 + 8 * mt;
      if (p < a.count) {
        double* dst = a.Ab + p * a.EP + e0;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
          __stcs(reinterpret_cast<double2*>(dst + nt * 8), make_double2(T[mt][nt][0], T[mt][nt][1]));
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) T[mt][nt][0] = T[mt][nt][1] = 0.0;
    }
    t = tn;
    slot = ns;
    phase = nph;
    cur = nxt;
    return have_next;
  };
  // one tile whose first record is in fa; returns false after the CTA's last tile.  An odd record count leaves the first
  // record of the next tile in fb, an even one in fa.
  auto tile = [&](ArFrag& fa, ArFrag& fb, bool odd) -> bool {
    int ks = 0;
    for (int i = nrec - 1; i >= 2; i -= 2) {
      body(fa, fb, cur + AR_REC_DOUBLES, ks + 1);
      body(fb, fa, cur + 2 * AR_REC_DOUBLES, ks + 2);
      cur += 2 * AR_REC_DOUBLES;
      ks += 2;
    }
    if (odd) return last(fa, fb);
    body(fa, fb, cur + AR_REC_DOUBLES, ks + 1);
    return last(fb, fa);
  };
  if (nrec & 1) {
    while (true) {
      if (!tile(f0, f1, true)) break;
      if (!tile(f1, f0, true)) break;
    }
  } else {
    while (tile(f0, f1, false)) {
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------------
void rb_asmrec_free(rb_ctx* ctx) {
  if (ctx->dAqRec) cudaFree(ctx->dAqRec);
  ctx->dAqRec = nullptr;
  ctx->asmrec_ok = false;
}

// called at the end of rb_set_system: packs ctx->dAq into records
int rb_asmrec_setup(rb_ctx* ctx) {
  rb_asmrec_free(ctx);
  if (const char* k = getenv("RB_ASM_KERNEL"))
    if (!strcmp(k, "old")) return RB_OK;
  if (ctx->N <= 0 || ctx->QlP <= 0) return RB_OK;
  const int nrec = ctx->QlP / 4;
  int KP = ctx->QlP;
  while (KP % 8 != 4) ++KP;
  int nslot = AR_MAXSLOT;
  while (nslot > 2 && asmrec_smem_bytes(nslot, nrec, KP) > 227 * 1024) --nslot;
  if (asmrec_smem_bytes(nslot, nrec, KP) > 227 * 1024) return RB_OK;  // too many terms: rb_assemble_kernel serves
  RB_CUDA(cudaMalloc(&ctx->dAqRec, (size_t)ctx->QlP * ctx->EP * sizeof(double)));
  rb_asmrec_pack_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->dAq, ctx->dAqRec, ctx->QlP, ctx->EP);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(rb_assemble_rec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)asmrec_smem_bytes(nslot, nrec, KP));
  if (e != cudaSuccess) {
    rb_asmrec_free(ctx);
    RB_FAIL(RB_ERR_CUDA, std::string("rb_set_system (assembly records): ") + cudaGetErrorString(e));
  }
  ctx->asmrec_KP = KP;
  ctx->asmrec_nslot = nslot;
  ctx->asmrec_ok = true;
  return RB_OK;
}

// Ab[0..n) = sum_q theta[p][q] A_q
int rb_asmrec_launch(rb_ctx* ctx, const double* theta, long long n) {
  AsmRecArgs a;
  a.theta = theta;
  a.rec = ctx->dAqRec;
  a.Ab = ctx->dAb;
  a.count = n;
  a.QT = ctx->QT;
  a.Ql = ctx->Ql;
  a.nrec = ctx->QlP / 4;
  a.KP = ctx->asmrec_KP;
  a.EP = ctx->EP;
  a.nslot = ctx->asmrec_nslot;
  // row blocks x entry-tile splits, one resident CTA per SM: the split that wastes the least of its last wave
  const unsigned nrb = (unsigned)((n + ASM_BM - 1) / ASM_BM);
  const unsigned slots = (unsigned)ctx->sm_count;
  const unsigned max_split = std::min<unsigned>((unsigned)(ctx->EP / ASM_BE), 16u);
  unsigned esplit = 1;
  double best_eff = 0.0;
  for (unsigned e = 1; e <= max_split; ++e) {
    const unsigned long long total = (unsigned long long)e * nrb;
    const unsigned long long waves = (total + slots - 1) / slots;
    double eff = (double)total / (double)(waves * slots) / (1.0 + 0.01 * (e - 1));  // every split reloads the Theta tile
    if (eff > best_eff) {
      best_eff = eff;
      esplit = e;
    }
  }
  dim3 grid(esplit, nrb);
  rb_assemble_rec_kernel<<<grid, AR_THREADS, asmrec_smem_bytes(a.nslot, a.nrec, a.KP), ctx->stream>>>(a);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}

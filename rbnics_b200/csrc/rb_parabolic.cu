// rb_parabolic.cu -- the POD-Greedy (parabolic) sweep as ONE kernel: per parameter one assembly and one LU factorisation
// of M/dt + A(mu), then K implicit-Euler steps (rhs, two triangular solves, residual dual norm, time quadrature) without
// leaving the SM.  SURVEY.md 8f row 2; replaces, for a whole training set,
//   rbnics/reduction_methods/base/time_dependent_rb_reduction.py:262-295 (the closure of _greedy),
//   rbnics/backends/online/numpy/time_stepping.py:104-112,226-238 (implicit Euler; the reference refactorises at every
//   step), rbnics/problems/parabolic/abstract_parabolic_rb_reduced_problem.py:29-92 (estimator over time) and
//   rbnics/backends/common/time_quadrature.py:22-29 (Simpson weights, passed in).
//
// Layout: one warp per parameter, 8 parameters per CTA.  Lane i owns row i of the NP x NP matrix in registers; pivoting is
// implicit during the factorisation (idamax tie rule), afterwards the rows are permuted ONCE into pivot order, so lane s
// holds row s of L\U and both triangular solves of every time step run in pivot-order space with one shuffle per
// column and no dynamic register index.  The residual norm of a step is eps2 = z^T G z, z = scale (.) V[src]
// (rb_set_estimator_ex); the 8 parameters of a CTA form the M = 8 dimension of DMMA m8n8k4: warp w multiplies column tile
// w of G (B fragments constant for the whole kernel: registers) with the z rows of all 8 parameters (shared memory).
// Served sizes: NP <= 32, Kz <= 128 (z rows padded to ZH = 2 or 4 chunks of 32 entries); anything else takes the
// multi-launch path of rb_sweep.cu.
#include "rb_common.cuh"
#include "rb_parabolic.cuh"

namespace {

constexpr int PB_WARPS = 8;
__host__ __device__ constexpr int pb_zp(int ZH) { return 32 * ZH + 4; }  // pitch of a z row (+ 4: conflict-free A-fragment loads)

// shared-memory carve-up (doubles)
struct PbLayout {
  int ms, rs, zs, vs, es, total;
  __host__ __device__ PbLayout(int NP, int Qm, int VP, int ZH) {
    ms = 0;                                  // [Qm][NP (col j)][NP (row i)]   M_q / dt
    rs = ms + Qm * NP * NP;                  // [8][NP (col j)][NP (row i)]    L\U of the 8 parameters, rows in place
    zs = rs + PB_WARPS * NP * NP;            // [2][8][pb_zp(ZH)]              z rows (double buffered over the steps)
    vs = zs + 2 * PB_WARPS * pb_zp(ZH);      // [8][VP]                        [u (NP) | udot (NP) | theta slice | 1]
    es = vs + PB_WARPS * VP;                 // [2][8 (tile)][8 (parameter)]   partial residual norms
    total = es + 2 * 8 * 8;
  }
};

template <int NP, int ZH>
__global__ void __launch_bounds__(32 * PB_WARPS, (ZH == 2 ? 3 : 2)) rb_parabolic_fused_kernel(PbArgs a) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int PB_ZP = pb_zp(ZH), KS = 8 * ZH, TPW = ZH / 2;  // k-steps of the z^T G z product, column tiles per warp
  extern __shared__ __align__(16) double pbs[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, q4 = lane & 3;
  const int N = a.N;
  const PbLayout lay(NP, a.Qm, a.VP, ZH);
  double* Ms = pbs + lay.ms;
  double* R = pbs + lay.rs + w * NP * NP;  // entry (row i, col j) at j * NP + i
  double* Zs = pbs + lay.zs;
  double* V = pbs + lay.vs + w * a.VP;
  double* Es = pbs + lay.es;

  for (int t = tid; t < a.Qm * NP * NP; t += 32 * PB_WARPS) {
    const int q = t / (NP * NP), r = t % (NP * NP);
    Ms[t] = __ldg(a.Aq + (long long)q * a.EP + r);
  }
  for (int t = tid; t < 2 * PB_WARPS * PB_ZP; t += 32 * PB_WARPS) Zs[t] = 0.0;

  const long long p = (long long)blockIdx.x * PB_WARPS + w;
  const bool live = p < a.n;
  const long long pc = live ? p : a.n - 1;  // idle warps of the last CTA shadow a valid parameter (they share barriers)
  const double* th = a.theta + pc * a.QT;
  const bool inrow = lane < NP;

  // ---- row `lane` of M/dt + A(mu), f(mu); lifting and padding rows are unit rows
  double f = 0.0;
  bool isbc = false;
  {
    bool unit = lane >= N;
    for (int t = 0; t < a.n_bc; ++t)
      if (__ldg(a.bc_rows + t) == lane) {
        unit = isbc = true;
        f = __ldg(a.thetaBC + pc * a.n_bc + t);
      }
    if (inrow) {
      if (unit) {
#pragma unroll
        for (int j = 0; j < NP; ++j) R[j * NP + lane] = (j == lane) ? 1.0 : 0.0;
      } else {
#pragma unroll 4
        for (int j = 0; j < NP; ++j) {
          double v = 0.0;
          for (int q = 0; q < a.Ql; ++q) v = fma(__ldg(th + q), __ldg(a.Aq + (long long)q * a.EP + j * NP + lane), v);
          R[j * NP + lane] = v;
        }
        for (int q = 0; q < a.Qr; ++q) f = fma(__ldg(th + a.Ql + q), __ldg(a.F + q * NP + lane), f);
      }
    }
  }
  __syncwarp();

  // ---- PA = LU in place, implicit partial pivoting (largest |a|, lowest row on ties: LAPACK idamax); rows never move
  bool done = !inrow;
  int myrow = 0;      // lane t: the row that became pivot t
  double rinv = 0.0;  // lane t: reciprocal of pivot t
  for (int t = 0; t < NP; ++t) {
    const double at = inrow ? R[t * NP + lane] : 0.0;
    const unsigned long long key = done ? 0ull : ((unsigned long long)__double_as_longlong(fabs(at)) + 1ull);
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(FULL, hi);
    const unsigned lo2 = (hi == mhi) ? lo : 0u;
    const unsigned mlo = __reduce_max_sync(FULL, lo2);
    const unsigned ball = __ballot_sync(FULL, hi == mhi && lo == mlo && !done);
    const int wl = ball ? (__ffs(ball) - 1) : t;  // (NP live rows for NP steps: there always is a candidate)
    const double rv = 1.0 / R[t * NP + wl];
    if (lane == t) {
      myrow = wl;
      rinv = rv;
    }
    if (lane == wl) {
      done = true;
    } else if (!done) {
      const double l = at * rv;
      R[t * NP + lane] = l;
      for (int j = t + 1; j < NP; ++j) R[j * NP + lane] = fma(-l, R[j * NP + wl], R[j * NP + lane]);
    }
    __syncwarp();
  }
  // from here on lane s works on equation myrow(s): pivot-order space without moving a row
  f = __shfl_sync(FULL, f, myrow);
  isbc = __shfl_sync(FULL, (int)isbc, myrow) != 0;
  const double* Rr = R + myrow;  // Rr[t * NP] = L[s][t] (t < s) | U[s][t] (t >= s)

  // ---- z map of this lane (entries lane + 32 h), remapped to V = [u (NP) | udot (NP) | theta slice | 1]
  int zsrc[ZH];
  double zsc[ZH];
#pragma unroll
  for (int h = 0; h < ZH; ++h) {
    const int k = lane + 32 * h;
    zsrc[h] = 0;
    zsc[h] = 0.0;
    if (k < a.Kz) {
      const int src = __ldg(a.zsrc + k);
      zsrc[h] = src < N ? src : (src < 2 * N ? NP + (src - N) : 2 * NP + (src - 2 * N));
      const int sc = __ldg(a.zscale + k);
      zsc[h] = sc < 0 ? 1.0 : __ldg(th + sc);
    }
  }
  double uprev = 0.0;
  if (lane < N && a.U0) uprev = a.U0[pc * N + lane];
  for (int t = lane; t < a.VP; t += 32) {
    double v = 0.0;
    if (t >= 2 * NP && t < 2 * NP + a.v_len) v = __ldg(th + a.v_off + (t - 2 * NP));
    if (t == 2 * NP + a.v_len) v = 1.0;
    V[t] = v;
  }
  __syncwarp();
  if (lane < N) V[lane] = uprev;
  double acc = a.E0 ? __ldg(a.weights) * fabs(a.E0[pc]) : 0.0;
  const double beta = __ldg(a.beta + pc);
  double* eh = (live && a.eps2_hist) ? a.eps2_hist + p * (a.K + 1) : nullptr;
  double* uh = (live && a.u_hist) ? a.u_hist + p * (long long)(a.K + 1) * N : nullptr;
  if (uh && lane < N) uh[lane] = uprev;
  const bool rhs_m = !isbc && myrow < N && inrow;
  // B fragments of column tile nt: k-step ks at Gfrag[(nt * KS + ks) * 32 + lane] (constant: L1 / L2 resident)
  const double* Gf = a.Gfrag + (long long)w * KS * 32 + lane;
  __syncthreads();  // Ms, the zero padding of Zs

  for (int k = 1; k <= a.K; ++k) {
    const int buf = k & 1;
    // ---- rhs of equation myrow: f + sum_q theta_q (M_q/dt u^{k-1})  (lifting rows: theta_bc)
    double y = f;
    if (rhs_m) {
      for (int q = 0; q < a.Qm; ++q) {
        const double* Mq = Ms + q * NP * NP + myrow;
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NP; j += 2) {
          const double2 u2 = *reinterpret_cast<const double2*>(V + j);  // V[N..NP) stays zero
          s = fma(Mq[j * NP], u2.x, s);
          s = fma(Mq[(j + 1) * NP], u2.y, s);
        }
        y = fma(__ldg(th + q), s, y);
      }
    }
    // ---- L y' = y, U x = y' in pivot order: one shuffle per column, the factors stream from shared memory
#pragma unroll
    for (int t = 0; t < NP - 1; ++t) {
      const double yt = __shfl_sync(FULL, y, t);
      if (lane > t) y = fma(-Rr[t * NP], yt, y);
    }
    double x = 0.0;
#pragma unroll
    for (int t = NP - 1; t >= 0; --t) {
      const double xt = __shfl_sync(FULL, y * rinv, t);
      if (lane == t) x = xt;
      if (lane < t) y = fma(-Rr[t * NP], xt, y);
    }
    // ---- V = [u^k | (u^k - u^{k-1}) / dt | ...], z
    __syncwarp();
    if (lane < N) {
      V[lane] = x;
      V[NP + lane] = (x - uprev) * a.inv_dt;
      if (uh) uh[(long long)k * N + lane] = x;
    }
    uprev = x;
    __syncwarp();
    double* Z = Zs + (buf * PB_WARPS + w) * PB_ZP;
#pragma unroll
    for (int h = 0; h < ZH; ++h) Z[lane + 32 * h] = zsc[h] * V[zsrc[h]];
    __syncthreads();
    // ---- W = Z G (8 parameters x the column tiles w, w + 8, ... of this warp), e = rowsum(W (.) Z)
    {
      const double* Zg = Zs + (buf * PB_WARPS + g) * PB_ZP;
      double c[TPW][2][2];
#pragma unroll
      for (int t = 0; t < TPW; ++t) c[t][0][0] = c[t][0][1] = c[t][1][0] = c[t][1][1] = 0.0;
#pragma unroll
      for (int ks = 0; ks < KS; ks += 2) {
        const double a0 = Zg[4 * ks + q4], a1 = Zg[4 * ks + 4 + q4];
#pragma unroll
        for (int t = 0; t < TPW; ++t) {
          dmma884(c[t][0][0], c[t][0][1], a0, __ldg(Gf + (t * 8 * KS + ks) * 32));
          dmma884(c[t][1][0], c[t][1][1], a1, __ldg(Gf + (t * 8 * KS + ks + 1) * 32));
        }
      }
      double part = 0.0;
#pragma unroll
      for (int t = 0; t < TPW; ++t) {
        const double2 zz = *reinterpret_cast<const double2*>(Zg + 8 * (w + 8 * t) + 2 * q4);
        part += fma(c[t][0][0] + c[t][1][0], zz.x, (c[t][0][1] + c[t][1][1]) * zz.y);
      }
      part += __shfl_xor_sync(FULL, part, 1);
      part += __shfl_xor_sync(FULL, part, 2);
      if (q4 == 0) Es[(buf * 8 + w) * 8 + g] = part;
    }
    __syncthreads();
    {
      double e = 0.0;
#pragma unroll
      for (int t = 0; t < 8; ++t) e += Es[(buf * 8 + t) * 8 + w];  // fixed order: deterministic
      acc = fma(__ldg(a.weights + k), fabs(e) / beta, acc);
      if (eh && lane == 0) eh[k] = e;
    }
  }
  if (live && lane == 0) a.acc_out[p] = acc;
}

template <int NP, int ZH>
int pb_launch(const PbArgs& a, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(rb_parabolic_fused_kernel<NP, ZH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const unsigned grid = (unsigned)((a.n + PB_WARPS - 1) / PB_WARPS);
  rb_parabolic_fused_kernel<NP, ZH><<<grid, 32 * PB_WARPS, smem, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace

int rb_parabolic_zh(int Kz) { return Kz <= 64 ? 2 : 4; }

bool rb_parabolic_fused_supported(int NP, int Kz, int Qm, int VP) {
  if (NP < 4 || NP > 32 || NP % 4 != 0 || Kz > 128 || Kz <= 0 || VP % 2 != 0) return false;
  return rb_parabolic_fused_smem(NP, Kz, Qm, VP) <= 160 * 1024;
}

size_t rb_parabolic_fused_smem(int NP, int Kz, int Qm, int VP) {
  return sizeof(double) * (size_t)PbLayout(NP, Qm, VP, rb_parabolic_zh(Kz)).total;
}

// B fragments of G for DMMA m8n8k4, zero padded to 32 ZH x 32 ZH: [column tile (4 ZH)][k-step (8 ZH)][lane (32)]
size_t rb_parabolic_pack_g_doubles(int Kz) {
  const int ZH = rb_parabolic_zh(Kz);
  return (size_t)(4 * ZH) * (8 * ZH) * 32;
}
void rb_parabolic_pack_g(const double* G, int Kz, double* out) {
  const int ZH = rb_parabolic_zh(Kz), KS = 8 * ZH, NTL = 4 * ZH;
  for (int nt = 0; nt < NTL; ++nt)
    for (int ks = 0; ks < KS; ++ks)
      for (int lane = 0; lane < 32; ++lane) {
        const int k = 4 * ks + (lane & 3), n = 8 * nt + (lane >> 2);
        out[((size_t)nt * KS + ks) * 32 + lane] = (k < Kz && n < Kz) ? G[(size_t)k * Kz + n] : 0.0;
      }
}

int rb_parabolic_fused_launch(const PbArgs& a, int NP, cudaStream_t st) {
  const size_t smem = rb_parabolic_fused_smem(NP, a.Kz, a.Qm, a.VP);
  const int ZH = rb_parabolic_zh(a.Kz);
  switch (NP) {
#define RB_PB_CASE(n) \
  case n: return ZH == 2 ? pb_launch<n, 2>(a, smem, st) : pb_launch<n, 4>(a, smem, st);
    RB_PB_CASE(4) RB_PB_CASE(8) RB_PB_CASE(12) RB_PB_CASE(16) RB_PB_CASE(20) RB_PB_CASE(24) RB_PB_CASE(28) RB_PB_CASE(32)
#undef RB_PB_CASE
    default: return (int)cudaErrorInvalidValue;
  }
}

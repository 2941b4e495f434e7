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
// Served sizes: NP <= 32, Kz <= 64; anything else takes the multi-launch path of rb_sweep.cu.
#include "rb_common.cuh"
#include "rb_parabolic.cuh"

namespace {

constexpr int PB_WARPS = 8;
constexpr int PB_ZP = 68;  // pitch of a z row (64 + 4: conflict-free A-fragment loads)

template <int NP>
__global__ void __launch_bounds__(32 * PB_WARPS, (NP <= 20 ? 2 : 1)) rb_parabolic_fused_kernel(PbArgs a) {
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) double pbs[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, q4 = lane & 3;
  const int N = a.N;
  const int VP = a.VP;  // pitch of a V row: [u (N) | udot (N) | theta slice | 1 | 0 ...]
  double* Ms = pbs;                                      // [Qm][NP (col j)][NP (row i)]
  double* Zs = Ms + a.Qm * NP * NP;                      // [2][8][PB_ZP]
  double* Vs = Zs + 2 * PB_WARPS * PB_ZP;                // [8][VP]
  double* Es = Vs + PB_WARPS * VP;                       // [2][8 (tile)][8 (parameter)]

  for (int t = tid; t < a.Qm * NP * NP; t += 32 * PB_WARPS) {
    const int q = t / (NP * NP), r = t % (NP * NP);
    Ms[t] = __ldg(a.Aq + (long long)q * a.EP + r);
  }
  for (int t = tid; t < 2 * PB_WARPS * PB_ZP; t += 32 * PB_WARPS) Zs[t] = 0.0;

  // B fragments of column tile w of G: entry (k = 4 ks + q4, n = 8 w + g), zero outside Kz
  double Bf[16];
#pragma unroll
  for (int ks = 0; ks < 16; ++ks) {
    const int k = 4 * ks + q4, n = 8 * w + g;
    Bf[ks] = (k < a.Kz && n < a.Kz) ? __ldg(a.G + (long long)k * a.Kz + n) : 0.0;
  }

  const long long p = (long long)blockIdx.x * PB_WARPS + w;
  const bool live = p < a.n;
  const long long pc = live ? p : a.n - 1;  // idle warps of the last CTA shadow a valid parameter (they share barriers)
  const double* th = a.theta + pc * a.QT;

  // ---- row `lane` of M/dt + A(mu), f(mu); lifting and padding rows are unit rows
  double row[NP];
  double f = 0.0;
  bool isbc = false;
  {
    const bool in = lane < N;
#pragma unroll
    for (int j = 0; j < NP; ++j) row[j] = 0.0;
    if (in) {
      for (int q = 0; q < a.Ql; ++q) {
        const double tq = __ldg(th + q);
        const double* Aq = a.Aq + (long long)q * a.EP + lane;
#pragma unroll
        for (int j = 0; j < NP; ++j) row[j] = fma(tq, __ldg(Aq + j * NP), row[j]);
      }
      for (int q = 0; q < a.Qr; ++q) f = fma(__ldg(th + a.Ql + q), __ldg(a.F + q * NP + lane), f);
    }
    bool unit = !in;
    for (int t = 0; t < a.n_bc; ++t)
      if (__ldg(a.bc_rows + t) == lane) {
        unit = isbc = true;
        f = __ldg(a.thetaBC + pc * a.n_bc + t);
      }
    if (unit) {
#pragma unroll
      for (int j = 0; j < NP; ++j) row[j] = (j == lane) ? 1.0 : 0.0;
      if (!isbc) f = 0.0;
    }
  }

  // ---- PA = LU, implicit partial pivoting (largest |a|, lowest row on ties: LAPACK idamax)
  bool done = lane >= NP;
  int piv_of = 0;     // lane t: the row that became pivot t
  double rinv = 0.0;  // reciprocal pivot of this row
#pragma unroll
  for (int t = 0; t < NP; ++t) {
    const double at = row[t];
    const unsigned long long key = done ? 0ull : ((unsigned long long)__double_as_longlong(fabs(at)) + 1ull);
    const double ri = 1.0 / at;  // speculative: only the winner's is used
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(FULL, hi);
    const unsigned lo2 = (hi == mhi) ? lo : 0u;
    const unsigned mlo = __reduce_max_sync(FULL, lo2);
    const unsigned ball = __ballot_sync(FULL, hi == mhi && lo == mlo && !done);
    const int wl = ball ? (__ffs(ball) - 1) : t;  // no live row left cannot happen (NP rows, NP steps)
    const double rv = __shfl_sync(FULL, ri, wl);
    const bool iwin = lane == wl;
    const double l = at * rv;
#pragma unroll
    for (int j = t + 1; j < NP; ++j) {
      const double pj = __shfl_sync(FULL, row[j], wl);
      if (!done && !iwin) row[j] = fma(-l, pj, row[j]);
    }
    if (!done && !iwin) row[t] = l;
    if (iwin) {
      done = true;
      rinv = ri;
    }
    if (lane == t) piv_of = wl;
  }
  // rows into pivot order: lane s <- row piv_of(s).  From here on lane s holds L[s][0..s) | U[s][s..NP)
#pragma unroll
  for (int j = 0; j < NP; ++j) row[j] = __shfl_sync(FULL, row[j], piv_of);
  rinv = __shfl_sync(FULL, rinv, piv_of);
  f = __shfl_sync(FULL, f, piv_of);
  isbc = __shfl_sync(FULL, (int)isbc, piv_of) != 0;
  const int myrow = piv_of;  // original row of this lane's equation (rhs assembly)

  // ---- z map of this lane: entries lane and lane + 32
  double* V = Vs + w * VP;
  int zsrc[2];
  double zsc[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int k = lane + 32 * h;
    zsrc[h] = 0;
    zsc[h] = 0.0;
    if (k < a.Kz) {
      zsrc[h] = __ldg(a.zsrc + k);
      const int sc = __ldg(a.zscale + k);
      zsc[h] = sc < 0 ? 1.0 : __ldg(th + sc);
    }
  }
  // V = [u^0 | 0 | theta slice | 1 | 0 ...]
  double uprev = 0.0;
  if (lane < N && a.U0) uprev = a.U0[pc * N + lane];
  for (int t = lane; t < VP; t += 32) {
    double v = 0.0;
    if (t >= 2 * N && t < 2 * N + a.v_len) v = __ldg(th + a.v_off + (t - 2 * N));
    if (t == 2 * N + a.v_len) v = 1.0;
    V[t] = v;
  }
  __syncwarp();
  if (lane < N) V[lane] = uprev;
  double acc = a.E0 ? __ldg(a.weights) * fabs(a.E0[pc]) : 0.0;
  const double beta = __ldg(a.beta + pc);
  double* eh = (live && a.eps2_hist) ? a.eps2_hist + p * (a.K + 1) : nullptr;
  double* uh = (live && a.u_hist) ? a.u_hist + p * (long long)(a.K + 1) * N : nullptr;
  if (uh && lane < N) uh[lane] = uprev;
  __syncthreads();  // Ms, the zero padding of Zs

  for (int k = 1; k <= a.K; ++k) {
    const int buf = k & 1;
    // ---- rhs of equation `myrow`: f + sum_q theta_q (M_q/dt u^{k-1})  (lifting rows: theta_bc)
    double y = f;
    if (!isbc && myrow < N) {
      for (int q = 0; q < a.Qm; ++q) {
        const double* Mq = Ms + q * NP * NP + myrow;
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NP; j += 2) {
          const double2 u2 = *reinterpret_cast<const double2*>(V + j);  // entries >= N meet zero columns of M_q
          s = fma(Mq[j * NP], (j < N) ? u2.x : 0.0, s);
          if (j + 1 < NP) s = fma(Mq[(j + 1) * NP], (j + 1 < N) ? u2.y : 0.0, s);
        }
        y = fma(__ldg(th + q), s, y);
      }
    }
    // ---- L y' = y, U x = y' in pivot order
#pragma unroll
    for (int t = 0; t < NP - 1; ++t) {
      const double yt = __shfl_sync(FULL, y, t);
      if (lane > t) y = fma(-row[t], yt, y);
    }
    double x = 0.0;
#pragma unroll
    for (int t = NP - 1; t >= 0; --t) {
      const double xt = __shfl_sync(FULL, y * rinv, t);
      if (lane == t) x = xt;
      if (lane < t) y = fma(-row[t], xt, y);
    }
    // ---- V = [u^k | (u^k - u^{k-1}) / dt | ...], z
    __syncwarp();
    if (lane < N) {
      V[lane] = x;
      V[N + lane] = (x - uprev) * a.inv_dt;
      if (uh) uh[(long long)k * N + lane] = x;
    }
    uprev = x;
    __syncwarp();
    double* Z = Zs + (buf * PB_WARPS + w) * PB_ZP;
    Z[lane] = zsc[0] * V[zsrc[0]];
    Z[lane + 32] = zsc[1] * V[zsrc[1]];
    __syncthreads();
    // ---- W = Z G (8 parameters x column tile w), e = rowsum(W (.) Z)
    {
      const double* Zg = Zs + (buf * PB_WARPS + g) * PB_ZP;
      double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int ks = 0; ks < 16; ks += 2) {
        dmma884(c0, c1, Zg[4 * ks + q4], Bf[ks]);
        dmma884(d0, d1, Zg[4 * ks + 4 + q4], Bf[ks + 1]);
      }
      const double2 zz = *reinterpret_cast<const double2*>(Zg + 8 * w + 2 * q4);
      double part = fma(c0 + d0, zz.x, (c1 + d1) * zz.y);
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

template <int NP>
int pb_launch(const PbArgs& a, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(rb_parabolic_fused_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const unsigned grid = (unsigned)((a.n + PB_WARPS - 1) / PB_WARPS);
  rb_parabolic_fused_kernel<NP><<<grid, 32 * PB_WARPS, smem, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace

bool rb_parabolic_fused_supported(int NP, int Kz, int Qm, int VP) {
  if (NP < 4 || NP > 32 || NP % 4 != 0 || Kz > 64 || Kz <= 0) return false;
  return rb_parabolic_fused_smem(NP, Qm, VP) <= 100 * 1024;
}

size_t rb_parabolic_fused_smem(int NP, int Qm, int VP) {
  return sizeof(double) * ((size_t)Qm * NP * NP + 2 * PB_WARPS * PB_ZP + (size_t)PB_WARPS * VP + 2 * 8 * 8);
}

int rb_parabolic_fused_launch(const PbArgs& a, int NP, cudaStream_t st) {
  const size_t smem = rb_parabolic_fused_smem(NP, a.Qm, a.VP);
  switch (NP) {
#define RB_PB_CASE(n) \
  case n: return pb_launch<n>(a, smem, st);
    RB_PB_CASE(4) RB_PB_CASE(8) RB_PB_CASE(12) RB_PB_CASE(16) RB_PB_CASE(20) RB_PB_CASE(24) RB_PB_CASE(28) RB_PB_CASE(32)
#undef RB_PB_CASE
    default: return (int)cudaErrorInvalidValue;
  }
}

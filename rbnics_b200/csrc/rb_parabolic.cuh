// rb_parabolic.cuh -- arguments of the fused POD-Greedy sweep kernel (rb_parabolic.cu)
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

struct PbArgs {
  const double* theta;    // [n][QT]: lhs thetas (the first Qm belong to M_q/dt), then rhs thetas
  const double* thetaBC;  // [n][n_bc] or null
  const double* beta;     // [n]
  const double* Aq;       // [Ql][EP]: entry (i, j) of term q at q*EP + j*NP + i
  const double* F;        // [Qr][NP]
  const int* bc_rows;     // [n_bc]
  const double* Gfrag;    // the quadratic form of the estimator as DMMA B fragments (rb_parabolic_pack_g)
  const int* zsrc;        // [Kz] index into V = [u | udot | theta slice | 1]
  const int* zscale;      // [Kz] theta column that scales the entry, or -1
  const double* weights;  // [K+1] time-quadrature weights (device)
  const double* U0;       // [n][N] reduced initial condition or null
  const double* E0;       // [n] squared initial error bound or null
  double* acc_out;        // [n] time-integrated squared error bound
  double* eps2_hist;      // [n][K+1] or null
  double* u_hist;         // [n][K+1][N] or null
  long long n;
  int N, EP, QT, Ql, Qr, Qm, n_bc, K, Kz, v_off, v_len, VP;
  double inv_dt;
};

bool rb_parabolic_fused_supported(int NP, int Kz, int Qm, int VP);
size_t rb_parabolic_fused_smem(int NP, int Kz, int Qm, int VP);
int rb_parabolic_fused_launch(const PbArgs& a, int NP, cudaStream_t st);
size_t rb_parabolic_pack_g_doubles(int Kz);
void rb_parabolic_pack_g(const double* G, int Kz, double* out);

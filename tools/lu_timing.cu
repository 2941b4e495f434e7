// Phase-level clock64 timing of the DMMA LU kernel (development probe; not part of the library).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DLUD_TIMING -I rbnics_b200/csrc -o tools/lu_timing tools/lu_timing.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "rb_lu_dmma.cuh"
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
int main(int argc, char** argv){
  const int N=100, NP=100, EP=10048; const long long count = argc>1 ? atoll(argv[1]) : 18944;
  std::vector<double> A((size_t)EP, 0.0);
  srand(1);
  for(int j=0;j<N;j++) for(int i=0;i<N;i++) A[(size_t)j*NP+i] = (double)rand()/RAND_MAX - 0.5 + (i==j ? 3.0 : 0.0);
  double *dAb, *dTheta, *dF, *dU; 
  CK(cudaMalloc(&dAb, sizeof(double)*EP*count)); CK(cudaMalloc(&dTheta, sizeof(double)*count)); CK(cudaMalloc(&dF, sizeof(double)*NP)); CK(cudaMalloc(&dU, sizeof(double)*N*count));
  for(long long s=0;s<count;s++) CK(cudaMemcpy(dAb+s*EP, A.data(), sizeof(double)*EP, cudaMemcpyHostToDevice));
  std::vector<double> ones(count,1.0), F(NP,1.0);
  CK(cudaMemcpy(dTheta, ones.data(), sizeof(double)*count, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dF, F.data(), sizeof(double)*NP, cudaMemcpyHostToDevice));
  LuArgs a; a.Ab=dAb; a.theta=dTheta; a.thetaBC=nullptr; a.F=dF; a.bc_rows=nullptr; a.U=dU; a.count=count; a.N=N; a.EP=EP; a.QT=1; a.Ql=0; a.Qr=1; a.n_bc=0; a.ldu=N;
  int rc = rb_lu_dmma_launch<13,13>(a, 0); CK(cudaDeviceSynchronize()); if(rc){printf("launch rc %d\n",rc); return 1;}
  unsigned long long z[32]={0}; CK(cudaMemcpyToSymbol(lud_timing, z, sizeof(z)));
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0)); rc = rb_lu_dmma_launch<13,13>(a, 0); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms,e0,e1));
  CK(cudaMemcpyFromSymbol(z, lud_timing, sizeof(z)));
  printf("count %lld  kernel %.3f ms  (%.1f clk/system at 1.965 GHz x 296 slots)\n", count, ms, ms*1e-3*1.965e9*296/count);
  const char* names[32]={"P: wait panel","P: factor B","P: wait end","P: backsub","","","","","","","T: load+storeP0","T: wait B","T: gather C + bar","T: C2","T: bar C2","T: D first + store","T: D rest(last)"};
  names[20]="S: keys+REDUX1"; names[21]="S: sel+div+STS"; names[22]="S: REDUX2+owner"; names[23]="S: syncwarp+LDS"; names[24]="S: update";
  for(int i=0;i<25;i++) if(z[i]) printf("  [%2d] %-22s %10.1f clk/system\n", i, names[i]?names[i]:"", (double)z[i]/count);
  return 0;
}

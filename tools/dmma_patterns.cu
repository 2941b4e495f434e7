// DMMA issue-rate probes for realistic operand patterns (outer-product register tiles, operands from shared memory).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// MA x NB outer product per k-step, operands in registers (no loads)
template<int MA, int NB>
__global__ void __launch_bounds__(1024) outer_reg(double* out, int iters, double seed){
  double c[MA][NB][2]; double a[MA], b[NB];
  for(int i=0;i<MA;i++) a[i]=seed+i+threadIdx.x*1e-9;
  for(int j=0;j<NB;j++) b[j]=seed*1e-3+j;
  for(int i=0;i<MA;i++) for(int j=0;j<NB;j++){c[i][j][0]=0;c[i][j][1]=0;}
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int i=0;i<MA;i++)
      #pragma unroll
      for(int j=0;j<NB;j++) dmma(c[i][j][0],c[i][j][1],a[i],b[j]);
  }
  double s=0; for(int i=0;i<MA;i++) for(int j=0;j<NB;j++) s+=c[i][j][0]+c[i][j][1];
  if(s==123.456) out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// same, operands re-loaded from shared memory every k-step
template<int MA, int NB>
__global__ void __launch_bounds__(1024) outer_smem(double* out, int iters, double seed){
  __shared__ double sm[4096];
  for(int i=threadIdx.x;i<4096;i+=blockDim.x) sm[i]=seed+i*1e-6;
  __syncthreads();
  double c[MA][NB][2];
  for(int i=0;i<MA;i++) for(int j=0;j<NB;j++){c[i][j][0]=0;c[i][j][1]=0;}
  int lane=threadIdx.x&31; int off=lane;
  for(int it=0;it<iters;++it){
    double a[MA], b[NB];
    #pragma unroll
    for(int i=0;i<MA;i++) a[i]=sm[(off+i*36)&4095];
    #pragma unroll
    for(int j=0;j<NB;j++) b[j]=sm[(off+2048+j*32)&4095];
    off+=4;
    #pragma unroll
    for(int i=0;i<MA;i++)
      #pragma unroll
      for(int j=0;j<NB;j++) dmma(c[i][j][0],c[i][j][1],a[i],b[j]);
  }
  double s=0; for(int i=0;i<MA;i++) for(int j=0;j<NB;j++) s+=c[i][j][0]+c[i][j][1];
  if(s==123.456) out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<typename F> double time_ms(F f){ cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); f(); CK(cudaDeviceSynchronize()); float best=1e30f;
  for(int r=0;r<5;r++){ CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best)best=ms;} return best; }
#define RUN(K, MA, NB, TPB) { double ms=time_ms([&]{ K<MA,NB><<<sms,TPB>>>(out,iters,1.0); }); double fl=2.0*256*MA*NB*iters*(double)(TPB/32)*sms; printf("%-10s MA=%d NB=%d warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", #K, MA, NB, TPB/32, ms, fl/ms*1e-9);} 
int main(){ cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sms=p.multiProcessorCount; double* out; CK(cudaMalloc(&out,8*sms*1024)); int iters=4000;
  RUN(outer_reg,4,4,256) RUN(outer_reg,4,2,512) RUN(outer_reg,4,2,256) RUN(outer_reg,2,2,512) RUN(outer_reg,4,4,128) RUN(outer_reg,1,8,256) RUN(outer_reg,8,1,256)
  RUN(outer_smem,4,4,256) RUN(outer_smem,4,2,512) RUN(outer_smem,4,2,256) RUN(outer_smem,2,2,512) RUN(outer_smem,4,4,128)
  return 0; }

// FP64 peak microbenchmark for B200 (sm_100a): DFMA vs DMMA.8x8x4 issue rates.
// Used once per pool to obtain the FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 entry).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

template<int NACC>
__global__ void __launch_bounds__(1024) dfma_kernel(double* out, int iters, double seed){
  double acc[NACC];
  double a = seed + threadIdx.x*1e-9, b = 1.0 - 1e-9*seed;
  #pragma unroll
  for(int i=0;i<NACC;i++) acc[i]=i*0.5;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++) acc[i]=fma(acc[i], b, a);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=acc[i];
  if(s==123.456) out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<int NACC>
__global__ void __launch_bounds__(1024) dmma_kernel(double* out, int iters, double seed){
  double c[NACC][2];
  double a = seed + threadIdx.x*1e-9, b = 1e-3*seed;
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=0;c[i][1]=0;}
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  if(s==123.456) out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<typename F>
double time_ms(F f, int reps){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<reps;r++){
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best) best=ms;
  }
  return best;
}

int main(int argc, char** argv){
  int dev=0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,dev));
  int sms=p.multiProcessorCount;
  printf("device %s sms=%d clock=%d kHz\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*8*1024));
  int iters = 20000;
  // DFMA: sweep threads per SM
  for(int tpb : {128, 256, 512, 1024}){
    for(int bps : {1, 2}){
      if(tpb*bps>2048) continue;
      double ms = time_ms([&]{ dfma_kernel<16><<<sms*bps, tpb>>>(out, iters, 1.0); }, 5);
      double flops = 2.0*16*iters*(double)tpb*bps*sms;
      printf("DFMA  tpb=%4d bps=%d  %.3f ms  %.2f TFLOP/s\n", tpb, bps, ms, flops/ms*1e-9);
    }
  }
  for(int tpb : {32, 64, 128, 256, 512, 1024}){
    for(int bps : {1, 2}){
      if(tpb*bps>2048) continue;
      double ms = time_ms([&]{ dmma_kernel<8><<<sms*bps, tpb>>>(out, iters, 1.0); }, 5);
      double flops = 2.0*256*8*iters*(double)(tpb/32)*bps*sms;
      printf("DMMA8 tpb=%4d bps=%d  %.3f ms  %.2f TFLOP/s\n", tpb, bps, ms, flops/ms*1e-9);
    }
  }
  for(int tpb : {128, 256, 512}){
      double ms = time_ms([&]{ dmma_kernel<2><<<sms, tpb>>>(out, iters, 1.0); }, 5);
      double flops = 2.0*256*2*iters*(double)(tpb/32)*sms;
      printf("DMMA2 tpb=%4d bps=1  %.3f ms  %.2f TFLOP/s\n", tpb, ms, flops/ms*1e-9);
      ms = time_ms([&]{ dmma_kernel<1><<<sms, tpb>>>(out, iters, 1.0); }, 5);
      flops = 2.0*256*1*iters*(double)(tpb/32)*sms;
      printf("DMMA1 tpb=%4d bps=1  %.3f ms  %.2f TFLOP/s (latency-bound chain)\n", tpb, ms, flops/ms*1e-9);
  }
  // sustained: 3 s of DMMA back to back
  {
    cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int n=0; CK(cudaEventRecord(e0));
    for(int r=0;r<60;r++){ dmma_kernel<8><<<sms*2, 512>>>(out, iters*4, 1.0); n++; }
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms,e0,e1));
    double flops = 2.0*256*8*(iters*4.0)*(512/32)*2*sms*n;
    printf("DMMA8 sustained %d launches %.1f ms %.2f TFLOP/s\n", n, ms, flops/ms*1e-9);
  }
  return 0;
}

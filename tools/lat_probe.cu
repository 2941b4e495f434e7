// Latency probes (single warp, dependent chains): DFMA, DMUL, 1.0/x, REDUX, SHFL, LDS round trip, DMMA chain.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} }while(0)
__global__ void probe(double* out, long long* clk, double x0, int n){
  __shared__ double sm[64];
  double x = x0 + threadIdx.x*1e-9, y = 1.0000001;
  long long t0, t1; int k=0;
  t0=clock64(); for(int i=0;i<n;i++) x = fma(x, y, 1e-9); t1=clock64(); if(threadIdx.x==0) clk[k]=t1-t0; k++;
  t0=clock64(); for(int i=0;i<n;i++) x = x * y; t1=clock64(); if(threadIdx.x==0) clk[k]=t1-t0; k++;
  t0=clock64(); for(int i=0;i<n;i++) x = 1.0 / x + 0.5; t1=clock64(); if(threadIdx.x==0) clk[k]=t1-t0; k++;
  unsigned u = (unsigned)__double_as_longlong(x);
  t0=clock64(); for(int i=0;i<n;i++) u = __reduce_max_sync(0xffffffffu, u + threadIdx.x) ; t1=clock64(); if(threadIdx.x==0) clk[k]=t1-t0; k++;
  t0=clock64(); for(int i=0;i<n;i++) x = __shfl_sync(0xffffffffu, x, (i*7)&31) ; t1=clock64(); if(threadIdx.x==0) clk[k]=t1-t0; k++;
  t0=clock64(); for(int i=0;i<n;i++){ sm[threadIdx.x]=x; __syncwarp(); x = sm[(threadIdx.x+1)&31]; __syncwarp(); } t1=clock64(); if(threadIdx.x==0) clk[k]=t1-t0; k++;
  float f = (float)x;
  t0=clock64(); for(int i=0;i<n;i++) f = fmaf(f, 1.0000001f, 1e-9f); t1=clock64(); if(threadIdx.x==0) clk[k]=t1-t0; k++;
  double c0=x, c1=y;
  t0=clock64(); for(int i=0;i<n;i++) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(y), "d"(y)); t1=clock64(); if(threadIdx.x==0) clk[k]=t1-t0; k++;
  int iv = (int)u;
  t0=clock64(); for(int i=0;i<n;i++) iv = iv*3+1; t1=clock64(); if(threadIdx.x==0) clk[k]=t1-t0; k++;
  double sel = x;
  t0=clock64(); for(int i=0;i<n;i++) sel = (iv & (1<<(i&7))) ? sel : y + sel*0; t1=clock64(); if(threadIdx.x==0) clk[k]=t1-t0; k++;
  out[threadIdx.x] = x + u + f + c0 + c1 + iv + sel;
}
int main(){ double* out; long long* clk; CK(cudaMalloc(&out, 8*64)); CK(cudaMalloc(&clk, 8*16)); int n=4096;
  probe<<<1,32>>>(out, clk, 1.5, n); CK(cudaDeviceSynchronize()); probe<<<1,32>>>(out, clk, 1.5, n); CK(cudaDeviceSynchronize());
  long long h[16]; CK(cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost));
  const char* nm[]={"DFMA dep","DMUL dep","1.0/x + DADD dep","REDUX.MAX dep","SHFL(f64) dep","STS+syncwarp+LDS+syncwarp","FFMA dep","DMMA dep","IMAD dep","select/DFMA"};
  for(int i=0;i<10;i++) printf("%-28s %.1f clk/iter\n", nm[i], (double)h[i]/n); return 0; }

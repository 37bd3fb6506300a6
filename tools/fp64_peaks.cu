// FP64 pipe micro-benchmarks for the roofline denominators (DFMA, DMMA m8n8k4, mixed, exp, sqrt).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

template<int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b){
  double acc[ILP];
  #pragma unroll
  for(int i=0;i<ILP;i++) acc[i]=threadIdx.x*1e-3+i;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<ILP;i++) acc[i]=fma(acc[i],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=acc[i];
  if(s==123.456) out[0]=s;
}

template<int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b){
  double c[NACC][2];
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=threadIdx.x;c[i][1]=i;}
  double A=a+threadIdx.x*1e-9, B=b;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(A), "d"(B));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  if(s==123.456) out[0]=s;
}

// mixed: NACC DMMA + NF DFMA per iteration
template<int NACC,int NF>
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b){
  double c[NACC][2]; double acc[NF];
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=threadIdx.x;c[i][1]=i;}
  #pragma unroll
  for(int i=0;i<NF;i++) acc[i]=threadIdx.x*1e-3+i;
  double A=a+threadIdx.x*1e-9, B=b;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(A), "d"(B));
    #pragma unroll
    for(int i=0;i<NF;i++) acc[i]=fma(acc[i],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  #pragma unroll
  for(int i=0;i<NF;i++) s+=acc[i];
  if(s==123.456) out[0]=s;
}

__global__ void __launch_bounds__(256) k_exp(double* out, int iters, double a){
  double x[4]; for(int i=0;i<4;i++) x[i]=-(threadIdx.x*1e-3+i*0.1);
  double s=0;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<4;i++){ s+=exp(x[i]); x[i]-=a; }
  }
  if(s==123.456) out[0]=s;
}
__global__ void __launch_bounds__(256) k_sqrt(double* out, int iters, double a){
  double x[4]; for(int i=0;i<4;i++) x[i]=(threadIdx.x*1e-3+i*0.1+1.0);
  double s=0;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<4;i++){ s+=sqrt(x[i]); x[i]+=a; }
  }
  if(s==123.456) out[0]=s;
}
__global__ void __launch_bounds__(256) k_write(double2* out, size_t n, double v){
  size_t i = blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t stride=(size_t)gridDim.x*blockDim.x;
  for(; i<n; i+=stride) out[i]=make_double2(v,v+1);
}

template<class F> float timeit(F f, int reps=5){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best=1e30f;
  for(int r=0;r<reps;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best) best=ms; }
  return best;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount; printf("device %s SMs %d clock %d kHz\n",p.name,sms,p.clockRate);
  double* out; CK(cudaMalloc(&out, 1<<20));
  int iters=20000;
  for(int bps : {1,2,4,8}){
    int grid=sms*bps;
    float ms=timeit([&]{k_dfma<8><<<grid,256>>>(out,iters,1.0000001,1e-9);});
    double fl=(double)grid*256*iters*8*2;
    printf("DFMA ilp8 blocks/SM=%d: %.3f ms  %.2f TFLOP/s\n",bps,ms,fl/ms*1e-9);
  }
  for(int bps : {1,2,4,8}){
    int grid=sms*bps;
    float ms=timeit([&]{k_dmma<8><<<grid,256>>>(out,iters,1.0000001,1e-9);});
    double fl=(double)grid*8/*warps*/*iters*8.0*512;
    printf("DMMA nacc8 blocks/SM=%d: %.3f ms  %.2f TFLOP/s\n",bps,ms,fl/ms*1e-9);
  }
  {
    int grid=sms*2;
    float ms=timeit([&]{k_dmma<16><<<grid,256>>>(out,iters,1.0000001,1e-9);});
    double fl=(double)grid*8*iters*16.0*512;
    printf("DMMA nacc16 blocks/SM=2: %.3f ms  %.2f TFLOP/s\n",ms,fl/ms*1e-9);
  }
  {
    int grid=sms*4;
    float msA=timeit([&]{k_dmma<4><<<grid,256>>>(out,iters,1.0000001,1e-9);});
    float msB=timeit([&]{k_dfma<32><<<grid,256>>>(out,iters,1.0000001,1e-9);});
    float msC=timeit([&]{k_mixed<4,32><<<grid,256>>>(out,iters,1.0000001,1e-9);});
    printf("mixed test (4 DMMA + 32 DFMA per iter, 4 blocks/SM): dmma-only %.3f ms, dfma-only %.3f ms, mixed %.3f ms  (sum %.3f, max %.3f)\n",msA,msB,msC,msA+msB,msA>msB?msA:msB);
    double fl=(double)grid*8*iters*4.0*512 + (double)grid*256*iters*32*2;
    printf("mixed combined rate %.2f TFLOP/s\n", fl/msC*1e-9);
  }
  {
    int grid=sms*8; int it2=5000;
    float ms=timeit([&]{k_exp<<<grid,256>>>(out,it2,1e-3);});
    printf("exp(double): %.3f ms  %.2f Gexp/s\n",ms,(double)grid*256*it2*4/ms*1e-6);
    ms=timeit([&]{k_sqrt<<<grid,256>>>(out,it2,1e-3);});
    printf("sqrt(double): %.3f ms  %.2f Gsqrt/s\n",ms,(double)grid*256*it2*4/ms*1e-6);
  }
  {
    size_t bytes=(size_t)4<<30; double2* buf; CK(cudaMalloc(&buf,bytes)); size_t n=bytes/16;
    for(int bps: {4,8,16}){
      float ms=timeit([&]{k_write<<<sms*bps,256>>>(buf,n,1.0);});
      printf("write-only 4GiB blocks/SM=%d: %.3f ms %.1f GB/s\n",bps,ms,bytes/ms*1e-6);
    }
    cudaFree(buf);
  }
  return 0;
}

// Measures the FP64 denominators MEASURED_PEAKS.json lacks:
// DFMA (vector pipe) and DMMA.8x8x4 (tensor pipe) issue-bound throughput on this B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template<int ILP>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters){
  double c0[ILP], c1[ILP];
  #pragma unroll
  for(int j=0;j<ILP;j++){c0[j]=0;c1[j]=0;}
  double a=threadIdx.x*1e-3, b=1.0+threadIdx.x*1e-4;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<ILP;j++) dmma884(c0[j],c1[j],a,b);
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<ILP;j++) s+=c0[j]+c1[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b){
  double c[ILP];
  #pragma unroll
  for(int j=0;j<ILP;j++) c[j]=threadIdx.x+j;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<ILP;j++) c[j]=fma(c[j],a,b);
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<ILP;j++) s+=c[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void __launch_bounds__(256) k_exp(double* out, int iters, double a){
  double s=0; double x=a*threadIdx.x*1e-3;
  for(int i=0;i<iters;i++){ s+=exp(-x); x+=1e-9; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double)*sms*8*256);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  int iters=20000;
  for(int bps=1;bps<=8;bps*=2){
    int grid=sms*bps;
    k_dmma<8><<<grid,256>>>(out,100); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_dmma<8><<<grid,256>>>(out,iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms,e0,e1);
    double fl = 2.0*256*8*(double)iters*8 /*warps*/ *grid;
    printf("{\"kernel\":\"dmma884\",\"blocks_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n",bps,fl/ms*1e-9,ms);
    k_dfma<8><<<grid,256>>>(out,100,1.0000001,1e-9); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_dfma<8><<<grid,256>>>(out,iters,1.0000001,1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms,e0,e1);
    fl = 2.0*8*(double)iters*256*grid;
    printf("{\"kernel\":\"dfma\",\"blocks_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n",bps,fl/ms*1e-9,ms);
  }
  {
    int grid=sms*8; int it=2000;
    k_exp<<<grid,256>>>(out,10,1.0); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_exp<<<grid,256>>>(out,it,1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms,e0,e1);
    printf("{\"kernel\":\"exp_f64\",\"gevals_per_s\":%.2f,\"ms\":%.3f}\n",(double)it*256*grid/ms*1e-6,ms);
  }
  printf("{\"sms\":%d}\n",sms);
  return 0;
}

// Latency microbenchmarks (single warp): dependent DFMA / FFMA chains in a hot loop vs straight-line (cold I-cache),
// double division, double sincos.  Explains why the per-Gaussian projection kernels take ~10 us.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_loop(double* o, double a, double b, long long* cyc, int n){ double x=a; long long t0=clock64(); for(int i=0;i<n;i++) x=fma(x,a,b); long long t1=clock64(); o[threadIdx.x]=x; if(threadIdx.x==0) *cyc=t1-t0; }
__global__ void ffma_loop(float* o, float a, float b, long long* cyc, int n){ float x=a; long long t0=clock64(); for(int i=0;i<n;i++) x=fmaf(x,a,b); long long t1=clock64(); o[threadIdx.x]=x; if(threadIdx.x==0) *cyc=t1-t0; }
template<int N> __global__ void dfma_straight(double* o, double a, double b, long long* cyc){ double x=a; long long t0=clock64();
#pragma unroll
  for(int i=0;i<N;i++) x=fma(x,a,b+i);
  long long t1=clock64(); o[threadIdx.x]=x; if(threadIdx.x==0) *cyc=t1-t0; }
template<int N> __global__ void dfma_straight_ilp4(double* o, double a, double b, long long* cyc){ double x0=a,x1=a+1,x2=a+2,x3=a+3; long long t0=clock64();
#pragma unroll
  for(int i=0;i<N;i++){ x0=fma(x0,a,b+i); x1=fma(x1,a,b+i); x2=fma(x2,a,b+i); x3=fma(x3,a,b+i);} 
  long long t1=clock64(); o[threadIdx.x]=x0+x1+x2+x3; if(threadIdx.x==0) *cyc=t1-t0; }
__global__ void ddiv_loop(double* o, double a, long long* cyc, int n){ double x=a; long long t0=clock64(); for(int i=0;i<n;i++) x=1.0/(x+1.5); long long t1=clock64(); o[threadIdx.x]=x; if(threadIdx.x==0) *cyc=t1-t0; }
__global__ void sincos_loop(double* o, double a, long long* cyc, int n){ double x=a; long long t0=clock64(); for(int i=0;i<n;i++){ double s,c; sincos(x,&s,&c); x=s+c;} long long t1=clock64(); o[threadIdx.x]=x; if(threadIdx.x==0) *cyc=t1-t0; }
int main(){ double* o; float* of; long long* c; cudaMalloc(&o,4096); cudaMalloc(&of,4096); cudaMallocManaged(&c,8);
  auto rep=[&](const char* n, double ops){ cudaDeviceSynchronize(); printf("%-34s %8lld cycles  %7.2f cycles/op\n", n, *c, *c/ops); };
  for(int r=0;r<2;r++){ dfma_loop<<<1,32>>>(o,1.0000001,0.5,c,4096); rep("DFMA dependent, loop (hot I$)",4096); }
  for(int r=0;r<2;r++){ ffma_loop<<<1,32>>>(of,1.0000001f,0.5f,c,4096); rep("FFMA dependent, loop",4096); }
  for(int r=0;r<3;r++){ dfma_straight<2048><<<1,32>>>(o,1.0000001,0.5,c); rep(r==0?"DFMA dependent, 2048 straight COLD":"DFMA dependent, 2048 straight warm",2048); }
  for(int r=0;r<3;r++){ dfma_straight_ilp4<512><<<1,32>>>(o,1.0000001,0.5,c); rep(r==0?"DFMA ILP4, 2048 straight COLD":"DFMA ILP4, 2048 straight warm",2048); }
  for(int r=0;r<2;r++){ ddiv_loop<<<1,32>>>(o,1.3,c,256); rep("double 1/x dependent",256); }
  for(int r=0;r<2;r++){ sincos_loop<<<1,32>>>(o,0.7,c,256); rep("double sincos dependent",256); }
  // launch latency of an empty-ish kernel, back to back
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0); for(int i=0;i<200;i++) ffma_loop<<<1,32>>>(of,1.f,0.5f,c,1); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); printf("tiny kernel back-to-back: %.2f us per launch\n", ms*1000/200);
  return 0; }

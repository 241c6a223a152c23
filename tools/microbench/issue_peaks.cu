// Microbenchmarks of the issue/pipe peaks that bound the splat kernels on sm_100a:
// FFMA, FADD, FMUL, FFMA2 (fma.rn.f32x2), MUFU.EX2, and FFMA:EX2 mixes.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__device__ __forceinline__ float ex2(float x){float y; asm volatile("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
template<int MODE> __global__ void __launch_bounds__(256) k(float* out, float a, float b){
  float v[16];
  #pragma unroll
  for(int i=0;i<16;i++) v[i]=threadIdx.x*0.001f+i;
  for(int it=0; it<ITERS; ++it){
    if(MODE==0){ // FFMA x16
      #pragma unroll
      for(int i=0;i<16;i++) v[i]=fmaf(v[i],a,b);
    } else if(MODE==1){ // FADD
      #pragma unroll
      for(int i=0;i<16;i++) v[i]=v[i]+a;
    } else if(MODE==2){ // FMUL
      #pragma unroll
      for(int i=0;i<16;i++) v[i]=v[i]*a;
    } else if(MODE==3){ // EX2 x16
      #pragma unroll
      for(int i=0;i<16;i++) v[i]=ex2(v[i]);
    } else if(MODE==4){ // FFMA2 x8 (16 lanes-ops)
      #pragma unroll
      for(int i=0;i<16;i+=2){
        unsigned long long x,y,z,r;
        asm volatile("mov.b64 %0, {%1,%2};":"=l"(x):"f"(v[i]),"f"(v[i+1]));
        asm volatile("mov.b64 %0, {%1,%1};":"=l"(y):"f"(a));
        asm volatile("mov.b64 %0, {%1,%1};":"=l"(z):"f"(b));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(x),"l"(y),"l"(z));
        asm volatile("mov.b64 {%0,%1}, %2;":"=f"(v[i]),"=f"(v[i+1]):"l"(r));
      }
    } else if(MODE==9 || MODE==10 || MODE==11){ // FMUL2 / FADD2 / FFMA2 with a zero addend (a packed multiply on the FMA path)
      #pragma unroll
      for(int i=0;i<16;i+=2){
        unsigned long long x,y,z,r;
        asm volatile("mov.b64 %0, {%1,%2};":"=l"(x):"f"(v[i]),"f"(v[i+1]));
        asm volatile("mov.b64 %0, {%1,%1};":"=l"(y):"f"(a));
        if(MODE==9) asm volatile("mul.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(x),"l"(y));
        else if(MODE==10) asm volatile("add.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(x),"l"(y));
        else { asm volatile("mov.b64 %0, {%1,%1};":"=l"(z):"f"(0.0f)); asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(x),"l"(y),"l"(z)); }
        asm volatile("mov.b64 {%0,%1}, %2;":"=f"(v[i]),"=f"(v[i+1]):"l"(r));
      }
    } else if(MODE==12){ // the backward strip mix: 20 FMUL2 + 25 FFMA2 + 1 FADD2 per 4 MUFU
      #pragma unroll
      for(int i=0;i<16;i+=2){
        unsigned long long x,y,z,r;
        asm volatile("mov.b64 %0, {%1,%2};":"=l"(x):"f"(v[i]),"f"(v[i+1]));
        asm volatile("mov.b64 %0, {%1,%1};":"=l"(y):"f"(a));
        asm volatile("mov.b64 %0, {%1,%1};":"=l"(z):"f"(b));
        if((i&2)==0) asm volatile("mul.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(x),"l"(y));
        else asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(x),"l"(y),"l"(z));
        asm volatile("mov.b64 {%0,%1}, %2;":"=f"(v[i]),"=f"(v[i+1]):"l"(r));
      }
    } else if(MODE==5){ // 4 FFMA : 1 EX2 (x4)
      #pragma unroll
      for(int i=0;i<16;i+=4){ v[i]=fmaf(v[i],a,b); v[i+1]=fmaf(v[i+1],a,b); v[i+2]=fmaf(v[i+2],a,b); v[i+3]=ex2(fmaf(v[i+3],a,b)); }
    } else if(MODE==6){ // 8 FFMA : 1 EX2 (x2), 9 slots per EX2 like the algorithmic fwd count
      #pragma unroll
      for(int i=0;i<16;i+=8){
        #pragma unroll
        for(int j=0;j<7;j++) v[i+j]=fmaf(v[i+j],a,b);
        v[i+7]=ex2(fmaf(v[i+7],a,b)); }
    } else if(MODE==7){ // 6 : 1
      #pragma unroll
      for(int i=0;i<12;i+=6){
        #pragma unroll
        for(int j=0;j<5;j++) v[i+j]=fmaf(v[i+j],a,b);
        v[i+5]=ex2(fmaf(v[i+5],a,b)); }
    } else if(MODE==8){ // fwd inner loop shape: FADD, FMUL, FFMA, EX2, FADD
      #pragma unroll
      for(int i=0;i<16;i+=2){ float t=v[i]-a; float g=ex2(fmaf(b*t,t,a)); v[i+1]+=g; v[i]=t; }
    }
  }
  float s=0; for(int i=0;i<16;i++) s+=v[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int MODE> void run(const char* name, double ops_per_iter_thread, double ex2_per_iter_thread){
  float* out; cudaMalloc(&out, 148*8*256*4*2);
  int blocks=148*8; // 8 CTAs of 256 = 2048 threads/SM
  k<MODE><<<blocks,256>>>(out,1.0001f,0.5f); cudaDeviceSynchronize();
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); for(int r=0;r<5;r++) k<MODE><<<blocks,256>>>(out,1.0001f,0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); ms/=5;
  double thr=(double)blocks*256*ITERS;
  printf("%-28s %8.3f ms  %7.2f T lane-instr/s  (%5.2f warp-instr/clk/SM @1965MHz)  ex2 %6.2f T/s\n", name, ms, thr*ops_per_iter_thread/ms/1e9, thr*ops_per_iter_thread/ms/1e9*1e12/32/148/1.965e9, thr*ex2_per_iter_thread/ms/1e9);
  cudaFree(out);
}
int main(){
  run<0>("FFMA",16,0); run<1>("FADD",16,0); run<2>("FMUL",16,0); run<3>("MUFU.EX2",16,16); run<4>("FFMA2 (lane-ops=2/instr)",16,0);
  run<9>("FMUL2 (lane-ops=2/instr)",16,0); run<10>("FADD2 (lane-ops=2/instr)",16,0); run<11>("FFMA2 with zero addend",16,0); run<12>("FMUL2 : FFMA2 = 1 : 1",16,0);
  run<5>("4 FFMA : 1 EX2 (5 slots)",20,4); run<7>("6 FFMA: 1 EX2 (7 slots)",14,2); run<6>("8 FFMA : 1 EX2 (9 slots)",18,2); run<8>("fwd-loop shape (5 slots/px)",40,8);
  return 0;
}

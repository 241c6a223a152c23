// Does FFMA2 keep its full rate when its three 64-bit operands are distinct registers (register-file read
// bandwidth / bank conflicts), as in the backward splat's  S += G[j] * W[j]  accumulation?
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c){ f32x2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b){ f32x2 r; asm volatile("mul.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r; }
__device__ __forceinline__ f32x2 pack2(float lo, float hi){ f32x2 r; asm("mov.b64 %0, {%1, %2};":"=l"(r):"f"(lo),"f"(hi)); return r; }
#define ITERS 2048
// MODE 0: s_m += G[j] * W[m][j]  (24 FFMA2, all operands distinct registers: 8 G + 24 W + 3 S pairs)
// MODE 1: same but W shared across m (operand reuse)      MODE 2: scalar FFMA version of MODE 0 (48 FFMA)
// MODE 3: MODE 0 plus the 13-step g/r recurrence (FMUL2 chain) feeding G
template<int MODE> __global__ void __launch_bounds__(256) k(float* out, float a){
  f32x2 G[8], W[3][8], S[3];
  float gs[16], ws[3][16], ss[3];
  for(int j=0;j<8;j++){ G[j]=pack2(a+j, a-j); for(int m=0;m<3;m++) W[m][j]=pack2(a*m+j, a*j-m); }
  for(int j=0;j<16;j++){ gs[j]=a+j; for(int m=0;m<3;m++) ws[m][j]=a*m+j; }
  for(int m=0;m<3;m++){ S[m]=pack2(0.f,0.f); ss[m]=0.f; }
  f32x2 g=pack2(a,a*0.5f), r=pack2(1.0001f,0.9999f), c=pack2(1.00001f,1.00001f);
  for(int it=0; it<ITERS; ++it){
    if(MODE==3){
      #pragma unroll
      for(int j=0;j<8;j++){ G[j]=g; if(j<7) g=mul2(g,r); if(j<6) r=mul2(r,c); }
    }
    if(MODE==0 || MODE==3){
      #pragma unroll
      for(int j=0;j<8;j++){
        #pragma unroll
        for(int m=0;m<3;m++) S[m]=fma2(G[j],W[m][j],S[m]);
      }
    } else if(MODE==1){
      #pragma unroll
      for(int j=0;j<8;j++){
        #pragma unroll
        for(int m=0;m<3;m++) S[m]=fma2(G[j],W[0][j],S[m]);
      }
    } else if(MODE==2){
      #pragma unroll
      for(int j=0;j<16;j++){
        #pragma unroll
        for(int m=0;m<3;m++) ss[m]=fmaf(gs[j],ws[m][j],ss[m]);
      }
    }
  }
  float s=ss[0]+ss[1]+ss[2];
  for(int m=0;m<3;m++){ float lo,hi; asm("mov.b64 {%0,%1}, %2;":"=f"(lo),"=f"(hi):"l"(S[m])); s+=lo+hi; }
  { float lo,hi; asm("mov.b64 {%0,%1}, %2;":"=f"(lo),"=f"(hi):"l"(g)); s+=lo+hi; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int MODE> void run(const char* name, double laneops, int ctas_per_sm){
  float* out; cudaMalloc(&out, 148*8*256*4);
  int blocks=148*ctas_per_sm;
  k<MODE><<<blocks,256>>>(out,1.0001f); cudaDeviceSynchronize();
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); for(int r=0;r<5;r++) k<MODE><<<blocks,256>>>(out,1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); ms/=5;
  double thr=(double)blocks*256*ITERS;
  printf("%-52s %d CTA/SM %8.3f ms  %6.2f T lane-ops/s (%4.2f warp-lane-instr/clk/SM)\n", name, ctas_per_sm, ms, thr*laneops/ms/1e9, thr*laneops/ms/1e9*1e12/32/148/1.965e9);
  cudaFree(out);
}
int main(){
  for(int c: {3, 8}){
    run<0>("24 FFMA2, distinct G/W/S registers", 48, c);
    run<1>("24 FFMA2, W shared across the 3 sums", 48, c);
    run<2>("48 scalar FFMA, distinct registers", 48, c);
    run<3>("13 FMUL2 recurrence + 24 FFMA2", 74, c);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// Does a packed FP32x2 instruction (2 pipe cycles) leave the issue port free for a non-FMA instruction in its second
// cycle?  Per iteration: 8 independent fma.rn.f32x2 + M instructions of another pipe (ALU: lop3 / LSU: lds / MUFU: ex2).
// 8 warps per scheduler.  If the other instructions hide under the FFMA2s the time stays at the M = 0 value until
// M = 8; if FFMA2 holds the issue port for both cycles it grows by one cycle per extra instruction from the start.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o coissue coissue.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
typedef unsigned long long u64;
template <int KIND, int M, bool SCALAR>
__global__ void __launch_bounds__(256, 4) k(float *out, float a, float b, unsigned m) {
  __shared__ float sm[256];
  sm[threadIdx.x] = a;
  __syncthreads();
  u64 v[8];
  float w[16];
  unsigned q[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) asm volatile("mov.b64 %0, {%1,%2};" : "=l"(v[i]) : "f"(threadIdx.x * 0.001f + i), "f"(b + i));
#pragma unroll
  for (int i = 0; i < 16; ++i) { q[i] = threadIdx.x + i; w[i] = threadIdx.x * 0.01f + i; }
  u64 ya, zb;
  asm volatile("mov.b64 %0, {%1,%1};" : "=l"(ya) : "f"(a));
  asm volatile("mov.b64 %0, {%1,%1};" : "=l"(zb) : "f"(b));
  for (int it = 0; it < ITERS; ++it) {
    if (SCALAR) {
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = fmaf(w[i], a, b);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v[i]) : "l"(v[i]), "l"(ya), "l"(zb));
    }
#pragma unroll
    for (int i = 0; i < M; ++i) {
      if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(m), "r"(q[(i + 1) & 15]));
      else if (KIND == 1) { float t; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"((unsigned)(((q[i] & 255u)) * 4u))); q[i] += __float_as_uint(t) & 1u; }
      else if (KIND == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(w[i]));
      else if (KIND == 3) asm volatile("st.shared.f32 [%0], %1;" :: "r"((unsigned)(threadIdx.x * 4u)), "f"(w[i]) : "memory");
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) { float lo, hi; asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i])); s += lo + hi; }
  for (int i = 0; i < 16; ++i) s += w[i] + q[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND, int M, bool SCALAR>
void run(const char *name) {
  float *out;
  cudaMalloc(&out, 148 * 4 * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<KIND, M, SCALAR><<<148 * 4, 256>>>(out, 0.999f, 0.001f, 0x55u);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<KIND, M, SCALAR><<<148 * 4, 256>>>(out, 0.999f, 0.001f, 0x55u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 5;
  // cycles per iteration per scheduler: 8 warps per scheduler share it
  const double cyc = ms * 1e-3 * 1.965e9 / ITERS / 8.0;
  printf("%-34s M=%2d  %.3f ms  %.2f cycles per (warp, iteration) per scheduler\n", name, M, ms, cyc);
  cudaFree(out);
}
#define SWEEP(KIND, SC, NAME) run<KIND, 0, SC>(NAME); run<KIND, 2, SC>(NAME); run<KIND, 4, SC>(NAME); run<KIND, 8, SC>(NAME); run<KIND, 12, SC>(NAME); run<KIND, 16, SC>(NAME);
int main() {
  SWEEP(0, false, "8 FFMA2 + M LOP3")
  SWEEP(0, true, "16 FFMA (scalar) + M LOP3")
  SWEEP(1, false, "8 FFMA2 + M (LDS + IADD)")
  SWEEP(3, false, "8 FFMA2 + M STS")
  run<2, 0, false>("8 FFMA2 + M EX2"); run<2, 1, false>("8 FFMA2 + M EX2"); run<2, 2, false>("8 FFMA2 + M EX2"); run<2, 4, false>("8 FFMA2 + M EX2");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

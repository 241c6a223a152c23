// gg_common.cuh -- shared device helpers and host-side launch descriptors (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gaussigan.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libgaussigan_sm100 is written for sm_100a (B200) only"
#endif

namespace gg {

constexpr int kThreads = 256;          // CTA size of every splat kernel (8 warps)
constexpr int kWarps = kThreads / 32;
constexpr int kStrip = 8;              // pixels per thread-strip along x (two 128-bit accesses)
constexpr int kKBlock = 64;            // Gaussians handled per CTA along grid.z in the backward

// Staged per-Gaussian record, GG_PARAM_STRIDE floats, written by the projection kernels:
//   [0] mu_x  [1] mu_y  [2] nA = -a*log2(e)  [3] nr = -(b/a)  [4] nE = -(c - b*b/a)*log2(e)
//   [5..7] a, b, c (kept for debugging / future use)
// so that  g(x, y) = exp2( nA * t^2 + nE * dy^2 ),  dy = y - mu_y,  t = x - (mu_x + nr*dy)
// which is exp(-(a dx^2 + 2 b dx dy + c dy^2)) after completing the square along x.
struct ScaleDesc {
  int n;           // image side
  int spr;         // strips (threads) per row  = ceil(ceil(n/4) / 2); a strip = quads s and s + spr
  int sprc;        // strips per row per chunk  = min(spr, kThreads)
  int rpp;         // rows per pass of one CTA  = max(1, kThreads / spr)
  int chunks;      // column chunks             = ceil(spr / kThreads)
  int rows_per_tile;
  int tile_begin;  // first tile index of this scale inside the launch
  int tiles;       // tiles of this scale per batch element
  int contig;      // 1: a strip is 8 CONSECUTIVE pixels (gg_splat_fast.cu); 0: two half-row quads (gg_strip.cuh)
  int cta;         // threads per CTA this geometry was planned for (kThreads except in gg_splat_fast.cu)
  float step;      // float32 (2 / (n-1)), TF 1.12 LinSpace step (mask_gen_dynamic.py:121-122)
  int cstride;     // NHWC maps: floats from one pixel to the next (K when dense; C_total when the K channels live at
                   // channel offset c0 of a wider concat buffer -- `maps` then points at channel c0 of pixel 0)
  float *maps;     // forward outputs / backward upstream gradients (const in the backward)
  float *mask;
};

struct SplatArgs {
  const float *params;   // [B,K,GG_PARAM_STRIDE]
  float *partials;       // backward: [B,T,K,GG_MOMENTS]
  int B, K;
  int T;                 // tiles per batch element in THIS launch (splits blockIdx.x)
  int slots;             // backward: total partial slots per batch element (partials stride)
  int nscales;
  int layout;
  int tile_offset;       // backward: first partial slot written by this launch
  // fused silhouette-loss kernel only (gg_silhouette_loss_fused):
  const void *target;    // [B,n,n] uint8 (raw) or float32 ([-1,1])
  int target_kind;
  float inv_n;           // 1 / (B*n*n)
  float *loss_partials;  // [B*T] per-CTA sums of |residual|
  ScaleDesc sc[GG_MAX_SCALES];
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Pixel-centre coordinate exactly as TF 1.12 LinSpace evaluates it in float32:
// out[j] = start + step * j, product and sum each rounded (no FMA contraction).
__device__ __forceinline__ float grid_coord(int j, float step) {
  return __fadd_rn(-1.0f, __fmul_rn(step, (float)j));
}

// ---- packed FP32x2 arithmetic (sm_100a: FFMA2 / FADD2 / FMUL2).  One instruction = two IEEE-rn float32
// operations on a 64-bit register pair: same lane throughput as the scalar forms (measured,
// tools/microbench/issue_peaks.cu) at half the issue slots -- which is what an issue-bound loop needs.
#ifndef GG_USE_F32X2
#define GG_USE_F32X2 1
#endif
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// ---- programmatic dependent launch (PDL).  Every kernel of the library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may become resident while the previous kernel of the
// stream drains, run their index arithmetic, and block in pdl_wait() until that kernel has completed and its
// memory is visible.  No global memory is touched before pdl_wait(); pdl_launch_dependents() right after it lets
// the NEXT kernel do the same with respect to this one.  Hides the launch gap and the first wave's prologue in the
// 3-4 kernel chains of a step (each kernel is only 10-30 us at the reference's batch sizes).  GG_PDL=0 disables.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_launch_dependents(); }
// same, but the wait is made to depend on already-computed values so the compiler cannot sink the index
// arithmetic that produced them below it
__device__ __forceinline__ void pdl_enter_after(int i0, int i1, float f0, float f1) {
  asm volatile("griddepcontrol.wait;" ::"r"(i0), "r"(i1), "f"(f0), "f"(f1) : "memory");
  pdl_launch_dependents();
}

bool pdl_enabled();   // gg_capi.cu
int set_pdl(int enable);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// keep a value in its register: stops the compiler from re-deriving cheap loop invariants (a grid coordinate is an
// I2F + constant-bank load + FMUL + FADD) inside the hot loop to save a register
__device__ __forceinline__ float pin(float v) { asm volatile("" : "+f"(v)); return v; }
__device__ __forceinline__ f32x2 pin2(f32x2 v) { asm volatile("" : "+l"(v)); return v; }

__device__ __forceinline__ void st_f4(float *p, float a, float b, float c, float d) {
  *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}

}  // namespace gg

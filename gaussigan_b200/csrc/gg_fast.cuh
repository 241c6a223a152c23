// gg_fast.cuh -- pieces shared by the forward-differenced kernels (gg_splat_fast.cu: silhouette; gg_splat_tma.cu:
// NHWC maps backward): the per-(CTA, Gaussian) constants of the recurrence and the guard that decides where it is used.
// The derivation and the accuracy study are in the header of gg_splat_fast.cu.
#pragma once
#include "gg_common.cuh"

namespace gg {

#ifndef GG_MAX_VISIBLE_SLOPE
#define GG_MAX_VISIBLE_SLOPE 3.0f
#endif
#ifndef GG_MAX_SHARPNESS
#define GG_MAX_SHARPNESS 512.0f
#endif
#ifndef GG_MAX_SLOPE
#define GG_MAX_SLOPE 10.0f
#endif
// largest |d(exponent)/d(pixel)| the strip forms are trusted with (range safety): the backward's Horner sums run in
// units of the strip's first pixel, so they reach 2^(7 * 10) times the upstream gradient -- finite up to |Gbar| = 2^55
constexpr float kMaxSlope = GG_MAX_SLOPE;
constexpr float kMaxVisibleSlope = GG_MAX_VISIBLE_SLOPE;   // ... and where the Gaussian is visible (accuracy)
constexpr float kMaxSharpness = GG_MAX_SHARPNESS;    // largest |nA| = a log2(e): bounds the error of stepping by exactly h

__device__ __forceinline__ float exp2_small_f32(float delta) {   // exp2(delta), |delta| <= 0.125, ~half an ulp
  const float z = delta * 0.693147180559945f;
  float q = fmaf(z, 1.0f / 6.0f, 1.0f);
  q = fmaf(z * 0.2f, q, 1.0f);
  q = fmaf(z * 0.25f, q, 1.0f);
  q = fmaf(z * (1.0f / 3.0f), q, 1.0f);
  q = fmaf(z * 0.5f, q, 1.0f);
  return fmaf(z, q, 1.0f);
}

// Per-Gaussian constants of the recurrence at grid step h:
//   c0 = mu_x, mu_y, nA, -nr      c1 = nE, 2 nA h, nA h^2, c = exp2(2 nA h^2)      steep: evaluate directly instead
// `steep` is decided per (tile, Gaussian) from the tile's corners: t = dx + r dy is linear in the pixel position,
// so max |t| over the tile is attained at a corner.
__device__ __forceinline__ void fast_consts_one(const float4 *__restrict__ gp, int k, float h, float x_lo, float x_hi,
                                                float y_lo, float y_hi, float4 &c0, float4 &c1, bool &steep) {
  const float4 lo = __ldg(gp + 2 * k);
  const float nE = __ldg(reinterpret_cast<const float *>(gp + 2 * k + 1));
  const float nAh = __fmul_rn(lo.z, h);
  const float nAhh = __fmul_rn(nAh, h);
  // c = exp2(2 nA h^2), needed to ~half an ulp (its error is amplified 21x over a strip).  The argument is tiny at
  // the reference's sizes (|delta| <= 0.03 at n = 256): 1 + p with a degree-6 Taylor polynomial p in float32 is
  // as good as the float64 value rounded; larger arguments (coarse grids, thin Gaussians) take the float64 path.
  const float delta = 2.0f * nAhh;
  float c;
  if (fabsf(delta) <= 0.125f) {
    c = exp2_small_f32(delta);
  } else {
    c = (float)exp2(2.0 * (double)lo.z * (double)h * (double)h);
  }
  float tmax = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float xx = (i & 1) ? x_hi : x_lo, yy = (i & 2) ? y_hi : y_lo;
    tmax = fmaxf(tmax, fabsf((xx - lo.x) - lo.w * (yy - lo.y)));
  }
  // exponent slope over a strip: |b_j| <= |2 nA h| (|t| + 8 h)
  const float slope = fabsf(2.0f * nAh) * (tmax + 8.0f * h);
  // accuracy: where the Gaussian is not negligible (exponent > -14) the slope is at most 2 h sqrt(14 |nA|); the
  // rounding of the ratio's exponent grows with it ...
  const float slope_vis = 2.0f * h * sqrtf(14.0f * fabsf(lo.z));
  // ... and the recurrence steps by exactly h while the reference's float32 grid deviates from uniform by up to
  // 1.8e-7: the resulting error peaks at 0.6 sqrt(|nA|) * 1.8e-7, i.e. 3e-6 at |nA| = 512 (the reference's range
  // ends near 210)
  steep = !(slope <= kMaxSlope) || !(slope_vis <= kMaxVisibleSlope) || !(fabsf(lo.z) <= kMaxSharpness);
  c0 = make_float4(lo.x, lo.y, lo.z, -lo.w);
  c1 = make_float4(nE, 2.0f * nAh, nAhh, c);
}

// One CTA's Gaussians in shared memory, float4 records:
//   [0] c0 = mu_x, mu_y, nA, -nr
//   [1] nE, 2 nA h, nA h^2, c            forward  (g_{j+1} = g_j r_j, r_0 = exp2(2 nA h t_0 + nA h^2), r_{j+1} = r_j c)
//   [.] nE, 2 nA h, 13 nA h^2, 1/c       backward (Horner form, ratios from the far end of the strip:
//                                                  r_6 = exp2(2 nA h t_0 + 13 nA h^2), r_{j-1} = r_j / c)
// The last word is 0 for a guarded Gaussian (c and 1/c are positive otherwise): evaluate directly.
// WHICH = 1: forward record only (2 float4 per Gaussian), 2: backward only (2), 3: both (3: c0, forward, backward).
template <int WHICH>
__device__ __forceinline__ void stage_fast_consts(const float *__restrict__ gparams, int kn, float h, float x_lo,
                                                  float x_hi, float y_lo, float y_hi, float4 *sc) {
  constexpr int REC = WHICH == 3 ? 3 : 2;
  const float4 *gp = reinterpret_cast<const float4 *>(gparams);
  for (int k = threadIdx.x; k < kn; k += blockDim.x) {
    float4 c0, c1;
    bool steep;
    fast_consts_one(gp, k, h, x_lo, x_hi, y_lo, y_hi, c0, c1, steep);
    sc[REC * k] = c0;
    if (WHICH & 1) sc[REC * k + 1] = make_float4(c1.x, c1.y, c1.z, steep ? 0.f : c1.w);
    if (WHICH & 2) {
      const float delta = -2.0f * c1.z;
      float ci = 0.f;
      if (!steep)
        ci = fabsf(delta) <= 0.125f ? exp2_small_f32(delta) : (float)exp2(-2.0 * (double)c0.z * (double)h * (double)h);
      sc[REC * k + (WHICH == 3 ? 2 : 1)] = make_float4(c1.x, c1.y, __fmul_rn(13.0f, c1.z), ci);
    }
  }
}

// The same constants interleaved for kernels whose FP32x2 lanes are two consecutive GAUSSIANS (gg_splat_tma.cu):
// word w of Gaussian k goes to float index (k / 2) * 16 + 2 w + (k & 1) (w = 0..3: c0, 4..7: c1), so a lane pair is
// loaded as aligned 64-bit halves of four LDS.128; the steep flags follow at float index 8 Kp + k, Kp = kn rounded
// up to even.  For odd kn the missing partner of the last Gaussian gets all-zero constants (g = 1, finite).
__device__ __forceinline__ void stage_fast_consts_paired(const float *__restrict__ gparams, int kn, float h, float x_lo,
                                                         float x_hi, float y_lo, float y_hi, float *sc) {
  const float4 *gp = reinterpret_cast<const float4 *>(gparams);
  const int Kp = (kn + 1) & ~1;
  for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
    float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
    bool steep = false;
    if (k < kn) fast_consts_one(gp, k, h, x_lo, x_hi, y_lo, y_hi, c0, c1, steep);
    float *d = sc + (k >> 1) * 16 + (k & 1);
    d[0] = c0.x; d[2] = c0.y; d[4] = c0.z; d[6] = c0.w;
    d[8] = c1.x; d[10] = c1.y; d[12] = c1.z; d[14] = c1.w;
    sc[8 * Kp + k] = steep ? 1.f : 0.f;
  }
}

__device__ __forceinline__ f32x2 bc(float v) { return pack2(v, v); }
__device__ __forceinline__ f32x2 ex2_2(f32x2 v) {
  float lo, hi;
  unpack2(v, lo, hi);
  return pack2(ex2_approx(lo), ex2_approx(hi));
}

}  // namespace gg

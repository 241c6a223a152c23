// gg_splat_fast.cu -- mask-only splat kernels with the Gaussians FORWARD-DIFFERENCED along each pixel strip.
//
// Same contract as the strip kernels of gg_splat.cu in their mask-only modes (sum composite
// mask/mask_gen_dynamic.py:785-786, its gradient, and the fused m_recon_loss_bis :785-787), different
// arithmetic.  The direct form pays one MUFU.EX2 per (pixel, Gaussian) and the MUFU pipe (16 lanes/clk/SM)
// is what bounds it.  Along a row the pixel centres are equally spaced (step h), so with
//     a_j = nA t_j^2 + e,   t_j = t_0 + j h          (log2 of the Gaussian at pixel j of the strip)
//     g_{j+1} = g_j r_j,    r_j = exp2(nA h (2 t_j + h)),    r_{j+1} = r_j c,    c = exp2(2 nA h^2)
// a strip of 8 pixels needs TWO exponentials (g_0, r_0) instead of eight; c is computed once per
// (CTA, Gaussian) to half an ulp (gg_fast.cuh).  MUFU work drops 4x and the kernels become FP32-pipe bound.
//
// Lanes.  A thread owns one 8-pixel strip of TWO rows (row, row + rpp).  Every FP32x2 register holds the
// same quantity for the two rows, so the whole per-(Gaussian,row) set-up (dy, t0, the two exponents), the
// recurrence and the moment folding run as packed FFMA2/FMUL2/FADD2: half the issue slots of the scalar form.
//
// Accuracy (tests/studies/recurrence_accuracy.py, float32 emulation against the float64 oracle, reference
// activation ranges): silhouette max|err|/max = 8e-7 (direct form: 1e-7; tolerance 1e-5).
//
// Guard.  The strip forms are only used where the per-pixel change of the exponent is small,
// |nA h (2 t + h)| <= 10 over the whole tile: then nothing can overflow and a strip whose first pixel
// underflows (g_0 < 2^-126) stays below 2^-56 everywhere; and <= 3 wherever the Gaussian is visible (the rounding
// error of the ratio grows with its exponent; the reference's range stays below 0.8); and only for |nA| <= 512,
// because the recurrence steps by exactly h while the reference's float32 pixel grid deviates from uniform by up to
// 1.8e-7, an error that grows like sqrt(|nA|) (the reference's range ends near 210).  The test is made once per
// (CTA, Gaussian) from the tile's corners (t is linear in the pixel position); Gaussians that fail it (narrower than
// ~2 pixels, or very distant) are evaluated directly, one exponential per pixel on the true grid, by the whole CTA.
//
// Backward.  Per strip and Gaussian three sums about the strip's first pixel are needed,
//     S0 = sum_j W_j g_j,   S1 = sum_j xi_j W_j g_j,   S2 = sum_j xi_j^2 W_j g_j       (W = dL/dmask, xi_j = x_j - x_0
// on the reference's float32 grid), re-centred afterwards to the sheared coordinate t = t0 + xi per row.  Term by term
// that is g_j (2 multiplies for the two recurrences), u_j = g_j W_j and three multiply-adds: 6 lane-ops per pixel and
// Gaussian (round 1).  Round 2 evaluates each sum in HORNER form in the ratios r_j = g_{j+1} / g_j,
//     S = g_0 (C_0 + r_0 (C_1 + r_1 (C_2 + ... + r_6 C_7)))         C_j = W_j,  xi_j W_j,  xi_j^2 W_j
// whose coefficients do not depend on the Gaussian and are formed once per thread: one multiply for the ratio
// (r_{j-1} = r_j / c, started from r_6 = exp2(2 nA h t_0 + 13 nA h^2)) and three multiply-adds, 4 lane-ops.  The 24
// coefficient pairs cost registers: 80 per thread, 3 CTAs per SM instead of 4 at 64.  Measured at B = 64, K = 16,
// 256 x 256: 24.3 -> 22.4 us per launch; SASS 54 -> 43 packed FP32x2 instructions per (Gaussian, 2 rows x 8 pixels).
// The Horner sums run in units of g_0, so they reach 2^(7 x 10) times the upstream gradient: finite for |W| < 2^55.
// Guarded Gaussians are summed term by term from the direct values (CENTRE kernels, used for launches with a scale
// below 64 pixels: about their own centre line instead of the strip origin, see the kernel).
#include "gg_common.cuh"
#include "gg_strip.cuh"
#include "gg_fast.cuh"
#include "gg_launch.h"

namespace gg {

namespace {

#ifndef GG_BWD_KUNROLL
#define GG_BWD_KUNROLL 1      // Gaussians processed per main-loop iteration of the backward (ILP vs registers)
#endif
#ifndef GG_BWD_MINBLOCKS
#define GG_BWD_MINBLOCKS 3    // resident 256-thread CTAs per SM the backward is compiled for
#endif
#ifndef GG_FWD_KUNROLL
#define GG_FWD_KUNROLL 0      // 0 = leave it to the compiler
#endif
constexpr int kBwdKUnroll = GG_BWD_KUNROLL;

// One Gaussian on one strip of a ROW PAIR: lanes of every f32x2 are the two rows.
//   dy = y - mu_y;  t0 = (x_0 - mu_x) - nr dy;  e = nE dy^2   (per lane)
// Guarded (narrow or very distant) Gaussian: one exponential per pixel on the reference's own float32 grid coordinates
// (not x_0 + j h: at this sharpness the 1e-7 non-uniformity of the grid matters), exactly as gg_splat.cu does.
__device__ __forceinline__ void eval_strip_direct(const float4 &c0, f32x2 e, f32x2 dy, int col0, float h, f32x2 (&G)[8]) {
  const f32x2 nA = bc(c0.z);
  const f32x2 sh = fma2(bc(-c0.w), dy, bc(c0.x));          // mu_x + nr dy
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const f32x2 t = sub2(bc(grid_coord(col0 + j, h)), sh);
    G[j] = ex2_2(fma2(nA, mul2(t, t), e));
  }
}

// Forward: the 8 values of a strip by the recurrence g_{j+1} = g_j r_j, r_{j+1} = r_j c: 2 exponentials and 13 multiplies.
// (The closed form g_j = g_0 rho^j Q_j with per-Gaussian tables Q_j = exp2(nA (j h)^2) needs 7 multiplies and 7
// multiply-adds into the accumulators but two more 128-bit shared-memory loads per Gaussian: measured 12.66 us against
// 12.99 us at B = 64, K = 16, 256 x 256 -- not kept, the fused-loss kernel has no shared memory left for the tables and
// its silhouette has to equal this kernel's bit for bit.)
__device__ __forceinline__ void eval_strip(const float4 &c0, const float4 &c1, bool steep, f32x2 t0, f32x2 e, f32x2 dy,
                                           int col0, float h, f32x2 (&G)[8]) {
  if (steep) {
    eval_strip_direct(c0, e, dy, col0, h, G);
    return;
  }
  f32x2 g = ex2_2(fma2(mul2(bc(c0.z), t0), t0, e));
  f32x2 r = ex2_2(fma2(bc(c1.y), t0, bc(c1.z)));
  const f32x2 c = bc(c1.w);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    G[j] = g;
    if (j < 7) g = mul2(g, r);
    if (j < 6) r = mul2(r, c);
  }
}

// Backward: the three strip sums  S0 = sum W_j g_j,  S1 = sum xi_j W_j g_j,  S2 = sum xi_j^2 W_j g_j  (W = upstream
// gradient, xi_j = x_j - x_0 on the reference's float32 grid) as three Horner evaluations in the ratios r_j = g_{j+1}/g_j:
//   S = g_0 (C_0 + r_0 (C_1 + r_1 (C_2 + ... r_6 C_7)))      with C_j = W_j, xi_j W_j, xi_j^2 W_j
// The coefficients do not depend on the Gaussian: they are formed once per thread.  Per pixel and Gaussian this is one
// multiply for the ratio and three multiply-adds (the term-by-term form needs g_j, u_j = g_j W_j and three sums: six).
__device__ __forceinline__ void horner_strip(const float4 &c0, const float4 &c1, f32x2 t0, f32x2 e,
                                             const f32x2 (&W)[8], const f32x2 (&A)[8], const f32x2 (&Bc)[8],
                                             f32x2 &S0, f32x2 &S1, f32x2 &S2) {
  const f32x2 g0 = ex2_2(fma2(mul2(bc(c0.z), t0), t0, e));
  f32x2 r = ex2_2(fma2(bc(c1.y), t0, bc(c1.z)));        // r_6
  const f32x2 ci = bc(c1.w);
  f32x2 T = W[7], U = A[7], V = Bc[7];
#pragma unroll
  for (int m = 6; m >= 1; --m) {
    T = fma2(r, T, W[m]);
    U = fma2(r, U, A[m]);
    V = fma2(r, V, Bc[m]);
    r = mul2(r, ci);
  }
  T = fma2(r, T, W[0]);
  U = mul2(r, U);         // xi_0 = 0
  V = mul2(r, V);
  S0 = mul2(g0, T);
  S1 = mul2(g0, U);
  S2 = mul2(g0, V);
}

__device__ __forceinline__ void store8c(float *rowp, int col0, int n, const float *f) {
  if ((n & 3) == 0) {
    if (col0 < n) st_f4(rowp + col0, f[0], f[1], f[2], f[3]);
    if (col0 + 4 < n) st_f4(rowp + col0 + 4, f[4], f[5], f[6], f[7]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (col0 + j < n) rowp[col0 + j] = f[j];
  }
}

__device__ __forceinline__ void load8c(const float *rowp, int col0, int n, float *v) {
  if ((n & 3) == 0) {
    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
    if (col0 < n) lo = __ldg(reinterpret_cast<const float4 *>(rowp + col0));
    if (col0 + 4 < n) hi = __ldg(reinterpret_cast<const float4 *>(rowp + col0 + 4));
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
    v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (col0 + j < n) ? __ldg(rowp + col0 + j) : 0.f;
  }
}

// tile extent in grid coordinates, for the steepness test
__device__ __forceinline__ void tile_extent(const ScaleDesc &s, int tile_local, float &x_lo, float &x_hi, float &y_lo,
                                            float &y_hi) {
  const int rt = tile_local / s.chunks, ch = tile_local - rt * s.chunks;
  const int c_lo = ch * s.cta * 8, c_hi = min(s.n, c_lo + s.cta * 8) - 1;
  const int r_lo = rt * s.rows_per_tile, r_hi = min(s.n, r_lo + s.rows_per_tile) - 1;
  x_lo = grid_coord(c_lo, s.step); x_hi = grid_coord(c_hi, s.step);
  y_lo = grid_coord(r_lo, s.step); y_hi = grid_coord(r_hi, s.step);
}

}  // namespace

// ================================================================================== forward (mask)
// NP = row pairs per thread (rows per thread = 2 NP, spaced s.rpp apart)
template <int NP, int NW>
__global__ void __launch_bounds__(NW * 32) splat_fwd_mask_fast_kernel(const __grid_constant__ SplatArgs a) {
  extern __shared__ float4 sc[];   // [K][2]: c0, forward record
  const int b = blockIdx.x / a.T;
  const int tile = blockIdx.x - b * a.T;
  const ScaleDesc &s = a.sc[find_scale(a, tile)];
  const int K = a.K, n = s.n;
  const float h = s.step;
  const StripPos pos = strip_pos(s, tile - s.tile_begin);   // contiguous strips: colB = colA + 4
  const float x0 = pin(grid_coord(pos.colA, h));
  f32x2 y[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p)
    y[p] = pin2(pack2(grid_coord(pos.row0 + (2 * p) * s.rpp, h), grid_coord(pos.row0 + (2 * p + 1) * s.rpp, h)));
  float x_lo, x_hi, y_lo, y_hi;
  tile_extent(s, tile - s.tile_begin, x_lo, x_hi, y_lo, y_hi);

  {  // everything above is index arithmetic; global memory is first touched below
    float ya, yb;
    unpack2(y[NP - 1], ya, yb);
    pdl_enter_after(pos.colA, pos.row0, x0 + x_hi, ya + yb + y_lo);
  }
  stage_fast_consts<1>(a.params + (size_t)b * K * GG_PARAM_STRIDE, K, h, x_lo, x_hi, y_lo, y_hi, sc);
  __syncthreads();
  if (!pos.valid) return;

  f32x2 acc[NP][8];
#pragma unroll
  for (int p = 0; p < NP; ++p)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = pack2(0.f, 0.f);

  for (int k = 0; k < K; ++k) {
    const float4 c0 = sc[2 * k], c1 = sc[2 * k + 1];
    const bool steep = c1.w == 0.f;
    const float xm = x0 - c0.x;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const f32x2 dy = sub2(y[p], bc(c0.y));
      const f32x2 t0 = fma2(bc(c0.w), dy, bc(xm));
      const f32x2 e = mul2(mul2(bc(c1.x), dy), dy);
      f32x2 G[8];
      eval_strip(c0, c1, steep, t0, e, dy, pos.colA, h, G);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[p][j] = add2(acc[p][j], G[j]);
    }
  }
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    float fa[8], fb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) unpack2(acc[p][j], fa[j], fb[j]);
    const int rowA = pos.row0 + (2 * p) * s.rpp, rowB = rowA + s.rpp;
    if (rowA < n) store8c(s.mask + ((size_t)b * n + rowA) * n, pos.colA, n, fa);
    if (rowB < n) store8c(s.mask + ((size_t)b * n + rowB) * n, pos.colA, n, fb);
  }
}

// ================================================================================= backward (mask)
// FUSED = false: upstream gradient dL/dmask read from s.mask.
// FUSED = true : m_recon_loss_bis in one launch -- a forward pass builds the silhouette in registers, the residual
//                against the target gives dL/dmask = -sign(0.5 (t + 1) - m) / N (mask/mask_gen_dynamic.py:785-787)
//                and the loss partial; then the same CTA runs the backward.  K <= kKBlock.
// NW = warps per CTA; the register cap keeps 80 registers per thread at every CTA size
// CENTRE: guarded Gaussians accumulate their moments about their own centre line (launches with a scale below 64
//         pixels, where a guarded Gaussian can be far sub-pixel; costs 4 % of the kernel, so not at the large scales)
template <int NP, bool FUSED, int NW, bool CENTRE>
__global__ void __launch_bounds__(NW * 32, NP == 1 ? (GG_BWD_MINBLOCKS * 8) / NW : 1) splat_bwd_mask_fast_kernel(const __grid_constant__ SplatArgs a) {
  constexpr int REC = FUSED ? 3 : 2;   // c0, [forward record,] backward record
  __shared__ float4 sc[kKBlock * REC];
  __shared__ __align__(16) float stage[NW][kStageCols][kStagePad];
  __shared__ float warpsum[NW][kKBlock * GG_MOMENTS];
  float *lut = &stage[0][0][0];   // FUSED: byte -> target value table; dead before the transpose buffer is first written
  static_assert(NW * kStageCols * kStagePad >= 256, "lut");

  const int b = blockIdx.x / a.T;
  const int tile = blockIdx.x - b * a.T;
  const ScaleDesc &s = a.sc[find_scale(a, tile)];
  const int K = a.K, n = s.n;
  const float h = s.step;
  const int k0 = blockIdx.y * kKBlock;
  const int kn = min(kKBlock, K - k0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const StripPos pos = strip_pos(s, tile - s.tile_begin);
  const float x0 = pin(grid_coord(pos.colA, h));
  f32x2 y[NP];
  int row[2 * NP];
  bool rv[2 * NP];
#pragma unroll
  for (int r = 0; r < 2 * NP; ++r) {
    row[r] = pos.row0 + r * s.rpp;
    rv[r] = pos.valid && row[r] < n;
  }
#pragma unroll
  for (int p = 0; p < NP; ++p) y[p] = pin2(pack2(grid_coord(row[2 * p], h), grid_coord(row[2 * p + 1], h)));

  float x_lo, x_hi, y_lo, y_hi;
  tile_extent(s, tile - s.tile_begin, x_lo, x_hi, y_lo, y_hi);

  {  // everything above is index arithmetic; global memory is first touched below
    float ya, yb;
    unpack2(y[NP - 1], ya, yb);
    pdl_enter_after(pos.colA, row[2 * NP - 1], x0 + x_hi, ya + yb + y_lo);
  }
  float gm[2 * NP][8];     // upstream gradient dL/dmask of this thread's pixels
  float lsum = 0.f;
  if (!FUSED) {   // issued before the constants are staged: the two global round trips overlap
#pragma unroll
    for (int r = 0; r < 2 * NP; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) gm[r][j] = 0.f;
      if (rv[r]) load8c(s.mask + ((size_t)b * n + row[r]) * n, pos.colA, n, gm[r]);
    }
  }
  stage_fast_consts<FUSED ? 3 : 2>(a.params + ((size_t)b * K + k0) * GG_PARAM_STRIDE, kn, h, x_lo, x_hi, y_lo, y_hi, sc);
  if (FUSED && (a.target_kind == GG_TARGET_U8 || a.target_kind == GG_TARGET_STRIPS)) {
    // 0.5 * ((raw / 127.5 - 1) + 1), op by op in float32 as the reference (:672, :787): one division per
    // thread instead of one per pixel
    for (int i = threadIdx.x; i < 256; i += NW * 32)
      lut[i] = __fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn((float)i, 127.5f), 1.0f), 1.0f), 0.5f);
  }
  __syncthreads();

  if (FUSED) {
    // ---- forward pass: the silhouette of this thread's strips
    f32x2 acc[NP][8];
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[p][j] = pack2(0.f, 0.f);
    for (int kk = 0; kk < kn; ++kk) {
      const float4 c0 = sc[3 * kk], c1 = sc[3 * kk + 1];
      const bool steep = c1.w == 0.f;
      const float xm = x0 - c0.x;
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const f32x2 dy = sub2(y[p], bc(c0.y));
        const f32x2 t0 = fma2(bc(c0.w), dy, bc(xm));
        const f32x2 e = mul2(mul2(bc(c1.x), dy), dy);
        f32x2 G[8];
        eval_strip(c0, c1, steep, t0, e, dy, pos.colA, h, G);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[p][j] = add2(acc[p][j], G[j]);
      }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
      for (int j = 0; j < 8; ++j) unpack2(acc[p][j], gm[2 * p][j], gm[2 * p + 1][j]);
    // ---- residual -> loss partial and dL/dmask
#pragma unroll
    for (int r = 0; r < 2 * NP; ++r) {
      if (!rv[r]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) gm[r][j] = 0.f;
        continue;
      }
      const size_t off = ((size_t)b * n + row[r]) * n;
      if (s.mask) store8c(s.mask + off, pos.colA, n, gm[r]);        // optional silhouette output
      float tg[8];
      if (a.target_kind == GG_TARGET_U8) {
        const unsigned char *tp = reinterpret_cast<const unsigned char *>(a.target) + off;
        if ((n & 7) == 0) {
          const uint2 u = pos.colA < n ? __ldg(reinterpret_cast<const uint2 *>(tp + pos.colA)) : make_uint2(0u, 0u);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            tg[j] = lut[(u.x >> (8 * j)) & 0xffu];
            tg[j + 4] = lut[(u.y >> (8 * j)) & 0xffu];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) tg[j] = (pos.colA + j < n) ? lut[__ldg(tp + pos.colA + j)] : 0.f;
        }
      } else if (a.target_kind == GG_TARGET_STRIPS) {
        // lossless strip packing (include/gaussigan.h): two header bits per flat strip, 8 raw bytes otherwise
        const unsigned char *base = reinterpret_cast<const unsigned char *>(a.target);
        const unsigned *offs = reinterpret_cast<const unsigned *>(base + 16);
        const size_t payload = (16 + (size_t)a.B * n * 4 + 7) & ~(size_t)7;
        const int W = (((n + 7) >> 3) + 31) >> 5;
        const unsigned char *rec = base + payload + __ldg(offs + (size_t)b * n + row[r]);
        const unsigned *hdr = reinterpret_cast<const unsigned *>(rec);
        const int strip = pos.colA >> 3, word = strip >> 5, bit = strip & 31;
        const unsigned M = __ldg(hdr + word);
        if ((M >> bit) & 1u) {
          int idx = __popc(M & ((1u << bit) - 1u));
          for (int w = 0; w < word; ++w) idx += __popc(__ldg(hdr + w));
          const uint2 u = __ldg(reinterpret_cast<const uint2 *>(rec + 8 * W) + idx);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            tg[j] = lut[(u.x >> (8 * j)) & 0xffu];
            tg[j + 4] = lut[(u.y >> (8 * j)) & 0xffu];
          }
        } else {
          const float v = ((__ldg(hdr + W + word) >> bit) & 1u) ? lut[255] : lut[0];
#pragma unroll
          for (int j = 0; j < 8; ++j) tg[j] = v;
        }
      } else if (a.target_kind == GG_TARGET_BITS) {
        // one byte = this thread's 8 pixels (strips start at multiples of 8); bit set -> 0.5 * (1 + 1) = 1
        const unsigned char *tp = reinterpret_cast<const unsigned char *>(a.target) +
                                  ((size_t)b * n + row[r]) * (size_t)((n + 7) >> 3);
        const unsigned bits = pos.colA < n ? (unsigned)__ldg(tp + (pos.colA >> 3)) : 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) tg[j] = ((bits >> j) & 1u) ? 1.0f : 0.0f;
      } else {
        load8c(reinterpret_cast<const float *>(a.target) + off, pos.colA, n, tg);
#pragma unroll
        for (int j = 0; j < 8; ++j) tg[j] = __fmul_rn(__fadd_rn(tg[j], 1.0f), 0.5f);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float res = tg[j] - gm[r][j];
        const bool in = pos.colA + j < n;
        lsum += in ? fabsf(res) : 0.f;
        gm[r][j] = !in ? 0.f : (res > 0.f ? -a.inv_n : (res < 0.f ? a.inv_n : 0.f));
      }
    }
    // per-CTA loss partial (fixed order: lanes by butterfly, warps serially)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
    if (lane == 0) warpsum[warp][0] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) tot += warpsum[w][0];
      a.loss_partials[(size_t)b * a.T + tile] = tot;
    }
    __syncthreads();
  }
  // the upstream gradient of a pixel is shared by all Gaussians, and so are its products with the pixel's offset from
  // the strip origin on the reference's float32 grid (j h to ~1e-7): the Horner coefficients, packed once per row pair
  f32x2 W0[NP][8], W1[NP][8], W2[NP][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float xi1 = grid_coord(pos.colA + j, h) - x0;
    const float xi2 = xi1 * xi1;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      W0[p][j] = pack2(gm[2 * p][j], gm[2 * p + 1][j]);
      W1[p][j] = mul2(W0[p][j], bc(xi1));
      W2[p][j] = mul2(W0[p][j], bc(xi2));
    }
  }

#pragma unroll kBwdKUnroll
  for (int kk = 0; kk < kn; ++kk) {
    const float4 c0 = sc[REC * kk], c1 = sc[REC * kk + REC - 1];
    const bool steep = c1.w == 0.f;
    const float xm = x0 - c0.x;
    f32x2 P20 = pack2(0.f, 0.f), P11 = P20, P02 = P20, P10 = P20, P01 = P20;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const f32x2 dy = sub2(y[p], bc(c0.y));
      const f32x2 t0 = fma2(bc(c0.w), dy, bc(xm));
      const f32x2 e = mul2(mul2(bc(c1.x), dy), dy);
      f32x2 S0, m1, m2;
      if (!steep) {
        f32x2 S1, S2;
        horner_strip(c0, c1, t0, e, W0[p], W1[p], W2[p], S0, S1, S2);
        // strip-local moments (about x_0) -> moments of t = t0 + xi
        m1 = fma2(t0, S0, S1);
        m2 = fma2(t0, add2(S1, m1), S2);
      } else {
        f32x2 G[8];
        eval_strip_direct(c0, e, dy, pos.colA, h, G);
        if (!CENTRE) {
          S0 = mul2(G[0], W0[p][0]);
          f32x2 S1 = mul2(G[1], W1[p][1]), S2 = mul2(G[1], W2[p][1]);
#pragma unroll
          for (int j = 1; j < 8; ++j) {
            S0 = fma2(G[j], W0[p][j], S0);
            if (j > 1) {
              S1 = fma2(G[j], W1[p][j], S1);
              S2 = fma2(G[j], W2[p][j], S2);
            }
          }
          m1 = fma2(t0, S0, S1);
          m2 = fma2(t0, add2(S1, m1), S2);
        } else {
          // CENTRE kernels, guarded Gaussian (a needle much narrower than the strip): sums about its own centre line,
          // with the t of the direct evaluation -- re-centring from x_0 cancels (strip length / sigma)^2 of the second
          // moment, which only matters on coarse grids, where such Gaussians are sub-pixel
          const f32x2 sh = fma2(bc(-c0.w), dy, bc(c0.x));
          S0 = pack2(0.f, 0.f);
          m1 = S0;
          m2 = S0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const f32x2 u = mul2(G[j], W0[p][j]);
            const f32x2 tj = sub2(bc(grid_coord(pos.colA + j, h)), sh);
            const f32x2 v = mul2(u, tj);
            S0 = add2(S0, u);
            m1 = add2(m1, v);
            m2 = fma2(v, tj, m2);
          }
        }
      }
      // ... then the row weights
      const f32x2 d0 = mul2(dy, S0);
      if (p == 0) {
        P20 = m2; P10 = m1; P11 = mul2(dy, m1); P01 = d0; P02 = mul2(dy, d0);
      } else {
        P20 = add2(P20, m2); P10 = add2(P10, m1); P11 = fma2(dy, m1, P11); P01 = add2(P01, d0); P02 = fma2(dy, d0, P02);
      }
    }
    float lo, hi;
    const int cc = (kk % kKC) * GG_MOMENTS;
    unpack2(P20, lo, hi); stage[warp][cc + 0][lane] = lo + hi;
    unpack2(P11, lo, hi); stage[warp][cc + 1][lane] = lo + hi;
    unpack2(P02, lo, hi); stage[warp][cc + 2][lane] = lo + hi;
    unpack2(P10, lo, hi); stage[warp][cc + 3][lane] = lo + hi;
    unpack2(P01, lo, hi); stage[warp][cc + 4][lane] = lo + hi;
    if ((kk % kKC) == kKC - 1 || kk == kn - 1) {
      __syncwarp();
      const int ncols = cc + GG_MOMENTS;
      const int cbase = (kk / kKC) * kStageCols;
      if (lane < ncols) {   // 8 LDS.128, four independent chains, fixed order
        const float4 *rowp = reinterpret_cast<const float4 *>(&stage[warp][lane][0]);
        float4 acc = rowp[0];
#pragma unroll
        for (int l = 1; l < 8; ++l) {
          const float4 v = rowp[l];
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        warpsum[warp][cbase + lane] = (acc.x + acc.y) + (acc.z + acc.w);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  float *out = a.partials + (((size_t)b * a.slots + a.tile_offset + tile) * K + k0) * GG_MOMENTS;
  for (int c = threadIdx.x; c < kn * GG_MOMENTS; c += NW * 32) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) acc += warpsum[w][c];
    out[c] = acc;
  }
}

// ====================================================================================== launchers
// rpt = rows per thread of the plan (2 or 4); the CTA size comes from the plan's geometry (sc[0].cta)
// K * 32 bytes of staged constants exceed the 48 KB default above K = 1536 (the C ABI's limit): kept for safety (the C ABI accepts K <= 1536): opt in
template <typename Kern>
static cudaError_t launch_fwd_fast_one(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const SplatArgs &a) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  launch_pdl(kern, grid, block, smem, st, a);
  return cudaGetLastError();
}

cudaError_t launch_fwd_mask_fast(const SplatArgs &a, int B, int rpt, cudaStream_t st) {
  dim3 grid((unsigned)((size_t)B * a.T));
  size_t smem = (size_t)a.K * 2 * sizeof(float4);
  const int cta = a.sc[0].cta;
  if (rpt == 4) return launch_fwd_fast_one(splat_fwd_mask_fast_kernel<2, 8>, grid, dim3(256), smem, st, a);
  if (cta == 128) return launch_fwd_fast_one(splat_fwd_mask_fast_kernel<1, 4>, grid, dim3(128), smem, st, a);
  if (cta == 160) return launch_fwd_fast_one(splat_fwd_mask_fast_kernel<1, 5>, grid, dim3(160), smem, st, a);
  if (cta == 192) return launch_fwd_fast_one(splat_fwd_mask_fast_kernel<1, 6>, grid, dim3(192), smem, st, a);
  return launch_fwd_fast_one(splat_fwd_mask_fast_kernel<1, 8>, grid, dim3(256), smem, st, a);
}

template <bool FUSED, bool CENTRE>
static void launch_bwd_fast_t(const SplatArgs &a, dim3 grid, int rpt, cudaStream_t st) {
  const int cta = a.sc[0].cta;
  if (rpt == 4) launch_pdl(splat_bwd_mask_fast_kernel<2, FUSED, 8, CENTRE>, grid, dim3(256), 0, st, a);
  else if (cta == 128) launch_pdl(splat_bwd_mask_fast_kernel<1, FUSED, 4, CENTRE>, grid, dim3(128), 0, st, a);
  else if (cta == 160) launch_pdl(splat_bwd_mask_fast_kernel<1, FUSED, 5, CENTRE>, grid, dim3(160), 0, st, a);
  else if (cta == 192) launch_pdl(splat_bwd_mask_fast_kernel<1, FUSED, 6, CENTRE>, grid, dim3(192), 0, st, a);
  else launch_pdl(splat_bwd_mask_fast_kernel<1, FUSED, 8, CENTRE>, grid, dim3(256), 0, st, a);
}

cudaError_t launch_bwd_mask_fast(const SplatArgs &a, bool fused, int B, int rpt, cudaStream_t st) {
  dim3 grid((unsigned)((size_t)B * a.T), (unsigned)((a.K + kKBlock - 1) / kKBlock));
  bool coarse = false;      // a scale below 64 pixels: guarded Gaussians can be far sub-pixel there
  for (int i = 0; i < a.nscales; ++i) coarse = coarse || a.sc[i].n < 64;
  if (fused) { if (coarse) launch_bwd_fast_t<true, true>(a, grid, rpt, st); else launch_bwd_fast_t<true, false>(a, grid, rpt, st); }
  else { if (coarse) launch_bwd_fast_t<false, true>(a, grid, rpt, st); else launch_bwd_fast_t<false, false>(a, grid, rpt, st); }
  return cudaGetLastError();
}

}  // namespace gg

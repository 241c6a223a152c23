// gg_splat.cu -- the per-pixel kernels: evaluate K projected Gaussians at every pixel (forward) and
// reduce per-pixel upstream gradients into five moments per Gaussian (backward).
//
// Replaces get_gaussian_maps_2d's raster, the float32 cast, the 4-scale loop and the sum composite
// (mask/mask_gen_dynamic.py:121-141, 315-319, 785-786) and their tf.gradients.
//
// Arithmetic.  With the conic completed along x (see gg_common.cuh),
//     g = exp2(nA * t^2 + nE * dy^2),   dy = y - mu_y,   t = x - (mu_x + nr * dy)
// so a thread that owns a strip of 8 pixels of one row pays ~4 instructions per (row, Gaussian) for
// (dy, shift, nE*dy^2) and then FADD + FMUL + FFMA + MUFU.EX2 (+ FADD into the mask) per pixel:
// 5 issue slots per (pixel, Gaussian) instead of the 9 of the textbook form.  The backward
// accumulates sum(w), sum(w t), sum(w t^2) along the strip (w = g * Gbar) and folds them into the
// five moments {t^2, t dy, dy^2, t, dy} per row; the projection adjoint converts the sheared moments
// back (gg_project.cu: conic_adjoint).  No tensor cores: nothing here is a contraction.
//
// Two thread mappings:
//   strip kernels  -- thread = 8 consecutive pixels of a row, loop over K.  Used for mask-only
//                     output/gradient and for NCHW maps (128-bit stores along W); also the generic
//                     NHWC fallback for K that is not 4 * 2^m.
//   quad kernels   -- NHWC maps: thread = (pixel, 4 consecutive Gaussians), one 128-bit access per
//                     pixel-quad, consecutive lanes -> consecutive 16-byte chunks (fully coalesced).
//
// Backward reduction: per-thread registers -> per-warp transpose through padded shared memory
// (5 STS per Gaussian, conflict-free) -> 32-term serial sums by 30 lanes -> fixed-order sum over the
// 8 warps -> one [K,5] partial per CTA written to its own slot.  No atomics, no zero-fill,
// bit-reproducible; the O(T) sum over slots happens in float64 in the projection adjoint.
#include <atomic>
#include <cstdlib>
#include <type_traits>
#include "gg_common.cuh"
#include "gg_strip.cuh"
#include "gg_launch.h"

namespace gg {


// =============================================================================== forward, strips
// OUT: 0 = mask only, 1 = NCHW maps (+ mask when given), 2 = NHWC maps, scalar stores (+ mask)
template <int RPT, int OUT>
__global__ void __launch_bounds__(kThreads) splat_fwd_strip_kernel(const __grid_constant__ SplatArgs a) {
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  extern __shared__ float4 sp[];   // [K][2]
  const int b = blockIdx.x / a.T;  // a.T = tiles per batch element in this launch
  const int tile = blockIdx.x - b * a.T;
  const ScaleDesc &s = a.sc[find_scale(a, tile)];
  const int K = a.K, n = s.n;

  const float4 *gp = reinterpret_cast<const float4 *>(a.params + (size_t)b * K * GG_PARAM_STRIDE);
  for (int i = threadIdx.x; i < 2 * K; i += kThreads) sp[i] = __ldg(gp + i);
  __syncthreads();

  const StripPos pos = strip_pos(s, tile - s.tile_begin);
  if (!pos.valid) return;

  float x[kStrip];
#pragma unroll
  for (int j = 0; j < kStrip; ++j) x[j] = grid_coord(pos.col(j), s.step);
  float y[RPT];
  bool rv[RPT];
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int row = pos.row0 + r * s.rpp;
    rv[r] = row < n;
    y[r] = grid_coord(row, s.step);
  }
  float acc[RPT][kStrip];
#pragma unroll
  for (int r = 0; r < RPT; ++r)
#pragma unroll
    for (int j = 0; j < kStrip; ++j) acc[r][j] = 0.f;

  for (int k = 0; k < K; ++k) {
    const float4 p = sp[2 * k];          // mu_x, mu_y, nA, nr
    const float nE = sp[2 * k + 1].x;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const float dy = y[r] - p.y;
      const float sh = fmaf(p.w, dy, p.x);
      const float e = (nE * dy) * dy;
      float g[kStrip];
#if GG_USE_F32X2
      {
        const f32x2 sh2 = pack2(sh, sh), nA2 = pack2(p.z, p.z), e2 = pack2(e, e);
#pragma unroll
        for (int j = 0; j < kStrip; j += 2) {
          const f32x2 t2 = sub2(pack2(x[j], x[j + 1]), sh2);
          float a0, a1;
          unpack2(fma2(nA2, mul2(t2, t2), e2), a0, a1);
          g[j] = ex2_approx(a0);
          g[j + 1] = ex2_approx(a1);
          unpack2(add2(pack2(acc[r][j], acc[r][j + 1]), pack2(g[j], g[j + 1])), acc[r][j], acc[r][j + 1]);
        }
      }
#else
#pragma unroll
      for (int j = 0; j < kStrip; ++j) {
        const float t = x[j] - sh;
        g[j] = ex2_approx(fmaf(p.z, t * t, e));
        acc[r][j] += g[j];
      }
#endif
      if (OUT == 1) {
        if (rv[r]) {
          int row = pos.row0 + r * s.rpp;
          store8(s.maps + (((size_t)b * K + k) * n + row) * n, pos, n, g);
        }
      } else if (OUT == 2) {
        if (rv[r]) {
          int row = pos.row0 + r * s.rpp;
          float *q = s.maps + ((size_t)b * n + row) * n * s.cstride + k;
#pragma unroll
          for (int j = 0; j < kStrip; ++j)
            if (pos.col(j) < n) q[(size_t)pos.col(j) * s.cstride] = g[j];
        }
      }
    }
  }
  if (s.mask) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      if (!rv[r]) continue;
      int row = pos.row0 + r * s.rpp;
      store8(s.mask + ((size_t)b * n + row) * n, pos, n, acc[r]);
    }
  }
}

// ============================================================================== backward, strips
// GIN: 0 = gradient w.r.t. mask only, 1 = NCHW gmaps (+ gmask), 2 = NHWC gmaps, scalar loads (+ gmask),
//      3 = FUSED silhouette loss: a forward pass over the K Gaussians builds the silhouette in registers,
//          the residual against the target gives dL/dmask = -sign(0.5*(t+1) - m)/N on the spot
//          (mask/mask_gen_dynamic.py:785-787), and the same CTA runs the backward -- the mask and its gradient
//          never touch HBM (optionally the mask is written out through s.mask).  Requires K <= kKBlock.
template <int RPT, int GIN>
__global__ void __launch_bounds__(kThreads) splat_bwd_strip_kernel(const __grid_constant__ SplatArgs a) {
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  __shared__ float4 sp[kKBlock * 2];
  __shared__ float stage[kWarps][kStageCols][33];
  __shared__ float warpsum[kWarps][kKBlock * GG_MOMENTS];

  const int b = blockIdx.x / a.T;   // a.T here = tiles per batch element in THIS launch
  const int tile = blockIdx.x - b * a.T;
  const ScaleDesc &s = a.sc[find_scale(a, tile)];
  const int K = a.K, n = s.n;
  const int k0 = blockIdx.y * kKBlock;
  const int kn = min(kKBlock, K - k0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const float4 *gp = reinterpret_cast<const float4 *>(a.params + ((size_t)b * K + k0) * GG_PARAM_STRIDE);
  for (int i = threadIdx.x; i < 2 * kn; i += kThreads) sp[i] = __ldg(gp + i);
  __syncthreads();

  const StripPos pos = strip_pos(s, tile - s.tile_begin);
  const float *gmask = s.mask;
  const float *gmaps = s.maps;

  float x[kStrip];
#pragma unroll
  for (int j = 0; j < kStrip; ++j) x[j] = grid_coord(pos.col(j), s.step);
  float y[RPT];
  bool rv[RPT];
  float gm[RPT][kStrip];
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    int row = pos.row0 + r * s.rpp;
    rv[r] = pos.valid && row < n;
    y[r] = grid_coord(row, s.step);
#pragma unroll
    for (int j = 0; j < kStrip; ++j) gm[r][j] = 0.f;
    if (GIN != 3 && rv[r] && gmask) load8(gmask + ((size_t)b * n + row) * n, pos, n, gm[r]);
  }

  if (GIN == 3) {
    // ---- forward pass: gm <- silhouette
    for (int kk = 0; kk < kn; ++kk) {
      const float4 p = sp[2 * kk];
      const float nE = sp[2 * kk + 1].x;
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const float dy = y[r] - p.y;
        const float sh = fmaf(p.w, dy, p.x);
        const float e = (nE * dy) * dy;
        const f32x2 sh2 = pack2(sh, sh), nA2 = pack2(p.z, p.z), e2 = pack2(e, e);
#pragma unroll
        for (int j = 0; j < kStrip; j += 2) {
          const f32x2 t2 = sub2(pack2(x[j], x[j + 1]), sh2);
          float a0, a1;
          unpack2(fma2(nA2, mul2(t2, t2), e2), a0, a1);
          unpack2(add2(pack2(gm[r][j], gm[r][j + 1]), pack2(ex2_approx(a0), ex2_approx(a1))), gm[r][j], gm[r][j + 1]);
        }
      }
    }
    // ---- residual -> loss partial and dL/dmask, in place
    float lsum = 0.f;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      if (!rv[r]) continue;
      const int row = pos.row0 + r * s.rpp;
      const size_t off = ((size_t)b * n + row) * n;
      if (gmask) store8(s.mask + off, pos, n, gm[r]);              // optional silhouette output
      float tg[kStrip];
      if (a.target_kind == GG_TARGET_U8) {
        const unsigned char *tp = reinterpret_cast<const unsigned char *>(a.target) + off;
        if ((n & 3) == 0) {
          const unsigned ua = pos.colA < n ? __ldg(reinterpret_cast<const unsigned *>(tp + pos.colA)) : 0u;
          const unsigned ub = pos.colB < n ? __ldg(reinterpret_cast<const unsigned *>(tp + pos.colB)) : 0u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            tg[j] = (float)((ua >> (8 * j)) & 0xffu);
            tg[j + 4] = (float)((ub >> (8 * j)) & 0xffu);
          }
        } else {
#pragma unroll
          for (int j = 0; j < kStrip; ++j) tg[j] = (pos.col(j) < n) ? (float)__ldg(tp + pos.col(j)) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < kStrip; ++j)   // (raw/127.5 - 1 + 1) * 0.5, op by op as the reference (:672, :787)
          tg[j] = __fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn(tg[j], 127.5f), 1.0f), 1.0f), 0.5f);
      } else {
        load8(reinterpret_cast<const float *>(a.target) + off, pos, n, tg);
#pragma unroll
        for (int j = 0; j < kStrip; ++j) tg[j] = __fmul_rn(__fadd_rn(tg[j], 1.0f), 0.5f);
      }
#pragma unroll
      for (int j = 0; j < kStrip; ++j) {
        const float res = tg[j] - gm[r][j];
        const bool in = pos.col(j) < n;
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
      for (int w = 0; w < kWarps; ++w) tot += warpsum[w][0];
      a.loss_partials[(size_t)b * a.T + tile] = tot;
    }
    __syncthreads();
  }

  // NCHW upstream gradients are software-pipelined one Gaussian ahead: the 128-bit loads of k+1 are in
  // flight while k is being evaluated (the loop is otherwise a load -> use chain per Gaussian).
  float Gn[GIN == 1 ? RPT : 1][kStrip];
  if (GIN == 1) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
#pragma unroll
      for (int j = 0; j < kStrip; ++j) Gn[r][j] = 0.f;
      if (rv[r]) load8(gmaps + (((size_t)b * K + k0) * n + pos.row0 + r * s.rpp) * n, pos, n, Gn[r]);
    }
  }

  for (int kk = 0; kk < kn; ++kk) {
    const float4 p = sp[2 * kk];
    const float nE = sp[2 * kk + 1].x;
    float T20 = 0.f, T11 = 0.f, T02 = 0.f, T10 = 0.f, T01 = 0.f;
    float Gc[GIN == 1 ? RPT : 1][kStrip];
    if (GIN == 1) {
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
#pragma unroll
        for (int j = 0; j < kStrip; ++j) Gc[r][j] = Gn[r][j] + gm[r][j];
        if (rv[r] && kk + 1 < kn)
          load8(gmaps + (((size_t)b * K + k0 + kk + 1) * n + pos.row0 + r * s.rpp) * n, pos, n, Gn[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      if (!rv[r]) continue;
      float G[kStrip];
      if (GIN == 0 || GIN == 3) {
#pragma unroll
        for (int j = 0; j < kStrip; ++j) G[j] = gm[r][j];
      } else if (GIN == 1) {
#pragma unroll
        for (int j = 0; j < kStrip; ++j) G[j] = Gc[r][j];
      } else {
        int row = pos.row0 + r * s.rpp;
        const float *q = gmaps + ((size_t)b * n + row) * n * s.cstride + k0 + kk;
#pragma unroll
        for (int j = 0; j < kStrip; ++j)
          G[j] = ((pos.col(j) < n) ? __ldg(q + (size_t)pos.col(j) * s.cstride) : 0.f) + gm[r][j];
      }
      const float dy = y[r] - p.y;
      const float sh = fmaf(p.w, dy, p.x);
      const float e = (nE * dy) * dy;
      float S0, S1, S2;
#if GG_USE_F32X2
      {
        const f32x2 sh2 = pack2(sh, sh), nA2 = pack2(p.z, p.z), e2 = pack2(e, e);
        f32x2 s0 = pack2(0.f, 0.f), s1 = s0, s2 = s0;
#pragma unroll
        for (int j = 0; j < kStrip; j += 2) {
          const f32x2 t2 = sub2(pack2(x[j], x[j + 1]), sh2);
          const f32x2 u2 = mul2(t2, t2);
          float a0, a1;
          unpack2(fma2(nA2, u2, e2), a0, a1);
          const f32x2 w2 = mul2(pack2(ex2_approx(a0), ex2_approx(a1)), pack2(G[j], G[j + 1]));
          s0 = add2(s0, w2);
          s1 = fma2(w2, t2, s1);
          s2 = fma2(w2, u2, s2);
        }
        float lo, hi;
        unpack2(s0, lo, hi); S0 = lo + hi;
        unpack2(s1, lo, hi); S1 = lo + hi;
        unpack2(s2, lo, hi); S2 = lo + hi;
      }
#else
      S0 = S1 = S2 = 0.f;
#pragma unroll
      for (int j = 0; j < kStrip; ++j) {
        const float t = x[j] - sh;
        const float u = t * t;
        const float w = ex2_approx(fmaf(p.z, u, e)) * G[j];
        S0 += w;
        S1 = fmaf(w, t, S1);
        S2 = fmaf(w, u, S2);
      }
#endif
      T20 += S2;
      T10 += S1;
      T11 = fmaf(dy, S1, T11);
      const float d0 = dy * S0;
      T01 += d0;
      T02 = fmaf(dy, d0, T02);
    }
    const int c0 = (kk % kKC) * GG_MOMENTS;
    stage[warp][c0 + 0][lane] = T20;
    stage[warp][c0 + 1][lane] = T11;
    stage[warp][c0 + 2][lane] = T02;
    stage[warp][c0 + 3][lane] = T10;
    stage[warp][c0 + 4][lane] = T01;
    if ((kk % kKC) == kKC - 1 || kk == kn - 1) {
      __syncwarp();
      const int ncols = c0 + GG_MOMENTS;
      const int cbase = (kk / kKC) * kStageCols;
      if (lane < ncols) {
        float acc = 0.f;
#pragma unroll
        for (int l = 0; l < 32; ++l) acc += stage[warp][lane][l];
        warpsum[warp][cbase + lane] = acc;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  float *out = a.partials + (((size_t)b * a.slots + a.tile_offset + tile) * K + k0) * GG_MOMENTS;
  for (int c = threadIdx.x; c < kn * GG_MOMENTS; c += kThreads) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) acc += warpsum[w][c];
    out[c] = acc;
  }
}

// ================================================================================ forward, quads
// NHWC maps, K = 4 * KQ with KQ a power of two <= 32.  Thread = (pixel, 4 Gaussians).
template <bool WITH_MASK>
__global__ void __launch_bounds__(kThreads) splat_fwd_quad_kernel(const __grid_constant__ SplatArgs a,
                                                                  int kq_shift) {
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  const int b = blockIdx.x / a.T;
  const int tile = blockIdx.x - b * a.T;
  const ScaleDesc &s = a.sc[find_scale(a, tile)];
  const int K = a.K, n = s.n;
  const int KQ = 1 << kq_shift;
  const int kq = threadIdx.x & (KQ - 1);

  float mux[4], muy[4], nA[4], nr[4], nE[4];
  {
    const float4 *gp = reinterpret_cast<const float4 *>(a.params + ((size_t)b * K + 4 * kq) * GG_PARAM_STRIDE);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 lo = __ldg(gp + 2 * i), hi = __ldg(gp + 2 * i + 1);
      mux[i] = lo.x; muy[i] = lo.y; nA[i] = lo.z; nr[i] = lo.w; nE[i] = hi.x;
    }
  }
  const int row_begin = (tile - s.tile_begin) * s.rows_per_tile;
  const int row_end = min(n, row_begin + s.rows_per_tile);
  const int quads_per_row = n << kq_shift;

  for (int row = row_begin; row < row_end; ++row) {
    const float y = grid_coord(row, s.step);
    float sh[4], e[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float dy = y - muy[i];
      sh[i] = fmaf(nr[i], dy, mux[i]);
      e[i] = (nE[i] * dy) * dy;
    }
    float *mrow = s.maps + ((size_t)b * n + row) * n * s.cstride + 4 * kq;
    for (int q0 = 0; q0 < quads_per_row; q0 += kThreads) {
      const int qq = q0 + threadIdx.x;
      const bool active = qq < quads_per_row;
      const int col = min(qq >> kq_shift, n - 1);
      const float x = grid_coord(col, s.step);
      float g[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float t = x - sh[i];
        g[i] = ex2_approx(fmaf(nA[i] * t, t, e[i]));
      }
      if (active) st_f4(mrow + (size_t)col * s.cstride, g[0], g[1], g[2], g[3]);
      if (WITH_MASK) {
        float sum = (g[0] + g[1]) + (g[2] + g[3]);
        for (int off = 1; off < KQ; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (active && kq == 0 && s.mask) s.mask[((size_t)b * n + row) * n + col] = sum;
      }
    }
  }
}

// ====================================================================== forward, quads, any K % 4 == 0
// The quad kernel without the power-of-two restriction (the reference's K = 12): the first (256 / KQ) KQ threads of
// the CTA are used, so that a thread keeps its quad of Gaussians (constants in registers) from pass to pass and
// consecutive threads still write consecutive 16-byte chunks.  No fused silhouette: the lanes of a pixel do not form
// an aligned group of a warp; the launcher gives the silhouette to the mask kernel.
// NC: columns a thread keeps in registers (its whole share of a row: n <= NC * pixels per pass); 0: column loop
template <int NC>
__global__ void __launch_bounds__(kThreads) splat_fwd_quadg_kernel(const __grid_constant__ SplatArgs a, int KQ) {
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  const int b = blockIdx.x / a.T;
  const int tile = blockIdx.x - b * a.T;
  const ScaleDesc &s = a.sc[find_scale(a, tile)];
  const int K = a.K, n = s.n;
  const int cpp = kThreads / KQ;                    // pixels per pass
  const int c0 = threadIdx.x / KQ, kq = threadIdx.x - c0 * KQ;
  if (c0 >= cpp) return;

  float mux[4], muy[4], nA[4], nr[4], nE[4];
  {
    const float4 *gp = reinterpret_cast<const float4 *>(a.params + ((size_t)b * K + 4 * kq) * GG_PARAM_STRIDE);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 lo = __ldg(gp + 2 * i), hi = __ldg(gp + 2 * i + 1);
      mux[i] = lo.x; muy[i] = lo.y; nA[i] = lo.z; nr[i] = lo.w; nE[i] = hi.x;
    }
  }
  const int row_begin = (tile - s.tile_begin) * s.rows_per_tile;
  const int row_end = min(n, row_begin + s.rows_per_tile);
  if (NC > 0) {
    constexpr int kColsPerThread = NC > 0 ? NC : 1;
    // A thread owns at most kColsPerThread columns (c0, c0 + cpp, ...: 4 at n = 256, K = 12): their coordinates and
    // store pointers are formed once, so a row costs its 8 setup operations plus four operations, one exponential
    // and a quarter of a 128-bit store per value (the column loop below re-derives coordinate and address per
    // element and ran at 70 % of the issue slots).  Same arithmetic, bit-identical values.
    float xs[kColsPerThread];
    float *ptr[kColsPerThread];
    bool on[kColsPerThread];
#pragma unroll
    for (int j = 0; j < kColsPerThread; ++j) {
      const int col = c0 + j * cpp;
      on[j] = col < n;
      xs[j] = grid_coord(on[j] ? col : c0, s.step);
      ptr[j] = s.maps + ((size_t)b * n + row_begin) * n * s.cstride + 4 * kq + (size_t)(on[j] ? col : c0) * s.cstride;
    }
    const size_t row_stride = (size_t)n * s.cstride;
    // one copy of the row loop per number of columns the WARP uses (fewer on the coarse scales of a pyramid launch and
    // in the warps behind the last column of a row): no branches inside, stores predicated per lane
    auto rows = [&](auto mc) {
      constexpr int M = decltype(mc)::value;
      for (int row = row_begin; row < row_end; ++row) {
        const float y = grid_coord(row, s.step);
        float sh[4], e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float dy = y - muy[i];
          sh[i] = fmaf(nr[i], dy, mux[i]);
          e[i] = (nE[i] * dy) * dy;
        }
#pragma unroll
        for (int j = 0; j < M; ++j) {
          float g[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float t = xs[j] - sh[i];
            g[i] = ex2_approx(fmaf(nA[i] * t, t, e[i]));
          }
          if (on[j]) st_f4(ptr[j], g[0], g[1], g[2], g[3]);
          ptr[j] += row_stride;
        }
      }
    };
    int used = 0;
#pragma unroll
    for (int j = 0; j < kColsPerThread; ++j) used += on[j] ? 1 : 0;
    used = __reduce_max_sync(__activemask(), used);   // warp-uniform: the lanes of a warp run the same copy of the loop
                                                      // (the lanes past the last unit of a pass have left already)
    switch (used) {
      case 1: rows(std::integral_constant<int, 1>{}); break;
      case 2: rows(std::integral_constant<int, 2>{}); break;
      case 3: rows(std::integral_constant<int, 3>{}); break;
      case 4: rows(std::integral_constant<int, 4>{}); break;
      case 5: rows(std::integral_constant<int, (kColsPerThread >= 5 ? 5 : 1)>{}); break;
      case 6: rows(std::integral_constant<int, (kColsPerThread >= 6 ? 6 : 1)>{}); break;
      case 7: rows(std::integral_constant<int, (kColsPerThread >= 7 ? 7 : 1)>{}); break;
      case 8: rows(std::integral_constant<int, (kColsPerThread >= 8 ? 8 : 1)>{}); break;
      default: break;
    }
    return;
  }
  for (int row = row_begin; row < row_end; ++row) {
    const float y = grid_coord(row, s.step);
    float sh[4], e[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float dy = y - muy[i];
      sh[i] = fmaf(nr[i], dy, mux[i]);
      e[i] = (nE[i] * dy) * dy;
    }
    float *mrow = s.maps + ((size_t)b * n + row) * n * s.cstride + 4 * kq;
    for (int col = c0; col < n; col += cpp) {
      const float x = grid_coord(col, s.step);
      float g[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float t = x - sh[i];
        g[i] = ex2_approx(fmaf(nA[i] * t, t, e[i]));
      }
      st_f4(mrow + (size_t)col * s.cstride, g[0], g[1], g[2], g[3]);
    }
  }
}

// ========================================================================== forward, pairs, even K
// The same for K = 2 KP that is not a multiple of 4 (the reference's K = 6): thread = (pixel, 2 Gaussians), one
// 64-bit store, consecutive threads -> consecutive 8-byte chunks.
template <int NC>
__global__ void __launch_bounds__(kThreads) splat_fwd_pair_kernel(const __grid_constant__ SplatArgs a, int KP) {
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  const int b = blockIdx.x / a.T;
  const int tile = blockIdx.x - b * a.T;
  const ScaleDesc &s = a.sc[find_scale(a, tile)];
  const int K = a.K, n = s.n;
  const int cpp = kThreads / KP;                    // pixels per pass
  const int c0 = threadIdx.x / KP, kp = threadIdx.x - c0 * KP;
  if (c0 >= cpp) return;

  float mux[2], muy[2], nA[2], nr[2], nE[2];
  {
    const float4 *gp = reinterpret_cast<const float4 *>(a.params + ((size_t)b * K + 2 * kp) * GG_PARAM_STRIDE);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 lo = __ldg(gp + 2 * i), hi = __ldg(gp + 2 * i + 1);
      mux[i] = lo.x; muy[i] = lo.y; nA[i] = lo.z; nr[i] = lo.w; nE[i] = hi.x;
    }
  }
  const int row_begin = (tile - s.tile_begin) * s.rows_per_tile;
  const int row_end = min(n, row_begin + s.rows_per_tile);
  if (NC > 0) {      // columns of a thread hoisted out of the row loop: see splat_fwd_quadg_kernel
    constexpr int kColsPerThread = NC > 0 ? NC : 1;
    float xs[kColsPerThread];
    float *ptr[kColsPerThread];
    bool on[kColsPerThread];
#pragma unroll
    for (int j = 0; j < kColsPerThread; ++j) {
      const int col = c0 + j * cpp;
      on[j] = col < n;
      xs[j] = grid_coord(on[j] ? col : c0, s.step);
      ptr[j] = s.maps + ((size_t)b * n + row_begin) * n * s.cstride + 2 * kp + (size_t)(on[j] ? col : c0) * s.cstride;
    }
    const size_t row_stride = (size_t)n * s.cstride;
    auto rows = [&](auto mc) {
      constexpr int M = decltype(mc)::value;
      for (int row = row_begin; row < row_end; ++row) {
        const float y = grid_coord(row, s.step);
        float sh[2], e[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float dy = y - muy[i];
          sh[i] = fmaf(nr[i], dy, mux[i]);
          e[i] = (nE[i] * dy) * dy;
        }
#pragma unroll
        for (int j = 0; j < M; ++j) {
          float g[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float t = xs[j] - sh[i];
            g[i] = ex2_approx(fmaf(nA[i] * t, t, e[i]));
          }
          if (on[j]) *reinterpret_cast<float2 *>(ptr[j]) = make_float2(g[0], g[1]);
          ptr[j] += row_stride;
        }
      }
    };
    int used = 0;
#pragma unroll
    for (int j = 0; j < kColsPerThread; ++j) used += on[j] ? 1 : 0;
    used = __reduce_max_sync(__activemask(), used);   // warp-uniform: the lanes of a warp run the same copy of the loop
                                                      // (the lanes past the last unit of a pass have left already)
    switch (used) {
      case 1: rows(std::integral_constant<int, 1>{}); break;
      case 2: rows(std::integral_constant<int, 2>{}); break;
      case 3: rows(std::integral_constant<int, 3>{}); break;
      case 4: rows(std::integral_constant<int, 4>{}); break;
      case 5: rows(std::integral_constant<int, (kColsPerThread >= 5 ? 5 : 1)>{}); break;
      case 6: rows(std::integral_constant<int, (kColsPerThread >= 6 ? 6 : 1)>{}); break;
      case 7: rows(std::integral_constant<int, (kColsPerThread >= 7 ? 7 : 1)>{}); break;
      case 8: rows(std::integral_constant<int, (kColsPerThread >= 8 ? 8 : 1)>{}); break;
      default: break;
    }
    return;
  }
  for (int row = row_begin; row < row_end; ++row) {
    const float y = grid_coord(row, s.step);
    float sh[2], e[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float dy = y - muy[i];
      sh[i] = fmaf(nr[i], dy, mux[i]);
      e[i] = (nE[i] * dy) * dy;
    }
    float *mrow = s.maps + ((size_t)b * n + row) * n * s.cstride + 2 * kp;
    for (int col = c0; col < n; col += cpp) {
      const float x = grid_coord(col, s.step);
      float g[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float t = x - sh[i];
        g[i] = ex2_approx(fmaf(nA[i] * t, t, e[i]));
      }
      *reinterpret_cast<float2 *>(mrow + (size_t)col * s.cstride) = make_float2(g[0], g[1]);
    }
  }
}

// ================================================================================= forward, flat
// NHWC maps for ANY K with n K % 4 == 0 (the reference's own K = 6 and 12 among them).  A row of the output is n K
// consecutive floats; thread = 4 of them = one 128-bit store, whichever (pixel, Gaussian) pairs they belong to.
// The row-dependent part of every Gaussian (sh = mu_x + nr dy, e = nE dy^2) is staged once per tile in shared
// memory, so an element costs one LDS.128, four FP32 instructions and one exponential, as in the quad kernel --
// same arithmetic, bit-identical values.
__global__ void __launch_bounds__(kThreads) splat_fwd_flat_kernel(const __grid_constant__ SplatArgs a, float inv_k) {
  extern __shared__ float4 rowc[];   // [rows of the tile][K]: sh, e, nA, -
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  const int b = blockIdx.x / a.T;
  const int tile = blockIdx.x - b * a.T;
  const ScaleDesc &s = a.sc[find_scale(a, tile)];
  const int K = a.K, n = s.n;
  const float h = s.step;
  const int row_begin = (tile - s.tile_begin) * s.rows_per_tile;
  const int rows = min(n, row_begin + s.rows_per_tile) - row_begin;
  const float4 *gp = reinterpret_cast<const float4 *>(a.params + (size_t)b * K * GG_PARAM_STRIDE);
  for (int idx = threadIdx.x; idx < rows * K; idx += kThreads) {
    const int r = idx / K, k = idx - r * K;
    const float4 lo = __ldg(gp + 2 * k);
    const float nE = __ldg(reinterpret_cast<const float *>(gp + 2 * k + 1));
    const float dy = grid_coord(row_begin + r, h) - lo.y;
    rowc[idx] = make_float4(fmaf(lo.w, dy, lo.x), (nE * dy) * dy, lo.z, 0.f);
  }
  __syncthreads();
  const int vpr = (n * K) >> 2;                       // 128-bit vectors per row
  const float inv_vpr = 1.0f / (float)vpr;
  float *mtile = s.maps + ((size_t)b * n + row_begin) * n * K;
  // Index arithmetic is kept off the conversion/transcendental pipe (the exponentials need it): (row, vector) and
  // (pixel, Gaussian) advance incrementally, the pixel index is carried as a float as well (exact), and the float
  // reciprocal divisions (exact for operands < 2^22, which the launcher checks, up to a correction step) only run
  // at the start and when a thread moves to another row.
  const int dp = (4 * kThreads) / K, dk = 4 * kThreads - dp * K;     // pixel / Gaussian step of one pass
  const float dpf = (float)dp;
  int r = __float2int_rd(((float)threadIdx.x + 0.5f) * inv_vpr);
  int q = (int)threadIdx.x - r * vpr;
  if (q < 0) { q += vpr; --r; } else if (q >= vpr) { q -= vpr; ++r; }
  int pix = 0, k = 0;
  float pixf = 0.f;
  auto locate = [&]() {   // (pixel, Gaussian) of the first float of vector q
    const int f = 4 * q;
    pix = __float2int_rd(((float)f + 0.5f) * inv_k);
    k = f - pix * K;
    if (k < 0) { k += K; --pix; } else if (k >= K) { k -= K; ++pix; }
    pixf = (float)pix;
  };
  locate();
  while (r < rows) {
    const float4 *rc = rowc + r * K;
    int kk = k;
    float xf = pixf;
    float x = __fadd_rn(-1.0f, __fmul_rn(h, xf));     // grid_coord of the pixel
    float g[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float4 c = rc[kk];
      const float t = x - c.x;
      g[e] = ex2_approx(fmaf(c.z * t, t, c.y));
      if (++kk == K) { kk = 0; xf += 1.0f; x = __fadd_rn(-1.0f, __fmul_rn(h, xf)); }
    }
    st_f4(mtile + (size_t)r * n * K + 4 * q, g[0], g[1], g[2], g[3]);
    q += kThreads;
    if (q < vpr) {
      pix += dp; pixf += dpf; k += dk;
      if (k >= K) { k -= K; ++pix; pixf += 1.0f; }
    } else {
      do { q -= vpr; ++r; } while (q >= vpr);
      locate();
    }
  }
}

// =============================================================================== backward, quads
template <bool WITH_GMASK>
__global__ void __launch_bounds__(kThreads, 2) splat_bwd_quad_kernel(const __grid_constant__ SplatArgs a,
                                                                  int kq_shift) {
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  __shared__ float wsum[kWarps][32 * 4 * GG_MOMENTS];   // [warp][kq][i][moment], K*5 <= 640 floats used

  const int b = blockIdx.x / a.T;
  const int tile = blockIdx.x - b * a.T;
  const ScaleDesc &s = a.sc[find_scale(a, tile)];
  const int K = a.K, n = s.n;
  const int KQ = 1 << kq_shift;
  const int kq = threadIdx.x & (KQ - 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  float mux[4], muy[4], nA[4], nr[4], nE[4];
  {
    const float4 *gp = reinterpret_cast<const float4 *>(a.params + ((size_t)b * K + 4 * kq) * GG_PARAM_STRIDE);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 lo = __ldg(gp + 2 * i), hi = __ldg(gp + 2 * i + 1);
      mux[i] = lo.x; muy[i] = lo.y; nA[i] = lo.z; nr[i] = lo.w; nE[i] = hi.x;
    }
  }
  const int row_begin = (tile - s.tile_begin) * s.rows_per_tile;
  const int row_end = min(n, row_begin + s.rows_per_tile);
  const int quads_per_row = n << kq_shift;

  float T[4][GG_MOMENTS];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < GG_MOMENTS; ++j) T[i][j] = 0.f;

  for (int row = row_begin; row < row_end; ++row) {
    const float y = grid_coord(row, s.step);
    float dy[4], sh[4], e[4], S0[4], S1[4], S2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      dy[i] = y - muy[i];
      sh[i] = fmaf(nr[i], dy[i], mux[i]);
      e[i] = (nE[i] * dy[i]) * dy[i];
      S0[i] = S1[i] = S2[i] = 0.f;
    }
    const float *grow = s.maps + ((size_t)b * n + row) * n * s.cstride + 4 * kq;
    const bool use_gmask = WITH_GMASK && s.mask != nullptr;
    const float *mrow = use_gmask ? s.mask + ((size_t)b * n + row) * n : nullptr;
    // batches of kUQ quads per thread: all 128-bit loads of a batch are issued before any use (MLP)
    constexpr int kUQ = 4;
    for (int q0 = threadIdx.x; q0 < quads_per_row; q0 += kUQ * kThreads) {
      float4 G[kUQ];
      float gmv[kUQ];
#pragma unroll
      for (int u = 0; u < kUQ; ++u) {
        const int qq = q0 + u * kThreads;
        G[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        gmv[u] = 0.f;
        if (qq < quads_per_row) {
          G[u] = __ldg(reinterpret_cast<const float4 *>(grow + (size_t)(qq >> kq_shift) * s.cstride));
          if (use_gmask) gmv[u] = __ldg(mrow + (qq >> kq_shift));
        }
      }
#pragma unroll
      for (int u = 0; u < kUQ; ++u) {
        const int qq = q0 + u * kThreads;
        // column clamped for the quads past the row's end (their G is 0, but a conic that is not positive definite
        // -- a blob at the camera plane -- grows away from its centre and inf * 0 would poison the sums)
        const float x = grid_coord(min(qq >> kq_shift, n - 1), s.step);
        const f32x2 x2 = pack2(x, x);
        const f32x2 gm2 = pack2(gmv[u], gmv[u]);
        const f32x2 Ga = add2(pack2(G[u].x, G[u].y), gm2), Gb = add2(pack2(G[u].z, G[u].w), gm2);
        // Gaussians (0,1) and (2,3) as packed pairs
        const f32x2 ta = sub2(x2, pack2(sh[0], sh[1])), tb = sub2(x2, pack2(sh[2], sh[3]));
        const f32x2 ua = mul2(ta, ta), ub = mul2(tb, tb);
        float a0, a1, a2, a3;
        unpack2(fma2(pack2(nA[0], nA[1]), ua, pack2(e[0], e[1])), a0, a1);
        unpack2(fma2(pack2(nA[2], nA[3]), ub, pack2(e[2], e[3])), a2, a3);
        const f32x2 wa = mul2(pack2(ex2_approx(a0), ex2_approx(a1)), Ga);
        const f32x2 wb = mul2(pack2(ex2_approx(a2), ex2_approx(a3)), Gb);
        f32x2 s0a = pack2(S0[0], S0[1]), s0b = pack2(S0[2], S0[3]);
        f32x2 s1a = pack2(S1[0], S1[1]), s1b = pack2(S1[2], S1[3]);
        f32x2 s2a = pack2(S2[0], S2[1]), s2b = pack2(S2[2], S2[3]);
        s0a = add2(s0a, wa); s0b = add2(s0b, wb);
        s1a = fma2(wa, ta, s1a); s1b = fma2(wb, tb, s1b);
        s2a = fma2(wa, ua, s2a); s2b = fma2(wb, ub, s2b);
        unpack2(s0a, S0[0], S0[1]); unpack2(s0b, S0[2], S0[3]);
        unpack2(s1a, S1[0], S1[1]); unpack2(s1b, S1[2], S1[3]);
        unpack2(s2a, S2[0], S2[1]); unpack2(s2b, S2[2], S2[3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      T[i][0] += S2[i];
      T[i][3] += S1[i];
      T[i][1] = fmaf(dy[i], S1[i], T[i][1]);
      const float d0 = dy[i] * S0[i];
      T[i][4] += d0;
      T[i][2] = fmaf(dy[i], d0, T[i][2]);
    }
  }
  // lanes that share kq (stride KQ) -> butterfly; then lanes 0..KQ-1 hold the warp totals
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < GG_MOMENTS; ++j) {
      float v = T[i][j];
      for (int off = KQ; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane < KQ) wsum[warp][(4 * kq + i) * GG_MOMENTS + j] = v;
    }
  __syncthreads();
  float *out = a.partials + (((size_t)b * a.slots + a.tile_offset + tile) * K) * GG_MOMENTS;
  for (int c = threadIdx.x; c < K * GG_MOMENTS; c += kThreads) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) acc += wsum[w][c];
    out[c] = acc;
  }
}

// ====================================================================================== planning
namespace {

int env_int(const char *name, int dflt) {
  const char *v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

void strip_geometry(ScaleDesc &d, int n, int rpt, bool contig = false, int cta = kThreads) {
  d.n = n;
  d.contig = contig ? 1 : 0;
  d.cta = cta;
  d.spr = ((n + 3) / 4 + 1) / 2;     // threads per row: half the quads, each thread takes quad s and s + spr
  d.sprc = d.spr < cta ? d.spr : cta;
  d.rpp = d.spr <= cta ? cta / d.spr : 1;
  d.chunks = (d.spr + cta - 1) / cta;
  d.rows_per_tile = d.rpp * rpt;
  d.tiles = ((n + d.rows_per_tile - 1) / d.rows_per_tile) * d.chunks;
  d.step = n > 1 ? (1.0f - (-1.0f)) / (float)(n - 1) : 0.0f;
}

void quad_geometry(ScaleDesc &d, int n, int KQ, int passes) {
  d.n = n;
  d.contig = 0;
  d.cta = kThreads;
  d.spr = d.sprc = d.rpp = d.chunks = 0;
  int per_row = (n * KQ + kThreads - 1) / kThreads;       // passes a CTA needs for one row
  d.rows_per_tile = passes / per_row > 0 ? passes / per_row : 1;
  if (d.rows_per_tile > n) d.rows_per_tile = n;
  d.tiles = (n + d.rows_per_tile - 1) / d.rows_per_tile;
  d.step = n > 1 ? (1.0f - (-1.0f)) / (float)(n - 1) : 0.0f;
}

// tiles of the persistent maps backward (gg_splat_tma.cu): whole rows, about GG_FAST_TILE_KB of upstream gradient
// each, whatever K is.  Measured at B = 64, 256 x 256 (profiles/r1_maps_mode_timings.txt): 128 KB tiles 49.9 us at
// K = 16 and 26.6 us at K = 6; 32 KB tiles 55.2 / 30.8 us (per-item bookkeeping), 512 KB tiles balance worse
void fast_bwd_geometry(ScaleDesc &d, int n, int K) {
  static const int tile_kb = env_int("GG_FAST_TILE_KB", 128);
  d.n = n;
  d.contig = 0;
  d.cta = kThreads;
  d.spr = d.sprc = d.rpp = d.chunks = 0;
  const long long row_bytes = (long long)n * K * 4;
  long long rows = ((long long)tile_kb * 1024 + row_bytes / 2) / row_bytes;
  rows = rows < 1 ? 1 : rows > n ? n : rows;
  // The 8 warps of a CTA take the tile's slices round-robin and meet at a barrier when an image's tiles are done:
  // prefer a tile height near the target whose slice count is a multiple of 8 (K = 6: 21 rows = 33.6 slices left two
  // warps with a fifth slice while six waited -- 15 % of the warp time in the barrier under ncu; 20 rows = 32 slices)
  {
    const bool quads = (K & 3) == 0 && (K >> 2) <= 32;
    const int KU = quads ? K >> 2 : (K + 1) >> 1;
    const int strips_per_slice = KU <= 32 ? (quads ? 1 : 2) * (32 / KU) : 0;
    const int strips_per_row = n >> 3;
    if (strips_per_slice > 0 && strips_per_row > 0) {
      long long best = rows;
      double best_cost = 1e30;
      for (long long r = rows - rows / 4; r <= rows + rows / 4 && r <= n; ++r) {
        if (r < 1) continue;
        const long long nsl = (r * strips_per_row + strips_per_slice - 1) / strips_per_slice;
        const long long per_warp = (nsl + kWarps - 1) / kWarps;
        const double waste = (double)(per_warp * kWarps - nsl) / (double)(per_warp * kWarps);   // idle warp-slices
        const double cost = waste + 0.02 * (double)(r > rows ? r - rows : rows - r) / (double)rows;
        if (cost < best_cost) { best_cost = cost; best = r; }
      }
      rows = best;
    }
  }
  d.rows_per_tile = (int)rows;
  d.tiles = (n + d.rows_per_tile - 1) / d.rows_per_tile;
  d.step = n > 1 ? (1.0f - (-1.0f)) / (float)(n - 1) : 0.0f;
}

bool quad_ok(int K, int *shift) {
  if (K < 4 || (K & 3)) return false;
  int KQ = K >> 2;
  if (KQ & (KQ - 1)) return false;
  if (KQ > 32) return false;
  int sft = 0;
  while ((1 << sft) < KQ) ++sft;
  *shift = sft;
  return true;
}

}  // namespace

// exported for gg_fused.cu
void make_strip_desc(ScaleDesc &d, int n, int rpt) { strip_geometry(d, n, rpt); }

int fast_cta(bool backward, int rpt);
cudaError_t launch_splat_bwd_quad_tma(const SplatArgs &a, int B, int kq_shift, bool any_mask, cudaStream_t st);
bool splat_bwd_quad_tma_ok(const SplatArgs &a, int kq_shift);
cudaError_t launch_splat_bwd_nhwc_fast(const SplatArgs &a, int B, bool any_mask, cudaStream_t st);
bool splat_bwd_nhwc_fast_ok(int K, int nscales, const int *sizes);
cudaError_t launch_fwd_mask_fast(const SplatArgs &a, int B, int rpt, cudaStream_t st);
cudaError_t launch_bwd_mask_fast(const SplatArgs &a, bool fused, int B, int rpt, cudaStream_t st);

// Mask-only launches use the forward-differenced kernels of gg_splat_fast.cu unless GG_PRECISE_MASK=1 (or
// gg_set_fast_mask(0)) asks for the direct one-exponential-per-pixel form.
static std::atomic<int> g_fast_mask{-1};
bool fast_mask() {
  int v = g_fast_mask.load(std::memory_order_relaxed);
  if (v < 0) {
    v = env_int("GG_PRECISE_MASK", 0) ? 0 : 1;
    int expected = -1;
    g_fast_mask.compare_exchange_strong(expected, v, std::memory_order_relaxed);
    v = g_fast_mask.load(std::memory_order_relaxed);
  }
  return v != 0;
}
int set_fast_mask(int enable) {
  const int prev = fast_mask() ? 1 : 0;
  g_fast_mask.store(enable ? 1 : 0, std::memory_order_relaxed);
  return prev;
}

int fwd_rpt() { static int v = env_int("GG_RPT_FWD", 2); return v; }
// CTA size of the forward-differenced kernels (threads): tuned so that the CTAs of the reference's batch (64
// images of 256 rows) fill whole waves of resident CTAs; GG_FAST_WARPS_FWD / GG_FAST_WARPS_BWD override (4, 5, 6, 8)
int fast_cta(bool backward, int rpt) {
  static int f = env_int("GG_FAST_WARPS_FWD", 8), b = env_int("GG_FAST_WARPS_BWD", 8);
  int w = backward ? b : f;
  if (rpt >= 4 || (w != 4 && w != 5 && w != 6)) w = 8;
  return 32 * w;
}
int bwd_rpt() { static int v = env_int("GG_RPT_BWD", 2); return v; }
int quad_passes(bool backward) {
  static int f = env_int("GG_QUAD_PASSES", 16), b = env_int("GG_QUAD_PASSES_BWD", 32);
  return backward ? b : f;
}

// forward-differenced NHWC maps backward (gg_splat_tma.cu): same switch as the silhouette kernels, GG_FAST_MAPS_BWD=0
// keeps the direct kernels
static bool fast_maps_bwd() {
  static const bool env_on = env_int("GG_FAST_MAPS_BWD", 1) != 0 && env_int("GG_TMA_BWD", 1) != 0;
  return env_on && fast_mask();
}

// flat forward kernel: 128-bit stores need whole vectors per row, the index arithmetic operands below 2^22, and the
// per-tile constants have to fit the default dynamic shared memory
static bool flat_fwd_ok(int K, int nscales, const int *sizes, const bool *has_maps) {
  for (int i = 0; i < nscales; ++i) {
    if (!has_maps[i]) continue;
    const long long nk = (long long)sizes[i] * K;
    if ((nk & 3) != 0 || nk >= (1 << 22)) return false;
  }
  return (size_t)16 * K * sizeof(float4) <= 48 * 1024;
}

// A launch plan: which scales go to which kernel family, with tile numbering.
enum MapsKind { MAPS_STRIP = 0, MAPS_QUAD, MAPS_FLAT };
struct Plan {
  SplatArgs strip_mask;    // scales with only a mask (or gmask); forward MAPS_FLAT: also the masks of map scales
  SplatArgs maps;          // scales with maps, mask fused (except forward MAPS_FLAT)
  int maps_kind;           // MAPS_QUAD: NHWC, K = 4 * 2^i (quad kernels); MAPS_FLAT: NHWC, any other K that the
                           // flat forward / forward-differenced backward handle; MAPS_STRIP: everything else
  bool maps_fast_bwd;      // backward of the maps through splat_bwd_nhwc_fast_kernel
  int kq_shift;
  bool maps_any_mask;
  bool strided;            // a maps scale sits inside a wider NHWC buffer (channel stride != K or channel offset != 0)
  int tiles_mask, tiles_maps;
};

// cstrides / coffsets (NHWC only, may be NULL = dense): scale i's K channels live at channel coffsets[i] of a buffer
// with cstrides[i] channels per pixel (the generator's concat buffer, mask/mask_gen_dynamic.py:523-529).  Vector
// accesses need the channel offset and stride to keep their alignment; otherwise the scalar strip kernels take over.
static bool make_plan(Plan &pl, int K, int nscales, const int *sizes, const bool *has_maps, const bool *has_mask,
                      int layout, bool backward, const int *cstrides = nullptr, const int *coffsets = nullptr) {
  pl = Plan();
  pl.kq_shift = 0;
  int map_sizes[GG_MAX_SCALES], nmaps = 0;
  bool strided = false, vec4 = true, vec2 = true;
  for (int i = 0; i < nscales; ++i)
    if (has_maps[i]) {
      map_sizes[nmaps++] = sizes[i];
      const int cs = cstrides ? cstrides[i] : K, c0 = coffsets ? coffsets[i] : 0;
      strided = strided || cs != K || c0 != 0;
      vec4 = vec4 && (cs & 3) == 0 && (c0 & 3) == 0;
      vec2 = vec2 && (cs & 1) == 0 && (c0 & 1) == 0;
    }
  pl.strided = strided;
  pl.maps_kind = MAPS_STRIP;
  if (layout == GG_LAYOUT_NHWC && nmaps > 0 && !strided) {
    pl.maps_fast_bwd = backward && fast_maps_bwd() && splat_bwd_nhwc_fast_ok(K, nmaps, map_sizes);
    if (quad_ok(K, &pl.kq_shift)) pl.maps_kind = MAPS_QUAD;
    else if (backward ? pl.maps_fast_bwd : flat_fwd_ok(K, nscales, sizes, has_maps)) pl.maps_kind = MAPS_FLAT;
  } else if (layout == GG_LAYOUT_NHWC && nmaps > 0) {
    // concat-direct: quads with 128-bit accesses (forward: any K % 4 == 0; backward: K = 4 * 2^m, register-staged or
    // 2-D bulk-tensor copies), pairs with 64-bit stores (forward, even K), scalar strips otherwise
    if (quad_ok(K, &pl.kq_shift) && vec4) pl.maps_kind = MAPS_QUAD;
    else if (!backward && (K & 3) == 0 && vec4 && (K >> 2) <= kThreads) pl.maps_kind = MAPS_FLAT;
    else if (!backward && (K & 1) == 0 && vec2 && (K >> 1) <= kThreads) pl.maps_kind = MAPS_FLAT;
  }
  int rpt = backward ? bwd_rpt() : fwd_rpt();
  int tb_mask = 0, tb_maps = 0;
  auto add_mask_scale = [&](int n) {
    ScaleDesc &d = pl.strip_mask.sc[pl.strip_mask.nscales++];
    d.cstride = K;
    if (fast_mask()) strip_geometry(d, n, rpt >= 4 ? 4 : 2, true, fast_cta(backward, rpt));   // row pairs
    else strip_geometry(d, n, rpt);
    d.tile_begin = tb_mask;
    tb_mask += d.tiles;
  };
  for (int i = 0; i < nscales; ++i) {
    if (has_maps[i]) {
      ScaleDesc &d = pl.maps.sc[pl.maps.nscales++];
      d.cstride = cstrides ? cstrides[i] : K;
      if (pl.maps_fast_bwd) fast_bwd_geometry(d, sizes[i], K);
      else if (pl.maps_kind != MAPS_STRIP) quad_geometry(d, sizes[i], (K + 3) >> 2, quad_passes(backward));
      else strip_geometry(d, sizes[i], rpt);
      d.tile_begin = tb_maps;
      tb_maps += d.tiles;
      if (has_mask[i]) {
        if (pl.maps_kind == MAPS_FLAT && !backward) add_mask_scale(sizes[i]);   // silhouette by its own kernel
        else pl.maps_any_mask = true;
      }
    } else if (has_mask[i]) {
      add_mask_scale(sizes[i]);
    }
  }
  pl.tiles_mask = tb_mask;
  pl.tiles_maps = tb_maps;
  return true;
}

int plan_bwd_tiles(int K, int nscales, const int *sizes, const int *has_gmaps, const int *has_gmasks, int layout,
                   const int *cstrides, const int *coffsets) {
  Plan pl;
  bool hm[GG_MAX_SCALES] = {}, hk[GG_MAX_SCALES] = {};
  for (int i = 0; i < nscales; ++i) { hm[i] = has_gmaps && has_gmaps[i]; hk[i] = has_gmasks && has_gmasks[i]; }
  make_plan(pl, K, nscales, sizes, hm, hk, layout, true, cstrides, coffsets);
  return pl.tiles_mask + pl.tiles_maps;
}

template <int RPT>
static cudaError_t launch_fwd_strip(const SplatArgs &a, int out_mode, int B, cudaStream_t st) {
  dim3 grid((unsigned)((size_t)B * a.T));
  size_t smem = (size_t)a.K * GG_PARAM_STRIDE * sizeof(float);
  switch (out_mode) {
    case 0: launch_pdl(splat_fwd_strip_kernel<RPT, 0>, dim3(grid), dim3(kThreads), smem, st, a); break;
    case 1: launch_pdl(splat_fwd_strip_kernel<RPT, 1>, dim3(grid), dim3(kThreads), smem, st, a); break;
    default: launch_pdl(splat_fwd_strip_kernel<RPT, 2>, dim3(grid), dim3(kThreads), smem, st, a); break;
  }
  return cudaGetLastError();
}

template <int RPT>
static cudaError_t launch_bwd_strip(const SplatArgs &a, int gin_mode, int B, cudaStream_t st) {
  dim3 grid((unsigned)((size_t)B * a.T), (unsigned)((a.K + kKBlock - 1) / kKBlock));
  switch (gin_mode) {
    case 0: launch_pdl(splat_bwd_strip_kernel<RPT, 0>, dim3(grid), dim3(kThreads), 0, st, a); break;
    case 1: launch_pdl(splat_bwd_strip_kernel<RPT, 1>, dim3(grid), dim3(kThreads), 0, st, a); break;
    case 3: launch_pdl(splat_bwd_strip_kernel<RPT, 3>, dim3(grid), dim3(kThreads), 0, st, a); break;
    default: launch_pdl(splat_bwd_strip_kernel<RPT, 2>, dim3(grid), dim3(kThreads), 0, st, a); break;
  }
  return cudaGetLastError();
}

// columns of a row one thread of the quadg / pair forwards writes (cpp = pixels per pass), largest scale of the launch
static int fwd_cols_per_thread(const SplatArgs &a, int cpp) {
  int cols = 1;
  for (int i = 0; i < a.nscales; ++i) cols = max(cols, (a.sc[i].n + cpp - 1) / cpp);
  return cols;
}

cudaError_t launch_splat_fwd(const float *params, int B, int K, int nscales, const int *sizes, float *const *maps,
                             float *const *masks, int layout, cudaStream_t st, const int *cstrides,
                             const int *coffsets) {
  bool hm[GG_MAX_SCALES] = {}, hk[GG_MAX_SCALES] = {};
  for (int i = 0; i < nscales; ++i) { hm[i] = maps && maps[i]; hk[i] = masks && masks[i]; }
  Plan pl;
  make_plan(pl, K, nscales, sizes, hm, hk, layout, false, cstrides, coffsets);
  int im = 0, ik = 0;
  for (int i = 0; i < nscales; ++i) {
    const bool own_mask_kernel = hm[i] && hk[i] && pl.maps_kind == MAPS_FLAT;
    if (hm[i]) {
      pl.maps.sc[im].maps = maps[i] + (coffsets ? coffsets[i] : 0);
      pl.maps.sc[im].mask = hk[i] && !own_mask_kernel ? masks[i] : nullptr;
      ++im;
    }
    if (hk[i] && (!hm[i] || own_mask_kernel)) { pl.strip_mask.sc[ik].maps = nullptr; pl.strip_mask.sc[ik].mask = masks[i]; ++ik; }
  }
  cudaError_t err = cudaSuccess;
  const int rpt = fwd_rpt();
  if (pl.strip_mask.nscales) {
    SplatArgs &a = pl.strip_mask;
    a.params = params; a.B = B; a.K = K; a.T = pl.tiles_mask; a.layout = layout;
    if (fast_mask()) err = launch_fwd_mask_fast(a, B, rpt, st);
    else err = rpt == 1 ? launch_fwd_strip<1>(a, 0, B, st) : rpt == 4 ? launch_fwd_strip<4>(a, 0, B, st)
                                                                        : launch_fwd_strip<2>(a, 0, B, st);
    if (err != cudaSuccess) return err;
  }
  if (pl.maps.nscales) {
    SplatArgs &a = pl.maps;
    a.params = params; a.B = B; a.K = K; a.T = pl.tiles_maps; a.layout = layout;
    static const bool quad_hoist = env_int("GG_QUAD_HOIST", 1) != 0;
    // (K = 4 / 8 without a fused silhouette: the column-hoisted kernel below, 14.3 -> 11.3 us at K = 4, B = 64, 256 x 256;
    // K >= 16 is bound by the HBM writes in either kernel)
    if (pl.maps_kind == MAPS_QUAD && (pl.maps_any_mask || !quad_hoist || K > 8 || fwd_cols_per_thread(a, kThreads / (K >> 2)) > 8)) {
      dim3 grid((unsigned)((size_t)B * a.T));
      if (pl.maps_any_mask) launch_pdl(splat_fwd_quad_kernel<true>, dim3(grid), dim3(kThreads), 0, st, a, pl.kq_shift);
      else launch_pdl(splat_fwd_quad_kernel<false>, dim3(grid), dim3(kThreads), 0, st, a, pl.kq_shift);
      err = cudaGetLastError();
    } else if ((pl.maps_kind == MAPS_QUAD || pl.maps_kind == MAPS_FLAT) && (K & 3) == 0 && (K >> 2) <= kThreads) {
      // (maps without a fused silhouette at K = 4 * 2^m take the column-hoisted kernel as well)
      const int cols = fwd_cols_per_thread(a, kThreads / (K >> 2));
      const dim3 grid((unsigned)((size_t)B * a.T));
      if (cols <= 4) launch_pdl(splat_fwd_quadg_kernel<4>, grid, dim3(kThreads), 0, st, a, K >> 2);
      else if (cols <= 8) launch_pdl(splat_fwd_quadg_kernel<8>, grid, dim3(kThreads), 0, st, a, K >> 2);
      else launch_pdl(splat_fwd_quadg_kernel<0>, grid, dim3(kThreads), 0, st, a, K >> 2);
      err = cudaGetLastError();
    } else if (pl.maps_kind == MAPS_FLAT && (K & 1) == 0 && (K >> 1) <= kThreads) {
      // (a variant with 128-bit stores over pixel PAIRS -- K / 2 whole vectors per two pixels -- was measured in round 2
      // and is slower than these 64-bit stores: K = 6: 4.41 vs 4.55 TB/s, K = 10: 5.33 vs 5.76, K = 14: 5.70 vs 6.28 at
      // B = 64; at this size the launch ramp, not the store width, is what separates K = 6 from the copy bandwidth)
      const int cols = fwd_cols_per_thread(a, kThreads / (K >> 1));
      const dim3 grid((unsigned)((size_t)B * a.T));
      if (cols <= 4) launch_pdl(splat_fwd_pair_kernel<4>, grid, dim3(kThreads), 0, st, a, K >> 1);
      else if (cols <= 8) launch_pdl(splat_fwd_pair_kernel<8>, grid, dim3(kThreads), 0, st, a, K >> 1);
      else launch_pdl(splat_fwd_pair_kernel<0>, grid, dim3(kThreads), 0, st, a, K >> 1);
      err = cudaGetLastError();
    } else if (pl.maps_kind == MAPS_FLAT) {
      int max_rows = 1;
      for (int i = 0; i < a.nscales; ++i) max_rows = a.sc[i].rows_per_tile > max_rows ? a.sc[i].rows_per_tile : max_rows;
      launch_pdl(splat_fwd_flat_kernel, dim3((unsigned)((size_t)B * a.T)), dim3(kThreads),
                 (size_t)max_rows * K * sizeof(float4), st, a, 1.0f / (float)K);
      err = cudaGetLastError();
    } else {
      int mode = layout == GG_LAYOUT_NCHW ? 1 : 2;
      err = rpt == 1 ? launch_fwd_strip<1>(a, mode, B, st) : rpt == 4 ? launch_fwd_strip<4>(a, mode, B, st)
                                                                       : launch_fwd_strip<2>(a, mode, B, st);
    }
  }
  return err;
}

cudaError_t launch_splat_bwd(const float *params, int B, int K, int nscales, const int *sizes,
                             const float *const *gmaps, const float *const *gmasks, int layout, float *partials,
                             int T, cudaStream_t st, const int *cstrides, const int *coffsets) {
  bool hm[GG_MAX_SCALES] = {}, hk[GG_MAX_SCALES] = {};
  for (int i = 0; i < nscales; ++i) { hm[i] = gmaps && gmaps[i]; hk[i] = gmasks && gmasks[i]; }
  Plan pl;
  make_plan(pl, K, nscales, sizes, hm, hk, layout, true, cstrides, coffsets);
  if (pl.tiles_mask + pl.tiles_maps != T) return cudaErrorInvalidValue;
  int im = 0, ik = 0;
  for (int i = 0; i < nscales; ++i) {
    if (hm[i]) {
      pl.maps.sc[im].maps = const_cast<float *>(gmaps[i]) + (coffsets ? coffsets[i] : 0);
      pl.maps.sc[im].mask = hk[i] ? const_cast<float *>(gmasks[i]) : nullptr;
      ++im;
    } else if (hk[i]) {
      pl.strip_mask.sc[ik].maps = nullptr;
      pl.strip_mask.sc[ik].mask = const_cast<float *>(gmasks[i]);
      ++ik;
    }
  }
  cudaError_t err = cudaSuccess;
  const int rpt = bwd_rpt();
  // SplatArgs.T is the tile count of THIS launch (splits blockIdx.x); SplatArgs.slots is the total
  // number of partial slots per batch element (the partials stride).
  if (pl.strip_mask.nscales) {
    SplatArgs &a = pl.strip_mask;
    a.params = params; a.partials = partials; a.B = B; a.K = K; a.T = pl.tiles_mask; a.slots = T; a.tile_offset = 0;
    a.layout = layout;
    if (fast_mask()) err = launch_bwd_mask_fast(a, false, B, rpt, st);
    else err = rpt == 1 ? launch_bwd_strip<1>(a, 0, B, st) : rpt == 4 ? launch_bwd_strip<4>(a, 0, B, st)
                                                                        : launch_bwd_strip<2>(a, 0, B, st);
    if (err != cudaSuccess) return err;
  }
  if (pl.maps.nscales) {
    SplatArgs &a = pl.maps;
    a.params = params; a.partials = partials; a.B = B; a.K = K; a.T = pl.tiles_maps; a.slots = T;
    a.tile_offset = pl.tiles_mask; a.layout = layout;
    if (pl.maps_fast_bwd) {
      err = launch_splat_bwd_nhwc_fast(a, B, pl.maps_any_mask, st);
    } else if (pl.maps_kind == MAPS_QUAD) {
      dim3 grid((unsigned)((size_t)B * a.T));
      // the quad kernels add gmask to every channel only when a scale carries one; scales without a
      // gmask are handled by giving the kernel a per-scale NULL check
      static const bool use_tma = env_int("GG_TMA_BWD", 1) != 0;
      if (use_tma && !pl.strided && splat_bwd_quad_tma_ok(a, pl.kq_shift)) {
        err = launch_splat_bwd_quad_tma(a, B, pl.kq_shift, pl.maps_any_mask, st);
      } else {
        if (pl.maps_any_mask) launch_pdl(splat_bwd_quad_kernel<true>, dim3(grid), dim3(kThreads), 0, st, a, pl.kq_shift);
        else launch_pdl(splat_bwd_quad_kernel<false>, dim3(grid), dim3(kThreads), 0, st, a, pl.kq_shift);
        err = cudaGetLastError();
      }
    } else {
      int mode = layout == GG_LAYOUT_NCHW ? 1 : 2;
      err = rpt == 1 ? launch_bwd_strip<1>(a, mode, B, st) : rpt == 4 ? launch_bwd_strip<4>(a, mode, B, st)
                                                                       : launch_bwd_strip<2>(a, mode, B, st);
    }
  }
  return err;
}

// fused silhouette + L1 loss + backward, one launch (K <= kKBlock)
int fused_loss_tiles(int K, int n) {
  const int sizes[1] = {n}, zero[1] = {0}, one[1] = {1};
  return plan_bwd_tiles(K, 1, sizes, zero, one, GG_LAYOUT_NHWC);
}

cudaError_t launch_silhouette_loss_fused(const float *params, const void *target, int target_kind, int B, int K, int n,
                                         float *mask_out, float *partials, float *loss_partials, int T,
                                         cudaStream_t st) {
  if (K > kKBlock) return cudaErrorInvalidValue;
  bool hm[1] = {false}, hk[1] = {true};
  const int sizes[1] = {n};
  Plan pl;
  make_plan(pl, K, 1, sizes, hm, hk, GG_LAYOUT_NHWC, true);
  if (pl.tiles_mask != T) return cudaErrorInvalidValue;
  SplatArgs &a = pl.strip_mask;
  a.params = params; a.partials = partials; a.B = B; a.K = K; a.T = T; a.slots = T; a.tile_offset = 0;
  a.layout = GG_LAYOUT_NHWC;
  a.sc[0].maps = nullptr;
  a.sc[0].mask = mask_out;
  a.target = target; a.target_kind = target_kind; a.inv_n = 1.0f / (float)((size_t)B * n * n);
  a.loss_partials = loss_partials;
  const int rpt = bwd_rpt();
  if (fast_mask()) return launch_bwd_mask_fast(a, true, B, rpt, st);
  return rpt == 1 ? launch_bwd_strip<1>(a, 3, B, st) : rpt == 4 ? launch_bwd_strip<4>(a, 3, B, st)
                                                                : launch_bwd_strip<2>(a, 3, B, st);
}

}  // namespace gg

// gg_fused.cu -- direct consumers of the silhouette, fused so the K-channel maps never exist.
//
// mask_l1_loss_bwd: the reference's `m_recon_loss_bis`
//     m_b  = reduce_sum(pose_embeddings[-1], axis=-1)            mask/mask_gen_dynamic.py:785
//     loss = reduce_mean(|0.5 * (mask + 1) - m_b|)               mask/mask_gen_dynamic.py:787
// with `mask = raw / 127.5 - 1` (:672) when the target is given as the 8-bit image the data loader
// produces.  Emits dL/dm (= -sign(residual) / N, TF's Abs gradient: sign(0) = 0) for the splat backward
// and one partial sum of |residual| per CTA (summed on the host in float64: deterministic, no atomics).
#include "gg_common.cuh"
#include "gg_strip.cuh"
#include "gg_launch.h"

namespace gg {

template <int KIND>   // 0: uint8 raw target, 1: float32 target already in [-1, 1]
__global__ void __launch_bounds__(256) mask_l1_loss_bwd_kernel(const float *__restrict__ mask,
                                                               const void *__restrict__ target, size_t N,
                                                               float inv_n, float *__restrict__ gmask,
                                                               float *__restrict__ partials) {
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  __shared__ float red[8];
  float acc = 0.f;
  const size_t N4 = N >> 2;
  if (blockIdx.x == 0 && threadIdx.x < (N & 3)) {   // the 1-3 elements past the last whole vector
    const size_t i = 4 * N4 + threadIdx.x;
    const float t = KIND == 0
        ? __fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn((float)__ldg(reinterpret_cast<const unsigned char *>(target) + i), 127.5f), 1.0f), 1.0f), 0.5f)
        : __fmul_rn(__fadd_rn(__ldg(reinterpret_cast<const float *>(target) + i), 1.0f), 0.5f);
    const float r = t - __ldg(mask + i);
    acc += fabsf(r);
    gmask[i] = r > 0.f ? -inv_n : (r < 0.f ? inv_n : 0.f);
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < N4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 m = __ldg(reinterpret_cast<const float4 *>(mask) + i);
    float t[4];
    if (KIND == 0) {
      const uchar4 u = __ldg(reinterpret_cast<const uchar4 *>(target) + i);
      // (raw / 127.5 - 1 + 1) * 0.5 evaluated as the reference does, op by op in float32
      t[0] = __fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn((float)u.x, 127.5f), 1.0f), 1.0f), 0.5f);
      t[1] = __fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn((float)u.y, 127.5f), 1.0f), 1.0f), 0.5f);
      t[2] = __fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn((float)u.z, 127.5f), 1.0f), 1.0f), 0.5f);
      t[3] = __fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn((float)u.w, 127.5f), 1.0f), 1.0f), 0.5f);
    } else {
      const float4 f = __ldg(reinterpret_cast<const float4 *>(target) + i);
      t[0] = __fmul_rn(__fadd_rn(f.x, 1.0f), 0.5f);
      t[1] = __fmul_rn(__fadd_rn(f.y, 1.0f), 0.5f);
      t[2] = __fmul_rn(__fadd_rn(f.z, 1.0f), 0.5f);
      t[3] = __fmul_rn(__fadd_rn(f.w, 1.0f), 0.5f);
    }
    const float mm[4] = {m.x, m.y, m.z, m.w};
    float g[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float r = t[j] - mm[j];
      acc += fabsf(r);
      g[j] = r > 0.f ? -inv_n : (r < 0.f ? inv_n : 0.f);
    }
    reinterpret_cast<float4 *>(gmask)[i] = make_float4(g[0], g[1], g[2], g[3]);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    partials[blockIdx.x] = s;
  }
}

cudaError_t launch_mask_l1_loss_bwd(const float *mask, const void *target, int kind, int B, int n, float *gmask,
                                    float *partials, int npartials, cudaStream_t st) {
  size_t N = (size_t)B * n * n;
  float inv_n = 1.0f / (float)N;
  if (kind == 0) launch_pdl(mask_l1_loss_bwd_kernel<0>, dim3(npartials), dim3(256), 0, st, mask, target, N, inv_n, gmask, partials);
  else launch_pdl(mask_l1_loss_bwd_kernel<1>, dim3(npartials), dim3(256), 0, st, mask, target, N, inv_n, gmask, partials);
  return cudaGetLastError();
}


// ======================================================================================== cycle L1
// gm_cycle_loss = reduce_mean(|pose_embeddings[-1] - pose_embeddings_cyc[-1]|)   mask/mask_gen_dynamic.py:791
// over [B,n,n,K].  Both renders are evaluated per (pixel, Gaussian) in registers; the sign of the difference
// is the upstream gradient of BOTH renders (+ for a, - for b), so loss and both sets of moments come out of one
// pass and the two [B,n,n,K] tensors (2 x 268 MB at B64,K16,256^2) are never written.
constexpr int kCycKBlock = 32;     // Gaussians per CTA along grid.y
constexpr int kCycKC = 3;          // Gaussians per warp flush: 3 x (2 sets x 5 moments) = 30 columns

struct CycleArgs {
  const float *params_a, *params_b;
  float *partials_a, *partials_b;   // [B,T,K,5]
  float *loss_partials;             // [B*T*KB]
  int B, K, T, KB;
  float inv_n;                      // 1 / (B*n*n*K)
  ScaleDesc sc;
};

template <int RPT>
__global__ void __launch_bounds__(kThreads) cycle_l1_fused_kernel(const __grid_constant__ CycleArgs a) {
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  __shared__ float4 spa[kCycKBlock * 2], spb[kCycKBlock * 2];
  __shared__ float stage[kWarps][kCycKC * 2 * GG_MOMENTS][33];
  __shared__ float warpsum[kWarps][kCycKBlock * 2 * GG_MOMENTS];
  __shared__ float lred[kWarps];

  const ScaleDesc &s = a.sc;
  const int b = blockIdx.x / a.T, tile = blockIdx.x - b * a.T;
  const int K = a.K, n = s.n;
  const int k0 = blockIdx.y * kCycKBlock, kn = min(kCycKBlock, K - k0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const float4 *ga = reinterpret_cast<const float4 *>(a.params_a + ((size_t)b * K + k0) * GG_PARAM_STRIDE);
    const float4 *gb = reinterpret_cast<const float4 *>(a.params_b + ((size_t)b * K + k0) * GG_PARAM_STRIDE);
    for (int i = threadIdx.x; i < 2 * kn; i += kThreads) { spa[i] = __ldg(ga + i); spb[i] = __ldg(gb + i); }
  }
  __syncthreads();
  const StripPos pos = strip_pos(s, tile);
  float x[kStrip], y[RPT], vm[RPT][kStrip];
#pragma unroll
  for (int j = 0; j < kStrip; ++j) x[j] = grid_coord(pos.col(j), s.step);
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int row = pos.row0 + r * s.rpp;
    y[r] = grid_coord(row, s.step);
#pragma unroll
    for (int j = 0; j < kStrip; ++j) vm[r][j] = (pos.valid && row < n && pos.col(j) < n) ? 1.f : 0.f;
  }
  float lsum = 0.f;
  for (int kk = 0; kk < kn; ++kk) {
    const float4 pa = spa[2 * kk], pb = spb[2 * kk];
    const float nEa = spa[2 * kk + 1].x, nEb = spb[2 * kk + 1].x;
    float Ta[GG_MOMENTS] = {0.f, 0.f, 0.f, 0.f, 0.f}, Tb[GG_MOMENTS] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const float dya = y[r] - pa.y, dyb = y[r] - pb.y;
      const float sha = fmaf(pa.w, dya, pa.x), shb = fmaf(pb.w, dyb, pb.x);
      const float ea = (nEa * dya) * dya, eb = (nEb * dyb) * dyb;
      const f32x2 sha2 = pack2(sha, sha), shb2 = pack2(shb, shb);
      const f32x2 nAa2 = pack2(pa.z, pa.z), nAb2 = pack2(pb.z, pb.z), ea2 = pack2(ea, ea), eb2 = pack2(eb, eb);
      f32x2 a0 = pack2(0.f, 0.f), a1 = a0, a2 = a0, b0 = a0, b1 = a0, b2 = a0;
#pragma unroll
      for (int j = 0; j < kStrip; j += 2) {
        const f32x2 x2 = pack2(x[j], x[j + 1]);
        const f32x2 ta = sub2(x2, sha2), tb = sub2(x2, shb2);
        const f32x2 ua = mul2(ta, ta), ub = mul2(tb, tb);
        float qa0, qa1, qb0, qb1;
        unpack2(fma2(nAa2, ua, ea2), qa0, qa1);
        unpack2(fma2(nAb2, ub, eb2), qb0, qb1);
        const f32x2 ga = pack2(ex2_approx(qa0), ex2_approx(qa1)), gb = pack2(ex2_approx(qb0), ex2_approx(qb1));
        float d0, d1;
        unpack2(mul2(sub2(ga, gb), pack2(vm[r][j], vm[r][j + 1])), d0, d1);
        lsum += fabsf(d0) + fabsf(d1);
        // TF's Abs gradient: sign(d) (0 at 0), scaled by 1/N of the mean
        const float s0 = d0 > 0.f ? a.inv_n : (d0 < 0.f ? -a.inv_n : 0.f);
        const float s1 = d1 > 0.f ? a.inv_n : (d1 < 0.f ? -a.inv_n : 0.f);
        const f32x2 wa = mul2(ga, pack2(s0, s1)), wb = mul2(gb, pack2(-s0, -s1));
        a0 = add2(a0, wa); a1 = fma2(wa, ta, a1); a2 = fma2(wa, ua, a2);
        b0 = add2(b0, wb); b1 = fma2(wb, tb, b1); b2 = fma2(wb, ub, b2);
      }
      float lo, hi, S0, S1, S2;
      unpack2(a0, lo, hi); S0 = lo + hi; unpack2(a1, lo, hi); S1 = lo + hi; unpack2(a2, lo, hi); S2 = lo + hi;
      Ta[0] += S2; Ta[3] += S1; Ta[1] = fmaf(dya, S1, Ta[1]);
      float d0a = dya * S0; Ta[4] += d0a; Ta[2] = fmaf(dya, d0a, Ta[2]);
      unpack2(b0, lo, hi); S0 = lo + hi; unpack2(b1, lo, hi); S1 = lo + hi; unpack2(b2, lo, hi); S2 = lo + hi;
      Tb[0] += S2; Tb[3] += S1; Tb[1] = fmaf(dyb, S1, Tb[1]);
      float d0b = dyb * S0; Tb[4] += d0b; Tb[2] = fmaf(dyb, d0b, Tb[2]);
    }
    const int c0 = (kk % kCycKC) * 2 * GG_MOMENTS;
#pragma unroll
    for (int j = 0; j < GG_MOMENTS; ++j) {
      stage[warp][c0 + j][lane] = Ta[j];
      stage[warp][c0 + GG_MOMENTS + j][lane] = Tb[j];
    }
    if ((kk % kCycKC) == kCycKC - 1 || kk == kn - 1) {
      __syncwarp();
      const int ncols = c0 + 2 * GG_MOMENTS;
      const int cbase = (kk / kCycKC) * (kCycKC * 2 * GG_MOMENTS);
      if (lane < ncols) {
        float acc = 0.f;
#pragma unroll
        for (int l = 0; l < 32; ++l) acc += stage[warp][lane][l];
        warpsum[warp][cbase + lane] = acc;        // column = kk*10 + set*5 + j
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
  if (lane == 0) lred[warp] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) tot += lred[w];
    a.loss_partials[((size_t)b * a.T + tile) * a.KB + blockIdx.y] = tot;
  }
  for (int c = threadIdx.x; c < kn * 2 * GG_MOMENTS; c += kThreads) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) acc += warpsum[w][c];
    const int kk = c / (2 * GG_MOMENTS), rem = c - kk * 2 * GG_MOMENTS;
    const int set = rem / GG_MOMENTS, j = rem - set * GG_MOMENTS;
    float *out = (set == 0 ? a.partials_a : a.partials_b) + (((size_t)b * a.T + tile) * K + k0 + kk) * GG_MOMENTS + j;
    *out = acc;
  }
}

int cycle_tiles(int n) {
  ScaleDesc d;
  make_strip_desc(d, n, bwd_rpt() == 4 ? 4 : (bwd_rpt() == 1 ? 1 : 2));
  return d.tiles;
}

cudaError_t launch_cycle_l1_fused(const float *params_a, const float *params_b, int B, int K, int n, float *partials_a,
                                  float *partials_b, float *loss_partials, int T, cudaStream_t st) {
  CycleArgs a;
  const int rpt = bwd_rpt() == 4 ? 4 : (bwd_rpt() == 1 ? 1 : 2);
  make_strip_desc(a.sc, n, rpt);
  a.sc.tile_begin = 0; a.sc.maps = nullptr; a.sc.mask = nullptr;
  if (a.sc.tiles != T) return cudaErrorInvalidValue;
  a.params_a = params_a; a.params_b = params_b; a.partials_a = partials_a; a.partials_b = partials_b;
  a.loss_partials = loss_partials; a.B = B; a.K = K; a.T = T; a.KB = (K + kCycKBlock - 1) / kCycKBlock;
  a.inv_n = 1.0f / ((float)B * (float)n * (float)n * (float)K);
  dim3 grid((unsigned)((size_t)B * T), (unsigned)a.KB);
  if (rpt == 1) launch_pdl(cycle_l1_fused_kernel<1>, dim3(grid), dim3(kThreads), 0, st, a);
  else if (rpt == 4) launch_pdl(cycle_l1_fused_kernel<4>, dim3(grid), dim3(kThreads), 0, st, a);
  else launch_pdl(cycle_l1_fused_kernel<2>, dim3(grid), dim3(kThreads), 0, st, a);
  return cudaGetLastError();
}

// ==================================================================================== viz composite
// colorize_landmark_maps: rgb[b,i,j,c] = max_k maps[b,i,j,k] * COLORS[k][c]     mask/mask_gen_dynamic.py:221-231,
// mask/infer_masks.py:174-184.  Two sources: staged Gaussians (maps never written) or an existing NHWC tensor.
// Output float32 [.,3] or, for frames (infer_masks.py:463: 255 - (rgb*255).astype(uint8)), uint8.
__device__ __forceinline__ void put_rgb(void *out, int out_u8, size_t pix, float r, float g, float bl) {
  if (out_u8) {
    unsigned char *o = reinterpret_cast<unsigned char *>(out) + pix * 3;
    // numpy's float -> uint8 cast truncates toward zero; colours are in [0,1] and maps in [0,1].
    // out_u8 == 2: the inverted frame the reference writes, 255 - uint8(rgb*255) (mask/infer_masks.py:463)
    unsigned char c0 = (unsigned char)(int)(r * 255.0f), c1 = (unsigned char)(int)(g * 255.0f),
                  c2 = (unsigned char)(int)(bl * 255.0f);
    if (out_u8 == 2) { c0 = 255 - c0; c1 = 255 - c1; c2 = 255 - c2; }
    o[0] = c0; o[1] = c1; o[2] = c2;
  } else {
    float *o = reinterpret_cast<float *>(out) + pix * 3;
    o[0] = r; o[1] = g; o[2] = bl;
  }
}

__global__ void __launch_bounds__(256) colorize_params_kernel(const float *__restrict__ params,
                                                              const float *__restrict__ colors, int K, int n,
                                                              float step, void *out, int out_u8) {
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  extern __shared__ float4 sm[];          // [K][2] params, then [K] colours (float4, w unused)
  float4 *sp = sm, *sc = sm + 2 * K;
  const int b = blockIdx.y;
  const float4 *gp = reinterpret_cast<const float4 *>(params + (size_t)b * K * GG_PARAM_STRIDE);
  for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) sp[i] = __ldg(gp + i);
  for (int i = threadIdx.x; i < K; i += blockDim.x)
    sc[i] = make_float4(__ldg(colors + 3 * i), __ldg(colors + 3 * i + 1), __ldg(colors + 3 * i + 2), 0.f);
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= n * n) return;
  const int row = pix / n, col = pix - row * n;
  const float x = grid_coord(col, step), y = grid_coord(row, step);
  float r = -INFINITY, g = -INFINITY, bl = -INFINITY;
  for (int k = 0; k < K; ++k) {
    const float4 p = sp[2 * k];
    const float nE = sp[2 * k + 1].x;
    const float dy = y - p.y;
    const float t = x - fmaf(p.w, dy, p.x);
    const float v = ex2_approx(fmaf(p.z, t * t, (nE * dy) * dy));
    const float4 c = sc[k];
    r = fmaxf(r, v * c.x); g = fmaxf(g, v * c.y); bl = fmaxf(bl, v * c.z);
  }
  put_rgb(out, out_u8, (size_t)b * n * n + pix, r, g, bl);
}

__global__ void __launch_bounds__(256) colorize_maps_kernel(const float *__restrict__ maps,
                                                            const float *__restrict__ colors, size_t npix, int K,
                                                            void *out, int out_u8) {
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const float *m = maps + pix * K;
  float r = -INFINITY, g = -INFINITY, bl = -INFINITY;
  for (int k = 0; k < K; ++k) {
    const float v = __ldg(m + k);
    r = fmaxf(r, v * __ldg(colors + 3 * k)); g = fmaxf(g, v * __ldg(colors + 3 * k + 1));
    bl = fmaxf(bl, v * __ldg(colors + 3 * k + 2));
  }
  put_rgb(out, out_u8, pix, r, g, bl);
}

cudaError_t launch_colorize_params(const float *params, const float *colors, int B, int K, int n, void *out,
                                   int out_u8, cudaStream_t st) {
  dim3 grid((unsigned)(((size_t)n * n + 255) / 256), (unsigned)B);
  const float step = n > 1 ? (1.0f - (-1.0f)) / (float)(n - 1) : 0.0f;
  launch_pdl(colorize_params_kernel, dim3(grid), dim3(256), (size_t)K * 3 * sizeof(float4), st, params, colors, K, n, step, out, out_u8);
  return cudaGetLastError();
}

cudaError_t launch_colorize_maps(const float *maps, const float *colors, size_t npix, int K, void *out, int out_u8,
                                 cudaStream_t st) {
  launch_pdl(colorize_maps_kernel, dim3((unsigned)((npix + 255) / 256)), dim3(256), 0, st, maps, colors, npix, K, out, out_u8);
  return cudaGetLastError();
}

}  // namespace gg

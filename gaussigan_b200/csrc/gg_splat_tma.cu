// gg_splat_tma.cu -- backward of the NHWC maps through a bulk-copy (TMA) pipeline.
//
// Same contract and arithmetic as splat_bwd_quad_kernel (gg_splat.cu): upstream gradients w.r.t. the
// [B,n,n,K] maps (mask/mask_gen_dynamic.py:141, the tensors the generator consumes at :523-529) are reduced to
// five moments per Gaussian.  The kernel is HBM-read bound (4 K bytes per pixel, ~45 instructions per 16 bytes),
// and with plain loads a thread can only keep a few 128-bit loads in flight in registers: measured 2.6 TB/s of the
// 6.5 TB/s the part delivers.  Here the rows of a tile -- contiguous in NHWC -- are streamed into shared-memory
// rings by cp.async.bulk (completion on mbarriers), 64 KB per CTA in flight independent of the register file; the
// compute threads read their quad with one LDS.128.
#include "gg_common.cuh"
#include "gg_strip.cuh"
#include "gg_launch.h"

namespace gg {

namespace {

#ifndef GG_TMA_STAGES
#define GG_TMA_STAGES 4
#endif
#ifndef GG_TMA_MINBLOCKS
#define GG_TMA_MINBLOCKS 2
#endif
constexpr int kQStages = GG_TMA_STAGES;        // ring depth

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned); completes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace

// One ring PER WARP: a warp streams its own 128-quad (2 KB) slices of the tile, its lane 0 issues the bulk copies and
// the warp waits on its own mbarriers -- no CTA-wide synchronisation inside the loop, warps drift freely.
// The tile (rows [row_begin, row_end) of one image) is one contiguous span of quads in NHWC; slice j of the span is
// handled by warp (j mod 8) and lies inside one image row because quads_per_row is a multiple of 128 (checked by
// the launcher; other shapes take the register-staged kernel).
constexpr int kWarpQuads = 128;                        // quads per warp slice = 4 per lane
constexpr int kWarpStageBytes = kWarpQuads * 16;       // 2 KB

template <bool WITH_GMASK>
__global__ void __launch_bounds__(kThreads, GG_TMA_MINBLOCKS) splat_bwd_quad_tma_kernel(const __grid_constant__ SplatArgs a,
                                                                         int kq_shift) {
  extern __shared__ __align__(128) unsigned char ring[];       // [kWarps][kQStages][kWarpStageBytes]
  __shared__ __align__(8) uint64_t full[kWarps][kQStages];
  __shared__ float wsum[kWarps][32 * 4 * GG_MOMENTS];          // [warp][kq][i][moment], K*5 <= 640 floats used

  const int b = blockIdx.x / a.T;
  const int tile = blockIdx.x - b * a.T;
  const ScaleDesc &s = a.sc[find_scale(a, tile)];
  const int K = a.K, n = s.n;
  const int KQ = 1 << kq_shift;
  const int kq = threadIdx.x & (KQ - 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_begin = (tile - s.tile_begin) * s.rows_per_tile;
  const int row_end = min(n, row_begin + s.rows_per_tile);
  const int quads_per_row = n << kq_shift;
  const int slices_per_row = quads_per_row / kWarpQuads;
  const int nslices = (row_end - row_begin) * slices_per_row;           // of the whole tile
  const int my_slices = nslices > warp ? (nslices - warp + kWarps - 1) / kWarps : 0;   // slices warp, warp+8, ...

  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kQStages; ++i) mbar_init(&full[warp][i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_enter();   // global memory is first touched below
  __syncwarp();

  const float *gtile = s.maps + ((size_t)b * n + row_begin) * n * K;     // the tile's contiguous span
  unsigned char *wring = ring + (size_t)warp * kQStages * kWarpStageBytes;
  auto issue = [&](int i) {   // lane 0 only: i-th slice of this warp
    const int st = i % kQStages;
    mbar_expect_tx(&full[warp][st], kWarpStageBytes);
    bulk_g2s(wring + (size_t)st * kWarpStageBytes, gtile + (size_t)(warp + i * kWarps) * kWarpQuads * 4, kWarpStageBytes,
             &full[warp][st]);
  };
  if (lane == 0) {
    for (int i = 0; i < min(kQStages, my_slices); ++i) issue(i);
  }

  float mux[4], muy[4], nA[4], nr[4], nE[4];
  {
    const float4 *gp = reinterpret_cast<const float4 *>(a.params + ((size_t)b * K + 4 * kq) * GG_PARAM_STRIDE);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 lo = __ldg(gp + 2 * i), hi = __ldg(gp + 2 * i + 1);
      mux[i] = lo.x; muy[i] = lo.y; nA[i] = lo.z; nr[i] = lo.w; nE[i] = hi.x;
    }
  }
  float T[4][GG_MOMENTS];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < GG_MOMENTS; ++j) T[i][j] = 0.f;

  const bool use_gmask = WITH_GMASK && s.mask != nullptr;
  float dy[4], sh[4], e[4], S0[4], S1[4], S2[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) dy[g] = sh[g] = e[g] = S0[g] = S1[g] = S2[g] = 0.f;
  int cur_row = -1;
  auto fold_row = [&]() {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      T[g][0] += S2[g];
      T[g][3] += S1[g];
      T[g][1] = fmaf(dy[g], S1[g], T[g][1]);
      const float d0 = dy[g] * S0[g];
      T[g][4] += d0;
      T[g][2] = fmaf(dy[g], d0, T[g][2]);
    }
  };
  for (int i = 0; i < my_slices; ++i) {
    const int slice = warp + i * kWarps;
    const int rrel = slice / slices_per_row;
    const int row = row_begin + rrel;
    const int q0 = (slice - rrel * slices_per_row) * kWarpQuads;      // first quad of the slice inside its row
    if (row != cur_row) {
      if (cur_row >= 0) fold_row();
      cur_row = row;
      const float y = grid_coord(row, s.step);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        dy[g] = y - muy[g];
        sh[g] = fmaf(nr[g], dy[g], mux[g]);
        e[g] = (nE[g] * dy[g]) * dy[g];
        S0[g] = S1[g] = S2[g] = 0.f;
      }
    }
    const float *mrow = use_gmask ? s.mask + ((size_t)b * n + row) * n : nullptr;
    float gmv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) gmv[u] = use_gmask ? __ldg(mrow + ((q0 + u * 32 + lane) >> kq_shift)) : 0.f;
    const int st = i % kQStages;
    mbar_wait(&full[warp][st], (unsigned)((i / kQStages) & 1));
    const float4 *sq = reinterpret_cast<const float4 *>(wring + (size_t)st * kWarpStageBytes);
    float4 Gq[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) Gq[u] = sq[u * 32 + lane];
    __syncwarp();                                    // every lane has its data in registers: the stage is free
    if (lane == 0 && i + kQStages < my_slices) issue(i + kQStages);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int qq = q0 + u * 32 + lane;
      const float4 G = Gq[u];
      const float x = grid_coord(qq >> kq_shift, s.step);
      const f32x2 x2 = pack2(x, x);
      const f32x2 gm2 = pack2(gmv[u], gmv[u]);
      const f32x2 Ga = add2(pack2(G.x, G.y), gm2), Gb = add2(pack2(G.z, G.w), gm2);
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
  if (cur_row >= 0) fold_row();
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

// shapes the bulk-copy kernel handles: every warp slice (128 quads) inside one row
bool splat_bwd_quad_tma_ok(const SplatArgs &a, int kq_shift) {
  for (int i = 0; i < a.nscales; ++i)
    if (((a.sc[i].n << kq_shift) % kWarpQuads) != 0) return false;
  return true;
}

cudaError_t launch_splat_bwd_quad_tma(const SplatArgs &a, int B, int kq_shift, bool any_mask, cudaStream_t st) {
  static bool configured_dev[64] = {};
  const size_t smem = (size_t)kWarps * kQStages * kWarpStageBytes;
  int dev = 0;
  cudaGetDevice(&dev);
  bool &configured = configured_dev[dev & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(splat_bwd_quad_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(splat_bwd_quad_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((unsigned)((size_t)B * a.T));
  if (any_mask) launch_pdl(splat_bwd_quad_tma_kernel<true>, grid, dim3(kThreads), smem, st, a, kq_shift);
  else launch_pdl(splat_bwd_quad_tma_kernel<false>, grid, dim3(kThreads), smem, st, a, kq_shift);
  return cudaGetLastError();
}

}  // namespace gg

// gg_splat_tma.cu -- backward of the NHWC maps through a bulk-copy (TMA) pipeline.
//
// Same contract and arithmetic as splat_bwd_quad_kernel (gg_splat.cu): upstream gradients w.r.t. the
// [B,n,n,K] maps (mask/mask_gen_dynamic.py:141, the tensors the generator consumes at :523-529) are reduced to
// five moments per Gaussian.  The kernel is HBM-read bound (4 K bytes per pixel, ~45 instructions per 16 bytes),
// and with plain loads a thread can only keep a few 128-bit loads in flight in registers: measured 2.6 TB/s of the
// 6.5 TB/s the part delivers.  Here the rows of a tile -- contiguous in NHWC -- are streamed into a 4-stage shared
// memory ring by cp.async.bulk (one elected thread, completion on an mbarrier), 16 KB per stage, so 64 KB per CTA
// are in flight independent of the register file; the compute threads read their quad with one LDS.128.
#include "gg_common.cuh"
#include "gg_strip.cuh"
#include "gg_launch.h"

namespace gg {

namespace {

constexpr int kQStages = 4;                    // ring depth
constexpr int kQBatch = 4 * kThreads;          // quads (16-byte chunks: one pixel x 4 Gaussians) per stage = 16 KB
constexpr int kStageBytes = kQBatch * 16;

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

template <bool WITH_GMASK>
__global__ void __launch_bounds__(kThreads, 2) splat_bwd_quad_tma_kernel(const __grid_constant__ SplatArgs a,
                                                                         int kq_shift) {
  extern __shared__ __align__(128) unsigned char ring[];       // [kQStages][kStageBytes]
  __shared__ __align__(8) uint64_t full[kQStages];
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
  const int nb = (quads_per_row + kQBatch - 1) / kQBatch;      // batches per row
  const int nbatches = (row_end - row_begin) * nb;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kQStages; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_enter();   // global memory is first touched below
  __syncthreads();

  const float *gbase = s.maps + (size_t)b * n * n * K;
  auto issue = [&](int i) {   // thread 0 only
    const int row = row_begin + i / nb;
    const int q0 = (i - (i / nb) * nb) * kQBatch;
    const unsigned bytes = (unsigned)min(kQBatch, quads_per_row - q0) * 16u;
    const int st = i % kQStages;
    mbar_expect_tx(&full[st], bytes);
    bulk_g2s(ring + (size_t)st * kStageBytes, gbase + ((size_t)row * n * K + (size_t)q0 * 4), bytes, &full[st]);
  };
  if (threadIdx.x == 0) {
    for (int i = 0; i < min(kQStages, nbatches); ++i) issue(i);
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
  for (int i = 0; i < nbatches; ++i) {
    const int rb = i / nb, ib = i - rb * nb;
    const int row = row_begin + rb;
    const int q0 = ib * kQBatch;
    if (ib == 0) {
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
    for (int u = 0; u < 4; ++u) {
      const int qq = q0 + u * kThreads + (int)threadIdx.x;
      gmv[u] = (use_gmask && qq < quads_per_row) ? __ldg(mrow + (qq >> kq_shift)) : 0.f;
    }
    const int st = i % kQStages;
    mbar_wait(&full[st], (unsigned)((i / kQStages) & 1));
    const float4 *sq = reinterpret_cast<const float4 *>(ring + (size_t)st * kStageBytes);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int qq = q0 + u * kThreads + (int)threadIdx.x;
      float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
      if (qq < quads_per_row) G = sq[u * kThreads + threadIdx.x];
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
    if (ib == nb - 1) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        T[g][0] += S2[g];
        T[g][3] += S1[g];
        T[g][1] = fmaf(dy[g], S1[g], T[g][1]);
        const float d0 = dy[g] * S0[g];
        T[g][4] += d0;
        T[g][2] = fmaf(dy[g], d0, T[g][2]);
      }
    }
    __syncthreads();                                   // every thread is done with stage `st`
    if (threadIdx.x == 0 && i + kQStages < nbatches) issue(i + kQStages);
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

cudaError_t launch_splat_bwd_quad_tma(const SplatArgs &a, int B, int kq_shift, bool any_mask, cudaStream_t st) {
  static bool configured_dev[64] = {};
  const size_t smem = (size_t)kQStages * kStageBytes;
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

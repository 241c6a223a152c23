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
#include "gg_fast.cuh"
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

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_u32(uint32_t dst, const void *src, unsigned bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
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

// =============================================================== forward-differenced NHWC maps backward
// Same contract again for ANY K, a quarter of the exponentials and half the issue slots: thread = (strip of 8
// consecutive pixels of a row, unit of consecutive Gaussians) -- a unit is a quad (K % 4 == 0: one LDS.128 per
// pixel, NP = 2 lane pairs) or a pair (any other K, the reference's own K = 6 among them: NP = 1; an odd K leaves
// the last pair half empty).  Along the strip each Gaussian follows the recurrence of gg_splat_fast.cu
// (g_{j+1} = g_j r_j, r_{j+1} = r_j c: two exponentials per strip instead of eight); the two lanes of every FP32x2
// register are two GAUSSIANS of the unit, which is how the gradient arrives, so nothing is repacked.
//
// A warp slice is 32 / KU strips (KU = units per pixel; twice that for pair units, two strips per thread) of
// contiguous upstream gradient, at most 4 KB, one bulk copy; strips of a slice may lie in different rows (n % 8 == 0 keeps a strip inside one row), so any n that is a
// multiple of 8 works, including the 32 x 32 scale of the pyramid.  Bank conflicts: the 8 lanes of an LDS.128 phase
// are 8/KQ strips x KQ quads reading the same pixel index j, and strips are 128 KQ bytes apart -- the same banks.
// For power-of-two KQ <= 4 odd strips therefore run the recurrence BACKWARDS (from pixel 7, step -h: only the sign
// of one constant changes): at step i even strips read pixel i, odd strips pixel 7 - i, which differ in parity,
// i.e. in the 64-byte half of the bank window (conflict-free at K = 16; KQ >= 8 is conflict-free as it is).
//
// The kernel is PERSISTENT: every CTA (2 per SM) takes a contiguous range of the launch's (image, tile) items; a
// warp's ring keeps running across items, the accumulators across the tiles of an image -- one reduction and one
// partial-sum slot per (CTA, image), the slots of its other tiles are zero-filled -- and the per-Gaussian constants
// are re-staged only where the image or the scale changes (double-buffered, staged one segment ahead).
constexpr int kFastStageBytes = 4096;
#ifndef GG_FAST_STAGES
#define GG_FAST_STAGES 2
#endif
constexpr int kFastStages = GG_FAST_STAGES;     // ring depth per warp (3 measured slower: fewer bytes per barrier wait)

__device__ __forceinline__ f32x2 lds64(uint32_t addr) {
  f32x2 v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// the unit of one pixel: NP lane pairs.  `kodd` (odd K, NP = 1): pixels are not 8-byte aligned and the partner of
// the last Gaussian does not exist -- two 32-bit loads, the missing lane multiplied away by vm = (1, 0).
template <int NP>
__device__ __forceinline__ void load_unit(uint32_t addr, bool kodd, f32x2 vm, f32x2 (&G)[NP]) {
  if (NP == 2) {
    const float4 v = lds128(addr);
    G[0] = pack2(v.x, v.y);
    G[NP - 1] = pack2(v.z, v.w);
  } else if (!kodd) {
    G[0] = lds64(addr);
  } else {
    const float lo = lds32(addr), hi = lds32(addr + 4);
    G[0] = mul2(pack2(lo, hi), vm);
  }
}

// One strip (8 pixels) of one unit of Gaussians by the recurrence, per lane pair p (lanes = two Gaussians):
//   A = sum u_j,  Bs = sum (8 - j) u_j,  Cs = sum (8 - j)(9 - j)/2 u_j,   u_j = g_j Gbar_j
// as running sums (A += u; Bs += A; Cs += Bs: three two-operand adds per pixel, no per-pixel coordinates in
// registers).  They are the strip's moments about the point one step past its last pixel, on a uniform grid -- the
// same assumption the recurrence itself makes, and the guard (|nA| <= 512, i.e. sigma >= 4.8 pixels at n = 256)
// keeps the 1.8e-7 non-uniformity of the reference's float32 grid below 5e-6 of any first moment.
// sp = shared address of the unit of the strip's first pixel in the order of traversal, dstride = +-4 K bytes.
template <int NP, bool WITH_GMASK>
__device__ __forceinline__ void strip_sums_recurrence(uint32_t sp, int dstride, bool kodd, f32x2 vm,
                                                      const float (&gm)[8], f32x2 (&gg)[NP], f32x2 (&rr)[NP],
                                                      const f32x2 (&cc)[NP], f32x2 (&A)[NP], f32x2 (&Bs)[NP],
                                                      f32x2 (&Cs)[NP]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    f32x2 G[NP];
    load_unit<NP>(sp + (uint32_t)(j * dstride), kodd, vm, G);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      if (WITH_GMASK) G[p] = add2(G[p], bc(gm[j]));
      const f32x2 u = mul2(gg[p], G[p]);
      if (j == 0) {
        A[p] = Bs[p] = Cs[p] = u;
      } else {
        A[p] = add2(A[p], u);
        Bs[p] = add2(Bs[p], A[p]);
        Cs[p] = add2(Cs[p], Bs[p]);
      }
      if (j < 7) {
        gg[p] = mul2(gg[p], rr[p]);
        if (j < 6) rr[p] = mul2(rr[p], cc[p]);
      }
    }
  }
}

// The same strip with one exponential per pixel on the reference's own float32 grid (units that contain a narrow or
// very distant Gaussian): S0 = sum u, S1 = sum u t, S2 = sum u t^2 with t_j = x_j - (mu_x + nr dy) formed exactly as
// the direct kernels do (nsh = -(mu_x + nr dy)) and accumulated about the Gaussian's own centre line -- a needle
// much narrower than the strip would lose (strip length / sigma)^2 of its second moment's precision to cancellation
// if the sums were taken about the strip's end as above.
template <int NP, bool WITH_GMASK>
__device__ __forceinline__ void strip_sums_direct(uint32_t sp, int dstride, bool kodd, f32x2 vm, const float (&gm)[8],
                                                  int cstart, int dir, float h, const f32x2 (&nA)[NP],
                                                  const f32x2 (&e)[NP], const f32x2 (&nsh)[NP], f32x2 (&S0)[NP],
                                                  f32x2 (&S1)[NP], f32x2 (&S2)[NP]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    f32x2 G[NP];
    load_unit<NP>(sp + (uint32_t)(j * dstride), kodd, vm, G);
    const f32x2 xj = bc(grid_coord(cstart + dir * j, h));
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      if (WITH_GMASK) G[p] = add2(G[p], bc(gm[j]));
      const f32x2 t = add2(xj, nsh[p]);
      const f32x2 u = mul2(ex2_2(fma2(nA[p], mul2(t, t), e[p])), G[p]);
      const f32x2 v = mul2(u, t);
      S0[p] = j == 0 ? u : add2(S0[p], u);
      S1[p] = j == 0 ? v : add2(S1[p], v);
      S2[p] = j == 0 ? mul2(v, t) : fma2(v, t, S2[p]);
    }
  }
}

// the tile of item (b, tile): its contiguous span of upstream gradients and its size
struct FastTile {
  const float *gtile;
  int row_begin;
  int nstrips;               // 8-pixel strips in the tile
};

// `span` consecutive tiles of one scale of one image: they are contiguous in memory, and a CTA whose range holds several
// of them walks them as ONE span (the per-item bookkeeping is paid once)
__device__ __forceinline__ FastTile fast_tile(const SplatArgs &a, const ScaleDesc &s, int b, int tile, int span = 1) {
  FastTile t;
  t.row_begin = (tile - s.tile_begin) * s.rows_per_tile;
  const int row_end = min(s.n, t.row_begin + span * s.rows_per_tile);
  t.gtile = s.maps + ((size_t)b * s.n + t.row_begin) * s.n * a.K;
  t.nstrips = (row_end - t.row_begin) * (s.n >> 3);
  return t;
}

// constants of one (image, scale) segment; the guard is evaluated over the whole image
__device__ __forceinline__ void stage_segment_consts(const SplatArgs &a, const ScaleDesc &s, int b, float *sc) {
  const float lo = grid_coord(0, s.step), hi = grid_coord(s.n - 1, s.step);
  stage_fast_consts_paired(a.params + (size_t)b * a.K * GG_PARAM_STRIDE, a.K, s.step, lo, hi, lo, hi, sc);
}

// dynamic shared memory: [2][12 Kp] float constants (pair-interleaved, gg_fast.cuh; Kp = K rounded up to even) |
// [kWarps][K][5] float warp sums | [kWarps][kFastStages] stages
__host__ __device__ inline size_t fast_consts_floats(int K) { return (size_t)12 * ((K + 1) & ~1); }
__host__ __device__ inline size_t fast_consts_bytes(int K) { return ((2 * fast_consts_floats(K) * 4) + 127) & ~(size_t)127; }
__host__ __device__ inline size_t fast_wsum_bytes(int K) {
  return (((size_t)kWarps * K * GG_MOMENTS * 4) + 127) & ~(size_t)127;
}

// NP = lane pairs per thread (2: a quad of Gaussians, K % 4 == 0; 1: a pair); KU = units per pixel (<= 32)
// KODD (pair units only): K is odd -- decided at compile time, the even-K pair kernel (the reference's K = 6) carries no
// predicated-off copies of the odd-K loads
template <int NP, bool WITH_GMASK, bool KODD>
__global__ void __launch_bounds__(kThreads, 2) splat_bwd_nhwc_fast_kernel(const __grid_constant__ SplatArgs a, int KU,
                                                                          int nwork) {
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ __align__(8) uint64_t full[kWarps][kFastStages];

  const int K = a.K;
  const int Kp = (K + 1) & ~1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sl = lane / KU;                  // this thread's strip inside a warp slice ...
  const int ku = lane - sl * KU;             // ... and its unit of Gaussians (fixed for the whole kernel)
  const int kbase = 2 * NP * ku;             // first Gaussian of the unit
  const int spw = 32 / KU;                   // strips per warp slice; lanes beyond spw * KU idle
  const bool lane_ok = sl < spw;
  const bool pow2 = (KU & (KU - 1)) == 0;
  constexpr int SPT = NP == 2 ? 1 : 2;       // strips per thread and slice: pair units take two (4 KB copies again)
  const int slice_bytes = SPT * spw * 32 * K;   // a strip is 8 pixels x K floats
  const bool flip = NP == 2 && pow2 && KU <= 4 && (sl & 1);     // odd strips run from pixel 7 down to pixel 0
  const int dir = flip ? -1 : 1;
  constexpr bool kodd = KODD;
  const f32x2 vm = pack2(1.0f, kbase + 1 < K ? 1.0f : 0.0f);
  float *sc_base = reinterpret_cast<float *>(dyn);
  float *wsum = reinterpret_cast<float *>(dyn + fast_consts_bytes(K));
  // 32-bit shared addresses of this warp's ring and barriers, and of this thread's first unit inside a stage (in
  // its order of traversal); pinned so that they are not re-derived inside the loop
  const uint32_t ring_u32 = pin_u32(smem_u32(dyn + fast_consts_bytes(K) + fast_wsum_bytes(K)) +
                                    (uint32_t)(warp * kFastStages * kFastStageBytes));
  const uint32_t bar_u32 = pin_u32(smem_u32(&full[warp][0]));
  const uint32_t sp_u32 = pin_u32(ring_u32 + (uint32_t)(((sl * 8 + (flip ? 7 : 0)) * K + kbase) * 4));
  const int dstride = dir * 4 * K;
  const int w_begin = (int)((long long)blockIdx.x * nwork / gridDim.x);
  const int w_end = (int)((long long)(blockIdx.x + 1) * nwork / gridDim.x);

  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kFastStages; ++i) mbar_init(&full[warp][i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_enter();   // global memory is first touched below
  __syncwarp();

  // ---- issue cursor: runs kFastStages slices ahead of the consume cursor, across item boundaries.  Byte offsets
  // inside the item's tile: this warp's slices start at warp * slice_bytes and are 8 slices apart.
  int iss_b = w_begin / a.T, iss_tile = w_begin - iss_b * a.T, iss_left = w_end - w_begin;
  int iss_off = 0, iss_total = 0, iss_stage = 0, iss_span = 1;
  const char *iss_src = nullptr;
  auto iss_open = [&]() {   // the span of tiles that starts at item (iss_b, iss_tile): to the end of its scale or of the range
    const ScaleDesc &is = a.sc[find_scale(a, iss_tile)];
    iss_span = min(iss_left, is.tile_begin + is.tiles - iss_tile);
    const FastTile t = fast_tile(a, is, iss_b, iss_tile, iss_span);
    iss_src = reinterpret_cast<const char *>(t.gtile);
    iss_total = t.nstrips * 32 * K;
    iss_off = warp * slice_bytes;
  };
  if (iss_left > 0) iss_open();
  auto issue_next = [&]() {
    while (iss_left > 0 && iss_off >= iss_total) {
      iss_left -= iss_span;
      iss_tile += iss_span;
      if (iss_tile == a.T) { iss_tile = 0; ++iss_b; }
      if (iss_left > 0) iss_open();
    }
    if (iss_left <= 0) return;
    if (lane == 0) {
      const unsigned bytes = (unsigned)min(slice_bytes, iss_total - iss_off);
      const uint32_t bar = bar_u32 + (uint32_t)iss_stage * 8u;
      mbar_expect_tx_u32(bar, bytes);
      bulk_g2s_u32(ring_u32 + (uint32_t)iss_stage * kFastStageBytes, iss_src + iss_off, bytes, bar);
    }
    iss_off += kWarps * slice_bytes;
    iss_stage = iss_stage + 1 == kFastStages ? 0 : iss_stage + 1;
  };
#pragma unroll
  for (int i = 0; i < kFastStages; ++i) issue_next();

  int b = w_begin / a.T, tile = w_begin - b * a.T;
  if (w_begin < w_end) stage_segment_consts(a, a.sc[find_scale(a, tile)], b, sc_base);

  f32x2 mux[NP], muy[NP], nA[NP], nnr[NP], nE[NP], rb[NP], rc[NP], cc[NP];   // this thread's Gaussians, lane pairs
  bool steep = false;
  f32x2 T[NP][GG_MOMENTS];
#pragma unroll
  for (int p = 0; p < NP; ++p)
#pragma unroll
    for (int j = 0; j < GG_MOMENTS; ++j) T[p][j] = pack2(0.f, 0.f);
  int stage = 0;             // ring position and barrier phase of the consume cursor
  unsigned phase = 0;
  int seg_par = 0, prev_b = -1, prev_si = -1;

  for (int work = w_begin; work < w_end;) {
    const int si = find_scale(a, tile);
    const ScaleDesc &s = a.sc[si];
    const int n = s.n;
    const float h = s.step;
    const int span = min(w_end - work, s.tile_begin + s.tiles - tile);      // tiles of this image and scale walked as one
    if (b != prev_b || si != prev_si) {   // new (image, scale) segment: its constants were staged one segment ago
      prev_b = b; prev_si = si;
      __syncthreads();
      const float *sc = sc_base + (size_t)seg_par * fast_consts_floats(K);
      steep = false;
      if (lane_ok) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          // lane pair p = Gaussians kbase + 2 p, + 1: four LDS.128 whose 64-bit halves are the packed constants
          const float4 *q4 = reinterpret_cast<const float4 *>(sc + (NP * ku + p) * 16);
          const float4 w0 = q4[0], w1 = q4[1], w2 = q4[2], w3 = q4[3];
          mux[p] = pack2(w0.x, w0.y); muy[p] = pack2(w0.z, w0.w); nA[p] = pack2(w1.x, w1.y); nnr[p] = pack2(w1.z, w1.w);
          nE[p] = pack2(w2.x, w2.y); rc[p] = pack2(w3.x, w3.y); cc[p] = pack2(w3.z, w3.w);
          rb[p] = mul2(pack2(w2.z, w2.w), bc(flip ? -1.f : 1.f));     // 2 nA h, with h -> -h on reversed strips
          steep = steep || sc[8 * Kp + kbase + 2 * p] != 0.f || sc[8 * Kp + kbase + 2 * p + 1] != 0.f;
        }
      } else {
#pragma unroll
        for (int p = 0; p < NP; ++p)
          mux[p] = muy[p] = nA[p] = nnr[p] = nE[p] = rc[p] = cc[p] = rb[p] = pack2(0.f, 0.f);
      }
      seg_par ^= 1;
      // the next (image, scale) segment of this CTA's range starts right after this scale's tiles
      int nb = b, ntile = s.tile_begin + s.tiles;
      if (ntile == a.T) { ntile = 0; ++nb; }
      if (nb * a.T + ntile < w_end)
        stage_segment_consts(a, a.sc[find_scale(a, ntile)], nb, sc_base + (size_t)seg_par * fast_consts_floats(K));
    }
    const bool flush = work + span == w_end || tile + span == a.T;   // last tiles of this image in this CTA's range
    // the sums go into the slot of the span's last tile when the image ends here; every other slot of the span is zero
    float *slot0 = a.partials + (((size_t)b * a.slots + a.tile_offset + tile) * K) * GG_MOMENTS;
    float *out = slot0 + (size_t)(span - 1) * K * GG_MOMENTS;
    for (int c = threadIdx.x; c < (span - (flush ? 1 : 0)) * K * GG_MOMENTS; c += kThreads) slot0[c] = 0.f;

    const FastTile cur = fast_tile(a, s, b, tile, span);
    const int strips_per_row = n >> 3;
    const bool use_gmask = WITH_GMASK && s.mask != nullptr;
    const float hs = flip ? -h : h;                // signed step of this thread's traversal
    const f32x2 h8 = pin2(bc(8.0f * hs)), nhs = pin2(bc(-hs)), hh = pin2(bc(h * h));
    const int nslices = (cur.nstrips + SPT * spw - 1) / (SPT * spw);
    const int my_slices = nslices > warp ? (nslices - warp + kWarps - 1) / kWarps : 0;   // slices warp, warp+8, ...
    int g = warp * SPT * spw + sl;                 // first strip of this thread inside the tile
    int rrel = 0, cs = g;                          // its row inside the tile and its position inside the row
    while (cs >= strips_per_row) { cs -= strips_per_row; ++rrel; }
    for (int i = 0; i < my_slices; ++i) {
      bool joint_done = false;
      if (NP == 1 && __all_sync(0xffffffffu, !lane_ok || (!steep && g + spw < cur.nstrips))) {      // warp-uniform
        // pair units: the thread's TWO strips of the slice (g and g + spw) as two interleaved recurrences -- same
        // constants, different rows / columns / shared-memory addresses.  The straight-line code gives the scheduler two
        // independent dependency chains (the one-strip-after-the-other form below ran at half its issue slots in `wait`).
        int rs1 = rrel, cq1 = cs + spw;
        while (cq1 >= strips_per_row) { cq1 -= strips_per_row; ++rs1; }
        const int colA = cs * 8, colB = cq1 * 8;
        const int rowA = cur.row_begin + rrel, rowB = cur.row_begin + rs1;
        float gmA[8], gmB[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) gmA[j] = gmB[j] = 0.f;
        if (use_gmask && lane_ok) {
          const float4 *mpA = reinterpret_cast<const float4 *>(s.mask + ((size_t)b * n + rowA) * n + colA);
          const float4 *mpB = reinterpret_cast<const float4 *>(s.mask + ((size_t)b * n + rowB) * n + colB);
          const float4 a0 = __ldg(mpA), a1 = __ldg(mpA + 1), b0 = __ldg(mpB), b1 = __ldg(mpB + 1);
          gmA[0] = a0.x; gmA[1] = a0.y; gmA[2] = a0.z; gmA[3] = a0.w; gmA[4] = a1.x; gmA[5] = a1.y; gmA[6] = a1.z; gmA[7] = a1.w;
          gmB[0] = b0.x; gmB[1] = b0.y; gmB[2] = b0.z; gmB[3] = b0.w; gmB[4] = b1.x; gmB[5] = b1.y; gmB[6] = b1.z; gmB[7] = b1.w;
        }
        const f32x2 dyA = sub2(bc(grid_coord(rowA, h)), muy[0]), dyB = sub2(bc(grid_coord(rowB, h)), muy[0]);
        f32x2 trA = fma2(nnr[0], dyA, sub2(bc(grid_coord(colA, h)), mux[0]));
        f32x2 trB = fma2(nnr[0], dyB, sub2(bc(grid_coord(colB, h)), mux[0]));
        f32x2 ggA = ex2_2(fma2(mul2(nA[0], trA), trA, mul2(mul2(nE[0], dyA), dyA)));
        f32x2 ggB = ex2_2(fma2(mul2(nA[0], trB), trB, mul2(mul2(nE[0], dyB), dyB)));
        f32x2 rrA = ex2_2(fma2(rb[0], trA, rc[0])), rrB = ex2_2(fma2(rb[0], trB, rc[0]));
        mbar_wait_u32(bar_u32 + (uint32_t)stage * 8u, phase);
        // (idle lanes behind the last unit of a slice read the head of the stage: in bounds, result discarded)
        const uint32_t spA = (lane_ok ? sp_u32 : ring_u32) + (uint32_t)stage * kFastStageBytes;
        const uint32_t spB = spA + (lane_ok ? (uint32_t)(spw * 32 * K) : 0u);
        f32x2 A0, B0, C0, A1, B1, C1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f32x2 GA[1], GB[1];
          load_unit<1>(spA + (uint32_t)(j * dstride), kodd, vm, GA);
          load_unit<1>(spB + (uint32_t)(j * dstride), kodd, vm, GB);
          if (WITH_GMASK) { GA[0] = add2(GA[0], bc(gmA[j])); GB[0] = add2(GB[0], bc(gmB[j])); }
          const f32x2 uA = mul2(ggA, GA[0]), uB = mul2(ggB, GB[0]);
          if (j == 0) {
            A0 = B0 = C0 = uA; A1 = B1 = C1 = uB;
          } else {
            A0 = add2(A0, uA); A1 = add2(A1, uB);
            B0 = add2(B0, A0); B1 = add2(B1, A1);
            C0 = add2(C0, B0); C1 = add2(C1, B1);
          }
          if (j < 7) {
            ggA = mul2(ggA, rrA); ggB = mul2(ggB, rrB);
            if (j < 6) { rrA = mul2(rrA, cc[0]); rrB = mul2(rrB, cc[0]); }
          }
        }
        __syncwarp();                                  // every lane is done with the stage
        issue_next();
        trA = add2(trA, h8); trB = add2(trB, h8);
        const f32x2 M1A = mul2(B0, nhs), M1B = mul2(B1, nhs);
        const f32x2 M2A = mul2(sub2(add2(C0, C0), B0), hh), M2B = mul2(sub2(add2(C1, C1), B1), hh);
        const f32x2 m1A = fma2(trA, A0, M1A), m1B = fma2(trB, A1, M1B);
        const f32x2 m2A = fma2(trA, add2(M1A, m1A), M2A), m2B = fma2(trB, add2(M1B, m1B), M2B);
        const f32x2 d0A = mul2(dyA, A0), d0B = mul2(dyB, A1);
        // (the same order of additions as the one-strip form: strip g first, then g + spw; the idle lanes behind the
        // last unit of a slice have read stale but in-bounds shared memory: they keep their zero sums)
        if (lane_ok) {
          T[0][0] = add2(add2(T[0][0], m2A), m2B);
          T[0][3] = add2(add2(T[0][3], m1A), m1B);
          T[0][1] = fma2(dyB, m1B, fma2(dyA, m1A, T[0][1]));
          T[0][4] = add2(add2(T[0][4], d0A), d0B);
          T[0][2] = fma2(dyB, d0B, fma2(dyA, d0A, T[0][2]));
        }
        joint_done = true;
      }
#pragma unroll
      for (int ss = 0; ss < SPT; ++ss) {           // the thread's strips of this slice: g, g + spw
        if (joint_done) break;
        int rs = rrel, cq = cs;
        if (ss > 0) {
          cq += ss * spw;
          while (cq >= strips_per_row) { cq -= strips_per_row; ++rs; }
        }
        const bool valid = lane_ok && g + ss * spw < cur.nstrips;
        const int col0 = valid ? cq * 8 : 0;
        const int row = cur.row_begin + (valid ? rs : 0);
        const int cstart = flip ? col0 + 7 : col0;
        const float x0 = grid_coord(cstart, h);
        float gm[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) gm[j] = 0.f;
        if (use_gmask && valid) {
          const float4 *mp = reinterpret_cast<const float4 *>(s.mask + ((size_t)b * n + row) * n + col0);
          const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
          const float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) gm[j] = flip ? mm[7 - j] : mm[j];
        }
        const f32x2 y2 = bc(grid_coord(row, h)), x02 = bc(x0);
        f32x2 dy[NP], tr[NP], e[NP], gg[NP], rr[NP], M0[NP], M1[NP], M2[NP];   // tr = t at the moments' reference point
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          dy[p] = sub2(y2, muy[p]);
          tr[p] = fma2(nnr[p], dy[p], sub2(x02, mux[p]));
          e[p] = mul2(mul2(nE[p], dy[p]), dy[p]);
          gg[p] = ex2_2(fma2(mul2(nA[p], tr[p]), tr[p], e[p]));
          rr[p] = ex2_2(fma2(rb[p], tr[p], rc[p]));
        }
        if (ss == 0) mbar_wait_u32(bar_u32 + (uint32_t)stage * 8u, phase);
        const uint32_t sp = sp_u32 + (uint32_t)stage * kFastStageBytes + (uint32_t)(ss * spw * 32 * K);
        if (!valid) {
#pragma unroll
          for (int p = 0; p < NP; ++p) M0[p] = M1[p] = M2[p] = pack2(0.f, 0.f);
        } else if (!steep) {
          f32x2 Bs[NP], Cs[NP];
          strip_sums_recurrence<NP, WITH_GMASK>(sp, dstride, kodd, vm, gm, gg, rr, cc, M0, Bs, Cs);
#pragma unroll
          for (int p = 0; p < NP; ++p) {   // moments about the point one step past the strip: xi_j = -(8 - j) hs
            tr[p] = add2(tr[p], h8);
            M1[p] = mul2(Bs[p], nhs);
            M2[p] = mul2(sub2(add2(Cs[p], Cs[p]), Bs[p]), hh);
          }
        } else {
          f32x2 nsh[NP];
#pragma unroll
          for (int p = 0; p < NP; ++p) nsh[p] = fma2(nnr[p], dy[p], sub2(bc(0.f), mux[p]));
          strip_sums_direct<NP, WITH_GMASK>(sp, dstride, kodd, vm, gm, cstart, dir, h, nA, e, nsh, M0, M1, M2);
#pragma unroll
          for (int p = 0; p < NP; ++p) tr[p] = pack2(0.f, 0.f);      // the sums are already about t = 0
        }
        if (ss == SPT - 1) {
          __syncwarp();                                  // every lane is done with the stage
          issue_next();
        }
        // strip-local moments (about the reference point) -> moments of t = tr + xi, then the row weights
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const f32x2 m1 = fma2(tr[p], M0[p], M1[p]);
          const f32x2 m2 = fma2(tr[p], add2(M1[p], m1), M2[p]);
          const f32x2 d0 = mul2(dy[p], M0[p]);
          T[p][0] = add2(T[p][0], m2);
          T[p][3] = add2(T[p][3], m1);
          T[p][1] = fma2(dy[p], m1, T[p][1]);
          T[p][4] = add2(T[p][4], d0);
          T[p][2] = fma2(dy[p], d0, T[p][2]);
        }
      }
      g += kWarps * SPT * spw;
      cs += kWarps * SPT * spw;
      while (cs >= strips_per_row) { cs -= strips_per_row; ++rrel; }
      if (++stage == kFastStages) { stage = 0; phase ^= 1u; }
    }
    if (flush) {
      // lanes that share a unit (stride KU) are summed in a fixed order: butterfly when KU is a power of two, else
      // lane ku adds the lanes ku + KU, ku + 2 KU, ... one by one; lanes 0..KU-1 then hold the warp totals
#pragma unroll
      for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int j = 0; j < GG_MOMENTS; ++j) {
          float lo, hi;
          unpack2(T[p][j], lo, hi);
          if (pow2) {
            for (int off = KU; off < 32; off <<= 1) {
              lo += __shfl_xor_sync(0xffffffffu, lo, off);
              hi += __shfl_xor_sync(0xffffffffu, hi, off);
            }
          } else {
            const float lo0 = lo, hi0 = hi;
            for (int t = 1; t < spw; ++t) {
              lo += __shfl_sync(0xffffffffu, lo0, (lane + t * KU) & 31);
              hi += __shfl_sync(0xffffffffu, hi0, (lane + t * KU) & 31);
            }
          }
          if (lane < KU) {
            const int k = kbase + 2 * p;
            if (k < K) wsum[((size_t)warp * K + k) * GG_MOMENTS + j] = lo;
            if (k + 1 < K) wsum[((size_t)warp * K + k + 1) * GG_MOMENTS + j] = hi;
          }
          T[p][j] = pack2(0.f, 0.f);
        }
      __syncthreads();
      for (int c = threadIdx.x; c < K * GG_MOMENTS; c += kThreads) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) acc += wsum[(size_t)w * K * GG_MOMENTS + c];
        out[c] = acc;
      }
    }
    work += span;
    tile += span;
    if (tile == a.T) { tile = 0; ++b; }
  }
}

// units per pixel of the kernel above (0: shape not handled)
int splat_bwd_nhwc_fast_units(int K) {
  if (K >= 4 && (K & 3) == 0 && (K >> 2) <= 32) return K >> 2;     // quads
  if (K >= 1 && (K + 1) / 2 <= 32) return (K + 1) / 2;             // pairs
  return 0;
}

bool splat_bwd_nhwc_fast_ok(int K, int nscales, const int *sizes) {
  if (splat_bwd_nhwc_fast_units(K) == 0) return false;
  for (int i = 0; i < nscales; ++i)
    if ((sizes[i] & 7) != 0) return false;
  return true;
}

cudaError_t launch_splat_bwd_nhwc_fast(const SplatArgs &a, int B, bool any_mask, cudaStream_t st) {
  static int sm_count[64] = {};
  const size_t smem = fast_consts_bytes(a.K) + fast_wsum_bytes(a.K) + (size_t)kWarps * kFastStages * kFastStageBytes;
  int dev = 0;
  cudaGetDevice(&dev);
  int &sms = sm_count[dev & 63];
  if (sms == 0) {   // first launch on this device
    const int max_smem = (int)(fast_consts_bytes(128) + fast_wsum_bytes(128) + (size_t)kWarps * kFastStages * kFastStageBytes);
    cudaError_t e = cudaSuccess;
#define GG_SET_SMEM(NP, GM, KO)                                                                                       \
  if (e == cudaSuccess)                                                                                               \
    e = cudaFuncSetAttribute(splat_bwd_nhwc_fast_kernel<NP, GM, KO>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)
    GG_SET_SMEM(1, false, false); GG_SET_SMEM(1, true, false); GG_SET_SMEM(1, false, true); GG_SET_SMEM(1, true, true);
    GG_SET_SMEM(2, false, false); GG_SET_SMEM(2, true, false);
#undef GG_SET_SMEM
    int count = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    sms = count;
  }
  const size_t nwork = (size_t)B * a.T;
  if (nwork == 0) return cudaSuccess;
  if (nwork > 0x7fffffffu) return cudaErrorInvalidValue;
  const int KU = splat_bwd_nhwc_fast_units(a.K);
  if (KU == 0) return cudaErrorInvalidValue;
  const bool quads = (a.K & 3) == 0;
  // persistent: 2 CTAs per SM (registers and shared memory both allow exactly two)
  const unsigned grid = (unsigned)(nwork < (size_t)2 * sms ? nwork : (size_t)2 * sms);
  const bool kodd = (a.K & 1) != 0;
#define GG_LAUNCH_FAST(NP, GM, KO) \
  launch_pdl(splat_bwd_nhwc_fast_kernel<NP, GM, KO>, dim3(grid), dim3(kThreads), smem, st, a, KU, (int)nwork)
  if (quads) {
    if (any_mask) GG_LAUNCH_FAST(2, true, false); else GG_LAUNCH_FAST(2, false, false);
  } else if (kodd) {
    if (any_mask) GG_LAUNCH_FAST(1, true, true); else GG_LAUNCH_FAST(1, false, true);
  } else {
    if (any_mask) GG_LAUNCH_FAST(1, true, false); else GG_LAUNCH_FAST(1, false, false);
  }
#undef GG_LAUNCH_FAST
  return cudaGetLastError();
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

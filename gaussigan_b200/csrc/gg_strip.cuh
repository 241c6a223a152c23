// gg_strip.cuh -- strip geometry and 8-pixel load/store helpers shared by the per-pixel kernels
// (gg_splat.cu, gg_fused.cu).
#pragma once
#include "gg_common.cuh"

namespace gg {


static constexpr int kKC = 6;                      // Gaussians staged per warp-flush (5*kKC <= 32 columns)
static constexpr int kStageCols = kKC * GG_MOMENTS;
static constexpr int kStagePad = 36;               // row stride of the per-warp transpose buffer: 16-byte aligned rows,
                                                   // conflict-free for the column writes and for LDS.128 row reads

__device__ __forceinline__ int find_scale(const SplatArgs &a, int tile) {
  int si = 0;
#pragma unroll
  for (int i = 1; i < GG_MAX_SCALES; ++i)
    if (i < a.nscales && tile >= a.sc[i].tile_begin) si = i;
  return si;
}

// thread -> strip geometry shared by the strip kernels.  A strip is TWO quads of one row: pixels
// [colA, colA+4) and [colB, colB+4) with colB = colA + 4*ceil(Q/2), Q = quads per row -- so that the 32
// lanes of a warp-level 128-bit access touch one contiguous 512-byte run instead of every other 16 bytes.
struct StripPos {
  int colA, colB;
  int row0;       // row of pass r is row0 + r * rpp
  bool valid;
  __device__ __forceinline__ int col(int j) const { return j < 4 ? colA + j : colB + (j - 4); }
};

__device__ __forceinline__ StripPos strip_pos(const ScaleDesc &s, int tile_local) {
  StripPos p;
  int rt = tile_local / s.chunks;
  int ch = tile_local - rt * s.chunks;
  int rp = threadIdx.x / s.sprc;
  int sc = threadIdx.x - rp * s.sprc;
  int strip = ch * s.cta + sc;
  p.colA = s.contig ? strip * 8 : strip * 4;
  p.colB = s.contig ? strip * 8 + 4 : (strip + s.spr) * 4;
  p.row0 = rt * s.rows_per_tile + rp;
  p.valid = (rp < s.rpp) && (strip < s.spr);
  return p;
}

// rowp = address of column 0 of the row
__device__ __forceinline__ void store8(float *rowp, const StripPos &pos, int n, const float *v) {
  if ((n & 3) == 0) {
    if (pos.colA < n) st_f4(rowp + pos.colA, v[0], v[1], v[2], v[3]);
    if (pos.colB < n) st_f4(rowp + pos.colB, v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int j = 0; j < kStrip; ++j)
      if (pos.col(j) < n) rowp[pos.col(j)] = v[j];
  }
}

__device__ __forceinline__ void load8(const float *rowp, const StripPos &pos, int n, float *v) {
  if ((n & 3) == 0) {
    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
    if (pos.colA < n) lo = __ldg(reinterpret_cast<const float4 *>(rowp + pos.colA));
    if (pos.colB < n) hi = __ldg(reinterpret_cast<const float4 *>(rowp + pos.colB));
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
    v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
  } else {
#pragma unroll
    for (int j = 0; j < kStrip; ++j) v[j] = (pos.col(j) < n) ? __ldg(rowp + pos.col(j)) : 0.f;
  }
}


}  // namespace gg

// gg_fused.cu -- direct consumers of the silhouette, fused so the K-channel maps never exist.
//
// mask_l1_loss_bwd: the reference's `m_recon_loss_bis`
//     m_b  = reduce_sum(pose_embeddings[-1], axis=-1)            mask/mask_gen_dynamic.py:785
//     loss = reduce_mean(|0.5 * (mask + 1) - m_b|)               mask/mask_gen_dynamic.py:787
// with `mask = raw / 127.5 - 1` (:672) when the target is given as the 8-bit image the data loader
// produces.  Emits dL/dm (= -sign(residual) / N, TF's Abs gradient: sign(0) = 0) for the splat backward
// and one partial sum of |residual| per CTA (summed on the host in float64: deterministic, no atomics).
#include "gg_common.cuh"
#include "gg_launch.h"

namespace gg {

template <int KIND>   // 0: uint8 raw target, 1: float32 target already in [-1, 1]
__global__ void __launch_bounds__(256) mask_l1_loss_bwd_kernel(const float *__restrict__ mask,
                                                               const void *__restrict__ target, size_t N4,
                                                               float inv_n, float *__restrict__ gmask,
                                                               float *__restrict__ partials) {
  __shared__ float red[8];
  float acc = 0.f;
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
  if (N & 3) return cudaErrorInvalidValue;
  float inv_n = 1.0f / (float)N;
  if (kind == 0) mask_l1_loss_bwd_kernel<0><<<npartials, 256, 0, st>>>(mask, target, N / 4, inv_n, gmask, partials);
  else mask_l1_loss_bwd_kernel<1><<<npartials, 256, 0, st>>>(mask, target, N / 4, inv_n, gmask, partials);
  return cudaGetLastError();
}

}  // namespace gg

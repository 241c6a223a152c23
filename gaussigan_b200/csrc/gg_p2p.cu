// gg_p2p.cu -- the training-time exchange as ONE kernel over NVLink peer memory.
//
// The canonical Gaussians are broadcast over the batch (mask/mask_gen_dynamic.py:701-702), so their gradient is the
// sum of the per-sample gradients over the local batch AND over the ranks.  The payload is 12 K
// numbers (1.5 KB at K = 16): pure latency.  NCCL's all_reduce costs ~40-60 us per eager call at this size; here a
// single CTA per GPU does the local batch sum, stores its partial vector straight into every peer's receive slot
// through NVLink/NVSwitch peer pointers (symmetric memory: the same buffer layout on every rank), raises a flag on
// each peer with a system-scope release store, waits for the peers' flags with acquire loads and adds the slots in
// rank order -- the same order on every rank, so the result is bit-identical everywhere.
//
// Buffer layout (per rank, `gg_p2p_buffer_bytes`): [0] call counter, [1] error flag, [16 .. 16+W) arrival flags
// (uint32 each), then at byte 1024 the receive slots double[2][W][12 K].  Two slot sets: a rank can be at most one
// call ahead of the slowest peer (it cannot finish call e+1 before every peer has entered it, i.e. finished reading
// call e), so writes of call e+1 never touch what a peer still reads for call e.  The call counter lives on the
// device and is advanced by the kernel itself, which keeps the launch replayable from a CUDA graph.
// The wait is bounded: if a peer never arrives the kernel gives up after ~1 s, sets the error word and returns.
#include "gg_common.cuh"
#include "gg_launch.h"

namespace gg {

namespace {

constexpr int kP2PThreads = 256;
constexpr size_t kP2PDataOffset = 1024;

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double *p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

}  // namespace

__global__ void __launch_bounds__(kP2PThreads) shared_grad_allreduce_kernel(
    const float *__restrict__ dmus, const double *__restrict__ dsigma, int B, int K,
    const unsigned long long *__restrict__ peers, int world, int rank, double *__restrict__ out) {
  __shared__ unsigned epoch_s;
  __shared__ int fail_s;
  pdl_enter();
  const int L = 12 * K;
  unsigned char *mine = reinterpret_cast<unsigned char *>(peers[rank]);
  unsigned *hdr = reinterpret_cast<unsigned *>(mine);
  if (threadIdx.x == 0) {
    const unsigned e = hdr[0] + 1u;       // only this kernel touches the counter, launches are stream-ordered
    hdr[0] = e;
    epoch_s = e;
    fail_s = 0;
  }
  __syncthreads();
  const unsigned epoch = epoch_s;
  const int slot = (int)(epoch & 1u);

  // (1) local batch sum, fixed order; (2) the partial goes straight into every rank's slot [slot][rank]
  for (int c = threadIdx.x; c < L; c += kP2PThreads) {
    double acc = 0.0;   // fixed order; loads issued eight at a time so the B round trips overlap
    if (c < 3 * K) {
      for (int b0 = 0; b0 < B; b0 += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = b0 + u < B ? __ldg(dmus + (size_t)(b0 + u) * 3 * K + c) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += (double)v[u];
      }
    } else {
      const int cc = c - 3 * K;
      for (int b0 = 0; b0 < B; b0 += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = b0 + u < B ? __ldg(dsigma + (size_t)(b0 + u) * 9 * K + cc) : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
      }
    }
    for (int p = 0; p < world; ++p) {
      double *slots = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(peers[p]) + kP2PDataOffset);
      slots[((size_t)slot * world + rank) * L + c] = acc;
    }
  }
  __threadfence_system();
  __syncthreads();
  // (3) tell every peer this rank's slot is complete
  if ((int)threadIdx.x < world) {
    unsigned *pf = reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(peers[threadIdx.x])) + 16;
    st_release_sys(pf + rank, epoch);
  }
  // (4) wait for every peer (bounded)
  if ((int)threadIdx.x < world) {
    const unsigned *f = hdr + 16 + threadIdx.x;
    long long spins = 0;
    while ((int)(ld_acquire_sys(f) - epoch) < 0) {     // a peer may already be one call ahead: >= epoch is enough
      __nanosleep(64);
      if (++spins > 12000000ll) { fail_s = 1; break; }
    }
  }
  __syncthreads();
  if (fail_s) {
    if (threadIdx.x == 0) hdr[1] = epoch;
    return;
  }
  // (5) sum the slots in rank order
  const double *slots = reinterpret_cast<const double *>(mine + kP2PDataOffset) + (size_t)slot * world * L;
  for (int c = threadIdx.x; c < L; c += kP2PThreads) {
    double acc = 0.0;
    for (int r = 0; r < world; ++r) acc += ld_volatile_f64(slots + (size_t)r * L + c);
    out[c] = acc;
  }
}

size_t p2p_buffer_bytes(int K, int world) { return kP2PDataOffset + (size_t)2 * world * 12 * K * sizeof(double); }

cudaError_t launch_shared_grad_allreduce(const float *dmus, const double *dsigma, int B, int K,
                                         const unsigned long long *peers, int world, int rank, double *out,
                                         cudaStream_t st) {
  launch_pdl(shared_grad_allreduce_kernel, dim3(1), dim3(kP2PThreads), 0, st, dmus, dsigma, B, K, peers, world, rank, out);
  return cudaGetLastError();
}

}  // namespace gg

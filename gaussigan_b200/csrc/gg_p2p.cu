// gg_p2p.cu -- the training-time exchange over NVLink peer memory, split into POST and WAIT so that it overlaps.
//
// The canonical Gaussians are broadcast over the batch (mask/mask_gen_dynamic.py:701-702), so their gradient is the
// sum of the per-sample gradients over the local batch AND over the ranks.  The payload is 12 K
// numbers (1.5 KB at K = 16): pure latency.  NCCL's all_reduce costs ~40-60 us per eager call at this size.
//
//   post : one CTA of 1024 threads = 256 columns x 4 batch slices; every thread sums its slice with sixteen loads in
//          flight (one round trip for B <= 64), the slices are added in a fixed order in shared memory and the column
//          sums go straight into every rank's receive slot through NVLink/NVSwitch peer pointers (symmetric memory:
//          the same buffer layout on every rank); one system-scope fence, then this rank's flag is raised on every
//          peer with a system-scope release store.
//   wait : one CTA.  Waits for the peers' flags with acquire loads (bounded by a wall-clock timeout, default 60 s),
//          then adds the slots in rank order -- the same order on every rank, so the result is bit-identical
//          everywhere.  On a timeout it fills `out` with NaN and latches the call number in the header: the failure
//          reaches the gradients instead of leaving the previous call's sums there.
//
// `post` is launched right behind the projection adjoint; `wait` can follow immediately (round-1 behaviour, one call)
// or be enqueued later -- after the next forward pass, on a side stream, just before the optimiser step -- so that only
// the flag latency is ever exposed.
//
// Buffer layout (per rank, `gg_p2p_buffer_bytes`): u32 words [0] calls posted, [1] error latch (call number),
// [2] calls waited, [16 .. 16+W) arrival flags; at byte 1024 the receive
// slots double[2][W][12 K].  Two slot sets: a rank can be at most one call ahead of the slowest peer (it cannot finish
// waiting for call e+1 before every peer has posted it, i.e. finished waiting for call e), so writes of call e+1 never
// touch what a peer still reads for call e.  The counters live on the device and are advanced by the kernels
// themselves, which keeps the launches replayable from a CUDA graph.
#include <cstdlib>
#include "gg_common.cuh"
#include "gg_launch.h"

namespace gg {

namespace {

constexpr int kPostCols = 256, kPostSlices = 4, kPostThreads = kPostCols * kPostSlices;
constexpr int kWaitThreads = 256;
constexpr size_t kP2PDataOffset = 1024;
enum { H_POSTED = 0, H_ERROR = 1, H_WAITED = 2, H_ARRIVE = 3, H_FLAGS = 16 };

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
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

}  // namespace

__global__ void __launch_bounds__(kPostThreads) shared_grad_post_kernel(
    const float *__restrict__ dmus, const double *__restrict__ dsigma, int B, int K,
    const unsigned long long *__restrict__ peers, int world, int rank) {
  __shared__ double part[kPostSlices][kPostCols];
  pdl_enter();
  const int L = 12 * K;
  unsigned *hdr = reinterpret_cast<unsigned *>(peers[rank]);
  const unsigned epoch = hdr[H_POSTED] + 1u;     // posts are stream-ordered; advanced by thread 0 below
  const int slot = (int)(epoch & 1u);
  const int col = threadIdx.x & (kPostCols - 1), sl = threadIdx.x / kPostCols;
  // this thread's batch slice [b0, b1), fixed order; sixteen loads in flight (one round trip for B <= 64)
  const int per = (B + kPostSlices - 1) / kPostSlices;
  const int b0 = sl * per, b1 = min(B, b0 + per);
  for (int c0 = 0; c0 < L; c0 += kPostCols) {
    const int c = c0 + col;
    double acc = 0.0;
    if (c < L) {
      if (c < 3 * K) {
        for (int b = b0; b < b1; b += 16) {
          float v[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) v[u] = b + u < b1 ? __ldg(dmus + (size_t)(b + u) * 3 * K + c) : 0.f;
#pragma unroll
          for (int u = 0; u < 16; ++u) acc += (double)v[u];
        }
      } else {
        const int cc = c - 3 * K;
        for (int b = b0; b < b1; b += 16) {
          double v[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) v[u] = b + u < b1 ? __ldg(dsigma + (size_t)(b + u) * 9 * K + cc) : 0.0;
#pragma unroll
          for (int u = 0; u < 16; ++u) acc += v[u];
        }
      }
    }
    __syncthreads();                 // part[] of the previous column block has been consumed
    part[sl][col] = acc;
    __syncthreads();
    // slices in order, then the column sum goes straight into every rank's slot [slot][rank]
    if (sl == 0 && c < L) {
      double t = 0.0;
#pragma unroll
      for (int s = 0; s < kPostSlices; ++s) t += part[s][col];
      for (int p = 0; p < world; ++p) {
        double *slots = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(peers[p]) + kP2PDataOffset);
        slots[((size_t)slot * world + rank) * L + c] = t;
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  // tell every peer that this rank's slot is complete
  if ((int)threadIdx.x < world) {
    unsigned *pf = reinterpret_cast<unsigned *>(peers[threadIdx.x]) + H_FLAGS;
    st_release_sys(pf + rank, epoch);
  }
  if (threadIdx.x == 0) hdr[H_POSTED] = epoch;
}

__global__ void __launch_bounds__(kWaitThreads) shared_grad_wait_kernel(
    int K, const unsigned long long *__restrict__ peers, int world, int rank, double *__restrict__ out,
    unsigned long long timeout_ns) {
  __shared__ int fail_s;
  pdl_enter();
  const int L = 12 * K;
  unsigned char *mine = reinterpret_cast<unsigned char *>(peers[rank]);
  unsigned *hdr = reinterpret_cast<unsigned *>(mine);
  const unsigned epoch = hdr[H_WAITED] + 1u;     // waits are stream-ordered; advanced by thread 0 below
  const int slot = (int)(epoch & 1u);
  if (threadIdx.x == 0) fail_s = 0;
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const unsigned *f = hdr + H_FLAGS + threadIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    while ((int)(ld_acquire_sys(f) - epoch) < 0) {     // a peer may already be one call ahead: >= epoch is enough
      __nanosleep(64);
      if (globaltimer_ns() - t0 > timeout_ns) { fail_s = 1; break; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) hdr[H_WAITED] = epoch;
  if (fail_s) {
    // make the failure visible in the data: NaN gradients, and the call number latched for the host to read
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (int c = threadIdx.x; c < L; c += kWaitThreads) out[c] = nan;
    if (threadIdx.x == 0 && hdr[H_ERROR] == 0u) hdr[H_ERROR] = epoch;
    return;
  }
  const double *slots = reinterpret_cast<const double *>(mine + kP2PDataOffset) + (size_t)slot * world * L;
  for (int c = threadIdx.x; c < L; c += kWaitThreads) {
    double acc = 0.0;
    for (int r = 0; r < world; ++r) acc += ld_volatile_f64(slots + (size_t)r * L + c);
    out[c] = acc;
  }
}

size_t p2p_buffer_bytes(int K, int world) { return kP2PDataOffset + (size_t)2 * world * 12 * K * sizeof(double); }

static unsigned long long p2p_timeout_ns() {
  static const unsigned long long v = [] {
    const char *e = std::getenv("GG_P2P_TIMEOUT_MS");
    const long long ms = e ? std::atoll(e) : 60000ll;
    return (unsigned long long)(ms > 0 ? ms : 60000ll) * 1000000ull;
  }();
  return v;
}

cudaError_t launch_shared_grad_post(const float *dmus, const double *dsigma, int B, int K,
                                    const unsigned long long *peers, int world, int rank, cudaStream_t st) {
  const int L = 12 * K;
  (void)L;
  launch_pdl(shared_grad_post_kernel, dim3(1), dim3(kPostThreads), 0, st, dmus, dsigma, B, K, peers, world, rank);
  return cudaGetLastError();
}

cudaError_t launch_shared_grad_wait(int K, const unsigned long long *peers, int world, int rank, double *out,
                                    cudaStream_t st) {
  launch_pdl(shared_grad_wait_kernel, dim3(1), dim3(kWaitThreads), 0, st, K, peers, world, rank, out, p2p_timeout_ns());
  return cudaGetLastError();
}

}  // namespace gg

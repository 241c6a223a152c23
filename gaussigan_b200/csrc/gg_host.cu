// gg_host.cu -- the host-buffer entry point: one call = H2D copies + the whole fwd/loss/bwd chain +
// D2H copies, all asynchronous on the caller's stream.  This is the form a host-language binding of the
// reference would use (the reference feeds numpy arrays through session.run, mask/infer_masks.py:435-456).
#include <cstdlib>
#include <cstring>
#include "gg_common.cuh"
#include "gg_launch.h"

namespace gg {

struct StepLayout {
  size_t mus, sigma, rot, target, params, mask, gmask, partials, loss, dmus, dsigma, drot, total;
  int T;
  int nloss;   // loss partials written by the step (fused path: one per CTA; unfused: GG_LOSS_PARTIALS)
};

static size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

constexpr unsigned kStripsMagic = 0x54534747u;   // "GGST"

static size_t strips_payload_offset(int B, int n) { return (16 + (size_t)B * n * 4 + 7) & ~size_t(7); }
static size_t strips_bound(int B, int n) {
  const size_t strips = (size_t)(n + 7) / 8, W = (strips + 31) / 32;
  return strips_payload_offset(B, n) + (size_t)B * n * (8 * W + 8 * strips);
}

static size_t target_bytes(int B, int n, int target_kind) {
  if (target_kind == GG_TARGET_STRIPS) return strips_bound(B, n);        // capacity; the bytes in use are in the header
  if (target_kind == GG_TARGET_BITS) return (size_t)B * n * ((n + 7) / 8);
  return (size_t)B * n * n * (target_kind == GG_TARGET_U8 ? 1 : 4);
}

StepLayout step_layout(int B, int K, int n, int target_kind) {
  StepLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
  const int sizes[1] = {n}, zero[1] = {0}, one[1] = {1};
  L.T = plan_bwd_tiles(K, 1, sizes, zero, one, GG_LAYOUT_NHWC);
  L.mus = take((size_t)B * K * 3 * 4);
  L.sigma = take((size_t)B * K * 9 * 8);
  L.rot = take((size_t)B * 3 * 4);
  L.target = take(target_bytes(B, n, target_kind));
  L.params = take((size_t)B * K * GG_PARAM_STRIDE * 4);
  L.mask = take((size_t)B * n * n * 4);
  L.gmask = take((size_t)B * n * n * 4);
  L.partials = take((size_t)B * L.T * K * GG_MOMENTS * 4);
  L.nloss = K <= 64 ? B * L.T : GG_LOSS_PARTIALS;
  L.loss = take((size_t)L.nloss * 4);
  L.dmus = take((size_t)B * K * 3 * 4);
  L.dsigma = take((size_t)B * K * 9 * 8);
  L.drot = take((size_t)B * 3 * 4);
  L.total = off;
  return L;
}

// device-usable alias of a pinned host range [host, host + bytes) (UVA), or NULL when any of it is pageable /
// unknown: both ends must resolve to pinned memory with consecutive device addresses (a partly registered buffer
// would otherwise fault in the kernel)
template <typename T>
static T *mapped_or_null(T *host, size_t bytes) {
  if (bytes == 0) return nullptr;
  cudaPointerAttributes at, at_end;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) return nullptr;
  const char *last = reinterpret_cast<const char *>(host) + bytes - 1;
  if (cudaPointerGetAttributes(&at_end, last) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (at_end.type != cudaMemoryTypeHost || at_end.devicePointer == nullptr) return nullptr;
  if (reinterpret_cast<const char *>(at_end.devicePointer) - reinterpret_cast<const char *>(at.devicePointer) !=
      (ptrdiff_t)(bytes - 1))
    return nullptr;
  return reinterpret_cast<T *>(const_cast<void *>(static_cast<const void *>(at.devicePointer)));
}

// GG_ZEROCOPY: which of the small arrays the projection kernels access in place in pinned host memory --
// "out" the gradients (default), "in" the Gaussians and the camera, "1" both, "0" none.  Measured at batch 64, K = 16,
// 256 x 256, six steps in flight (tools/e2e_host_cost.py): out 34.8 us per step, none 35.1, in 42.0, both 44.0 -- reads
// over PCIe sit on the projection's dependent chain, posted writes do not.
static int zero_copy_mode() {   // read on every call (a getenv): tests and callers may switch it between steps
  const char *v = std::getenv("GG_ZEROCOPY");
  if (!v || v[0] == 'o') return 2;
  if (v[0] == '0') return 0;
  if (v[0] == 'i') return 1;
  return 3;
}

}  // namespace gg

using namespace gg::api;
using gg::mapped_or_null;
using gg::zero_copy_mode;

extern "C" {

size_t gg_target_strips_bound(int B, int n) {
  if (B <= 0 || n < 1) return 0;
  return gg::strips_bound(B, n);
}

long long gg_encode_target_strips(const unsigned char *image_host, int B, int n, void *out_host, size_t capacity) {
  if (B <= 0 || n < 1 || n > 16384) return fail(GG_E_SHAPE, "gg_encode_target_strips: bad B=%d n=%d", B, n);
  if (!image_host || !out_host) return fail(GG_E_NULL, "gg_encode_target_strips: NULL pointer");
  if (reinterpret_cast<uintptr_t>(out_host) & 7) return fail(GG_E_ALIGN, "gg_encode_target_strips: out must be 8-byte aligned");
  const int strips = (n + 7) / 8, W = (strips + 31) / 32;
  const size_t pay0 = gg::strips_payload_offset(B, n);
  if (capacity < pay0) return fail(GG_E_WORKSPACE, "gg_encode_target_strips: capacity %zu too small", capacity);
  unsigned char *out = reinterpret_cast<unsigned char *>(out_host);
  unsigned *hdr = reinterpret_cast<unsigned *>(out);
  unsigned *offs = hdr + 4;
  size_t used = 0;                                      // bytes of payload written
  for (size_t r = 0; r < (size_t)B * n; ++r) {
    const unsigned char *src = image_host + r * n;
    if (pay0 + used + 8 * (size_t)W + 8 * (size_t)strips > capacity)
      return fail(GG_E_WORKSPACE, "gg_encode_target_strips: capacity %zu too small (use gg_target_strips_bound)", capacity);
    unsigned char *rec = out + pay0 + used;
    unsigned *M = reinterpret_cast<unsigned *>(rec), *V = M + W;
    for (int w = 0; w < W; ++w) M[w] = V[w] = 0u;
    unsigned char *lit = rec + 8 * W;
    int nlit = 0;
    for (int s_ = 0; s_ < strips; ++s_) {
      const int c0 = s_ * 8, cnt = n - c0 < 8 ? n - c0 : 8;
      unsigned long long x = 0;
      std::memcpy(&x, src + c0, (size_t)cnt);
      const unsigned long long ones = cnt == 8 ? ~0ull : ((1ull << (8 * cnt)) - 1ull);
      if (x == 0ull) continue;                            // flat 0: M = 0, V = 0
      if (x == ones) { V[s_ >> 5] |= 1u << (s_ & 31); continue; }   // flat 255
      M[s_ >> 5] |= 1u << (s_ & 31);
      std::memcpy(lit + 8 * (size_t)nlit, &x, 8);
      ++nlit;
    }
    offs[r] = (unsigned)used;
    used += 8 * (size_t)W + 8 * (size_t)nlit;
    if (used > 0xffffffffull) return fail(GG_E_SHAPE, "gg_encode_target_strips: payload exceeds 4 GiB");
  }
  hdr[0] = gg::kStripsMagic; hdr[1] = (unsigned)n; hdr[2] = (unsigned)B; hdr[3] = (unsigned)(pay0 + used);
  return (long long)(pay0 + used);
}

size_t gg_silhouette_step_workspace_bytes(int B, int K, int n, int target_kind) {
  if (B <= 0 || K <= 0 || n < 1) return 0;
  return gg::step_layout(B, K, n, target_kind).total;
}

int gg_silhouette_step_loss_count(int B, int K, int n) {
  if (B <= 0 || K <= 0 || n < 1) return fail(GG_E_SHAPE, "bad B/K/n");
  return gg::step_layout(B, K, n, GG_TARGET_U8).nloss;
}

int gg_silhouette_step_host(const float *mus_host, const double *sigma_host, const float *rotx_host,
                            const float *roty_host, const float *rotz_host, const void *target_host, int target_kind,
                            double tx, double ty, double tz, double focal, int B, int K, int n,
                            float *loss_partials_host, int loss_partials_capacity, float *dmus_host,
                            double *dsigma_host, float *drot_host, float *mask_host, void *workspace,
                            size_t workspace_bytes, int device, void *stream) {
  if (int r = check_bk(B, K)) return r;
  const int sizes[1] = {n};
  if (int r = check_scales(K, 1, sizes)) return r;
  if (B == 0) return GG_OK;
  if (target_kind != GG_TARGET_U8 && target_kind != GG_TARGET_F32 && target_kind != GG_TARGET_BITS &&
      target_kind != GG_TARGET_STRIPS)
    return fail(GG_E_SHAPE, "unknown target_kind %d", target_kind);
  if ((target_kind == GG_TARGET_BITS || target_kind == GG_TARGET_STRIPS) && (K > 64 || !gg::fast_mask()))
    return fail(GG_E_SHAPE, "GG_TARGET_BITS / GG_TARGET_STRIPS: K <= 64 and the forward-differenced kernels only");
  if (!mus_host || !sigma_host || !rotx_host || !roty_host || !rotz_host || !target_host || !loss_partials_host ||
      !dmus_host || !dsigma_host || !drot_host || !workspace)
    return fail(GG_E_NULL, "gg_silhouette_step_host: NULL pointer");
  const gg::StepLayout L = gg::step_layout(B, K, n, target_kind);
  if (workspace_bytes < L.total)
    return fail(GG_E_WORKSPACE, "gg_silhouette_step_host: workspace %zu < required %zu bytes", workspace_bytes, L.total);
  if (loss_partials_capacity < L.nloss)
    return fail(GG_E_WORKSPACE, "gg_silhouette_step_host: loss_partials_host holds %d floats, this configuration writes %d "
                "(gg_silhouette_step_loss_count; it depends on the arithmetic mode)", loss_partials_capacity, L.nloss);
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(GG_E_ALIGN, "workspace must be 256-byte aligned");
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaStream_t st = (cudaStream_t)stream;
  char *w = reinterpret_cast<char *>(workspace);
  float *d_mus = reinterpret_cast<float *>(w + L.mus);
  double *d_sigma = reinterpret_cast<double *>(w + L.sigma);
  float *d_rot = reinterpret_cast<float *>(w + L.rot);
  void *d_target = w + L.target;
  float *d_params = reinterpret_cast<float *>(w + L.params);
  float *d_mask = reinterpret_cast<float *>(w + L.mask);
  float *d_gmask = reinterpret_cast<float *>(w + L.gmask);
  float *d_part = reinterpret_cast<float *>(w + L.partials);
  float *d_loss = reinterpret_cast<float *>(w + L.loss);
  float *d_dmus = reinterpret_cast<float *>(w + L.dmus);
  double *d_dsig = reinterpret_cast<double *>(w + L.dsigma);
  float *d_drot = reinterpret_cast<float *>(w + L.drot);
#define GG_TRY(call, what)                                  \
  do {                                                      \
    cudaError_t e__ = (call);                               \
    if (e__ != cudaSuccess) return cuda_fail(e__, what);    \
  } while (0)
  const size_t px = (size_t)B * n * n;
  // The Gaussians, the camera and the gradients are ~90 KB per step in nine arrays.  Pinned buffers (cudaHostAlloc /
  // cudaHostRegister, device-accessible through UVA) can be accessed by the two projection kernels in place over PCIe
  // (zero_copy_mode above); everything else, and every pageable buffer, takes the copy path.
  const size_t nbk = (size_t)B * K;
  const float *k_mus = nullptr, *k_rx = nullptr, *k_ry = nullptr, *k_rz = nullptr;
  const double *k_sigma = nullptr;
  if (zero_copy_mode() & 1) {
    k_mus = mapped_or_null(mus_host, nbk * 3 * 4);
    k_sigma = mapped_or_null(sigma_host, nbk * 9 * 8);
    k_rx = mapped_or_null(rotx_host, (size_t)B * 4);
    k_ry = mapped_or_null(roty_host, (size_t)B * 4);
    k_rz = mapped_or_null(rotz_host, (size_t)B * 4);
  }
  const bool in_place_in = k_mus && k_sigma && k_rx && k_ry && k_rz;
  // one block on the host with the workspace's own spacing (mus, sigma and rotx each at the next multiple of 256 bytes,
  // roty and rotz right behind rotx -- gaussigan_b200.host.HostSilhouetteStep.make_host_inputs builds it): one copy
  const char *hb = reinterpret_cast<const char *>(mus_host);
  const bool packed_in = reinterpret_cast<const char *>(sigma_host) == hb + (L.sigma - L.mus) &&
                         reinterpret_cast<const char *>(rotx_host) == hb + (L.rot - L.mus) &&
                         roty_host == rotx_host + B && rotz_host == roty_host + B;
  if (!in_place_in && packed_in) {
    GG_TRY(cudaMemcpyAsync(d_mus, mus_host, (L.rot - L.mus) + (size_t)3 * B * 4, cudaMemcpyHostToDevice, st), "H2D inputs");
    k_mus = d_mus; k_sigma = d_sigma; k_rx = d_rot; k_ry = d_rot + B; k_rz = d_rot + 2 * (size_t)B;
  } else if (!in_place_in) {
    GG_TRY(cudaMemcpyAsync(d_mus, mus_host, (size_t)B * K * 3 * 4, cudaMemcpyHostToDevice, st), "H2D mus");
    GG_TRY(cudaMemcpyAsync(d_sigma, sigma_host, (size_t)B * K * 9 * 8, cudaMemcpyHostToDevice, st), "H2D sigma");
    GG_TRY(cudaMemcpyAsync(d_rot, rotx_host, (size_t)B * 4, cudaMemcpyHostToDevice, st), "H2D rotx");
    GG_TRY(cudaMemcpyAsync(d_rot + B, roty_host, (size_t)B * 4, cudaMemcpyHostToDevice, st), "H2D roty");
    GG_TRY(cudaMemcpyAsync(d_rot + 2 * (size_t)B, rotz_host, (size_t)B * 4, cudaMemcpyHostToDevice, st), "H2D rotz");
    k_mus = d_mus; k_sigma = d_sigma; k_rx = d_rot; k_ry = d_rot + B; k_rz = d_rot + 2 * (size_t)B;
  }
  // Experimental (GG_ZEROCOPY_TARGET=1): the splat kernel reads the target mask from pinned host memory as well
  const void *k_target = d_target;
  {
    static const bool zc_target = [] { const char *v = std::getenv("GG_ZEROCOPY_TARGET"); return v && v[0] == '1'; }();
    const void *m = zc_target ? mapped_or_null(target_host, gg::target_bytes(B, n, target_kind)) : nullptr;
    if (m) k_target = m;
    else {
      size_t tbytes = gg::target_bytes(B, n, target_kind);
      if (target_kind == GG_TARGET_STRIPS) {       // the packed size is data dependent: read it from the buffer's header
        const unsigned *h = reinterpret_cast<const unsigned *>(target_host);
        if ((reinterpret_cast<uintptr_t>(target_host) & 7) || h[0] != gg::kStripsMagic || h[1] != (unsigned)n ||
            h[2] != (unsigned)B || h[3] > tbytes)
          return fail(GG_E_SHAPE, "gg_silhouette_step_host: target_host is not a GG_TARGET_STRIPS buffer for B=%d n=%d", B, n);
        tbytes = h[3];
      }
      GG_TRY(cudaMemcpyAsync(d_target, target_host, tbytes, cudaMemcpyHostToDevice, st), "H2D target");
    }
  }
  GG_TRY(gg::launch_project_fwd(k_mus, k_sigma, 1, k_rx, k_ry, k_rz, tx, ty, tz, focal, B, K,
                                nullptr, nullptr, d_params, st), "project_fwd");
  if (K <= 64) {
    GG_TRY(gg::launch_silhouette_loss_fused(d_params, k_target, target_kind, B, K, n, mask_host ? d_mask : nullptr,
                                            d_part, d_loss, L.T, st), "silhouette_loss_fused");
  } else {
    float *masks[1] = {d_mask};
    GG_TRY(gg::launch_splat_fwd(d_params, B, K, 1, sizes, nullptr, masks, GG_LAYOUT_NHWC, st), "splat_fwd");
    GG_TRY(gg::launch_mask_l1_loss_bwd(d_mask, k_target, target_kind, B, n, d_gmask, d_loss, GG_LOSS_PARTIALS, st),
           "mask_l1_loss_bwd");
    const float *gmasks[1] = {d_gmask};
    GG_TRY(gg::launch_splat_bwd(d_params, B, K, 1, sizes, nullptr, gmasks, GG_LAYOUT_NHWC, d_part, L.T, st), "splat_bwd");
  }
  float *k_dmus = nullptr, *k_drot = nullptr;
  double *k_dsig = nullptr;
  if (zero_copy_mode() & 2) {
    k_dmus = mapped_or_null(dmus_host, nbk * 3 * 4);
    k_drot = mapped_or_null(drot_host, (size_t)B * 3 * 4);
    k_dsig = mapped_or_null(dsigma_host, nbk * 9 * 8);
  }
  const bool in_place_out = k_dmus && k_dsig && k_drot;
  GG_TRY(gg::launch_project_bwd(k_mus, k_sigma, 1, k_rx, k_ry, k_rz, tx, ty, tz, focal, B, K, d_part, L.T, nullptr,
                                nullptr, in_place_out ? k_dmus : d_dmus, in_place_out ? k_dsig : d_dsig,
                                in_place_out ? k_drot : d_drot, st), "project_bwd");
  GG_TRY(cudaMemcpyAsync(loss_partials_host, d_loss, (size_t)L.nloss * 4, cudaMemcpyDeviceToHost, st), "D2H loss");
  if (!in_place_out) {
    GG_TRY(cudaMemcpyAsync(dmus_host, d_dmus, (size_t)B * K * 3 * 4, cudaMemcpyDeviceToHost, st), "D2H dmus");
    GG_TRY(cudaMemcpyAsync(dsigma_host, d_dsig, (size_t)B * K * 9 * 8, cudaMemcpyDeviceToHost, st), "D2H dsigma");
    GG_TRY(cudaMemcpyAsync(drot_host, d_drot, (size_t)B * 3 * 4, cudaMemcpyDeviceToHost, st), "D2H drot");
  }
  if (mask_host) GG_TRY(cudaMemcpyAsync(mask_host, d_mask, px * 4, cudaMemcpyDeviceToHost, st), "D2H mask");
#undef GG_TRY
  return GG_OK;
}

}  // extern "C"

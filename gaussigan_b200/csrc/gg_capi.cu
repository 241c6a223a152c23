// gg_capi.cu -- the extern "C" surface declared in include/gaussigan.h: argument validation,
// device selection, error reporting.  No compute here.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "gg_common.cuh"
#include "gg_launch.h"

namespace gg {

// process-wide switch; an atomic, so concurrent readers/writers are well defined (see gaussigan.h for the thread rules)
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char *e = std::getenv("GG_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
    int expected = -1;
    g_pdl.compare_exchange_strong(expected, v, std::memory_order_relaxed);
    v = g_pdl.load(std::memory_order_relaxed);
  }
  return v != 0;
}
int set_pdl(int enable) {
  const int prev = pdl_enabled() ? 1 : 0;
  g_pdl.store(enable ? 1 : 0, std::memory_order_relaxed);
  return prev;
}

namespace api {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char *where) {
  return fail(GG_E_CUDA, "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Makes `device` current for the duration of one entry point and restores the caller's device on exit: in a
// single-process multi-GPU program the thread's current device (which PyTorch allocations and stream lookups follow)
// must not change under the caller's feet.
DeviceGuard::DeviceGuard(int device) : prev(-1), status(GG_OK) {
  int current = -1;
  if (cudaGetDevice(&current) != cudaSuccess) { cudaGetLastError(); current = -1; }
  if (current != device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { status = cuda_fail(e, "cudaSetDevice"); return; }
    prev = current;
  }
}
DeviceGuard::~DeviceGuard() {
  if (prev >= 0) cudaSetDevice(prev);
}

int check_bk(int B, int K) {
  if (B < 0 || K <= 0) return fail(GG_E_SHAPE, "B must be >= 0 and K > 0 (got B=%d, K=%d)", B, K);
  if ((long long)B * K > (1ll << 30)) return fail(GG_E_SHAPE, "B*K too large (%lld)", (long long)B * K);
  return GG_OK;
}

int check_scales(int K, int nscales, const int *sizes) {
  if (nscales < 1 || nscales > GG_MAX_SCALES)
    return fail(GG_E_SHAPE, "nscales must be in 1..%d (got %d)", GG_MAX_SCALES, nscales);
  if (!sizes) return fail(GG_E_NULL, "sizes_host is NULL");
  for (int i = 0; i < nscales; ++i)
    if (sizes[i] < 1 || sizes[i] > 16384) return fail(GG_E_SHAPE, "scale %d: n=%d out of range 1..16384", i, sizes[i]);
  if (K > 1536) return fail(GG_E_SHAPE, "K=%d exceeds the staged-parameter shared-memory budget (1536)", K);
  return GG_OK;
}

}  // namespace api
}  // namespace gg

using namespace gg::api;

extern "C" {

int gg_version(void) { return GG_VERSION; }

const char *gg_last_error(void) { return g_err; }

int gg_set_fast_mask(int enable) { return gg::set_fast_mask(enable); }

int gg_set_pdl(int enable) { return gg::set_pdl(enable); }

int gg_project_fwd(const float *mus, const void *sigma, int sigma_is_f64, const float *rotx, const float *roty,
                   const float *rotz, double tx, double ty, double tz, double focal, int B, int K, double *mu2d,
                   double *sigma2d, float *params, int device, void *stream) {
  if (int r = check_bk(B, K)) return r;
  if (B == 0) return GG_OK;
  if (!mus || !sigma || !rotx || !roty || !rotz || !params) return fail(GG_E_NULL, "gg_project_fwd: NULL input");
  if (!aligned16(params)) return fail(GG_E_ALIGN, "gg_project_fwd: params must be 16-byte aligned");
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_project_fwd(mus, sigma, sigma_is_f64, rotx, roty, rotz, tx, ty, tz, focal, B, K, mu2d,
                                         sigma2d, params, (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_project_fwd");
}

int gg_project_bwd(const float *mus, const void *sigma, int sigma_is_f64, const float *rotx, const float *roty,
                   const float *rotz, double tx, double ty, double tz, double focal, int B, int K,
                   const float *partials, int T, const double *gmu2d, const double *gsigma2d, float *dmus,
                   double *dsigma, float *drot, int device, void *stream) {
  if (int r = check_bk(B, K)) return r;
  if (B == 0) return GG_OK;
  if (!mus || !sigma || !rotx || !roty || !rotz) return fail(GG_E_NULL, "gg_project_bwd: NULL input");
  if (T < 0 || (T > 0 && !partials)) return fail(GG_E_WORKSPACE, "gg_project_bwd: T=%d but partials is NULL", T);
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_project_bwd(mus, sigma, sigma_is_f64, rotx, roty, rotz, tx, ty, tz, focal, B, K,
                                         partials, T, gmu2d, gsigma2d, dmus, dsigma, drot, (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_project_bwd");
}

int gg_project_fwd_views(const float *mus, const void *sigma, int sigma_is_f64, const float *rotx, const float *roty,
                         const float *rotz, double tx, double ty, double tz, double focal, int Bg, int V, int K,
                         double *mu2d, double *sigma2d, float *params, int device, void *stream) {
  if (V < 1 || V > 64) return fail(GG_E_SHAPE, "V must be in 1..64 (got %d)", V);
  if (int r = check_bk(Bg * V, K)) return r;
  if (Bg == 0) return GG_OK;
  if (!mus || !sigma || !rotx || !roty || !rotz || !params) return fail(GG_E_NULL, "gg_project_fwd_views: NULL input");
  if (!aligned16(params)) return fail(GG_E_ALIGN, "gg_project_fwd_views: params must be 16-byte aligned");
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_project_fwd(mus, sigma, sigma_is_f64, rotx, roty, rotz, tx, ty, tz, focal, Bg * V, K, mu2d,
                                         sigma2d, params, (cudaStream_t)stream, V);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_project_fwd_views");
}

int gg_project_bwd_views(const float *mus, const void *sigma, int sigma_is_f64, const float *rotx, const float *roty,
                         const float *rotz, double tx, double ty, double tz, double focal, int Bg, int V, int K,
                         const float *partials, int T, const double *gmu2d, const double *gsigma2d, float *dmus,
                         double *dsigma, float *drot, int device, void *stream) {
  if (V < 1 || V > 64) return fail(GG_E_SHAPE, "V must be in 1..64 (got %d)", V);
  if (int r = check_bk(Bg * V, K)) return r;
  if (Bg == 0) return GG_OK;
  if (!mus || !sigma || !rotx || !roty || !rotz) return fail(GG_E_NULL, "gg_project_bwd_views: NULL input");
  if (T < 0 || (T > 0 && !partials)) return fail(GG_E_WORKSPACE, "gg_project_bwd_views: T=%d but partials is NULL", T);
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_project_bwd(mus, sigma, sigma_is_f64, rotx, roty, rotz, tx, ty, tz, focal, Bg * V, K,
                                         partials, T, gmu2d, gsigma2d, dmus, dsigma, drot, (cudaStream_t)stream, V);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_project_bwd_views");
}

int gg_conic_fwd(const void *mu2d, const void *sigma2d, int is_f64, int B, int K, float *params, int device,
                 void *stream) {
  if (int r = check_bk(B, K)) return r;
  if (B == 0) return GG_OK;
  if (!mu2d || !sigma2d || !params) return fail(GG_E_NULL, "gg_conic_fwd: NULL input");
  if (!aligned16(params)) return fail(GG_E_ALIGN, "gg_conic_fwd: params must be 16-byte aligned");
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_conic_fwd(mu2d, sigma2d, is_f64, B * K, params, (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_conic_fwd");
}

int gg_conic_bwd(const void *mu2d, const void *sigma2d, int is_f64, int B, int K, const float *partials, int T,
                 double *dmu2d, double *dsigma2d, int device, void *stream) {
  if (int r = check_bk(B, K)) return r;
  if (B == 0) return GG_OK;
  if (!mu2d || !sigma2d) return fail(GG_E_NULL, "gg_conic_bwd: NULL input");
  if (T <= 0 || !partials) return fail(GG_E_WORKSPACE, "gg_conic_bwd: needs T > 0 partials (T=%d)", T);
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_conic_bwd(mu2d, sigma2d, is_f64, B, K, partials, T, dmu2d, dsigma2d,
                                       (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_conic_bwd");
}

int gg_splat_fwd(const float *params, int B, int K, int nscales, const int *sizes_host, float *const *maps_host,
                 float *const *masks_host, int layout, int device, void *stream) {
  if (int r = check_bk(B, K)) return r;
  if (int r = check_scales(K, nscales, sizes_host)) return r;
  if (layout != GG_LAYOUT_NHWC && layout != GG_LAYOUT_NCHW) return fail(GG_E_SHAPE, "unknown layout %d", layout);
  if (B == 0) return GG_OK;
  if (!params) return fail(GG_E_NULL, "gg_splat_fwd: params is NULL");
  if (!aligned16(params)) return fail(GG_E_ALIGN, "gg_splat_fwd: params must be 16-byte aligned");
  bool any = false;
  for (int i = 0; i < nscales; ++i) {
    float *m = maps_host ? maps_host[i] : nullptr;
    float *k = masks_host ? masks_host[i] : nullptr;
    if (m || k) any = true;
    if ((m && !aligned16(m)) || (k && !aligned16(k)))
      return fail(GG_E_ALIGN, "gg_splat_fwd: scale %d output not 16-byte aligned", i);
    if ((long long)B * sizes_host[i] * sizes_host[i] > (1ll << 40)) return fail(GG_E_SHAPE, "scale %d too large", i);
  }
  if (!any) return fail(GG_E_NULL, "gg_splat_fwd: no output requested");
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_splat_fwd(params, B, K, nscales, sizes_host, maps_host, masks_host, layout,
                                       (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_splat_fwd");
}

static int check_concat(int K, int nscales, const int *cstrides, const int *coffsets, const char *who) {
  if (!cstrides || !coffsets) return fail(GG_E_NULL, "%s: cstrides_host / coffsets_host is NULL", who);
  for (int i = 0; i < nscales; ++i)
    if (coffsets[i] < 0 || cstrides[i] < K || coffsets[i] + K > cstrides[i])
      return fail(GG_E_SHAPE, "%s: scale %d: channels [%d, %d) do not fit a buffer of %d channels", who, i, coffsets[i],
                  coffsets[i] + K, cstrides[i]);
  return GG_OK;
}

int gg_splat_fwd_concat(const float *params, int B, int K, int nscales, const int *sizes_host, float *const *bufs_host,
                        const int *cstrides_host, const int *coffsets_host, float *const *masks_host, int device,
                        void *stream) {
  if (int r = check_bk(B, K)) return r;
  if (int r = check_scales(K, nscales, sizes_host)) return r;
  if (int r = check_concat(K, nscales, cstrides_host, coffsets_host, "gg_splat_fwd_concat")) return r;
  if (B == 0) return GG_OK;
  if (!params || !bufs_host) return fail(GG_E_NULL, "gg_splat_fwd_concat: params / bufs_host is NULL");
  if (!aligned16(params)) return fail(GG_E_ALIGN, "gg_splat_fwd_concat: params must be 16-byte aligned");
  bool any = false;
  for (int i = 0; i < nscales; ++i) {
    float *m = bufs_host[i];
    float *k = masks_host ? masks_host[i] : nullptr;
    any = any || m || k;
    if ((m && !aligned16(m)) || (k && !aligned16(k)))
      return fail(GG_E_ALIGN, "gg_splat_fwd_concat: scale %d buffer not 16-byte aligned", i);
  }
  if (!any) return fail(GG_E_NULL, "gg_splat_fwd_concat: no output requested");
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_splat_fwd(params, B, K, nscales, sizes_host, bufs_host, masks_host, GG_LAYOUT_NHWC,
                                       (cudaStream_t)stream, cstrides_host, coffsets_host);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_splat_fwd_concat");
}

int gg_splat_bwd_tiles_concat(int K, int nscales, const int *sizes_host, const int *has_gmaps_host,
                              const int *has_gmasks_host, const int *cstrides_host, const int *coffsets_host) {
  if (K <= 0) return fail(GG_E_SHAPE, "K must be > 0");
  if (int r = check_scales(K, nscales, sizes_host)) return r;
  if (int r = check_concat(K, nscales, cstrides_host, coffsets_host, "gg_splat_bwd_tiles_concat")) return r;
  return gg::plan_bwd_tiles(K, nscales, sizes_host, has_gmaps_host, has_gmasks_host, GG_LAYOUT_NHWC, cstrides_host,
                            coffsets_host);
}

int gg_splat_bwd_concat(const float *params, int B, int K, int nscales, const int *sizes_host,
                        const float *const *gbufs_host, const int *cstrides_host, const int *coffsets_host,
                        const float *const *gmasks_host, float *partials, int T, int device, void *stream) {
  if (int r = check_bk(B, K)) return r;
  if (int r = check_scales(K, nscales, sizes_host)) return r;
  if (int r = check_concat(K, nscales, cstrides_host, coffsets_host, "gg_splat_bwd_concat")) return r;
  if (B == 0) return GG_OK;
  if (!params || !partials) return fail(GG_E_NULL, "gg_splat_bwd_concat: params/partials is NULL");
  if (!aligned16(params)) return fail(GG_E_ALIGN, "gg_splat_bwd_concat: params must be 16-byte aligned");
  int hm[GG_MAX_SCALES], hk[GG_MAX_SCALES];
  for (int i = 0; i < nscales; ++i) {
    const float *m = gbufs_host ? gbufs_host[i] : nullptr;
    const float *k = gmasks_host ? gmasks_host[i] : nullptr;
    hm[i] = m != nullptr;
    hk[i] = k != nullptr;
    if ((m && !aligned16(m)) || (k && !aligned16(k)))
      return fail(GG_E_ALIGN, "gg_splat_bwd_concat: scale %d gradient buffer not 16-byte aligned", i);
  }
  const int need = gg::plan_bwd_tiles(K, nscales, sizes_host, hm, hk, GG_LAYOUT_NHWC, cstrides_host, coffsets_host);
  if (need != T || T <= 0)
    return fail(GG_E_WORKSPACE, "gg_splat_bwd_concat: T=%d but this configuration writes %d partial tiles", T, need);
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_splat_bwd(params, B, K, nscales, sizes_host, gbufs_host, gmasks_host, GG_LAYOUT_NHWC, partials,
                                       T, (cudaStream_t)stream, cstrides_host, coffsets_host);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_splat_bwd_concat");
}

int gg_mask_l1_loss_bwd(const float *mask, const void *target, int target_kind, int B, int n, float *gmask,
                        float *loss_partials, int device, void *stream) {
  if (B < 0 || n < 1) return fail(GG_E_SHAPE, "gg_mask_l1_loss_bwd: bad B=%d n=%d", B, n);
  if (B == 0) return GG_OK;
  if (target_kind != GG_TARGET_U8 && target_kind != GG_TARGET_F32) return fail(GG_E_SHAPE, "unknown target_kind %d", target_kind);
  if (!mask || !target || !gmask || !loss_partials) return fail(GG_E_NULL, "gg_mask_l1_loss_bwd: NULL pointer");
  if (!aligned16(mask) || !aligned16(gmask) || (reinterpret_cast<uintptr_t>(target) & 3))
    return fail(GG_E_ALIGN, "gg_mask_l1_loss_bwd: mask/gmask need 16-byte, target 4-byte alignment");
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_mask_l1_loss_bwd(mask, target, target_kind, B, n, gmask, loss_partials,
                                              GG_LOSS_PARTIALS, (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_mask_l1_loss_bwd");
}

int gg_silhouette_loss_fused(const float *params, const void *target, int target_kind, int B, int K, int n,
                             float *mask_out, float *partials, int T, float *loss_partials, int device,
                             void *stream) {
  if (int r = check_bk(B, K)) return r;
  const int sizes[1] = {n};
  if (int r = check_scales(K, 1, sizes)) return r;
  if (K > 64) return fail(GG_E_SHAPE, "gg_silhouette_loss_fused: K=%d > 64 (use the unfused entry points)", K);
  if (target_kind != GG_TARGET_U8 && target_kind != GG_TARGET_F32 && target_kind != GG_TARGET_BITS &&
      target_kind != GG_TARGET_STRIPS)
    return fail(GG_E_SHAPE, "unknown target_kind %d", target_kind);
  if ((target_kind == GG_TARGET_BITS || target_kind == GG_TARGET_STRIPS) && !gg::fast_mask())
    return fail(GG_E_SHAPE, "GG_TARGET_BITS / GG_TARGET_STRIPS need the forward-differenced kernels (gg_set_fast_mask(1))");
  if (target_kind == GG_TARGET_STRIPS && (reinterpret_cast<uintptr_t>(target) & 7))
    return fail(GG_E_ALIGN, "gg_silhouette_loss_fused: a GG_TARGET_STRIPS buffer must be 8-byte aligned");
  if (B == 0) return GG_OK;
  if (!params || !target || !partials || !loss_partials) return fail(GG_E_NULL, "gg_silhouette_loss_fused: NULL pointer");
  if (!aligned16(params) || (mask_out && !aligned16(mask_out)) || (target_kind != GG_TARGET_BITS && target_kind != GG_TARGET_STRIPS && !aligned16(target)))
    return fail(GG_E_ALIGN, "gg_silhouette_loss_fused: params/target/mask_out must be 16-byte aligned");
  if (T != gg::fused_loss_tiles(K, n)) return fail(GG_E_WORKSPACE, "gg_silhouette_loss_fused: T=%d, expected %d", T, gg::fused_loss_tiles(K, n));
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_silhouette_loss_fused(params, target, target_kind, B, K, n, mask_out, partials,
                                                   loss_partials, T, (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_silhouette_loss_fused");
}

int gg_cycle_tiles(int n) {
  if (n < 1 || n > 16384) return fail(GG_E_SHAPE, "n=%d out of range", n);
  return gg::cycle_tiles(n);
}

int gg_cycle_loss_count(int B, int K, int n) {
  if (B < 0 || K <= 0 || n < 1 || n > 16384) return fail(GG_E_SHAPE, "bad B/K/n");
  return B * gg::cycle_tiles(n) * ((K + 31) / 32);
}

int gg_cycle_l1_fused(const float *params_a, const float *params_b, int B, int K, int n, float *partials_a,
                      float *partials_b, int T, float *loss_partials, int device, void *stream) {
  if (int r = check_bk(B, K)) return r;
  const int sizes[1] = {n};
  if (int r = check_scales(K, 1, sizes)) return r;
  if (B == 0) return GG_OK;
  if (!params_a || !params_b || !partials_a || !partials_b || !loss_partials)
    return fail(GG_E_NULL, "gg_cycle_l1_fused: NULL pointer");
  if (!aligned16(params_a) || !aligned16(params_b)) return fail(GG_E_ALIGN, "gg_cycle_l1_fused: params must be 16-byte aligned");
  if (T != gg::cycle_tiles(n)) return fail(GG_E_WORKSPACE, "gg_cycle_l1_fused: T=%d, expected %d", T, gg::cycle_tiles(n));
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_cycle_l1_fused(params_a, params_b, B, K, n, partials_a, partials_b, loss_partials, T,
                                            (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_cycle_l1_fused");
}

int gg_colorize(const float *params, const float *maps, const float *colors, int B, int K, int n, void *out,
                int out_mode, int device, void *stream) {
  if (int r = check_bk(B, K)) return r;
  if (n < 1 || n > 16384) return fail(GG_E_SHAPE, "n=%d out of range", n);
  if (out_mode < 0 || out_mode > 2) return fail(GG_E_SHAPE, "unknown out_mode %d", out_mode);
  if (B == 0) return GG_OK;
  if ((params == nullptr) == (maps == nullptr)) return fail(GG_E_NULL, "gg_colorize: exactly one of params / maps");
  if (!colors || !out) return fail(GG_E_NULL, "gg_colorize: NULL pointer");
  if (params && (!aligned16(params) || K > 1024)) return fail(GG_E_ALIGN, "gg_colorize: params alignment / K <= 1024");
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = params ? gg::launch_colorize_params(params, colors, B, K, n, out, out_mode, (cudaStream_t)stream)
                         : gg::launch_colorize_maps(maps, colors, (size_t)B * n * n, K, out, out_mode,
                                                    (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_colorize");
}

int gg_sigma_from_frame_fwd(const float *v0, const float *v1, const float *u, int N, double *sigma, int device,
                            void *stream) {
  if (N < 0) return fail(GG_E_SHAPE, "N < 0");
  if (N == 0) return GG_OK;
  if (!v0 || !v1 || !u || !sigma) return fail(GG_E_NULL, "gg_sigma_from_frame_fwd: NULL pointer");
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_sigma_from_frame_fwd(v0, v1, u, N, sigma, (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_sigma_from_frame_fwd");
}

int gg_sigma_from_frame_bwd(const float *v0, const float *v1, const float *u, const double *gsigma, int N, float *dv0,
                            float *dv1, float *du, int device, void *stream) {
  if (N < 0) return fail(GG_E_SHAPE, "N < 0");
  if (N == 0) return GG_OK;
  if (!v0 || !v1 || !u || !gsigma || !dv0 || !dv1 || !du) return fail(GG_E_NULL, "gg_sigma_from_frame_bwd: NULL pointer");
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_sigma_from_frame_bwd(v0, v1, u, gsigma, N, dv0, dv1, du, (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_sigma_from_frame_bwd");
}

size_t gg_p2p_buffer_bytes(int K, int world) {
  if (K <= 0 || world <= 0) return 0;
  return gg::p2p_buffer_bytes(K, world);
}

int gg_shared_grad_post_p2p(const float *dmus, const double *dsigma, int B, int K, const void *peers_dev, int world,
                            int rank, int device, void *stream) {
  if (int r = check_bk(B, K)) return r;
  if (world < 1 || world > 64 || rank < 0 || rank >= world) return fail(GG_E_SHAPE, "bad world/rank (%d/%d)", world, rank);
  if (!dmus || !dsigma || !peers_dev) return fail(GG_E_NULL, "gg_shared_grad_post_p2p: NULL pointer");
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_shared_grad_post(dmus, dsigma, B, K, reinterpret_cast<const unsigned long long *>(peers_dev),
                                              world, rank, (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_shared_grad_post_p2p");
}

int gg_shared_grad_wait_p2p(int K, const void *peers_dev, int world, int rank, double *out, int device, void *stream) {
  if (K <= 0) return fail(GG_E_SHAPE, "K must be > 0");
  if (world < 1 || world > 64 || rank < 0 || rank >= world) return fail(GG_E_SHAPE, "bad world/rank (%d/%d)", world, rank);
  if (!peers_dev || !out) return fail(GG_E_NULL, "gg_shared_grad_wait_p2p: NULL pointer");
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_shared_grad_wait(K, reinterpret_cast<const unsigned long long *>(peers_dev), world, rank, out,
                                              (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_shared_grad_wait_p2p");
}

int gg_shared_grad_allreduce_p2p(const float *dmus, const double *dsigma, int B, int K, const void *peers_dev, int world,
                                 int rank, double *out, int device, void *stream) {
  if (int r = gg_shared_grad_post_p2p(dmus, dsigma, B, K, peers_dev, world, rank, device, stream)) return r;
  return gg_shared_grad_wait_p2p(K, peers_dev, world, rank, out, device, stream);
}

int gg_svd3_fwd(const double *A, int N, double *U, double *S, double *V, int device, void *stream) {
  if (N < 0) return fail(GG_E_SHAPE, "N < 0");
  if (N == 0) return GG_OK;
  if (!A || !U || !S) return fail(GG_E_NULL, "gg_svd3_fwd: NULL pointer");
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_svd3_fwd(A, N, U, S, V, (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_svd3_fwd");
}

int gg_svd3_bwd(const double *U, const double *S, const double *V, const double *gU, const double *gS, int N, double *gA,
                int device, void *stream) {
  if (N < 0) return fail(GG_E_SHAPE, "N < 0");
  if (N == 0) return GG_OK;
  if (!U || !S || !V || !gA) return fail(GG_E_NULL, "gg_svd3_bwd: NULL pointer");
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_svd3_bwd(U, S, V, gU, gS, N, gA, (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_svd3_bwd");
}

int gg_deform_fwd(const float *mu_t, const float *scale, const float *ang, const float *cmu, const double *U,
                  const double *S, const float *theta_raw, int B, int K, float *mus, double *sigmas, float *theta,
                  int device, void *stream) {
  if (int r = check_bk(B, K)) return r;
  if (B == 0) return GG_OK;
  if (!mu_t || !scale || !ang || !cmu || !U || !S || !mus || !sigmas) return fail(GG_E_NULL, "gg_deform_fwd: NULL pointer");
  if ((theta == nullptr) != (theta_raw == nullptr)) return fail(GG_E_NULL, "gg_deform_fwd: theta and theta_raw go together");
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_deform_fwd(mu_t, scale, ang, cmu, U, S, theta_raw, B, K, mus, sigmas, theta,
                                        (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_deform_fwd");
}

int gg_deform_bwd(const float *scale, const float *ang, const double *U, const double *S, const float *gmus,
                  const double *gsig, const float *gtheta, int B, int K, float *dtheta_raw, float *dmu_t,
                  float *dscale, float *dang, float *dcmu, double *dU, double *dS, int device, void *stream) {
  if (int r = check_bk(B, K)) return r;
  if (B == 0) return GG_OK;
  if (!scale || !ang || !U || !S || !dmu_t || !dscale || !dang || !dcmu || !dU || !dS)
    return fail(GG_E_NULL, "gg_deform_bwd: NULL pointer");
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_deform_bwd(scale, ang, U, S, gmus, gsig, gtheta, B, K, dtheta_raw, dmu_t, dscale, dang,
                                        dcmu, dU, dS, (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_deform_bwd");
}

int gg_splat_bwd_tiles(int K, int nscales, const int *sizes_host, const int *has_gmaps_host,
                       const int *has_gmasks_host, int layout) {
  if (K <= 0) return fail(GG_E_SHAPE, "K must be > 0");
  if (int r = check_scales(K, nscales, sizes_host)) return r;
  if (layout != GG_LAYOUT_NHWC && layout != GG_LAYOUT_NCHW) return fail(GG_E_SHAPE, "unknown layout %d", layout);
  return gg::plan_bwd_tiles(K, nscales, sizes_host, has_gmaps_host, has_gmasks_host, layout);
}

int gg_splat_bwd(const float *params, int B, int K, int nscales, const int *sizes_host,
                 const float *const *gmaps_host, const float *const *gmasks_host, int layout, float *partials, int T,
                 int device, void *stream) {
  if (int r = check_bk(B, K)) return r;
  if (int r = check_scales(K, nscales, sizes_host)) return r;
  if (layout != GG_LAYOUT_NHWC && layout != GG_LAYOUT_NCHW) return fail(GG_E_SHAPE, "unknown layout %d", layout);
  if (B == 0) return GG_OK;
  if (!params || !partials) return fail(GG_E_NULL, "gg_splat_bwd: params/partials is NULL");
  if (!aligned16(params)) return fail(GG_E_ALIGN, "gg_splat_bwd: params must be 16-byte aligned");
  int hm[GG_MAX_SCALES], hk[GG_MAX_SCALES];
  for (int i = 0; i < nscales; ++i) {
    const float *m = gmaps_host ? gmaps_host[i] : nullptr;
    const float *k = gmasks_host ? gmasks_host[i] : nullptr;
    hm[i] = m != nullptr;
    hk[i] = k != nullptr;
    if ((m && !aligned16(m)) || (k && !aligned16(k)))
      return fail(GG_E_ALIGN, "gg_splat_bwd: scale %d gradient not 16-byte aligned", i);
  }
  int need = gg::plan_bwd_tiles(K, nscales, sizes_host, hm, hk, layout);
  if (need != T || T <= 0)
    return fail(GG_E_WORKSPACE, "gg_splat_bwd: T=%d but this configuration writes %d partial tiles", T, need);
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  cudaError_t e = gg::launch_splat_bwd(params, B, K, nscales, sizes_host, gmaps_host, gmasks_host, layout, partials,
                                       T, (cudaStream_t)stream);
  return e == cudaSuccess ? GG_OK : cuda_fail(e, "gg_splat_bwd");
}

}  // extern "C"

// gg_launch.h -- host-side launchers implemented in gg_project.cu / gg_splat.cu, called by gg_capi.cu.
#pragma once
#include <cuda_runtime.h>

namespace gg {

cudaError_t launch_project_fwd(const float *mus, const void *sigma, int f64, const float *rx, const float *ry,
                               const float *rz, double tx, double ty, double tz, double focal, int B, int K,
                               double *mu2d, double *sigma2d, float *params, cudaStream_t st, int nviews = 1);
cudaError_t launch_project_bwd(const float *mus, const void *sigma, int f64, const float *rx, const float *ry,
                               const float *rz, double tx, double ty, double tz, double focal, int B, int K,
                               const float *partials, int T, const double *gmu2d, const double *gsigma2d,
                               float *dmus, double *dsigma, float *drot, cudaStream_t st, int nviews = 1);
cudaError_t launch_conic_fwd(const void *mu2d, const void *sigma2d, int f64, int N, float *params, cudaStream_t st);
cudaError_t launch_conic_bwd(const void *mu2d, const void *sigma2d, int f64, int B, int K, const float *partials,
                             int T, double *dmu2d, double *dsigma2d, cudaStream_t st);

// cstrides / coffsets (NHWC, may be NULL = dense [B,n,n,K]): channels per pixel of the buffer a scale's maps live in and
// the channel at which they start (concat-direct, mask/mask_gen_dynamic.py:523-529)
int plan_bwd_tiles(int K, int nscales, const int *sizes, const int *has_gmaps, const int *has_gmasks, int layout,
                   const int *cstrides = nullptr, const int *coffsets = nullptr);
cudaError_t launch_splat_fwd(const float *params, int B, int K, int nscales, const int *sizes, float *const *maps,
                             float *const *masks, int layout, cudaStream_t st, const int *cstrides = nullptr,
                             const int *coffsets = nullptr);
cudaError_t launch_splat_bwd(const float *params, int B, int K, int nscales, const int *sizes,
                             const float *const *gmaps, const float *const *gmasks, int layout, float *partials,
                             int T, cudaStream_t st, const int *cstrides = nullptr, const int *coffsets = nullptr);

cudaError_t launch_mask_l1_loss_bwd(const float *mask, const void *target, int kind, int B, int n, float *gmask,
                                    float *partials, int npartials, cudaStream_t st);

int fused_loss_tiles(int K, int n);
cudaError_t launch_silhouette_loss_fused(const float *params, const void *target, int target_kind, int B, int K, int n,
                                         float *mask_out, float *partials, float *loss_partials, int T,
                                         cudaStream_t st);

struct ScaleDesc;
void make_strip_desc(ScaleDesc &d, int n, int rpt);
int bwd_rpt();
int fwd_rpt();
bool fast_mask();
int set_fast_mask(int enable);

// fused cycle L1 between two renders (mask/mask_gen_dynamic.py:791) and the viz composite (:221-231)
int cycle_tiles(int n);
cudaError_t launch_cycle_l1_fused(const float *params_a, const float *params_b, int B, int K, int n, float *partials_a,
                                  float *partials_b, float *loss_partials, int T, cudaStream_t st);
cudaError_t launch_colorize_params(const float *params, const float *colors, int B, int K, int n, void *out,
                                   int out_u8, cudaStream_t st);
cudaError_t launch_colorize_maps(const float *maps, const float *colors, size_t npix, int K, void *out, int out_u8,
                                 cudaStream_t st);

cudaError_t launch_sigma_from_frame_fwd(const float *v0, const float *v1, const float *u, int N, double *sigma,
                                        cudaStream_t st);
cudaError_t launch_sigma_from_frame_bwd(const float *v0, const float *v1, const float *u, const double *gsigma, int N,
                                        float *dv0, float *dv1, float *du, cudaStream_t st);
cudaError_t launch_svd3_fwd(const double *A, int N, double *U, double *S, double *V, cudaStream_t st);
cudaError_t launch_svd3_bwd(const double *U, const double *S, const double *V, const double *gU, const double *gS, int N,
                            double *gA, cudaStream_t st);
cudaError_t launch_deform_fwd(const float *mu_t, const float *sc, const float *ang, const float *cmu, const double *U,
                              const double *S, const float *theta_raw, int B, int K, float *mus, double *sigmas,
                              float *theta, cudaStream_t st);
cudaError_t launch_deform_bwd(const float *sc, const float *ang, const double *U, const double *S, const float *gmus,
                              const double *gsig, const float *gtheta, int B, int K, float *dtheta_raw, float *dmu_t,
                              float *dsc, float *dang, float *dcmu, double *dU, double *dS, cudaStream_t st);

size_t p2p_buffer_bytes(int K, int world);
cudaError_t launch_shared_grad_post(const float *dmus, const double *dsigma, int B, int K,
                                    const unsigned long long *peers, int world, int rank, cudaStream_t st);
cudaError_t launch_shared_grad_wait(int K, const unsigned long long *peers, int world, int rank, double *out,
                                    cudaStream_t st);

namespace api {   // error/validation helpers shared by the extern "C" translation units (gg_capi.cu)
int fail(int code, const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *where);
bool aligned16(const void *p);
struct DeviceGuard {   // RAII: make `device` current, restore the caller's device on scope exit
  int prev, status;
  explicit DeviceGuard(int device);
  ~DeviceGuard();
  DeviceGuard(const DeviceGuard &) = delete;
  DeviceGuard &operator=(const DeviceGuard &) = delete;
};
int check_bk(int B, int K);
int check_scales(int K, int nscales, const int *sizes);
}  // namespace api

}  // namespace gg

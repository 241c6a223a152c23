/*
 * gaussigan.h -- C ABI of libgaussigan_sm100.so: the B200-native (sm_100a) drop-in for
 * GaussiGAN's differentiable splat-to-silhouette renderer.
 *
 * The reference has no FFI: the path is two plain Python functions evaluated at TF
 * graph-construction time.  Each entry point below names the reference interface it
 * replaces (paths relative to the reference repository root):
 *
 *   get_landmarks(mus, sigma, rotx, roty, rotz, tx, ty, tz, focal)
 *                                   mask/mask_gen_dynamic.py:233-323  (5 more copies)
 *   get_gaussian_maps_2d(mu, sigma, shape_hw)
 *                                   mask/mask_gen_dynamic.py:112-143, mask/infer_masks.py:345-373
 *   sum composite / L1 consumers    mask/mask_gen_dynamic.py:785-791
 *   their backward                  implicit tf.gradients, mask/GAN.py:113-119
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *   - the caller owns all memory; the library never allocates or frees device memory;
 *   - every launch entry takes the CUDA device ordinal and a cudaStream_t (as void*); calls are
 *     asynchronous on that stream and re-entrant, with no implicit synchronisation; the thread's current
 *     CUDA device is restored before the call returns;
 *   - the only state is two process-wide switches, gg_set_fast_mask and gg_set_pdl (atomics: safe to flip
 *     from any thread).  The tile count T of the backward depends on the arithmetic mode: a thread that
 *     flips gg_set_fast_mask between another thread's gg_splat_bwd_tiles query and its launch makes that
 *     launch fail with GG_E_WORKSPACE (every entry re-validates T and buffer capacities against the mode
 *     in force; nothing is ever written out of bounds).  Set the mode once, before planning;
 *   - return value: 0 = ok, < 0 = error (GG_E_*); gg_last_error() gives a thread-local message;
 *     nothing throws across the ABI;
 *   - layouts: mus [B,K,3] f32; sigma [B,K,3,3] f32 or f64 (row-major); rot* [B] f32 DEGREES;
 *     mu2d [B,K,2] (x = image column, y = row); sigma2d [B,K,2,2]; maps NHWC [B,n,n,K] or
 *     NCHW [B,K,n,n] f32; mask [B,n,n] f32; pixel centres = float32 linspace(-1,1,n).
 */
#ifndef GAUSSIGAN_H_
#define GAUSSIGAN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GG_API __attribute__((visibility("default")))
#else
#define GG_API
#endif

#define GG_VERSION 200            /* 0.2.0 */
#define GG_PARAM_STRIDE 8         /* floats per projected Gaussian in the staged `params` buffer */
#define GG_MOMENTS 5              /* gradient accumulators per Gaussian */
#define GG_MAX_SCALES 4           /* the reference's pyramid: 32, 64, 128, 256 (mask_gen_dynamic.py:315-316) */

#define GG_LAYOUT_NHWC 0          /* the reference's layout, [B,n,n,K] (mask_gen_dynamic.py:141) */
#define GG_LAYOUT_NCHW 1          /* [B,K,n,n], for PyTorch consumers */

#define GG_TARGET_U8 0            /* target mask as the raw 8-bit image; mask = raw/127.5 - 1 (mask_gen_dynamic.py:672) */
#define GG_TARGET_F32 1           /* target mask already in [-1, 1], float32 */
#define GG_TARGET_BITS 2          /* BINARY target mask, 1 bit per pixel: rows of ceil(n/8) bytes, pixel j of a row is bit
                                     (j & 7) of byte (j >> 3) (numpy.packbits(..., bitorder="little")); bit 1 = mask +1,
                                     bit 0 = mask -1.  Lossless only for two-level masks; 8x fewer bytes than GG_TARGET_U8
                                     on the host link.  gg_silhouette_loss_fused (forward-differenced kernels) and
                                     gg_silhouette_step_host with K <= 64 only. */
#define GG_TARGET_STRIPS 3        /* ANY 8-bit target mask, losslessly packed by 8-pixel strips (gg_encode_target_strips): a
                                     strip whose pixels are all 0 or all 255 costs two header bits, any other strip its 8 raw
                                     bytes.  Silhouette masks are mostly flat, so the buffer is ~7x smaller than the raw image
                                     on the host link; results are bit-identical to GG_TARGET_U8 on the same image.  Same
                                     entry points and restrictions as GG_TARGET_BITS. */
#define GG_LOSS_PARTIALS 296      /* per-CTA partial sums of |residual| written by gg_mask_l1_loss_bwd */

#define GG_OK 0
#define GG_E_NULL -1              /* required pointer is NULL */
#define GG_E_SHAPE -2             /* B/K/n out of range, bad scale count or layout */
#define GG_E_ALIGN -3             /* pointer not aligned for vector access */
#define GG_E_CUDA -4              /* CUDA runtime error (message in gg_last_error) */
#define GG_E_WORKSPACE -5         /* partials buffer too small / inconsistent tile count */

GG_API int gg_version(void);
/* thread-local description of the last error returned on this thread ("" if none) */
GG_API const char *gg_last_error(void);

/* ---- projection: 3-D Gaussians + camera -> image ellipses -------------------------------------
 * Replaces the body of get_landmarks up to the `mu2d` / `sigma2d` identities
 * (mask/mask_gen_dynamic.py:240-314): float32 trig and R = RX RY RZ, then everything in float64.
 * Outputs (each may be NULL except params): mu2d, sigma2d as float64 (the tensors the frozen decoder
 * graph is fed, mask/infer_masks.py:394-395) and `params`, the float32 staging buffer
 * [B,K,GG_PARAM_STRIDE] the splat kernels read. */
GG_API int gg_project_fwd(const float *mus, const void *sigma, int sigma_is_f64,
                   const float *rotx, const float *roty, const float *rotz,
                   double tx, double ty, double tz, double focal, int B, int K,
                   double *mu2d, double *sigma2d, float *params, int device, void *stream);

/* Backward of gg_project_fwd.  `partials` [B,T,K,GG_MOMENTS] come from gg_splat_bwd (may be NULL
 * when T == 0); gmu2d / gsigma2d (float64, may be NULL) are upstream gradients w.r.t. the mu2d /
 * sigma2d outputs themselves.  Outputs: dmus [B,K,3] f32, dsigma [B,K,3,3] f64 (the raw, generally
 * non-symmetric gradient the reference's autodiff yields), drot [B,3] f32 (d/d rotx,roty,rotz, per
 * degree).  The forward is recomputed in registers; nothing is saved between passes.  Under programmatic dependent
 * launch that recompute runs BEFORE the kernel waits for its predecessor in the stream, so mus / sigma / rot* must not
 * be outputs of the kernel launched immediately before this one (partials, gmu2d and gsigma2d may be); they never
 * are in the sequences of this library (the predecessor is a splat or fused-loss kernel).  For K <= 21 the same window
 * also takes the adjoint itself: it is linear in d mu2d / d sigma2d, so six threads per Gaussian run it for the six
 * unit inputs before the wait and only 15 dot products of length 6 per Gaussian remain after it (GG_PROJ_BWD_LIN=0
 * selects the generic kernel; the two agree to float64 round-off). */
GG_API int gg_project_bwd(const float *mus, const void *sigma, int sigma_is_f64,
                   const float *rotx, const float *roty, const float *rotz,
                   double tx, double ty, double tz, double focal, int B, int K,
                   const float *partials, int T, const double *gmu2d, const double *gsigma2d,
                   float *dmus, double *dsigma, float *drot, int device, void *stream);

/* ---- several cameras per set of Gaussians --------------------------------------------------------
 * The reference renders the SAME (mus, sigmas) three times per training step -- pose, pose + random yaw, canonical
 * (mask/mask_gen_dynamic.py:707-709) -- as three get_landmarks calls.  Here V views of Bg sets are one launch:
 * mus [Bg,K,3], sigma [Bg,K,3,3]; rot* [V*Bg] view-major (image v*Bg + b = set b seen by camera v); params, mu2d,
 * sigma2d and every splat output are indexed by the V*Bg images (pass B = V*Bg to gg_splat_fwd / gg_splat_bwd: views
 * are consecutive slices of the batch dimension).  The backward accumulates dmus [Bg,K,3] and dsigma [Bg,K,3,3] over
 * the views (fixed view order, no atomics); drot [V*Bg,3] is per view.  V = 1 is gg_project_fwd / gg_project_bwd. */
GG_API int gg_project_fwd_views(const float *mus, const void *sigma, int sigma_is_f64,
                   const float *rotx, const float *roty, const float *rotz,
                   double tx, double ty, double tz, double focal, int Bg, int V, int K,
                   double *mu2d, double *sigma2d, float *params, int device, void *stream);
GG_API int gg_project_bwd_views(const float *mus, const void *sigma, int sigma_is_f64,
                   const float *rotx, const float *roty, const float *rotz,
                   double tx, double ty, double tz, double focal, int Bg, int V, int K,
                   const float *partials, int T, const double *gmu2d, const double *gsigma2d,
                   float *dmus, double *dsigma, float *drot, int device, void *stream);

/* ---- conic staging for callers that already hold mu2d / sigma2d -------------------------------
 * Replaces `invsigma = tf.linalg.inv(sigma)` of get_gaussian_maps_2d (mask/mask_gen_dynamic.py:131):
 * mu2d [B,K,2], sigma2d [B,K,2,2] (both f32 or both f64) -> params. */
GG_API int gg_conic_fwd(const void *mu2d, const void *sigma2d, int is_f64, int B, int K, float *params,
                 int device, void *stream);
/* partials -> dmu2d [B,K,2], dsigma2d [B,K,2,2] (float64). */
GG_API int gg_conic_bwd(const void *mu2d, const void *sigma2d, int is_f64, int B, int K,
                 const float *partials, int T, double *dmu2d, double *dsigma2d,
                 int device, void *stream);

/* ---- splat: evaluate every Gaussian at every pixel --------------------------------------------
 * Replaces get_gaussian_maps_2d's raster (mask/mask_gen_dynamic.py:121-141), the float32 cast
 * (:318), the 4-scale loop (:315-319) and, when `masks` is given, the unclamped sum over K
 * (:785-786).  sizes_host[nscales] (host array, 1..GG_MAX_SCALES entries, H == W == n);
 * maps_host[i] / masks_host[i] are host arrays of device pointers (entry or whole array may be
 * NULL): maps [B,n,n,K] (NHWC) or [B,K,n,n] (NCHW) f32, mask [B,n,n] f32 = sum_k maps. */
GG_API int gg_splat_fwd(const float *params, int B, int K, int nscales, const int *sizes_host,
                 float *const *maps_host, float *const *masks_host, int layout,
                 int device, void *stream);

/* Number of [K,GG_MOMENTS] partial blocks per batch element that gg_splat_bwd writes for this
 * configuration (T); < 0 on error.  has_gmaps_host / has_gmasks_host: per-scale 0/1 flags.  T depends
 * on the arithmetic mode (gg_set_fast_mask): query it after the mode is set. */
GG_API int gg_splat_bwd_tiles(int K, int nscales, const int *sizes_host, const int *has_gmaps_host,
                       const int *has_gmasks_host, int layout);

/* Backward of the raster: upstream gradients w.r.t. maps and/or masks -> per-tile partial moments
 * partials [B,T,K,GG_MOMENTS] (every slot is written; no zero-fill needed; deterministic -- no
 * atomics; the sum over the T slots of a batch element is what gg_project_bwd / gg_conic_bwd consume,
 * individual slots may be zero).  T must equal gg_splat_bwd_tiles(...) for the same arguments.
 * Every K and n is accepted in both layouts; NHWC gradients of the maps with n % 8 == 0 take the
 * bulk-copy kernel (any K <= 128 that is a multiple of 4, any other K <= 64). */
GG_API int gg_splat_bwd(const float *params, int B, int K, int nscales, const int *sizes_host,
                 const float *const *gmaps_host, const float *const *gmasks_host, int layout,
                 float *partials, int T, int device, void *stream);

/* ---- concat-direct: the maps written into / their gradients read out of a wider NHWC buffer -------------
 * The reference concatenates every scale's maps onto the generator's activations, tf.concat([x, pe[i]], -1)
 * (mask/mask_gen_dynamic.py:523-529, rgb/rgb_gen_dynamic.py:604-610).  These entry points take the concat buffer
 * itself: bufs_host[i] is the BASE of a [B,n,n,cstrides_host[i]] float32 NHWC tensor and the K maps of scale i occupy
 * its channels [coffsets_host[i], coffsets_host[i] + K) -- no dense [B,n,n,K] tensor and no second pass over it, in
 * either direction (the backward reads the K-channel slice of the concat gradient in place).  128-bit accesses are
 * kept when K, cstride and coffset are multiples of 4 (64-bit for multiples of 2); other shapes take scalar accesses.
 * Other channels of the buffer are never touched.  T comes from gg_splat_bwd_tiles_concat with the same arguments. */
GG_API int gg_splat_fwd_concat(const float *params, int B, int K, int nscales, const int *sizes_host,
                               float *const *bufs_host, const int *cstrides_host, const int *coffsets_host,
                               float *const *masks_host, int device, void *stream);
GG_API int gg_splat_bwd_tiles_concat(int K, int nscales, const int *sizes_host, const int *has_gmaps_host,
                                     const int *has_gmasks_host, const int *cstrides_host, const int *coffsets_host);
GG_API int gg_splat_bwd_concat(const float *params, int B, int K, int nscales, const int *sizes_host,
                               const float *const *gbufs_host, const int *cstrides_host, const int *coffsets_host,
                               const float *const *gmasks_host, float *partials, int T, int device, void *stream);

/* Arithmetic of the MASK-ONLY launches (gg_splat_fwd / gg_splat_bwd on scales that carry only a mask, and
 * gg_silhouette_loss_fused).  1 (default): Gaussians forward-differenced along each 8-pixel strip -- two
 * exponentials per strip instead of eight, FP32-pipe bound, silhouette error ~6e-7 of its maximum.  0: one
 * exponential per (pixel, Gaussian), ~1e-7.  Maps outputs always use the direct form.  Process-wide (atomic); returns
 * the previous setting.  Environment GG_PRECISE_MASK=1 selects 0 at load time.  T (gg_splat_bwd_tiles) and
 * gg_silhouette_step_loss_count depend on it: set it before planning. */
GG_API int gg_set_fast_mask(int enable);

/* Programmatic dependent launch.  1 (default): every kernel is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization and blocks in griddepcontrol.wait before its first global
 * memory access, so its CTAs become resident and run their index arithmetic while the previous kernel of the
 * stream drains (works eagerly and under stream capture; results are identical).  0: plain launches.
 * Process-wide; returns the previous setting.  Environment GG_PDL=0 selects 0 at load time. */
GG_API int gg_set_pdl(int enable);

/* ---- direct consumer: sum-composite L1 loss ------------------------------------------------------
 * Replaces `m_recon_loss_bis` (mask/mask_gen_dynamic.py:785-787): loss = mean|0.5*(target+1) - mask| and its
 * gradient w.r.t. mask.  target [B,n,n] is uint8 (GG_TARGET_U8, normalised as :672) or float32 in [-1,1].
 * Writes gmask [B,n,n] = dloss/dmask and loss_partials[GG_LOSS_PARTIALS] (loss = sum(partials)/(B*n*n),
 * summed by the caller: deterministic, no atomics). */
GG_API int gg_mask_l1_loss_bwd(const float *mask, const void *target, int target_kind, int B, int n,
                               float *gmask, float *loss_partials, int device, void *stream);

/* Fused form of gg_splat_fwd(mask) + gg_mask_l1_loss_bwd + gg_splat_bwd(gmask) for one scale: the silhouette and
 * its gradient stay in registers (mask/mask_gen_dynamic.py:785-787 and their gradients in ONE launch).  Requires
 * K <= 64.  T = gg_splat_bwd_tiles(K, 1, {n}, {0}, {1}, NHWC).  Outputs: partials [B,T,K,GG_MOMENTS] (feed
 * gg_project_bwd), loss_partials [B*T] (loss = sum / (B*n*n)), and, if mask_out != NULL, the silhouette [B,n,n]. */
GG_API int gg_silhouette_loss_fused(const float *params, const void *target, int target_kind, int B, int K, int n,
                                    float *mask_out, float *partials, int T, float *loss_partials,
                                    int device, void *stream);

/* ---- direct consumer: cycle L1 between two renders ----------------------------------------------------
 * Replaces `gm_cycle_loss = reduce_mean(|pose_embeddings[-1] - pose_embeddings_cyc[-1]|)`
 * (mask/mask_gen_dynamic.py:791) AND its gradient into both renders, without materialising either
 * [B,n,n,K] tensor.  params_a / params_b: two staged projections of the same B, K.  T = gg_cycle_tiles(n).
 * Outputs: partials_a, partials_b [B,T,K,GG_MOMENTS] (moments of d loss / d maps, feed gg_project_bwd of each
 * side) and loss_partials [gg_cycle_loss_count(B,K,n)] with loss = sum(partials) / (B*n*n*K). */
GG_API int gg_cycle_tiles(int n);
GG_API int gg_cycle_loss_count(int B, int K, int n);
GG_API int gg_cycle_l1_fused(const float *params_a, const float *params_b, int B, int K, int n,
                             float *partials_a, float *partials_b, int T, float *loss_partials,
                             int device, void *stream);

/* ---- viz composite -----------------------------------------------------------------------------------
 * Replaces colorize_landmark_maps (mask/mask_gen_dynamic.py:221-231, mask/infer_masks.py:174-184):
 * rgb = max_k maps[...,k] * colors[k].  Exactly one of `params` ([B,K,GG_PARAM_STRIDE], maps never written) and
 * `maps` (NHWC [B,n,n,K]) is non-NULL.  colors [K,3] f32 (device).  out [B,n,n,3]: GG_RGB_F32, GG_RGB_U8
 * (uint8(rgb*255), numpy truncation) or GG_RGB_U8_INV (255 - that: the frames of infer_masks.py:463). */
#define GG_RGB_F32 0
#define GG_RGB_U8 1
#define GG_RGB_U8_INV 2
GG_API int gg_colorize(const float *params, const float *maps, const float *colors, int B, int K, int n,
                       void *out, int out_mode, int device, void *stream);

/* ---- Gaussian construction and per-frame deformation ---------------------------------------------------
 * gg_sigma_from_frame_*: tail of get_musigma (mask/mask_gen_dynamic.py:383-401).  v0, v1 [N,3] f32 (tanh outputs),
 * u [N,3] f32 (pre-sigmoid) -> sigma [N,3,3] f64 = V diag(0.5*sigmoid(u)+0.01) V^T built in float32.
 * Backward: gsigma [N,3,3] f64 -> dv0, dv1, du [N,3] f32. */
GG_API int gg_sigma_from_frame_fwd(const float *v0, const float *v1, const float *u, int N, double *sigma,
                                   int device, void *stream);
GG_API int gg_sigma_from_frame_bwd(const float *v0, const float *v1, const float *u, const double *gsigma, int N,
                                   float *dv0, float *dv1, float *du, int device, void *stream);
/* gg_svd3_*: `S, U, _ = tf.linalg.svd(cannsigma)` (mask/mask_gen_dynamic.py:430) for 3x3 matrices A [N,3,3] f64 ->
 * U [N,3,3], S [N,3] (descending), V [N,3,3] (may be NULL) with A = U diag(S) V^T; column signs arbitrary, as in
 * LAPACK.  One-sided Jacobi in float64; no host synchronisation (unlike cuSOLVER-backed torch.linalg.svd), so the
 * dynamic-object step can be captured in a CUDA graph.  Backward (square full-rank, no gradient through V, as the
 * reference discards it): gU, gS (either may be NULL) -> gA [N,3,3]; terms with equal singular values are dropped. */
GG_API int gg_svd3_fwd(const double *A, int N, double *U, double *S, double *V, int device, void *stream);
GG_API int gg_svd3_bwd(const double *U, const double *S, const double *V, const double *gU, const double *gS, int N,
                       double *gA, int device, void *stream);
/* gg_deform_*: tail of transform_mu_sigma (mask/mask_gen_dynamic.py:426-458).  Head outputs AFTER tanh, f32:
 * mu_t [B,K,3], scale [B,K,3], ang [B,K,3] (x,y,z), theta_raw [B]; canonical cmu [K,3] f32 and the SVD factors
 * U [K,3,3], S [K,3] (float64) of the canonical covariances (the SVD itself, :430, stays with the caller).
 * -> mus [B,K,3] f32, sigmas [B,K,3,3] f64, theta [B] f32 (degrees).
 * Backward returns PER-SAMPLE gradients for the broadcast canonical parameters (dcmu [B,K,3] f32, dU [B,K,3,3],
 * dS [B,K,3] f64): the caller sums them over the batch.  gmus / gsig / gtheta may be NULL (treated as zero). */
GG_API int gg_deform_fwd(const float *mu_t, const float *scale, const float *ang, const float *cmu, const double *U,
                         const double *S, const float *theta_raw, int B, int K, float *mus, double *sigmas,
                         float *theta, int device, void *stream);
GG_API int gg_deform_bwd(const float *scale, const float *ang, const double *U, const double *S, const float *gmus,
                         const double *gsig, const float *gtheta, int B, int K, float *dtheta_raw, float *dmu_t,
                         float *dscale, float *dang, float *dcmu, double *dU, double *dS, int device, void *stream);

/* ---- training-time exchange over NVLink peer memory ------------------------------------------------------
 * Gradient of the batch-shared canonical Gaussians (mask/mask_gen_dynamic.py:701-702 broadcast): sum of the
 * per-sample gradients dmus [B,K,3] f32 and dsigma [B,K,3,3] f64 over the local batch and over all `world` ranks of
 * one node, in ONE single-CTA kernel: local sum, direct stores into every peer's receive slot, flag exchange with
 * system-scope release/acquire, rank-ordered final sum (bit-identical on all ranks).  `peers_dev` is a DEVICE array
 * of `world` base addresses of the ranks' exchange buffers (peer-mapped, e.g. torch symmetric memory), each
 * gg_p2p_buffer_bytes(K, world) bytes, zero-filled once before the first call (with a barrier across ranks).
 * out [12*K] f64 = dmu sums [K,3] followed by dsigma sums [K,3,3].  Every rank must make the same sequence of calls.
 * Replayable from a CUDA graph (the call counters live in the buffer).
 * Two halves, so that the exchange overlaps whatever the caller runs in between:
 *   gg_shared_grad_post_p2p  -- local batch sum (ceil(12 K / 32) CTAs, one load round trip), stores into every peer's
 *                               slot, flag on every peer.  Enqueue right behind the projection adjoint.
 *   gg_shared_grad_wait_p2p  -- waits for the peers' flags, adds the slots in rank order into `out`.  Enqueue any time
 *                               later on a stream ordered after the post (just before the optimiser step), but before
 *                               the NEXT post: at most one post may be outstanding.
 * gg_shared_grad_allreduce_p2p = post + wait back to back.
 * The wait is bounded by GG_P2P_TIMEOUT_MS (environment, default 60000): on expiry the kernel fills `out` with NaN --
 * the failure reaches the gradients -- and latches the call number in word 1 of the rank's own buffer; the exchange
 * group must then be rebuilt (as after an NCCL abort). */
GG_API size_t gg_p2p_buffer_bytes(int K, int world);
GG_API int gg_shared_grad_post_p2p(const float *dmus, const double *dsigma, int B, int K, const void *peers_dev,
                                   int world, int rank, int device, void *stream);
GG_API int gg_shared_grad_wait_p2p(int K, const void *peers_dev, int world, int rank, double *out, int device,
                                   void *stream);
GG_API int gg_shared_grad_allreduce_p2p(const float *dmus, const double *dsigma, int B, int K,
                                        const void *peers_dev, int world, int rank, double *out, int device,
                                        void *stream);

/* ---- whole training-style step on HOST buffers ---------------------------------------------------
 * What a host-language binding of the reference would call (the reference feeds numpy arrays through
 * session.run, mask/infer_masks.py:435-456): H2D of the inputs, projection, splat (silhouette), L1 loss and its
 * gradient, splat backward, projection backward, D2H of loss partials and gradients -- all enqueued on
 * `stream`, no synchronisation.  *_host pointers are HOST memory (pinned for true asynchrony); `workspace` is
 * DEVICE memory of at least gg_silhouette_step_workspace_bytes(...) bytes, 256-byte aligned, owned by the
 * caller and private to this call until the stream has drained.  mask_host may be NULL (silhouette not
 * copied back).  sigma_host is float64 [B,K,3,3].  loss_partials_host (capacity loss_partials_capacity floats,
 * checked) receives gg_silhouette_step_loss_count(B, K, n) floats; loss = sum(partials) / (B*n*n).  For K <= 64 the step is three
 * launches (projection, fused splat+loss+backward, projection adjoint), otherwise five.  The inputs (Gaussians, camera,
 * target mask) go through the copy engine; when mus_host, sigma_host and rotx_host lie in ONE host block, each at the
 * next multiple of 256 bytes behind the previous array, with roty and rotz right behind rotx, the five arrays are
 * copied with one call.  When the GRADIENT buffers (dmus, dsigma, drot) are pinned, the projection
 * adjoint writes them in place over PCIe instead of into the workspace followed by three copies (a buffer is used in
 * place only if its WHOLE extent is pinned).  GG_ZEROCOPY selects what is accessed in place: "out" (default) the
 * gradients, "in" the Gaussians and the camera, "1" both, "0" nothing.  Measured on B200 / PCIe 5 (batch 64, K = 16,
 * 256 x 256, six steps in flight): "out" 34.8 us per step, "0" 35.1, "in" 42.0, "1" 44.0 -- reading the inputs in
 * place puts PCIe round trips on the projection's dependent chain.  LIFETIME: the output buffers -- and, with "in" or
 * "1", the input buffers as well, which the projection adjoint re-reads at the END of the chain -- must stay untouched
 * until the stream has passed the whole call (record an event after it), not merely until the leading copies are
 * done. */
/* GG_TARGET_STRIPS buffer (host or device memory, 8-byte aligned):
 *   bytes 0..15   uint32 magic 0x54534747 ("GGST"), n, B, total_bytes
 *   then          uint32 row_offset[B*n]: byte offset of row (b, r)'s record from the start of the payload
 *   then (8-byte aligned) payload; record of one row, W = ceil(ceil(n/8) / 32):
 *                 uint32 M[W] (bit s set: strip s is stored raw), uint32 V[W] (bit s: a flat strip's value, 1 = 255),
 *                 then 8 bytes per raw strip in strip order (pixels past the row's end are zero).
 * gg_target_strips_bound: bytes that always suffice.  gg_encode_target_strips (host function, no CUDA): image
 * [B,n,n] uint8 -> buffer; returns the bytes used (also stored in the header) or < 0.  One pass over the image at
 * memory speed; a data loader calls it when it produces the mask. */
GG_API size_t gg_target_strips_bound(int B, int n);
GG_API long long gg_encode_target_strips(const unsigned char *image_host, int B, int n, void *out_host, size_t capacity);

GG_API size_t gg_silhouette_step_workspace_bytes(int B, int K, int n, int target_kind);
GG_API int gg_silhouette_step_loss_count(int B, int K, int n);
GG_API int gg_silhouette_step_host(const float *mus_host, const double *sigma_host, const float *rotx_host,
                                   const float *roty_host, const float *rotz_host, const void *target_host,
                                   int target_kind, double tx, double ty, double tz, double focal,
                                   int B, int K, int n, float *loss_partials_host, int loss_partials_capacity,
                                   float *dmus_host, double *dsigma_host, float *drot_host, float *mask_host,
                                   void *workspace, size_t workspace_bytes, int device, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GAUSSIGAN_H_ */

// gg_geometry.cu -- the two small per-Gaussian geometry stages that sit in front of the projection:
//
//   sigma_from_frame   tail of get_musigma            mask/mask_gen_dynamic.py:383-401
//       v0 = l2n(a), v1 = l2n(v0 x b), v2 = v1 x v0, V = [v0 v2 v1], u = 0.5*sigmoid(x) + 0.01,
//       Sigma = V diag(u) V^T   (float32 arithmetic, cast to float64 at the end)
//   deform             tail of transform_mu_sigma     mask/mask_gen_dynamic.py:426-458
//       mu_b = mu_c + 0.1*tanh(.), Sigma_b = (R_k U) diag(s * (1 + 0.05*tanh(.))^2) (R_k U)^T, theta = 180*tanh(.)
//       with S, U = svd(Sigma_c) left to the caller (torch.linalg.svd keeps its own gradient) and R_k the
//       per-Gaussian Euler rotation (same column convention as get_landmarks, radians = tanh * pi / 10).
//
// In the reference each is ~25 tiny TF ops (pure launch latency); here each direction is one launch, one thread
// per Gaussian.  The reverse passes are hand-derived adjoints evaluated in float64.
#include <type_traits>
#include "gg_common.cuh"
#include "gg_launch.h"

namespace gg {

namespace {

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 scale(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 cross32(F3 a, F3 b) {   // tf.linalg.cross, float32, products/differences rounded
  return {__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
          __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x))};
}
__device__ __forceinline__ float sumsq32(F3 a) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)), __fmul_rn(a.z, a.z));
}
// tf.math.l2_normalize: x * rsqrt(max(sum(x^2), 1e-12))
__device__ __forceinline__ F3 l2n32(F3 a, float *inv_out) {
  float inv = __frcp_rn(__fsqrt_rn(fmaxf(sumsq32(a), 1e-12f)));
  *inv_out = inv;
  return {__fmul_rn(a.x, inv), __fmul_rn(a.y, inv), __fmul_rn(a.z, inv)};
}
__device__ __forceinline__ float sigmoid32(float x) { return 1.0f / (1.0f + expf(-x)); }

// float32 Euler rotation with the reference's column-stacked matrices (:442-453), radians in, pinned order
__device__ void euler32(float ax, float ay, float az, float *RX, float *RY, float *RZ, float *RXY, float *R,
                        float *c, float *s) {
  const float a[3] = {ax, ay, az};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (a[i] == 0.0f) { c[i] = 1.0f; s[i] = 0.0f; }
    else { double sd, cd; sincos((double)a[i], &sd, &cd); c[i] = (float)cd; s[i] = (float)sd; }
  }
  const float rx[9] = {1.f, 0.f, 0.f, 0.f, c[0], s[0], 0.f, -s[0], c[0]};
  const float ry[9] = {c[1], 0.f, -s[1], 0.f, 1.f, 0.f, s[1], 0.f, c[1]};
  const float rz[9] = {c[2], s[2], 0.f, -s[2], c[2], 0.f, 0.f, 0.f, 1.f};
#pragma unroll
  for (int i = 0; i < 9; ++i) { RX[i] = rx[i]; RY[i] = ry[i]; RZ[i] = rz[i]; }
  auto mm = [](const float *A, const float *Bm, float *C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        C[i * 3 + j] = __fadd_rn(__fadd_rn(__fmul_rn(A[i * 3], Bm[j]), __fmul_rn(A[i * 3 + 1], Bm[3 + j])),
                                 __fmul_rn(A[i * 3 + 2], Bm[6 + j]));
  };
  mm(RX, RY, RXY);
  mm(RXY, RZ, R);
}

__device__ __forceinline__ void mm64(const double *A, const double *Bm, double *C, bool ta, bool tb) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) acc += (ta ? A[k * 3 + i] : A[i * 3 + k]) * (tb ? Bm[j * 3 + k] : Bm[k * 3 + j]);
      C[i * 3 + j] = acc;
    }
}

}  // namespace

// ------------------------------------------------------------------------------ sigma_from_frame
__global__ void __launch_bounds__(128) sigma_from_frame_fwd_kernel(const float *__restrict__ v0r,
                                                                   const float *__restrict__ v1r,
                                                                   const float *__restrict__ ur, int N,
                                                                   double *__restrict__ sigma) {
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const F3 a = {v0r[3 * i], v0r[3 * i + 1], v0r[3 * i + 2]}, b = {v1r[3 * i], v1r[3 * i + 1], v1r[3 * i + 2]};
  float inv;
  const F3 v0 = l2n32(a, &inv);                         // :388
  const F3 v1 = l2n32(cross32(v0, b), &inv);            // :389-390
  const F3 v2 = cross32(v1, v0);                        // :391
  const float V[9] = {v0.x, v2.x, v1.x, v0.y, v2.y, v1.y, v0.z, v2.z, v1.z};   // columns v0, v2, v1 (:396)
  float u[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) u[j] = __fadd_rn(__fmul_rn(sigmoid32(ur[3 * i + j]), 0.5f), 1e-2f);   // :397
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {                        // (V U) V^T in float32 (:400), then cast (:401)
      float acc = __fmul_rn(__fmul_rn(V[r * 3], u[0]), V[c * 3]);
      acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(V[r * 3 + 1], u[1]), V[c * 3 + 1]));
      acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(V[r * 3 + 2], u[2]), V[c * 3 + 2]));
      sigma[(size_t)i * 9 + r * 3 + c] = (double)acc;
    }
}

__global__ void __launch_bounds__(128) sigma_from_frame_bwd_kernel(const float *__restrict__ v0r,
                                                                   const float *__restrict__ v1r,
                                                                   const float *__restrict__ ur,
                                                                   const double *__restrict__ gsigma, int N,
                                                                   float *__restrict__ dv0, float *__restrict__ dv1,
                                                                   float *__restrict__ du) {
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const V3 a = {v0r[3 * i], v0r[3 * i + 1], v0r[3 * i + 2]}, b = {v1r[3 * i], v1r[3 * i + 1], v1r[3 * i + 2]};
  const double na2 = fmax(dot(a, a), 1e-12), ia = 1.0 / sqrt(na2);
  const V3 v0 = scale(a, ia);
  const V3 w = cross(v0, b);
  const double nw2 = fmax(dot(w, w), 1e-12), iw = 1.0 / sqrt(nw2);
  const V3 v1 = scale(w, iw);
  const V3 v2 = cross(v1, v0);
  const double V[9] = {v0.x, v2.x, v1.x, v0.y, v2.y, v1.y, v0.z, v2.z, v1.z};
  double u[3], sg[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { sg[j] = 1.0 / (1.0 + exp(-(double)ur[3 * i + j])); u[j] = 0.5 * sg[j] + 1e-2; }
  double G[9], Gs[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) G[j] = gsigma[(size_t)i * 9 + j];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) Gs[r * 3 + c] = G[r * 3 + c] + G[c * 3 + r];
  // Sigma = V U V^T :  Vbar = (G + G^T) V U,  ubar_k = (V^T G V)_kk
  double Vb[9], GV[9];
  mm64(Gs, V, Vb, false, false);
  mm64(G, V, GV, false, false);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double ub = V[k] * GV[k] + V[3 + k] * GV[3 + k] + V[6 + k] * GV[6 + k];
    du[3 * i + k] = (float)(ub * 0.5 * sg[k] * (1.0 - sg[k]));
#pragma unroll
    for (int r = 0; r < 3; ++r) Vb[r * 3 + k] *= u[k];
  }
  V3 v0b = {Vb[0], Vb[3], Vb[6]}, v2b = {Vb[1], Vb[4], Vb[7]}, v1b = {Vb[2], Vb[5], Vb[8]};
  // v2 = v1 x v0
  v1b = add(v1b, cross(v0, v2b));
  v0b = add(v0b, cross(v2b, v1));
  // v1 = w / |w|
  V3 wb = scale(add(v1b, scale(v1, -dot(v1, v1b))), iw);
  if (dot(w, w) < 1e-12) wb = scale(v1b, iw);            // clamped branch of l2_normalize
  // w = v0 x b
  v0b = add(v0b, cross(b, wb));
  const V3 bb = cross(wb, v0);
  // v0 = a / |a|
  V3 ab = scale(add(v0b, scale(v0, -dot(v0, v0b))), ia);
  if (dot(a, a) < 1e-12) ab = scale(v0b, ia);
  dv0[3 * i] = (float)ab.x; dv0[3 * i + 1] = (float)ab.y; dv0[3 * i + 2] = (float)ab.z;
  dv1[3 * i] = (float)bb.x; dv1[3 * i + 1] = (float)bb.y; dv1[3 * i + 2] = (float)bb.z;
}

// ---------------------------------------------------------------------------------------- deform
// inputs per sample b and Gaussian k (all AFTER the tanh of the dense heads, float32):
//   mu_t [B,K,3], sc [B,K,3], ang [B,K,3] (x,y,z), theta [B];  canonical: cmu [K,3] f32, U [K,3,3] f64, S [K,3] f64
__global__ void __launch_bounds__(128) deform_fwd_kernel(const float *__restrict__ mu_t, const float *__restrict__ sc,
                                                         const float *__restrict__ ang, const float *__restrict__ cmu,
                                                         const double *__restrict__ U, const double *__restrict__ S,
                                                         const float *__restrict__ theta_raw, int B, int K,
                                                         float *__restrict__ mus, double *__restrict__ sigmas,
                                                         float *__restrict__ theta) {
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B && theta) theta[i] = __fmul_rn(theta_raw[i], 180.0f);       // :458
  if (i >= B * K) return;
  const int k = i % K;
#pragma unroll
  for (int j = 0; j < 3; ++j)   // mus = cannmu * ones + tanh * 0.1 (:426-428), float32
    mus[3 * i + j] = __fadd_rn(__fmul_rn(cmu[3 * k + j], 1.0f), __fmul_rn(mu_t[3 * i + j], 0.1f));
  float RX[9], RY[9], RZ[9], RXY[9], Rf[9], c[3], s[3];
  const float pi32 = 3.14159274101257324219f;
  euler32(__fdiv_rn(__fmul_rn(ang[3 * i], pi32), 10.0f), __fdiv_rn(__fmul_rn(ang[3 * i + 1], pi32), 10.0f),
          __fdiv_rn(__fmul_rn(ang[3 * i + 2], pi32), 10.0f), RX, RY, RZ, RXY, Rf, c, s);   // :439-453
  double R[9], Uk[9], Nm[9], d[3];
#pragma unroll
  for (int j = 0; j < 9; ++j) { R[j] = (double)Rf[j]; Uk[j] = U[(size_t)k * 9 + j]; }
  mm64(R, Uk, Nm, false, false);                                   // newU = R U (:455)
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double scj = (double)__fadd_rn(__fmul_rn(sc[3 * i + j], 0.05f), 1.0f);   // :432 float32, then cast (:436)
    const double e = sqrt(S[(size_t)k * 3 + j]) * scj;                               // :431, :437
    d[j] = e * e;                                                                    // matmul(Sscale, Sscale) (:456)
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
      sigmas[(size_t)i * 9 + r * 3 + cc] = Nm[r * 3] * d[0] * Nm[cc * 3] + Nm[r * 3 + 1] * d[1] * Nm[cc * 3 + 1] +
                                           Nm[r * 3 + 2] * d[2] * Nm[cc * 3 + 2];
}

// per-sample gradients; the caller sums dcmu / dU / dS over the batch (they belong to broadcast parameters)
__global__ void __launch_bounds__(128) deform_bwd_kernel(const float *__restrict__ sc, const float *__restrict__ ang,
                                                         const double *__restrict__ U, const double *__restrict__ S,
                                                         const float *__restrict__ gmus,
                                                         const double *__restrict__ gsig,
                                                         const float *__restrict__ gtheta, int B, int K,
                                                         float *__restrict__ dtheta_raw,
                                                         float *__restrict__ dmu_t, float *__restrict__ dsc,
                                                         float *__restrict__ dang, float *__restrict__ dcmu,
                                                         double *__restrict__ dU, double *__restrict__ dS) {
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B && dtheta_raw) dtheta_raw[i] = gtheta ? __fmul_rn(gtheta[i], 180.0f) : 0.f;
  if (i >= B * K) return;
  const int k = i % K;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float g = gmus ? gmus[3 * i + j] : 0.f;
    dmu_t[3 * i + j] = __fmul_rn(g, 0.1f);
    dcmu[3 * i + j] = g;
  }
  float RX[9], RY[9], RZ[9], RXY[9], Rf[9], c[3], s[3];
  const float pi32 = 3.14159274101257324219f;
  euler32(__fdiv_rn(__fmul_rn(ang[3 * i], pi32), 10.0f), __fdiv_rn(__fmul_rn(ang[3 * i + 1], pi32), 10.0f),
          __fdiv_rn(__fmul_rn(ang[3 * i + 2], pi32), 10.0f), RX, RY, RZ, RXY, Rf, c, s);
  double R[9], Uk[9], Nm[9], d[3], scd[3], Sk[3];
#pragma unroll
  for (int j = 0; j < 9; ++j) { R[j] = (double)Rf[j]; Uk[j] = U[(size_t)k * 9 + j]; }
  mm64(R, Uk, Nm, false, false);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    scd[j] = (double)__fadd_rn(__fmul_rn(sc[3 * i + j], 0.05f), 1.0f);
    Sk[j] = S[(size_t)k * 3 + j];
    d[j] = Sk[j] * scd[j] * scd[j];
  }
  double G[9], Gs[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) G[j] = gsig ? gsig[(size_t)i * 9 + j] : 0.0;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) Gs[r * 3 + cc] = G[r * 3 + cc] + G[cc * 3 + r];
  // Sigma = N D N^T: Nbar = (G + G^T) N D, dbar_j = (N^T G N)_jj
  double Nb[9], GN[9];
  mm64(Gs, Nm, Nb, false, false);
  mm64(G, Nm, GN, false, false);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double db = Nm[j] * GN[j] + Nm[3 + j] * GN[3 + j] + Nm[6 + j] * GN[6 + j];
    dS[(size_t)i * 3 + j] = db * scd[j] * scd[j];                   // d = S * sc^2 (S enters through sqrt then square)
    dsc[3 * i + j] = (float)(db * 2.0 * Sk[j] * scd[j] * 0.05);
#pragma unroll
    for (int r = 0; r < 3; ++r) Nb[r * 3 + j] *= d[j];
  }
  // N = R U: Rbar = Nbar U^T, Ubar = R^T Nbar
  double Rb[9], Ub[9];
  mm64(Nb, Uk, Rb, false, true);
  mm64(R, Nb, Ub, true, false);
#pragma unroll
  for (int j = 0; j < 9; ++j) dU[(size_t)i * 9 + j] = Ub[j];
  // R = (RX RY) RZ in float32; adjoint in float64 on the same float32 values
  double RXd[9], RYd[9], RZd[9], RXYd[9], RYZ[9], RXb[9], RYb[9], RZb[9], tmp[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) { RXd[j] = RX[j]; RYd[j] = RY[j]; RZd[j] = RZ[j]; RXYd[j] = RXY[j]; }
  mm64(RYd, RZd, RYZ, false, false);
  mm64(Rb, RYZ, RXb, false, true);
  mm64(RXd, Rb, tmp, true, false);
  mm64(tmp, RZd, RYb, false, true);
  mm64(RXYd, Rb, RZb, true, false);
  const double cx = c[0], sx = s[0], cy = c[1], sy = s[1], cz = c[2], sz = s[2];
  const double gx = -sx * RXb[4] + cx * RXb[5] - cx * RXb[7] - sx * RXb[8];
  const double gy = -sy * RYb[0] - cy * RYb[2] + cy * RYb[6] - sy * RYb[8];
  const double gz = -sz * RZb[0] + cz * RZb[1] - cz * RZb[3] - sz * RZb[4];
  const double k10 = (double)pi32 / 10.0;                            // radians = tanh * pi / 10
  dang[3 * i] = (float)(gx * k10); dang[3 * i + 1] = (float)(gy * k10); dang[3 * i + 2] = (float)(gz * k10);
}

// -------------------------------------------------------------------------------------- launchers
// ---------------------------------------------------------------------------------------------- 3x3 SVD
// `S, U, _ = tf.linalg.svd(cannsigma)` (mask/mask_gen_dynamic.py:430).  One-sided (Hestenes) Jacobi in float64:
// right rotations orthogonalise the columns, A V = U diag(S); singular values descending as LAPACK / TF return them;
// column signs are arbitrary (they cancel in (R U) D (R U)^T).  General 3x3 input -- the reference's canonical
// covariances are symmetric only up to 4e-8 (they are built in float32), and the factorisation of exactly that matrix
// is what the oracle differentiates.  One thread per matrix; a kernel instead of cuSOLVER because torch.linalg.svd
// synchronises with the host (its status check), which stalls the stream and keeps the dynamic-object step out of
// CUDA graphs.
__global__ void __launch_bounds__(128) svd3_fwd_kernel(const double *__restrict__ A, int N, double *__restrict__ U,
                                                      double *__restrict__ S, double *__restrict__ Vout) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double Bm[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
  for (int j = 0; j < 9; ++j) Bm[j] = A[(size_t)i * 9 + j];
  for (int sweep = 0; sweep < 16; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      double al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        al += Bm[r * 3 + p] * Bm[r * 3 + p];
        be += Bm[r * 3 + q] * Bm[r * 3 + q];
        ga += Bm[r * 3 + p] * Bm[r * 3 + q];
      }
      // columns orthogonal to 1e-15 (a threshold below the float64 round-off of ga itself is never met, and the loop
      // ran all 16 sweeps: 19.5 us for 16 matrices; Jacobi converges quadratically, so this stops after 4-6 sweeps)
      if (ga * ga <= 1e-30 * al * be) continue;
      rotated = true;
      const double zeta = (be - al) / (2.0 * ga);
      const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double bp = Bm[r * 3 + p], bq = Bm[r * 3 + q];
        Bm[r * 3 + p] = c * bp - sn * bq;
        Bm[r * 3 + q] = sn * bp + c * bq;
        const double vp = V[r * 3 + p], vq = V[r * 3 + q];
        V[r * 3 + p] = c * vp - sn * vq;
        V[r * 3 + q] = sn * vp + c * vq;
      }
    }
    if (!rotated) break;
  }
  double sv[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) sv[c] = sqrt(Bm[c] * Bm[c] + Bm[3 + c] * Bm[3 + c] + Bm[6 + c] * Bm[6 + c]);
  // singular values in descending order: three compare-and-swap steps on whole columns with compile-time indices (a
  // run-time permutation index would push Bm and V -- and with them the whole Jacobi loop -- into local memory)
  auto cswap = [&](auto ca, auto cb) {
    constexpr int a_ = decltype(ca)::value, b_ = decltype(cb)::value;
    if (sv[a_] < sv[b_]) {
      double t = sv[a_]; sv[a_] = sv[b_]; sv[b_] = t;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        t = Bm[r * 3 + a_]; Bm[r * 3 + a_] = Bm[r * 3 + b_]; Bm[r * 3 + b_] = t;
        t = V[r * 3 + a_]; V[r * 3 + a_] = V[r * 3 + b_]; V[r * 3 + b_] = t;
      }
    }
  };
  cswap(std::integral_constant<int, 0>{}, std::integral_constant<int, 1>{});
  cswap(std::integral_constant<int, 1>{}, std::integral_constant<int, 2>{});
  cswap(std::integral_constant<int, 0>{}, std::integral_constant<int, 1>{});
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double sc = sv[c];
    const double inv = sc > 0.0 ? 1.0 / sc : 0.0;
    S[(size_t)i * 3 + c] = sc;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      U[(size_t)i * 9 + r * 3 + c] = Bm[r * 3 + c] * inv;
      if (Vout) Vout[(size_t)i * 9 + r * 3 + c] = V[r * 3 + c];
    }
  }
}

// Backward for square, full-rank A with no gradient through V (the reference discards it: `S, U, _ = svd`):
//   gA = U [ (skew(U^T gU) / E) diag(S) + diag(gS) ] V^T,   skew(X) = X - X^T,   E_ij = s_j^2 - s_i^2
// gU, gS may each be NULL.  Terms with equal singular values are dropped.
__global__ void __launch_bounds__(128) svd3_bwd_kernel(const double *__restrict__ U, const double *__restrict__ S,
                                                      const double *__restrict__ V, const double *__restrict__ gU,
                                                      const double *__restrict__ gS, int N, double *__restrict__ gA) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double Um[9], Vm[9], G[9], P[9], s[3];
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    Um[j] = U[(size_t)i * 9 + j];
    Vm[j] = V[(size_t)i * 9 + j];
    G[j] = gU ? gU[(size_t)i * 9 + j] : 0.0;
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) s[j] = S[(size_t)i * 3 + j];
  mm64(Um, G, P, true, false);                       // U^T gU
  double gP[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (r == c) gP[r * 3 + c] = gS ? gS[(size_t)i * 3 + r] : 0.0;
      else {
        const double e = s[c] * s[c] - s[r] * s[r];
        gP[r * 3 + c] = e != 0.0 ? (P[r * 3 + c] - P[c * 3 + r]) / e * s[c] : 0.0;
      }
    }
  double T1[9], Gm[9];
  mm64(Um, gP, T1, false, false);
  mm64(T1, Vm, Gm, false, true);                     // U gP V^T
#pragma unroll
  for (int j = 0; j < 9; ++j) gA[(size_t)i * 9 + j] = Gm[j];
}

cudaError_t launch_sigma_from_frame_fwd(const float *v0, const float *v1, const float *u, int N, double *sigma,
                                        cudaStream_t st) {
  launch_pdl(sigma_from_frame_fwd_kernel, dim3((N + 127) / 128), dim3(128), 0, st, v0, v1, u, N, sigma);
  return cudaGetLastError();
}
cudaError_t launch_sigma_from_frame_bwd(const float *v0, const float *v1, const float *u, const double *gsigma, int N,
                                        float *dv0, float *dv1, float *du, cudaStream_t st) {
  launch_pdl(sigma_from_frame_bwd_kernel, dim3((N + 127) / 128), dim3(128), 0, st, v0, v1, u, gsigma, N, dv0, dv1, du);
  return cudaGetLastError();
}
cudaError_t launch_deform_fwd(const float *mu_t, const float *sc, const float *ang, const float *cmu, const double *U,
                              const double *S, const float *theta_raw, int B, int K, float *mus, double *sigmas,
                              float *theta, cudaStream_t st) {
  launch_pdl(deform_fwd_kernel, dim3((B * K + 127) / 128), dim3(128), 0, st, mu_t, sc, ang, cmu, U, S, theta_raw, B, K, mus, sigmas, theta);
  return cudaGetLastError();
}
cudaError_t launch_deform_bwd(const float *sc, const float *ang, const double *U, const double *S, const float *gmus,
                              const double *gsig, const float *gtheta, int B, int K, float *dtheta_raw, float *dmu_t,
                              float *dsc, float *dang, float *dcmu, double *dU, double *dS, cudaStream_t st) {
  launch_pdl(deform_bwd_kernel, dim3((B * K + 127) / 128), dim3(128), 0, st, sc, ang, U, S, gmus, gsig, gtheta, B, K, dtheta_raw, dmu_t, dsc,
                                                         dang, dcmu, dU, dS);
  return cudaGetLastError();
}

cudaError_t launch_svd3_fwd(const double *A, int N, double *U, double *S, double *V, cudaStream_t st) {
  launch_pdl(svd3_fwd_kernel, dim3((N + 127) / 128), dim3(128), 0, st, A, N, U, S, V);
  return cudaGetLastError();
}
cudaError_t launch_svd3_bwd(const double *U, const double *S, const double *V, const double *gU, const double *gS, int N,
                            double *gA, cudaStream_t st) {
  launch_pdl(svd3_bwd_kernel, dim3((N + 127) / 128), dim3(128), 0, st, U, S, V, gU, gS, N, gA);
  return cudaGetLastError();
}

}  // namespace gg

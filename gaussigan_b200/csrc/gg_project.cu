// gg_project.cu -- per-Gaussian projection (3-D ellipsoid -> image ellipse -> staged conic) and its
// reverse pass, one thread per (batch element, Gaussian), float64 registers.
//
// Follows the reference formula literally (mask/mask_gen_dynamic.py:240-309), including its treatment
// of a non-symmetric Sigma (the reference builds Sigma in float32, :400-401, so it is asymmetric at
// ~4e-8 and that moves mu2d by ~5e-7): float32 trig and float32 R = RX RY RZ, then float64
//   Sigma' = R Sigma R^T,  Q = inv(Sigma'),  m = R mu + t,  v = Q m^T,  s = m v - 1,
//   M = v v^T - s Q,  sign-flip of the z row/column,  2x2 minors and det(M)  ->  mu2d, sigma2d.
// The reverse pass is the hand-derived adjoint of exactly those steps, so d Sigma is the raw
// (non-symmetric) gradient TF's autodiff produces.  B*K threads of O(300) flops: launch-latency bound.
#include <cstdlib>
#include "gg_common.cuh"

namespace gg {

namespace {

constexpr double kLog2e = 1.4426950408889634074;

struct Cam {
  float c[3], s[3];       // float32 cos / sin of (rotx, roty, rotz), correctly rounded
  float RX[9], RY[9], RZ[9], RXY[9];
  double R[9];
};

__device__ __forceinline__ void mm32(const float *A, const float *Bm, float *C) {
  // pinned order ((p0 + p1) + p2), every product and sum rounded to float32 (oracle: matmul3_f32)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float p0 = __fmul_rn(A[i * 3 + 0], Bm[0 * 3 + j]);
      float p1 = __fmul_rn(A[i * 3 + 1], Bm[1 * 3 + j]);
      float p2 = __fmul_rn(A[i * 3 + 2], Bm[2 * 3 + j]);
      C[i * 3 + j] = __fadd_rn(__fadd_rn(p0, p1), p2);
    }
}

__device__ __forceinline__ float deg2rad_f32(float deg) {
  // `rot * np.pi / 180` on a float32 tensor: f32 multiply by f32(pi), then f32 divide (:240-242)
  return __fdiv_rn(__fmul_rn(deg, 3.14159274101257324219f), 180.0f);
}

__device__ void make_cam(float rx_deg, float ry_deg, float rz_deg, Cam &cam) {
  float a[3] = {deg2rad_f32(rx_deg), deg2rad_f32(ry_deg), deg2rad_f32(rz_deg)};
#pragma unroll
  for (int i = 0; i < 3; ++i) {   // correctly rounded float32 trig (oracle: cos_f32 / sin_f32)
    if (a[i] == 0.0f) {           // every reference call site passes rotx = rotz = 0 (:706-709): exact, no trig
      cam.c[i] = 1.0f;
      cam.s[i] = 0.0f;
    } else {
      double sd, cd;
      sincos((double)a[i], &sd, &cd);   // ~200 cycles each (tools/microbench/latency.cu)
      cam.c[i] = (float)cd;
      cam.s[i] = (float)sd;
    }
  }
  const float cx = cam.c[0], sx = cam.s[0], cy = cam.c[1], sy = cam.s[1], cz = cam.c[2], sz = cam.s[2];
  // inner tf.stack lists are matrix COLUMNS (:246-252)
  float RX[9] = {1.f, 0.f, 0.f, 0.f, cx, sx, 0.f, -sx, cx};
  float RY[9] = {cy, 0.f, -sy, 0.f, 1.f, 0.f, sy, 0.f, cy};
  float RZ[9] = {cz, sz, 0.f, -sz, cz, 0.f, 0.f, 0.f, 1.f};
  float Rf[9];
  mm32(RX, RY, cam.RXY);
  mm32(cam.RXY, RZ, Rf);
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    cam.RX[i] = RX[i]; cam.RY[i] = RY[i]; cam.RZ[i] = RZ[i];
    cam.R[i] = (double)Rf[i];
  }
}

__device__ __forceinline__ void mat3_mul(const double *A, const double *Bm, double *C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = A[i * 3 + 0] * Bm[0 * 3 + j] + A[i * 3 + 1] * Bm[1 * 3 + j] + A[i * 3 + 2] * Bm[2 * 3 + j];
}
__device__ __forceinline__ void mat3_mul_bt(const double *A, const double *Bm, double *C) {   // A * B^T
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = A[i * 3 + 0] * Bm[j * 3 + 0] + A[i * 3 + 1] * Bm[j * 3 + 1] + A[i * 3 + 2] * Bm[j * 3 + 2];
}
__device__ __forceinline__ void mat3_mul_at(const double *A, const double *Bm, double *C) {   // A^T * B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = A[0 * 3 + i] * Bm[0 * 3 + j] + A[1 * 3 + i] * Bm[1 * 3 + j] + A[2 * 3 + i] * Bm[2 * 3 + j];
}

// cofactor matrix C (C[i][j] = d det / d A[i][j]) and determinant
__device__ __forceinline__ double cofactors(const double *A, double *C) {
  C[0] = A[4] * A[8] - A[5] * A[7];
  C[1] = -(A[3] * A[8] - A[5] * A[6]);
  C[2] = A[3] * A[7] - A[4] * A[6];
  C[3] = -(A[1] * A[8] - A[2] * A[7]);
  C[4] = A[0] * A[8] - A[2] * A[6];
  C[5] = -(A[0] * A[7] - A[1] * A[6]);
  C[6] = A[1] * A[5] - A[2] * A[4];
  C[7] = -(A[0] * A[5] - A[2] * A[3]);
  C[8] = A[0] * A[4] - A[1] * A[3];
  return A[0] * C[0] + A[1] * C[1] + A[2] * C[2];
}

struct Proj {               // forward intermediates of one Gaussian
  double S[9];              // input Sigma
  double mu[3];
  double Sp[9], Q[9], m[3], v[3], s;
  double M[9];              // after the sign flip
  double d33, d31, d23, dM;
  double mu2d[2], sig2d[4];
};

__device__ void project_one(const Cam &cam, const double t[3], double f, Proj &p) {
  double RS[9];
  mat3_mul(cam.R, p.S, RS);
  mat3_mul_bt(RS, cam.R, p.Sp);                                   // :275
  double C[9];
  double det = cofactors(p.Sp, C);
  double idet = 1.0 / det;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) p.Q[i * 3 + j] = C[j * 3 + i] * idet;   // :276
#pragma unroll
  for (int i = 0; i < 3; ++i)
    p.m[i] = cam.R[i * 3 + 0] * p.mu[0] + cam.R[i * 3 + 1] * p.mu[1] + cam.R[i * 3 + 2] * p.mu[2] + t[i];  // :278
#pragma unroll
  for (int i = 0; i < 3; ++i)
    p.v[i] = p.Q[i * 3 + 0] * p.m[0] + p.Q[i * 3 + 1] * p.m[1] + p.Q[i * 3 + 2] * p.m[2];      // Q m^T (:280)
  p.s = p.m[0] * p.v[0] + p.m[1] * p.v[1] + p.m[2] * p.v[2] - 1.0;                               // :282
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double mij = p.v[i] * p.v[j] - p.s * p.Q[i * 3 + j];          // :280-285
      bool flip = (i == 2) != (j == 2);                             // :287-289
      p.M[i * 3 + j] = flip ? -mij : mij;
    }
  const double *M = p.M;
  p.d33 = M[0] * M[4] - M[1] * M[3];                                // M33 = rows{0,1} cols{0,1}
  p.d31 = M[1] * M[5] - M[2] * M[4];                                // M31 = rows{0,1} cols{1,2}
  p.d23 = M[0] * M[7] - M[1] * M[6];                                // M23 = rows{0,2} cols{0,1}
  double Cm[9];
  p.dM = cofactors(M, Cm);
  double i33 = 1.0 / p.d33;
  p.mu2d[0] = f * p.d31 * i33;                                      // :299-303 (px = py = 0)
  p.mu2d[1] = -f * p.d23 * i33;
  double q = -p.dM * i33 * i33;                                     // sigma2d = -(dM/d33) inv(M33) (:305-309)
  p.sig2d[0] = q * M[4];
  p.sig2d[1] = -q * M[1];
  p.sig2d[2] = -q * M[3];
  p.sig2d[3] = q * M[0];
}

// (mu2d, sigma2d) -> staged float32 record
__device__ void stage_params(const double mu[2], const double sg[4], float *out) {
  double det2 = sg[0] * sg[3] - sg[1] * sg[2];
  double idet = 1.0 / det2;
  double a = sg[3] * idet, b = -0.5 * (sg[1] + sg[2]) * idet, c = sg[0] * idet;   // P = inv(sigma2d), :131
  double r = b / a;
  double e = c - b * r;
  float4 lo = make_float4((float)mu[0], (float)mu[1], (float)(-a * kLog2e), (float)(-r));
  float4 hi = make_float4((float)(-e * kLog2e), (float)a, (float)b, (float)c);
  reinterpret_cast<float4 *>(out)[0] = lo;
  reinterpret_cast<float4 *>(out)[1] = hi;
}

// fixed-order float64 sum of one Gaussian's T per-tile partial moments (deterministic); loads are issued 8 tiles at
// a time so the T round trips to L2 overlap instead of serialising
__device__ __forceinline__ void sum_partials(const float *partials, int T, size_t tile_stride, double Tm[GG_MOMENTS]) {
#pragma unroll
  for (int j = 0; j < GG_MOMENTS; ++j) Tm[j] = 0.0;
  for (int t0 = 0; t0 < T; t0 += 8) {
    float v[8][GG_MOMENTS];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int j = 0; j < GG_MOMENTS; ++j)
        v[u][j] = (t0 + u < T) ? __ldg(partials + (size_t)(t0 + u) * tile_stride + j) : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int j = 0; j < GG_MOMENTS; ++j) Tm[j] += (double)v[u][j];
  }
}

// summed moments (shear coordinates t = dx + r dy) -> d mu2d, d sigma2d (accumulated into gmu, gsg), in two parts: what
// depends on the Gaussian alone (P = inv(sigma2d) and the float32-rounded shear the pixel kernel used) ...
struct ConicConsts {
  double P[4], a, b, c, r;
};
__device__ __forceinline__ void conic_consts(const double sg[4], ConicConsts &cc) {
  double det2 = sg[0] * sg[3] - sg[1] * sg[2];
  double idet = 1.0 / det2;
  cc.P[0] = sg[3] * idet; cc.P[1] = -sg[1] * idet; cc.P[2] = -sg[2] * idet; cc.P[3] = sg[0] * idet;
  cc.a = cc.P[0]; cc.b = 0.5 * (cc.P[1] + cc.P[2]); cc.c = cc.P[3];
  cc.r = -(double)((float)(-(cc.b / cc.a)));          // the float32-rounded shear the pixel kernel used
}
// ... and the map from the moments to the gradients
__device__ __forceinline__ void conic_apply(const ConicConsts &cc, const double Tm[GG_MOMENTS], double gmu[2], double gsg[4]) {
  const double *P = cc.P;
  const double a = cc.a, b = cc.b, c = cc.c, r = cc.r;
  double M20 = Tm[0] - 2.0 * r * Tm[1] + r * r * Tm[2];
  double M11 = Tm[1] - r * Tm[2];
  double M02 = Tm[2];
  double M10 = Tm[3] - r * Tm[4];
  double M01 = Tm[4];
  // L = sum Gbar g, g = exp(-d):  dL/dP = -[[M20, M11],[M11, M02]],  dL/dmu = 2 P_sym [M10, M01]
  gmu[0] += 2.0 * (a * M10 + b * M01);
  gmu[1] += 2.0 * (b * M10 + c * M01);
  double Pb[4] = {-M20, -M11, -M11, -M02};
  // sigma2d = inv(P)  =>  d sigma2d = -P^T Pbar P^T
  double PtPb[4] = {P[0] * Pb[0] + P[2] * Pb[2], P[0] * Pb[1] + P[2] * Pb[3],
                    P[1] * Pb[0] + P[3] * Pb[2], P[1] * Pb[1] + P[3] * Pb[3]};
  gsg[0] -= PtPb[0] * P[0] + PtPb[1] * P[1];
  gsg[1] -= PtPb[0] * P[2] + PtPb[1] * P[3];
  gsg[2] -= PtPb[2] * P[0] + PtPb[3] * P[1];
  gsg[3] -= PtPb[2] * P[2] + PtPb[3] * P[3];
}
__device__ void conic_adjoint(const double mu[2], const double sg[4], const double Tm[GG_MOMENTS], double gmu[2],
                              double gsg[4]) {
  ConicConsts cc;
  conic_consts(sg, cc);
  conic_apply(cc, Tm, gmu, gsg);
}

__device__ __forceinline__ void load_sigma(const void *sigma, int is_f64, size_t idx, double *S) {
  if (is_f64) {
    const double *p = reinterpret_cast<const double *>(sigma) + idx * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) S[i] = p[i];
  } else {
    const float *p = reinterpret_cast<const float *>(sigma) + idx * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) S[i] = (double)p[i];
  }
}

// Adjoint of project_one for ONE Gaussian: (d mu2d, d sigma2d) -> d Sigma (raw, non-symmetric), d m (camera frame)
// and this Gaussian's contribution to d R.  Linear in (gmu, gsg).
__device__ __forceinline__ void project_adjoint_one(const Cam &cam, const Proj &p, double f, const double gmu[2],
                                                    const double gsg[4], double Sbar[9], double mbar[3],
                                                    double Rbar_add[9]) {
  // ---- adjoint of: sigma2d = q * adj(M33), q = -dM / d33^2; mu2d = f (d31, -d23) / d33
  const double *M = p.M;
  double Mb[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double i33 = 1.0 / p.d33;
  double q = -p.dM * i33 * i33;
  double qbar = gsg[0] * M[4] - gsg[1] * M[1] - gsg[2] * M[3] + gsg[3] * M[0];
  Mb[4] += q * gsg[0]; Mb[1] -= q * gsg[1]; Mb[3] -= q * gsg[2]; Mb[0] += q * gsg[3];
  double dMbar = -qbar * i33 * i33;
  double d33bar = 2.0 * p.dM * i33 * i33 * i33 * qbar;
  double d31bar = f * gmu[0] * i33;
  double d23bar = -f * gmu[1] * i33;
  d33bar += -f * p.d31 * gmu[0] * i33 * i33 + f * p.d23 * gmu[1] * i33 * i33;
  Mb[0] += d33bar * M[4]; Mb[4] += d33bar * M[0]; Mb[1] -= d33bar * M[3]; Mb[3] -= d33bar * M[1];
  Mb[1] += d31bar * M[5]; Mb[5] += d31bar * M[1]; Mb[2] -= d31bar * M[4]; Mb[4] -= d31bar * M[2];
  Mb[0] += d23bar * M[7]; Mb[7] += d23bar * M[0]; Mb[1] -= d23bar * M[6]; Mb[6] -= d23bar * M[1];
  double Cm[9];
  cofactors(M, Cm);
#pragma unroll
  for (int i = 0; i < 9; ++i) Mb[i] += dMbar * Cm[i];
  // undo the sign flip
  Mb[2] = -Mb[2]; Mb[5] = -Mb[5]; Mb[6] = -Mb[6]; Mb[7] = -Mb[7];
  // ---- M = v v^T - s Q
  double vbar[3], Qbar[9], sbar = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    vbar[i] = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      vbar[i] += (Mb[i * 3 + j] + Mb[j * 3 + i]) * p.v[j];
      sbar -= Mb[i * 3 + j] * p.Q[i * 3 + j];
      Qbar[i * 3 + j] = -p.s * Mb[i * 3 + j];
    }
  }
  // ---- s = m Q m^T - 1 ;  v = Q m^T
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    mbar[i] = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      mbar[i] += sbar * (p.Q[i * 3 + j] + p.Q[j * 3 + i]) * p.m[j] + p.Q[j * 3 + i] * vbar[j];
      Qbar[i * 3 + j] += sbar * p.m[i] * p.m[j] + vbar[i] * p.m[j];
    }
  }
  // ---- Q = inv(Sigma')  =>  Sigma'bar = -Q^T Qbar Q^T
  double tmp[9], Spbar[9];
  mat3_mul_at(p.Q, Qbar, tmp);
  mat3_mul_bt(tmp, p.Q, Spbar);
#pragma unroll
  for (int i = 0; i < 9; ++i) Spbar[i] = -Spbar[i];
  // ---- Sigma' = R Sigma R^T ;  m = R mu + t
  mat3_mul_at(cam.R, Spbar, tmp);
  mat3_mul(tmp, cam.R, Sbar);                                // R^T Sigma'bar R
  // d R += Sigma'bar R Sigma^T + Sigma'bar^T R Sigma + mbar mu^T
  double RS[9], RSt[9];
  mat3_mul(cam.R, p.S, RS);
  mat3_mul_bt(cam.R, p.S, RSt);
  double t1[9], t2[9];
  mat3_mul(Spbar, RSt, t1);
  mat3_mul_at(Spbar, RS, t2);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Rbar_add[i * 3 + j] = t1[i * 3 + j] + t2[i * 3 + j] + mbar[i] * p.mu[j];
}

// Adjoint of R = RX RY RZ (float32 in the forward; evaluated in float64 on the same float32 values): d R -> d angles
// (radians).  Linear in Rbar.
__device__ __forceinline__ void rotation_adjoint(const Cam &cam, const double Rbar[9], double &gx, double &gy, double &gz) {
  double RXd[9], RYd[9], RZd[9], RXYd[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { RXd[i] = cam.RX[i]; RYd[i] = cam.RY[i]; RZd[i] = cam.RZ[i]; RXYd[i] = cam.RXY[i]; }
  double RYZ[9], RXbar[9], RYbar[9], RZbar[9], tmp[9];
  mat3_mul(RYd, RZd, RYZ);
  mat3_mul_bt(Rbar, RYZ, RXbar);            // Rbar (RY RZ)^T
  mat3_mul_at(RXd, Rbar, tmp);
  mat3_mul_bt(tmp, RZd, RYbar);             // RX^T Rbar RZ^T
  mat3_mul_at(RXYd, Rbar, RZbar);           // (RX RY)^T Rbar
  const double cx = cam.c[0], sx = cam.s[0], cy = cam.c[1], sy = cam.s[1], cz = cam.c[2], sz = cam.s[2];
  // dRX/da = [[0,0,0],[0,-s,c],[0,-c,-s]], dRY/da = [[-s,0,-c],[0,0,0],[c,0,-s]], dRZ/da = [[-s,c,0],[-c,-s,0],[0,0,0]]
  gx = -sx * RXbar[4] + cx * RXbar[5] - cx * RXbar[7] - sx * RXbar[8];
  gy = -sy * RYbar[0] - cy * RYbar[2] + cy * RYbar[6] - sy * RYbar[8];
  gz = -sz * RZbar[0] + cz * RZbar[1] - cz * RZbar[3] - sz * RZbar[4];
}

}  // namespace

__global__ void __launch_bounds__(128) project_fwd_kernel(
    const float *__restrict__ mus, const void *__restrict__ sigma, int sigma_is_f64,
    const float *__restrict__ rotx, const float *__restrict__ roty, const float *__restrict__ rotz,
    double tx, double ty, double tz, double focal, int B, int K, int Bg,
    double *__restrict__ mu2d, double *__restrict__ sigma2d, float *__restrict__ params) {
  // B rendered images over Bg <= B sets of Gaussians: image b uses set b % Bg (multi-view: V = B / Bg cameras per set,
  // view-major; Bg == B is the ordinary one camera per set)
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * K) return;
  int b = idx / K;
  const size_t gidx = (size_t)(b % Bg) * K + (idx - b * K);
  // issue every global load before the first use (one round trip instead of two)
  const float rxd = __ldg(rotx + b), ryd = __ldg(roty + b), rzd = __ldg(rotz + b);
  Proj p;
  load_sigma(sigma, sigma_is_f64, gidx, p.S);
#pragma unroll
  for (int i = 0; i < 3; ++i) p.mu[i] = (double)__ldg(mus + gidx * 3 + i);
  Cam cam;
  make_cam(rxd, ryd, rzd, cam);
  const double t[3] = {tx, ty, tz};
  project_one(cam, t, (double)(float)focal, p);   // focal passes through float32 (:266, :271)
  if (mu2d) { mu2d[(size_t)idx * 2] = p.mu2d[0]; mu2d[(size_t)idx * 2 + 1] = p.mu2d[1]; }
  if (sigma2d) {
#pragma unroll
    for (int i = 0; i < 4; ++i) sigma2d[(size_t)idx * 4 + i] = p.sig2d[i];
  }
  stage_params(p.mu2d, p.sig2d, params + (size_t)idx * GG_PARAM_STRIDE);
}

// One CTA (128 threads) per batch element, Gaussians in chunks of 128.  Per chunk: (1) all threads sum the
// per-tile partial moments column by column ([T][K][5] in global memory: coalesced, T independent loads in flight
// per thread, fixed order) into shared memory, while the Gaussians' own inputs are already on their way;
// (2) one thread per Gaussian runs the float64 adjoint.  d(angles) is reduced over K in shared memory in a
// fixed order.  The kernel is a dependent float64 chain (~900 DFMA-class instructions per Gaussian at ~8 cycles):
// launch-latency bound, so the point of (1) is to take the global round trips and the T-sum off that chain.
__global__ void __launch_bounds__(128) project_bwd_kernel(
    const float *__restrict__ mus, const void *__restrict__ sigma, int sigma_is_f64,
    const float *__restrict__ rotx, const float *__restrict__ roty, const float *__restrict__ rotz,
    double tx, double ty, double tz, double focal, int B, int K, int Bg,
    const float *__restrict__ partials, int T, const double *__restrict__ gmu2d_in,
    const double *__restrict__ gsigma2d_in, float *__restrict__ dmus, double *__restrict__ dsigma,
    float *__restrict__ drot) {
  __shared__ double red[3][128];
  __shared__ double Tm_s[128 * GG_MOMENTS];
  // One CTA per SET of Gaussians; its V = B / Bg views (images bset, bset + Bg, ...) are handled one after the other:
  // d mus / d Sigma accumulate over the views (the Gaussians are shared), d rot is per view.
  //
  // Programmatic dependent launch: the Gaussians' own inputs (mus, sigma, rot*) are never written by the kernel that
  // precedes this one in a stream (that is a splat / fused-loss kernel; they were last written at least two launches
  // back), so warp 0 loads them and recomputes the whole forward projection BEFORE griddepcontrol.wait, i.e. while the
  // splat backward is still draining; only the partial sums and the adjoint proper run after it.  The other warps go
  // straight to the wait and then sum the partials' columns.
  const int bset = blockIdx.x;
  const int nviews = B / Bg;
  const bool colwarp = threadIdx.x >= 32;
  const int ncol = (int)blockDim.x - 32;           // threads that sum columns
  bool waited = false;
  const double t[3] = {tx, ty, tz};
  const double f = (double)(float)focal;
  for (int view = 0; view < nviews; ++view) {
  const int b = view * Bg + bset;
  const float rxd = __ldg(rotx + b), ryd = __ldg(roty + b), rzd = __ldg(rotz + b);
  double Rbar[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  Cam cam;
  bool have_cam = false;

  for (int kb = 0; kb < K; kb += 128) {
    const int kn = min(128, K - kb);
    const bool active = (int)threadIdx.x < kn;
    // Gaussians of this chunk: thread i takes Gaussian kb + i, in rounds of blockDim.x when the CTA is smaller than
    // the chunk (never the case for the launch shapes used: 64 threads for K <= 32, 128 otherwise)
    const int k = kb + (int)threadIdx.x;
    const size_t idx = (size_t)b * K + (active ? k : kb);             // per-image quantities (partials, gmu2d, gsigma2d)
    const size_t gidx = (size_t)bset * K + (active ? k : kb);         // the Gaussian itself and its gradients
    Proj p;
    load_sigma(sigma, sigma_is_f64, gidx, p.S);
#pragma unroll
    for (int i = 0; i < 3; ++i) p.mu[i] = (double)__ldg(mus + gidx * 3 + i);
    const bool have_partials = partials && T > 0;
    if (colwarp) {
      if (!waited) { pdl_enter(); waited = true; }
      if (have_partials) {
        const float *base = partials + ((size_t)b * T * K + kb) * GG_MOMENTS;
        const size_t tile_stride = (size_t)K * GG_MOMENTS;
        for (int c = (int)threadIdx.x - 32; c < kn * GG_MOMENTS; c += ncol) {
          double acc = 0.0;
          for (int t0 = 0; t0 < T; t0 += 16) {     // 16 independent loads in flight: one round trip for T <= 16
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = (t0 + u < T) ? __ldg(base + (size_t)(t0 + u) * tile_stride + c) : 0.f;
#pragma unroll
            for (int u = 0; u < 16; ++u) acc += (double)v[u];
          }
          Tm_s[c] = acc;
        }
      }
    }
    if (!have_cam) { make_cam(rxd, ryd, rzd, cam); have_cam = true; }
    if (active) project_one(cam, t, f, p);
    if (!waited) { pdl_enter(); waited = true; }
    double gmu[2] = {0, 0}, gsg[4] = {0, 0, 0, 0};
    if (gmu2d_in) { gmu[0] = gmu2d_in[idx * 2]; gmu[1] = gmu2d_in[idx * 2 + 1]; }
    if (gsigma2d_in) {
#pragma unroll
      for (int i = 0; i < 4; ++i) gsg[i] = gsigma2d_in[idx * 4 + i];
    }
    __syncthreads();
    if (active) {
    if (have_partials) {
      double Tm[GG_MOMENTS];
#pragma unroll
      for (int j = 0; j < GG_MOMENTS; ++j) Tm[j] = Tm_s[threadIdx.x * GG_MOMENTS + j];
      conic_adjoint(p.mu2d, p.sig2d, Tm, gmu, gsg);
    }

    double Sbar[9], mbar[3], Radd[9];
    project_adjoint_one(cam, p, f, gmu, gsg, Sbar, mbar, Radd);
    // the same thread owns this Gaussian in every view: plain read-modify-write, fixed view order
    if (dsigma) {
#pragma unroll
      for (int i = 0; i < 9; ++i) dsigma[gidx * 9 + i] = view == 0 ? Sbar[i] : dsigma[gidx * 9 + i] + Sbar[i];
    }
    if (dmus) {
      // float32, like the sum of the V per-view float32 gradients autograd would form
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float g = (float)(cam.R[0 * 3 + i] * mbar[0] + cam.R[1 * 3 + i] * mbar[1] + cam.R[2 * 3 + i] * mbar[2]);
        dmus[gidx * 3 + i] = view == 0 ? g : dmus[gidx * 3 + i] + g;
      }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) Rbar[i] += Radd[i];
    }   // active
    __syncthreads();   // Tm_s is rewritten by the next chunk
  }

  double gx, gy, gz;
  rotation_adjoint(cam, Rbar, gx, gy, gz);
  red[0][threadIdx.x] = gx; red[1][threadIdx.x] = gy; red[2][threadIdx.x] = gz;
  __syncthreads();
  if (threadIdx.x < 3 && drot) {
    double acc = 0.0;
    int nthr = min((int)blockDim.x, K);
    for (int i = 0; i < nthr; ++i) acc += red[threadIdx.x][i];
    // degrees -> radians was (deg * pi32) / 180 in float32
    drot[(size_t)b * 3 + threadIdx.x] = (float)(acc * (double)3.14159274101257324219f / 180.0);
  }
  __syncthreads();   // red[] is reused by the next view
  }   // views
}

// The same adjoint for the common shape -- one camera per set of Gaussians and 6 K <= 128 -- with the dependent float64
// chain taken OFF the critical path.  The adjoint is linear in its six inputs (d mu2d [2], d sigma2d [4]), and those are
// the only quantities that depend on the splat backward's partial sums.  So thread (seed i, Gaussian k) recomputes the
// forward projection of k and runs the whole adjoint for the unit input e_i BEFORE griddepcontrol.wait -- under the
// tail of the splat backward, where the chain measured 2.8 us of free time beyond the forward recompute
// (tools/chain_parts.py and the pre-wait spin experiment in DESIGN.md) -- leaving the 6 x 15 Jacobian block of k in
// shared memory (9 d Sigma, 3 d mu, 3 d angle).  After the wait: the T partial sums per column (one round of loads),
// the 2 x 2 conic map to the six inputs, and 15 dot products of length 6 per Gaussian.  Same inputs contract as
// project_bwd_kernel (mus / sigma / rot* at least two launches old).
constexpr int kLinOut = 15;
__global__ void __launch_bounds__(128) project_bwd_lin_kernel(
    const float *__restrict__ mus, const void *__restrict__ sigma, int sigma_is_f64,
    const float *__restrict__ rotx, const float *__restrict__ roty, const float *__restrict__ rotz,
    double tx, double ty, double tz, double focal, int B, int K,
    const float *__restrict__ partials, int T, const double *__restrict__ gmu2d_in,
    const double *__restrict__ gsigma2d_in, float *__restrict__ dmus, double *__restrict__ dsigma,
    float *__restrict__ drot) {
  extern __shared__ double lin_sm[];
  double *J = lin_sm;                        // [K][6][kLinOut]
  double *gs = J + (size_t)K * 6 * kLinOut;   // [K][6]   the six inputs of every Gaussian
  double *Tm_s = gs + (size_t)K * 6;          // [K][5]   summed moments
  double *red = Tm_s + (size_t)K * GG_MOMENTS;   // [3][K]   d angle per Gaussian
  const int b = blockIdx.x;
  const int tid = (int)threadIdx.x;
  const int seed = tid / K, k = tid - seed * K;
  const bool active = tid < 6 * K;
  const size_t gidx = (size_t)b * K + (active ? k : 0);
  const double t[3] = {tx, ty, tz};
  const double f = (double)(float)focal;

  // ---- before the wait: forward recompute and the adjoint of one unit input
  const float rxd = __ldg(rotx + b), ryd = __ldg(roty + b), rzd = __ldg(rotz + b);
  Proj p;
  load_sigma(sigma, sigma_is_f64, gidx, p.S);
#pragma unroll
  for (int i = 0; i < 3; ++i) p.mu[i] = (double)__ldg(mus + gidx * 3 + i);
  Cam cam;
  make_cam(rxd, ryd, rzd, cam);
  ConicConsts cc;
  if (active) {
    project_one(cam, t, f, p);
    conic_consts(p.sig2d, cc);
    double gmu[2] = {seed == 0 ? 1.0 : 0.0, seed == 1 ? 1.0 : 0.0};
    double gsg[4] = {seed == 2 ? 1.0 : 0.0, seed == 3 ? 1.0 : 0.0, seed == 4 ? 1.0 : 0.0, seed == 5 ? 1.0 : 0.0};
    double Sbar[9], mbar[3], Radd[9], gx, gy, gz;
    project_adjoint_one(cam, p, f, gmu, gsg, Sbar, mbar, Radd);
    rotation_adjoint(cam, Radd, gx, gy, gz);
    double *Jk = J + ((size_t)k * 6 + seed) * kLinOut;
#pragma unroll
    for (int i = 0; i < 9; ++i) Jk[i] = Sbar[i];
#pragma unroll
    for (int i = 0; i < 3; ++i)
      Jk[9 + i] = cam.R[0 * 3 + i] * mbar[0] + cam.R[1 * 3 + i] * mbar[1] + cam.R[2 * 3 + i] * mbar[2];
    Jk[12] = gx; Jk[13] = gy; Jk[14] = gz;
  }
  pdl_enter();

  // ---- after the wait
  const bool have_partials = partials && T > 0;
  if (have_partials) {
    const float *base = partials + (size_t)b * T * K * GG_MOMENTS;
    const size_t tile_stride = (size_t)K * GG_MOMENTS;
    for (int c = tid; c < K * GG_MOMENTS; c += (int)blockDim.x) {
      double acc = 0.0;
      for (int t0 = 0; t0 < T; t0 += 16) {     // 16 independent loads in flight: one round trip for T <= 16
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = (t0 + u < T) ? __ldg(base + (size_t)(t0 + u) * tile_stride + c) : 0.f;
#pragma unroll
        for (int u = 0; u < 16; ++u) acc += (double)v[u];
      }
      Tm_s[c] = acc;
    }
  }
  double gmu[2] = {0, 0}, gsg[4] = {0, 0, 0, 0};
  if (tid < K) {
    if (gmu2d_in) { gmu[0] = gmu2d_in[gidx * 2]; gmu[1] = gmu2d_in[gidx * 2 + 1]; }
    if (gsigma2d_in) {
#pragma unroll
      for (int i = 0; i < 4; ++i) gsg[i] = gsigma2d_in[gidx * 4 + i];
    }
  }
  __syncthreads();
  if (tid < K) {
    if (have_partials) {
      double Tm[GG_MOMENTS];
#pragma unroll
      for (int j = 0; j < GG_MOMENTS; ++j) Tm[j] = Tm_s[tid * GG_MOMENTS + j];
      conic_apply(cc, Tm, gmu, gsg);
    }
    gs[tid * 6 + 0] = gmu[0]; gs[tid * 6 + 1] = gmu[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) gs[tid * 6 + 2 + i] = gsg[i];
  }
  __syncthreads();
  for (int o = tid; o < K * kLinOut; o += (int)blockDim.x) {
    const int kk = o / kLinOut, j = o - kk * kLinOut;
    const double *Jk = J + (size_t)kk * 6 * kLinOut + j;
    const double *g = gs + kk * 6;
    double acc = Jk[0] * g[0];
#pragma unroll
    for (int i = 1; i < 6; ++i) acc += Jk[i * kLinOut] * g[i];
    const size_t go = (size_t)b * K + kk;
    if (j < 9) { if (dsigma) dsigma[go * 9 + j] = acc; }
    else if (j < 12) { if (dmus) dmus[go * 3 + (j - 9)] = (float)acc; }
    else red[(j - 12) * K + kk] = acc;
  }
  __syncthreads();
  if (tid < 3 && drot) {
    double acc = 0.0;
    for (int i = 0; i < K; ++i) acc += red[tid * K + i];
    // degrees -> radians was (deg * pi32) / 180 in float32
    drot[(size_t)b * 3 + tid] = (float)(acc * (double)3.14159274101257324219f / 180.0);
  }
}

__global__ void __launch_bounds__(128) conic_fwd_kernel(const void *__restrict__ mu2d,
                                                        const void *__restrict__ sigma2d, int is_f64,
                                                        int N, float *__restrict__ params) {
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N) return;
  double mu[2], sg[4];
  if (is_f64) {
    const double *m = reinterpret_cast<const double *>(mu2d) + (size_t)idx * 2;
    const double *s = reinterpret_cast<const double *>(sigma2d) + (size_t)idx * 4;
    mu[0] = m[0]; mu[1] = m[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) sg[i] = s[i];
  } else {
    const float *m = reinterpret_cast<const float *>(mu2d) + (size_t)idx * 2;
    const float *s = reinterpret_cast<const float *>(sigma2d) + (size_t)idx * 4;
    mu[0] = m[0]; mu[1] = m[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) sg[i] = s[i];
  }
  stage_params(mu, sg, params + (size_t)idx * GG_PARAM_STRIDE);
}

__global__ void __launch_bounds__(128) conic_bwd_kernel(const void *__restrict__ mu2d,
                                                        const void *__restrict__ sigma2d, int is_f64,
                                                        int B, int K, const float *__restrict__ partials,
                                                        int T, double *__restrict__ dmu2d,
                                                        double *__restrict__ dsigma2d) {
  pdl_enter();   // programmatic dependent launch: see gg_common.cuh
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * K) return;
  int b = idx / K, k = idx - b * K;
  double mu[2], sg[4];
  if (is_f64) {
    const double *m = reinterpret_cast<const double *>(mu2d) + (size_t)idx * 2;
    const double *s = reinterpret_cast<const double *>(sigma2d) + (size_t)idx * 4;
    mu[0] = m[0]; mu[1] = m[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) sg[i] = s[i];
  } else {
    const float *m = reinterpret_cast<const float *>(mu2d) + (size_t)idx * 2;
    const float *s = reinterpret_cast<const float *>(sigma2d) + (size_t)idx * 4;
    mu[0] = m[0]; mu[1] = m[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) sg[i] = s[i];
  }
  double gmu[2] = {0, 0}, gsg[4] = {0, 0, 0, 0};
  double Tm[GG_MOMENTS];
  sum_partials(partials + ((size_t)b * T * K + k) * GG_MOMENTS, T, (size_t)K * GG_MOMENTS, Tm);
  conic_adjoint(mu, sg, Tm, gmu, gsg);
  if (dmu2d) { dmu2d[(size_t)idx * 2] = gmu[0]; dmu2d[(size_t)idx * 2 + 1] = gmu[1]; }
  if (dsigma2d) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dsigma2d[(size_t)idx * 4 + i] = gsg[i];
  }
}

// ------------------------------------------------------------------------------ host launchers
cudaError_t launch_project_fwd(const float *mus, const void *sigma, int f64, const float *rx, const float *ry,
                               const float *rz, double tx, double ty, double tz, double focal, int B, int K,
                               double *mu2d, double *sigma2d, float *params, cudaStream_t st, int nviews) {
  int N = B * K;
  // small launches: 32 threads per CTA spread the Gaussians over more SMs (every thread is one dependent chain)
  static const int env_threads = [] { const char *e = std::getenv("GG_PROJ_FWD_THREADS"); return e ? std::atoi(e) : 0; }();
  const int threads = env_threads > 0 ? env_threads : (N <= 148 * 32 ? 32 : 128);
  launch_pdl(project_fwd_kernel, dim3((N + threads - 1) / threads), dim3(threads), 0, st, mus, sigma, f64, rx, ry, rz, tx,
             ty, tz, focal, B, K, B / nviews, mu2d, sigma2d, params);
  return cudaGetLastError();
}

cudaError_t launch_project_bwd(const float *mus, const void *sigma, int f64, const float *rx, const float *ry,
                               const float *rz, double tx, double ty, double tz, double focal, int B, int K,
                               const float *partials, int T, const double *gmu2d, const double *gsigma2d,
                               float *dmus, double *dsigma, float *drot, cudaStream_t st, int nviews) {
  // 128 threads: warp 0 holds the Gaussians of a K <= 32 model, the other three sum the partials' columns in one round
  // of loads (measured on the C2 step chain, tools/time_step.py: 128 threads 44.1 us/step, 64 threads 46.2)
  static const int env_threads = [] { const char *e = std::getenv("GG_PROJ_BWD_THREADS"); return e ? std::atoi(e) : 0; }();
  const char *lin_env = std::getenv("GG_PROJ_BWD_LIN");      // read per call: the tests compare the two kernels
  const bool lin = !lin_env || lin_env[0] != '0';
  if (lin && nviews == 1 && 6 * K <= 128) {
    // one camera per set, at most 21 Gaussians: the linearised adjoint (Jacobian before the wait)
    const size_t smem = (size_t)K * (6 * kLinOut + 6 + GG_MOMENTS + 3) * sizeof(double);
    launch_pdl(project_bwd_lin_kernel, dim3(B), dim3(128), smem, st, mus, sigma, f64, rx, ry, rz, tx, ty, tz, focal, B, K,
               partials, T, gmu2d, gsigma2d, dmus, dsigma, drot);
    return cudaGetLastError();
  }
  const int threads = env_threads > 0 ? env_threads : 128;
  launch_pdl(project_bwd_kernel, dim3(B / nviews), dim3(threads), 0, st, mus, sigma, f64, rx, ry, rz, tx, ty, tz, focal, B,
             K, B / nviews, partials, T, gmu2d, gsigma2d, dmus, dsigma, drot);
  return cudaGetLastError();
}

cudaError_t launch_conic_fwd(const void *mu2d, const void *sigma2d, int f64, int N, float *params, cudaStream_t st) {
  launch_pdl(conic_fwd_kernel, dim3((N + 127) / 128), dim3(128), 0, st, mu2d, sigma2d, f64, N, params);
  return cudaGetLastError();
}

cudaError_t launch_conic_bwd(const void *mu2d, const void *sigma2d, int f64, int B, int K, const float *partials,
                             int T, double *dmu2d, double *dsigma2d, cudaStream_t st) {
  int N = B * K;
  launch_pdl(conic_bwd_kernel, dim3((N + 127) / 128), dim3(128), 0, st, mu2d, sigma2d, f64, B, K, partials, T, dmu2d, dsigma2d);
  return cudaGetLastError();
}

}  // namespace gg

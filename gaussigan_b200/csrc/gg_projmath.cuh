// gg_projmath.cuh -- the float64 projection of one Gaussian and its adjoint as device functions, shared by the projection
// kernels (gg_project.cu) and the splat backward that runs the adjoint from its last-arriving CTA (gg_splat_fast.cu).
// Literal transcription of mask/mask_gen_dynamic.py:240-309 and its hand-derived reverse pass; see gg_project.cu.
#pragma once
#include "gg_common.cuh"

namespace gg {
namespace projmath {

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

static __device__ void make_cam(float rx_deg, float ry_deg, float rz_deg, Cam &cam) {
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

static __device__ void project_one(const Cam &cam, const double t[3], double f, Proj &p) {
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
static __device__ void stage_params(const double mu[2], const double sg[4], float *out) {
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
static __device__ void conic_adjoint(const double mu[2], const double sg[4], const double Tm[GG_MOMENTS], double gmu[2],
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


}  // namespace projmath
}  // namespace gg

// FP64 camera / SE3 helpers for the tracking kernels.  Same expressions, in the same order, as the
// reference's scalar code (cited per function); the translation unit is built with -fmad=false so no
// multiply-add is contracted.  atan is the correctly-rounded atan_cr (atan_dd.cuh); sin/cos/asin/acos of the
// per-stream pose algebra are CUDA's (<= 2 ulp from the host libm).
#pragma once
#include "vslam_internal.cuh"
#include "atan_dd.cuh"

struct CamCache { double x, y, r, factor; bool invalid; };

// ATANCamera::Project (jni/ATANCamera.cc:133-145) + rtrans_factor (jni/ATANCamera.h:136-142)
__device__ __forceinline__ void cam_project(const CamDev& c, double x, double y, double* im, CamCache& cc) {
  cc.x = x; cc.y = y;
  cc.r = sqrt(x * x + y * y);
  cc.invalid = cc.r > c.maxR;
  cc.factor = (cc.r < 0.001 || c.W == 0.0) ? 1.0 : (c.Winv * atan_cr(cc.r * c.twoTan) / cc.r);
  const double dx = x * cc.factor, dy = y * cc.factor;
  im[0] = c.cx + c.fx * dx;
  im[1] = c.cy + c.fy * dy;
}
// ATANCamera::GetProjectionDerivs_Eigen (jni/ATANCamera.cc:198-231); d row-major
__device__ __forceinline__ void cam_derivs(const CamDev& c, const CamCache& cc, double* d) {
  double fx_, fy_;
  const double k = c.twoTan, x = cc.x, y = cc.y, r = cc.r * c.distEnabled;
  if (r < 0.01) { fx_ = 0.0; fy_ = 0.0; }
  else {
    fx_ = c.Winv * (k * x) / (r * r * (1 + k * k * r * r)) - x * cc.factor / (r * r);
    fy_ = c.Winv * (k * y) / (r * r * (1 + k * k * r * r)) - y * cc.factor / (r * r);
  }
  d[0] = c.fx * (fx_ * x + cc.factor);
  d[2] = c.fy * (fx_ * y);
  d[1] = c.fx * (fy_ * x);
  d[3] = c.fy * (fy_ * y + cc.factor);
}

// pose: row-major 3x4
__device__ __forceinline__ void se3_apply(const double* P, const double* v, double* o) {
#pragma unroll
  for (int i = 0; i < 3; i++) { double s = P[4 * i] * v[0]; s += P[4 * i + 1] * v[1]; s += P[4 * i + 2] * v[2]; o[i] = P[4 * i + 3] + s; }
}
__device__ __forceinline__ void rot_apply(const double* P, const double* v, double* o) {
#pragma unroll
  for (int i = 0; i < 3; i++) { double s = P[4 * i] * v[0]; s += P[4 * i + 1] * v[1]; s += P[4 * i + 2] * v[2]; o[i] = s; }
}
// mySE3::operator* (jni/RT.h:275-282): out = a * b
__device__ inline void se3_mul(const double* a, const double* b, double* o) {
  double r[12];
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) { double s = a[4 * i] * b[j]; s += a[4 * i + 1] * b[4 + j]; s += a[4 * i + 2] * b[8 + j]; r[4 * i + j] = s; }
    double s = a[4 * i] * b[3]; s += a[4 * i + 1] * b[7]; s += a[4 * i + 2] * b[11]; r[4 * i + 3] = a[4 * i + 3] + s;
  }
  for (int i = 0; i < 12; i++) o[i] = r[i];
}
// mySE3::inverse (jni/RT.h:262-270)
__device__ inline void se3_inverse(const double* a, double* o) {
  double r[12];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r[4 * i + j] = a[4 * j + i];
  for (int i = 0; i < 3; i++) { double s = r[4 * i] * a[3]; s += r[4 * i + 1] * a[7]; s += r[4 * i + 2] * a[11]; r[4 * i + 3] = -s; }
  for (int i = 0; i < 12; i++) o[i] = r[i];
}
__device__ __forceinline__ void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double* a, const double* b) { double s = 0; s += a[0] * b[0]; s += a[1] * b[1]; s += a[2] * b[2]; return s; }
// rodrigues_so3_exp (jni/RT.h:98-127) into the rotation part of a 3x4
__device__ inline void rodrigues(const double* w, double A, double B, double* P) {
  const double wx2 = w[0] * w[0], wy2 = w[1] * w[1], wz2 = w[2] * w[2];
  P[0] = 1.0 - B * (wy2 + wz2); P[5] = 1.0 - B * (wx2 + wz2); P[10] = 1.0 - B * (wx2 + wy2);
  { const double a = A * w[2], b = B * (w[0] * w[1]); P[1] = b - a; P[4] = b + a; }
  { const double a = A * w[1], b = B * (w[0] * w[2]); P[2] = b + a; P[8] = b - a; }
  { const double a = A * w[0], b = B * (w[1] * w[2]); P[6] = b - a; P[9] = b + a; }
}
// mySE3::exp (jni/RT.h:318-352)
__device__ inline void se3_exp(const double* mu, double* P) {
  const double one_6th = 1.0 / 6.0, one_20th = 1.0 / 20.0;
  const double* w = mu + 3;
  const double theta_sq = dot3(w, w), theta = sqrt(theta_sq);
  double A, B, cr[3];
  cross3(w, mu, cr);
  if (theta_sq < 1e-8) {
    A = 1.0 - one_6th * theta_sq; B = 0.5;
    for (int i = 0; i < 3; i++) P[4 * i + 3] = mu[i] + 0.5 * cr[i];
  } else {
    double C;
    if (theta_sq < 1e-6) { C = one_6th * (1.0 - one_20th * theta_sq); A = 1.0 - theta_sq * C; B = 0.5 - 0.25 * one_6th * theta_sq; }
    else { const double it = 1.0 / theta; A = sin(theta) * it; B = (1 - cos(theta)) * (it * it); C = (1 - A) * (it * it); }
    double wc[3]; cross3(w, cr, wc);
    for (int i = 0; i < 3; i++) P[4 * i + 3] = (mu[i] + B * cr[i]) + C * wc[i];
  }
  rodrigues(w, A, B, P);
}
// mySO3::exp (jni/RT.h:132-164), 3x3 row-major
__device__ inline void so3_exp(const double* w, double* R) {
  const double one_6th = 1.0 / 6.0, one_20th = 1.0 / 20.0;
  const double theta_sq = dot3(w, w), theta = sqrt(theta_sq);
  double A, B;
  if (theta_sq < 1e-8) { A = 1.0 - one_6th * theta_sq; B = 0.5; }
  else if (theta_sq < 1e-6) { B = 0.5 - 0.25 * one_6th * theta_sq; A = 1.0 - theta_sq * one_6th * (1.0 - one_20th * theta_sq); }
  else { const double it = 1.0 / theta; A = sin(theta) * it; B = (1 - cos(theta)) * (it * it); }
  double P[12]; rodrigues(w, A, B, P);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[3 * i + j] = P[4 * i + j];
}
// mySO3::ln (jni/RT.h:166-215) on the rotation part of a 3x4
__device__ inline void so3_ln(const double* P, double* result) {
  const double cos_angle = (P[0] + P[5] + P[10] - 1.0) * 0.5;
  result[0] = (P[9] - P[6]) / 2; result[1] = (P[2] - P[8]) / 2; result[2] = (P[4] - P[1]) / 2;
  const double sin_angle_abs = sqrt(dot3(result, result));
  const double kSqrt1_2 = 0.70710678118654752440;
  if (cos_angle > kSqrt1_2) {
    if (sin_angle_abs > 0) { const double f = asin(sin_angle_abs) / sin_angle_abs; for (int i = 0; i < 3; i++) result[i] *= f; }
  } else if (cos_angle > -kSqrt1_2) {
    const double angle = acos(cos_angle), f = angle / sin_angle_abs; for (int i = 0; i < 3; i++) result[i] *= f;
  } else {
    const double angle = 3.14159265358979323846 - asin(sin_angle_abs);
    const double d0 = P[0] - cos_angle, d1 = P[5] - cos_angle, d2 = P[10] - cos_angle;
    double r2[3];
    if (d0 * d0 > d1 * d1 && d0 * d0 > d2 * d2) { r2[0] = d0; r2[1] = (P[4] + P[1]) / 2; r2[2] = (P[2] + P[8]) / 2; }
    else if (d1 * d1 > d2 * d2) { r2[0] = (P[4] + P[1]) / 2; r2[1] = d1; r2[2] = (P[9] + P[6]) / 2; }
    else { r2[0] = (P[2] + P[8]) / 2; r2[1] = (P[9] + P[6]) / 2; r2[2] = d2; }
    if (dot3(r2, result) < 0) for (int i = 0; i < 3; i++) r2[i] *= -1;
    const double n = sqrt(dot3(r2, r2)); for (int i = 0; i < 3; i++) r2[i] /= n;
    for (int i = 0; i < 3; i++) result[i] = angle * r2[i];
  }
}
// mySE3::ln (jni/RT.h:354-378)
__device__ inline void se3_ln(const double* P, double* out6) {
  double rot[3]; so3_ln(P, rot);
  const double theta = sqrt(dot3(rot, rot));
  double shtot = 0.5;
  if (theta > 0.00001) shtot = sin(theta / 2) / theta;
  const double hw[3] = {rot[0] * -0.5, rot[1] * -0.5, rot[2] * -0.5};
  double Hm[9]; so3_exp(hw, Hm);
  const double t[3] = {P[3], P[7], P[11]};
  double rt[3];
  for (int i = 0; i < 3; i++) { double s = Hm[3 * i] * t[0]; s += Hm[3 * i + 1] * t[1]; s += Hm[3 * i + 2] * t[2]; rt[i] = s; }
  if (theta > 0.001) { const double f = (dot3(t, rot)) * (1 - 2 * shtot) / (dot3(rot, rot)); for (int i = 0; i < 3; i++) rt[i] -= rot[i] * f; }
  else { const double f = (dot3(t, rot)) / 24; for (int i = 0; i < 3; i++) rt[i] -= rot[i] * f; }
  for (int i = 0; i < 3; i++) rt[i] /= (2 * shtot);
  for (int i = 0; i < 3; i++) { out6[i] = rt[i]; out6[3 + i] = rot[i]; }
}

// glibc rand() (TYPE_3 additive feedback) — the generator behind std::random_shuffle in the reference
// (jni/Tracker.cc:396-397,525).  State = the 31-word ring + two indices, one per stream.
__device__ __forceinline__ int glibc_rand_next(int* ring, int& f, int& b) {
  const uint32_t v = (uint32_t)ring[f] + (uint32_t)ring[b];
  ring[f] = (int)v;
  if (++f >= 31) f = 0;
  if (++b >= 31) b = 0;
  return (int)(v >> 1);
}
// n consecutive draws of the same generator into out[0..n) (shared memory), with the 31-word ring held in registers:
// the ring is read rotated so that the rear index is 0, one draw is r[(j+3)%31] += r[j] (front = rear + 3, glibc
// random_r.c), and after 31 draws the rear index is back at 0.  Updates ring / f / b exactly as n calls of glibc_rand_next.
__device__ inline void glibc_rand_fill(int* ring, int& f, int& b, int* out, int n) {
  uint32_t r[31];
  const int b0 = b;
#pragma unroll
  for (int j = 0; j < 31; j++) { int q = b0 + j; if (q >= 31) q -= 31; r[j] = (uint32_t)ring[q]; }
  int k = 0;
  for (; n - k >= 31; k += 31) {
#pragma unroll
    for (int j = 0; j < 31; j++) { r[(j + 3) % 31] += r[j]; out[k + j] = (int)(r[(j + 3) % 31] >> 1); }
  }
  const int rem = n - k;
#pragma unroll
  for (int j = 0; j < 31; j++) if (j < rem) { r[(j + 3) % 31] += r[j]; out[k + j] = (int)(r[(j + 3) % 31] >> 1); }
#pragma unroll
  for (int j = 0; j < 31; j++) { int q = b0 + j; if (q >= 31) q -= 31; ring[q] = (int)r[j]; }
  b = (b0 + n) % 31; f = (b + 3) % 31;
}

// CPU ORACLE — a plain C++ restatement of the reference's per-frame tracking front-end
// (ahcorde/visualSLAM_Android, jni/).  TEST INFRASTRUCTURE ONLY: nothing in the product
// (visualslam_android_b200/) links, loads or calls this file; only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs do, and there only as the checker.
//
// Pinning: the reference ships no tests or golden vectors.  This restatement is
// pinned against the reference's OWN sources compiled here (oracle/_ref, see build_ref.sh) in
// tests/test_oracle_vs_ref.py, and against the fixtures that build generated
// (tests/golden/*.npz, script tests/golden/make_golden.py) wherever oracle/_ref is absent.
//
// Every function cites the reference file:line it follows.  Floating-point expressions keep the
// reference's evaluation order (left to right as written; no FMA: build with -ffp-contract=off).
// No Eigen, no OpenCV: POD arrays only.  Poses are row-major 3x4 [R|t] (camera-from-world).
#include "shim/cv_resize_linear_u8.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <utility>
#include <vector>

namespace {

const int LEVELS = 4;  // jni/KeyFrame.h:31

// ------------------------------------------------------------------------------------------------
// glibc rand() restated (TYPE_3 additive feedback, r[i] = r[i-3] + r[i-31]; seed 1 when never seeded).
// The reference calls std::random_shuffle, which draws from the process-global rand()
// (jni/Tracker.cc:396-397,525; never seeded).  A per-tracker copy of the generator lets several
// trackers live in one process without perturbing each other.  Checked against libc rand() in tests.
struct GlibcRand {
  int32_t r[34];
  int32_t ring[31];
  int f, b;  // front / back indices into ring (glibc: fptr = &state[3], rptr = &state[0])
  void seed(unsigned s) {
    if (s == 0) s = 1;
    int32_t st[31];
    st[0] = (int32_t)s;
    for (int i = 1; i < 31; i++) {
      long hi = st[i - 1] / 127773, lo = st[i - 1] % 127773;
      long w = 16807 * lo - 2836 * hi;
      if (w < 0) w += 2147483647;
      st[i] = (int32_t)w;
    }
    memcpy(ring, st, sizeof(st));
    f = 3; b = 0;
    for (int i = 0; i < 310; i++) next();
  }
  int next() {
    uint32_t v = (uint32_t)ring[f] + (uint32_t)ring[b];
    ring[f] = (int32_t)v;
    int res = (int)(v >> 1);
    if (++f >= 31) f = 0;
    if (++b >= 31) b = 0;
    return res;
  }
};

// ------------------------------------------------------------------------------------------------
// Camera scalars (host side of jni/ATANCamera.cc:37-129): the 13 doubles of synth.Camera.scalars().
struct Cam {
  double fx, fy, cx, cy, W, Winv, twoTan, oneOver2Tan, distEnabled, largestRadius, maxR, width, height;
  // cache of the last Project() (jni/ATANCamera.h:98-105) — GetProjectionDerivs reads it
  double lastCamX, lastCamY, lastR, lastFactor;
  bool invalid;
};

Cam cam_from13(const double* s);
// jni/ATANCamera.h:136-142
double rtrans_factor(const Cam& c, double r) {
  if (r < 0.001 || c.W == 0.0) return 1.0;
  return (c.Winv * atan(r * c.twoTan) / r);
}
// jni/ATANCamera.h:145-150
double invrtrans(const Cam& c, double r) {
  if (c.W == 0.0) return r;
  return (tan(r * c.W) * c.oneOver2Tan);
}
// jni/ATANCamera.cc:133-145
void cam_project(Cam& c, double x, double y, double* im) {
  c.lastCamX = x; c.lastCamY = y;
  c.lastR = sqrt(x * x + y * y);
  c.invalid = (c.lastR > c.maxR);
  c.lastFactor = rtrans_factor(c, c.lastR);
  const double dx = x * c.lastFactor, dy = y * c.lastFactor;
  im[0] = c.cx + c.fx * dx;
  im[1] = c.cy + c.fy * dy;
}
// jni/ATANCamera.cc:149-164
void cam_unproject(Cam& c, const double* im, double* out) {
  const double dx = (im[0] - c.cx) * (1.0 / c.fx), dy = (im[1] - c.cy) * (1.0 / c.fy);
  const double distR = sqrt(dx * dx + dy * dy);
  c.lastR = invrtrans(c, distR);
  double factor = (distR > 0.01) ? c.lastR / distR : 1.0;
  c.lastFactor = 1.0 / factor;
  c.lastCamX = dx * factor; c.lastCamY = dy * factor;
  out[0] = c.lastCamX; out[1] = c.lastCamY;
}
// jni/ATANCamera.cc:198-231 ; d = row-major 2x2
void cam_derivs(const Cam& c, double* d) {
  double fracBydx, fracBydy;
  const double k = c.twoTan, x = c.lastCamX, y = c.lastCamY;
  const double r = c.lastR * c.distEnabled;
  if (r < 0.01) {
    fracBydx = 0.0; fracBydy = 0.0;
  } else {
    fracBydx = c.Winv * (k * x) / (r * r * (1 + k * k * r * r)) - x * c.lastFactor / (r * r);
    fracBydy = c.Winv * (k * y) / (r * r * (1 + k * k * r * r)) - y * c.lastFactor / (r * r);
  }
  d[0] = c.fx * (fracBydx * x + c.lastFactor);
  d[2] = c.fy * (fracBydx * y);
  d[1] = c.fx * (fracBydy * x);
  d[3] = c.fy * (fracBydy * y + c.lastFactor);
}

// ------------------------------------------------------------------------------------------------
// SE3 / SO3 (jni/RT.h)
struct SE3 { double R[9]; double t[3]; };
SE3 se3_identity() { SE3 s; memset(&s, 0, sizeof(s)); s.R[0] = s.R[4] = s.R[8] = 1.0; return s; }
SE3 se3_from12(const double* p) { SE3 s; for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) s.R[3 * i + j] = p[4 * i + j]; s.t[i] = p[4 * i + 3]; } return s; }
void se3_to12(const SE3& s, double* p) { for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) p[4 * i + j] = s.R[3 * i + j]; p[4 * i + 3] = s.t[i]; } }
void mat3_mul_vec(const double* R, const double* v, double* o) {
  for (int i = 0; i < 3; i++) { double s = R[3 * i] * v[0]; s += R[3 * i + 1] * v[1]; s += R[3 * i + 2] * v[2]; o[i] = s; }
}
void mat3_mul(const double* A, const double* B, double* o) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = A[3 * i] * B[j]; s += A[3 * i + 1] * B[3 + j]; s += A[3 * i + 2] * B[6 + j]; o[3 * i + j] = s; }
}
// jni/RT.h:380-388  (lhs.t + lhs.R * rhs)
void se3_apply(const SE3& s, const double* v, double* o) { double rv[3]; mat3_mul_vec(s.R, v, rv); for (int i = 0; i < 3; i++) o[i] = s.t[i] + rv[i]; }
// jni/RT.h:275-282
SE3 se3_mul(const SE3& a, const SE3& b) {
  SE3 r; mat3_mul(a.R, b.R, r.R); double rv[3]; mat3_mul_vec(a.R, b.t, rv); for (int i = 0; i < 3; i++) r.t[i] = a.t[i] + rv[i]; return r;
}
// jni/RT.h:262-270
SE3 se3_inverse(const SE3& a) {
  SE3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.R[3 * i + j] = a.R[3 * j + i];
  double rv[3]; mat3_mul_vec(r.R, a.t, rv); for (int i = 0; i < 3; i++) r.t[i] = -rv[i]; return r;
}
void cross3(const double* a, const double* b, double* o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }
double dot3(const double* a, const double* b) { double s = 0; s += a[0] * b[0]; s += a[1] * b[1]; s += a[2] * b[2]; return s; }
// jni/RT.h:98-127
void rodrigues(const double* w, double A, double B, double* R) {
  const double wx2 = w[0] * w[0], wy2 = w[1] * w[1], wz2 = w[2] * w[2];
  R[0] = 1.0 - B * (wy2 + wz2); R[4] = 1.0 - B * (wx2 + wz2); R[8] = 1.0 - B * (wx2 + wy2);
  { const double a = A * w[2], b = B * (w[0] * w[1]); R[1] = b - a; R[3] = b + a; }
  { const double a = A * w[1], b = B * (w[0] * w[2]); R[2] = b + a; R[6] = b - a; }
  { const double a = A * w[0], b = B * (w[1] * w[2]); R[5] = b - a; R[7] = b + a; }
}
// jni/RT.h:318-352
SE3 se3_exp(const double* mu) {
  static const double one_6th = 1.0 / 6.0, one_20th = 1.0 / 20.0;
  SE3 res;
  const double* w = mu + 3;
  const double theta_sq = dot3(w, w), theta = sqrt(theta_sq);
  double A, B, cr[3];
  cross3(w, mu, cr);
  if (theta_sq < 1e-8) {
    A = 1.0 - one_6th * theta_sq; B = 0.5;
    for (int i = 0; i < 3; i++) res.t[i] = mu[i] + 0.5 * cr[i];
  } else {
    double C;
    if (theta_sq < 1e-6) {
      C = one_6th * (1.0 - one_20th * theta_sq); A = 1.0 - theta_sq * C; B = 0.5 - 0.25 * one_6th * theta_sq;
    } else {
      const double inv_theta = 1.0 / theta;
      A = sin(theta) * inv_theta; B = (1 - cos(theta)) * (inv_theta * inv_theta); C = (1 - A) * (inv_theta * inv_theta);
    }
    double wc[3]; cross3(w, cr, wc);
    for (int i = 0; i < 3; i++) res.t[i] = (mu[i] + B * cr[i]) + C * wc[i];
  }
  rodrigues(w, A, B, res.R);
  return res;
}
// jni/RT.h:132-164
void so3_exp(const double* w, double* R) {
  static const double one_6th = 1.0 / 6.0, one_20th = 1.0 / 20.0;
  const double theta_sq = dot3(w, w), theta = sqrt(theta_sq);
  double A, B;
  if (theta_sq < 1e-8) { A = 1.0 - one_6th * theta_sq; B = 0.5; }
  else if (theta_sq < 1e-6) { B = 0.5 - 0.25 * one_6th * theta_sq; A = 1.0 - theta_sq * one_6th * (1.0 - one_20th * theta_sq); }
  else { const double inv_theta = 1.0 / theta; A = sin(theta) * inv_theta; B = (1 - cos(theta)) * (inv_theta * inv_theta); }
  rodrigues(w, A, B, R);
}
// jni/RT.h:166-215
void so3_ln(const double* M, double* result) {
  const double cos_angle = (M[0] + M[4] + M[8] - 1.0) * 0.5;
  result[0] = (M[7] - M[5]) / 2; result[1] = (M[2] - M[6]) / 2; result[2] = (M[3] - M[1]) / 2;
  double sin_angle_abs = sqrt(dot3(result, result));
  if (cos_angle > M_SQRT1_2) {
    if (sin_angle_abs > 0) { const double f = asin(sin_angle_abs) / sin_angle_abs; for (int i = 0; i < 3; i++) result[i] *= f; }
  } else if (cos_angle > -M_SQRT1_2) {
    const double angle = acos(cos_angle); const double f = angle / sin_angle_abs; for (int i = 0; i < 3; i++) result[i] *= f;
  } else {
    const double angle = M_PI - asin(sin_angle_abs);
    const double d0 = M[0] - cos_angle, d1 = M[4] - cos_angle, d2 = M[8] - cos_angle;
    double r2[3];
    if (d0 * d0 > d1 * d1 && d0 * d0 > d2 * d2) { r2[0] = d0; r2[1] = (M[3] + M[1]) / 2; r2[2] = (M[2] + M[6]) / 2; }
    else if (d1 * d1 > d2 * d2) { r2[0] = (M[3] + M[1]) / 2; r2[1] = d1; r2[2] = (M[7] + M[5]) / 2; }
    else { r2[0] = (M[2] + M[6]) / 2; r2[1] = (M[7] + M[5]) / 2; r2[2] = d2; }
    if (dot3(r2, result) < 0) for (int i = 0; i < 3; i++) r2[i] *= -1;
    const double n = sqrt(dot3(r2, r2)); for (int i = 0; i < 3; i++) r2[i] /= n;
    for (int i = 0; i < 3; i++) result[i] = angle * r2[i];
  }
}
// jni/RT.h:354-378
void se3_ln(const SE3& s, double* out6) {
  double rot[3]; so3_ln(s.R, rot);
  const double theta = sqrt(dot3(rot, rot));
  double shtot = 0.5;
  if (theta > 0.00001) shtot = sin(theta / 2) / theta;
  double hw[3] = {rot[0] * -0.5, rot[1] * -0.5, rot[2] * -0.5}, H[9];
  so3_exp(hw, H);
  double rottrans[3]; mat3_mul_vec(H, s.t, rottrans);
  if (theta > 0.001) { const double f = (dot3(s.t, rot)) * (1 - 2 * shtot) / (dot3(rot, rot)); for (int i = 0; i < 3; i++) rottrans[i] -= rot[i] * f; }
  else { const double f = (dot3(s.t, rot)) / 24; for (int i = 0; i < 3; i++) rottrans[i] -= rot[i] * f; }
  for (int i = 0; i < 3; i++) rottrans[i] /= (2 * shtot);
  for (int i = 0; i < 3; i++) { out6[i] = rottrans[i]; out6[3 + i] = rot[i]; }
}

// ------------------------------------------------------------------------------------------------
// Images, pyramid, FAST
struct Image { int w, h; std::vector<uint8_t> px; const uint8_t* row(int y) const { return &px[(size_t)y * w]; } };
struct Corner { int x, y; };
struct OLevel {
  Image im;
  std::vector<Corner> corners;      // Level::vCorners
  std::vector<int> lut;             // Level::vCornerRowLUT
  std::vector<Corner> maxCorners;   // Level::vMaxCorners
  std::vector<Corner> candidates;   // Level::vCandidates (positions)
  std::vector<double> candScores;   //   .. dSTScore
};
struct OKeyFrame { OLevel lev[LEVELS]; };

// cv::resize(prev, lev, size/2) for an exact 2:1 u8 image == (a+b+c+d+2)>>2 (jni/KeyFrame.cc:20-23)
void half_sample(const Image& s, Image& d) {
  d.w = s.w / 2; d.h = s.h / 2; d.px.resize((size_t)d.w * d.h);
  for (int y = 0; y < d.h; y++) {
    const uint8_t* a = s.row(2 * y); const uint8_t* b = s.row(2 * y + 1); uint8_t* o = &d.px[(size_t)y * d.w];
    for (int x = 0; x < d.w; x++) o[x] = (uint8_t)((a[2 * x] + a[2 * x + 1] + b[2 * x] + b[2 * x + 1] + 2) >> 2);
  }
}
// Ring of jni/vision/cvfast.cpp:6094-6111
const int RING_DX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int RING_DY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
bool has_run10(unsigned m) { m |= m << 16; unsigned r = m; for (int s = 1; s < 10; s++) r &= (m >> s); return (r & 0xffffu) != 0; }
// cvCornerFast_10 (jni/vision/cvfast.cpp:6088-9241): the generated decision tree is exactly the segment test
// ">= 10 contiguous ring pixels all > p+t or all < p-t"; raster order; y in [3,rows-4], x in [3,cols-4].
void fast10(const Image& im, int thr, std::vector<Corner>& out) {
  out.clear();
  for (int y = 3; y < im.h - 3; y++)
    for (int x = 3; x < im.w - 3; x++) {
      const int p = im.row(y)[x], cb = p + thr, c_b = p - thr;
      unsigned br = 0, dk = 0;
      for (int k = 0; k < 16; k++) {
        const int v = im.row(y + RING_DY[k])[x + RING_DX[k]];
        if (v > cb) br |= 1u << k;
        if (v < c_b) dk |= 1u << k;
      }
      if (has_run10(br) || has_run10(dk)) { Corner c = {x, y}; out.push_back(c); }
    }
}
// jni/KeyFrame.cc:41-49
void row_lut(const std::vector<Corner>& c, int rows, std::vector<int>& lut) {
  lut.clear(); unsigned v = 0;
  for (int y = 0; y < rows; y++) { while (v < c.size() && y > c[v].y) v++; lut.push_back((int)v); }
}
// KeyFrame::MakeKeyFrame_Lite (jni/KeyFrame.cc:5-51); thresholds 10/15/15/10 (:32-39)
void make_keyframe_lite(OKeyFrame& kf, const uint8_t* gray, int w, int h, int stride) {
  static const int thr[LEVELS] = {10, 15, 15, 10};
  Image& l0 = kf.lev[0].im; l0.w = w; l0.h = h; l0.px.resize((size_t)w * h);
  for (int y = 0; y < h; y++) memcpy(&l0.px[(size_t)y * w], gray + (size_t)y * stride, w);
  for (int i = 0; i < LEVELS; i++) {
    if (i) half_sample(kf.lev[i - 1].im, kf.lev[i].im);
    kf.lev[i].maxCorners.clear(); kf.lev[i].candidates.clear(); kf.lev[i].candScores.clear();
    fast10(kf.lev[i].im, thr[i], kf.lev[i].corners);
    row_lut(kf.lev[i].corners, kf.lev[i].im.h, kf.lev[i].lut);
  }
}
// old_style_corner_score (jni/vision/cvfast.cpp:9337-9369)
int fast_score(const Image& im, int x, int y, int barrier) {
  const int p = im.row(y)[x], cb = p + barrier, c_b = p - barrier; int sp = 0, sn = 0;
  for (int k = 0; k < 16; k++) { const int v = im.row(y + RING_DY[k])[x + RING_DX[k]]; if (v > cb) sp += v - cb; else if (v < c_b) sn += c_b - v; }
  return sp > sn ? sp : sn;
}
// nonmax_suppression (jni/vision/cvfast.cpp:9243-9335).  Quirk kept: the "check right" test also requires the
// PREVIOUS list entry to be on the same row (:9284 reads corners[i-1]); at i==0 that read is out of bounds in the
// reference — here it is treated as "not on the same row".
void nonmax(const std::vector<Corner>& c, const std::vector<int>& sc, std::vector<Corner>& out) {
  out.clear();
  const int sz = (int)c.size();
  if (sz < 1) return;
  const int last_row = c.back().y;
  std::vector<int> row_start(last_row + 1, -1);
  int prev_row = -1;
  for (int i = 0; i < sz; i++) if (c[i].y != prev_row) { row_start[c[i].y] = i; prev_row = c[i].y; }
  int point_above = 0, point_below = 0;
  for (int i = 0; i < sz; i++) {
    const int score = sc[i]; const int px = c[i].x, py = c[i].y;
    if (i > 0) if (c[i - 1].x == px - 1 && c[i - 1].y == py && sc[i - 1] > score) continue;
    if (i < sz - 1) if (c[i + 1].x == px + 1 && (i > 0 && c[i - 1].y == py) && sc[i + 1] > score) continue;
    bool suppressed = false;
    if (py != 0 && row_start[py - 1] != -1) {
      if (c[point_above].y < py - 1) point_above = row_start[py - 1];
      for (; c[point_above].y < py && c[point_above].x < px - 1; point_above++) {}
      for (int j = point_above; c[j].y < py && c[j].x <= px + 1; j++) {
        const int x = c[j].x;
        if ((x == px - 1 || x == px || x == px + 1) && sc[j] > score) { suppressed = true; break; }
      }
    }
    if (suppressed) continue;
    if (py != last_row && row_start[py + 1] != -1 && point_below < sz) {
      if (c[point_below].y < py + 1) point_below = row_start[py + 1];
      for (; point_below < sz && c[point_below].y == py + 1 && c[point_below].x < px - 1; point_below++) {}
      for (int j = point_below; j < sz && c[j].y == py + 1 && c[j].x <= px + 1; j++) {
        const int x = c[j].x;
        if ((x == px - 1 || x == px || x == px + 1) && sc[j] > score) { suppressed = true; break; }
      }
    }
    if (suppressed) continue;
    out.push_back(c[i]);
  }
}
// FindShiTomasiScoreAtPoint (jni/vision/ImageHandler.cpp:124-155)
double shi_tomasi(const Image& im, int nsize, int px, int py) {
  double dXX = 0, dYY = 0, dXY = 0;
  const int startx = px - nsize, starty = py - nsize, endx = px + nsize, endy = py + nsize;
  for (int cy = starty; cy <= endy; cy++)
    for (int cx = startx; cx <= endx; cx++) {
      const double dx = (double)(im.row(cy)[cx + 1] - im.row(cy)[cx - 1]);
      const double dy = (double)(im.row(cy + 1)[cx] - im.row(cy - 1)[cx]);
      dXX += dx * dx; dYY += dy * dy; dXY += dx * dy;
    }
  const int nPixels = (endx - startx + 1) * (endy - starty + 1);
  dXX = dXX / (2.0 * nPixels); dYY = dYY / (2.0 * nPixels); dXY = dXY / (2.0 * nPixels);
  return 0.5 * (dXX + dYY - sqrt((dXX + dYY) * (dXX + dYY) - 4 * (dXX * dYY - dXY * dXY)));
}
// KeyFrame::MakeKeyFrame_Rest without the SmallBlurryImage part (jni/KeyFrame.cc:53-95)
void make_keyframe_rest(OKeyFrame& kf) {
  const double minST = 70;
  for (int l = 0; l < LEVELS; l++) {
    OLevel& L = kf.lev[l];
    std::vector<int> scores(L.corners.size());
    for (size_t i = 0; i < L.corners.size(); i++) scores[i] = fast_score(L.im, L.corners[i].x, L.corners[i].y, 10);
    nonmax(L.corners, scores, L.maxCorners);
    const int border = 10;
    L.candidates.clear(); L.candScores.clear();
    for (size_t i = 0; i < L.maxCorners.size(); i++) {
      const Corner c = L.maxCorners[i];
      if (!(c.x >= border && c.y >= border && c.x < L.im.w - border && c.y < L.im.h - border)) continue;
      const double s = shi_tomasi(L.im, 3, c.x, c.y);
      if (s > minST) { L.candidates.push_back(c); L.candScores.push_back(s); }
    }
  }
}

// jni/vision/ImageHandler.cpp:120-122
bool in_image_with_border(const Image& im, int px, int py, int border) { return px >= border && py >= border && px < im.w - border && py < im.h - border; }
// jni/LevelHelpers.h:17-45
int LevelScale(int l) { return 1 << l; }
double LevelZeroPos(double p, int l) { return (p + 0.5) * LevelScale(l) - 0.5; }
double LevelNPos(double p, int l) { return (p + 0.5) / LevelScale(l) - 0.5; }

// ------------------------------------------------------------------------------------------------
// Template generation: transform_image + sample(u8) (jni/vision/ImageHandler.cpp:12-113)
// M = row-major 2x2, inOrig = irCenter, outOrig = (P/2, P/2).  Returns the number of samples outside.
int transform_image_u8(const Image& in, uint8_t* out, int P, const double* M, const double* inOrig, const double* outOrig) {
  const int w = P, h = P, iw = in.w, ih = in.h;
  const double across[2] = {M[0], M[2]}, down[2] = {M[1], M[3]};
  double p0[2];
  { double a = M[0] * outOrig[0]; a += M[1] * outOrig[1]; double b = M[2] * outOrig[0]; b += M[3] * outOrig[1]; p0[0] = inOrig[0] - a; p0[1] = inOrig[1] - b; }
  double min_x = p0[0], min_y = p0[1], max_x = min_x, max_y = min_y;
  if (across[0] < 0) min_x += w * across[0]; else max_x += w * across[0];
  if (down[0] < 0) min_x += h * down[0]; else max_x += h * down[0];
  if (across[1] < 0) min_y += w * across[1]; else max_y += w * across[1];
  if (down[1] < 0) min_y += h * down[1]; else max_y += h * down[1];
  const double cr[2] = {down[0] - w * across[0], down[1] - w * across[1]};
  const bool inside = (min_x >= 0 && min_y >= 0 && max_x < iw - 1 && max_y < ih - 1);
  const float x_bound = iw - 1, y_bound = ih - 1;
  int count = 0;
  double p[2] = {p0[0], p0[1]};
  for (int i = 0; i < h; ++i, p[0] += cr[0], p[1] += cr[1])
    for (int j = 0; j < w; ++j, p[0] += across[0], p[1] += across[1]) {
      if (inside || (0 <= p[0] && 0 <= p[1] && p[0] < x_bound && p[1] < y_bound)) {
        double x = p[0], y = p[1];
        const int lx = (int)x, ly = (int)y;
        x -= lx; y -= ly;
        const uint8_t* r0 = in.row(ly); const uint8_t* r1 = in.row(ly + 1);
        out[i * P + j] = (uint8_t)((1 - y) * ((1 - x) * r0[lx] + x * r0[lx + 1]) + y * ((1 - x) * r1[lx] + x * r1[lx + 1]));
      } else { out[i * P + j] = 0; ++count; }
    }
  return count;
}

// ------------------------------------------------------------------------------------------------
// PatchFinder state, one per (tracker, map point)  (jni/PatchFinder.h:97-127)
struct Finder {
  int P, maxSSD;
  std::vector<uint8_t> tmpl; int tsum, tsumsq;
  double warpInv[4];      // mm2WarpInverse, row-major
  int level;              // mnSearchLevel
  bool templateBad;       // mbTemplateBad (sticky, see MakeTemplateCoarseCont)
  bool haveLast;          // mpLastTemplateMapPoint == &p
  double lastWarp[4];     // mm2LastWarpMatrix (9999.9*I at construction, jni/PatchFinder.cc:23)
  double coarsePos[2], subPixPos[2], meanDiff;
  double hinv[9];         // mm3HInv
  std::vector<double> jx, jy;  // mimJacs[0/1], indexed (x-1)*(P-2)+(y-1)
  bool found;
  void init(int p) {
    P = p; maxSSD = P * P * 500; tmpl.assign(P * P, 0); tsum = tsumsq = 0; level = 0; templateBad = false; haveLast = false;
    lastWarp[0] = lastWarp[3] = 9999.9; lastWarp[1] = lastWarp[2] = 0; found = false; meanDiff = 0;
    memset(warpInv, 0, sizeof(warpInv)); memset(coarsePos, 0, sizeof(coarsePos)); memset(subPixPos, 0, sizeof(subPixPos)); memset(hinv, 0, sizeof(hinv));
  }
};

struct MapPointO {
  double world[3], right[3], down[3];  // v3WorldPos, v3PixelRight_W, v3PixelDown_W
  int irCenter[2], srcLevel;
  int outlierCount, inlierCount;
  const struct OKeyFrame* srcKF;   // MapPoint::pPatchSourceKF (jni/MapPoint.h:38)
};

// PatchFinder::CalcSearchLevelAndWarpMatrix (jni/PatchFinder.cc:31-68)
int calc_search_level_and_warp(Finder& f, const MapPointO& p, const SE3& pose, const double* D /*row-major 2x2*/) {
  double v3Cam[3]; se3_apply(pose, p.world, v3Cam);
  const double invz = 1.0 / v3Cam[2];
  double mr[3], md[3]; mat3_mul_vec(pose.R, p.right, mr); mat3_mul_vec(pose.R, p.down, md);
  double a[2], b[2];
  for (int k = 0; k < 2; k++) { a[k] = mr[k] - v3Cam[k] * mr[2] * invz; b[k] = md[k] - v3Cam[k] * md[2] * invz; }
  double aux1[2], aux2[2];
  for (int i = 0; i < 2; i++) {
    double s = D[2 * i] * a[0]; s += D[2 * i + 1] * a[1]; aux1[i] = s * invz;
    double t = D[2 * i] * b[0]; t += D[2 * i + 1] * b[1]; aux2[i] = t * invz;
  }
  f.warpInv[0] = aux1[0]; f.warpInv[1] = aux2[0]; f.warpInv[2] = aux1[1]; f.warpInv[3] = aux2[1];
  double dDet = f.warpInv[0] * f.warpInv[3] - f.warpInv[1] * f.warpInv[2];
  f.level = 0;
  while (dDet > 3 && f.level < LEVELS - 1) { f.level++; dDet *= 0.25; }
  if (dDet > 3 || dDet < 0.25) { f.templateBad = true; return -1; }
  return f.level;
}
// PatchFinder::MakeTemplateSums (jni/PatchFinder.cc:152-164)
void make_template_sums(Finder& f) { int s = 0, q = 0; for (int i = 0; i < f.P * f.P; i++) { int b = f.tmpl[i]; s += b; q += b * b; } f.tsum = s; f.tsumsq = q; }
// PatchFinder::MakeTemplateCoarseCont (jni/PatchFinder.cc:79-125).  Returns 1 if the template was regenerated.
int make_template_coarse_cont(Finder& f, const MapPointO& p, const OKeyFrame& srcKF) {
  // m2 = inverse(mm2WarpInverse) * LevelScale : 2x2 inverse = adjugate * (1/det) (oracle/shim/Eigen/Dense)
  const double* w = f.warpInv;
  const double invdet = 1.0 / (w[0] * w[3] - w[1] * w[2]);
  const int s = LevelScale(f.level);
  double m2[4] = {(w[3] * invdet) * s, (-w[1] * invdet) * s, (-w[2] * invdet) * s, (w[0] * invdet) * s};
  bool refresh = !f.haveLast;
  for (int i = 0; !refresh && i < 2; i++) {
    const double d0 = m2[i] - f.lastWarp[i], d1 = m2[2 + i] - f.lastWarp[2 + i];
    double dd = 0; dd += d0 * d0; dd += d1 * d1;
    const double lim = 0.07;
    if (dd > lim * lim) refresh = true;
  }
  if (!refresh) return 0;
  const double inOrig[2] = {(double)p.irCenter[0], (double)p.irCenter[1]};
  const double outOrig[2] = {(double)(f.P / 2), (double)(f.P / 2)};
  const int nOutside = transform_image_u8(srcKF.lev[p.srcLevel].im, &f.tmpl[0], f.P, m2, inOrig, outOrig);
  f.templateBad = nOutside != 0;
  make_template_sums(f);
  f.haveLast = true; memcpy(f.lastWarp, m2, sizeof(m2));
  return 1;
}
// PatchFinder::ZMSSDAtPoint (jni/PatchFinder.cc:352-380)
int zmssd_at_point(const Finder& f, const Image& img, int icol, int irow) {
  const int b = f.P / 2;
  if (!in_image_with_border(img, icol, irow, b)) return f.maxSSD + 1;
  const int bx = icol - b, by = irow - b;
  int sumsq = 0, sum = 0, cross = 0;
  for (int r = 0; r < f.P; r++) {
    const uint8_t* ip = img.row(by + r); const uint8_t* tp = &f.tmpl[r * f.P];
    for (int c = 0; c < f.P; c++) { const int n = ip[bx + c]; sum += n; sumsq += n * n; cross += n * tp[c]; }
  }
  const int SA = f.tsum, SB = sum, N = f.P * f.P;
  return ((2 * SA * SB - SA * SA - SB * SB) / N + sumsq + f.tsumsq - 2 * cross);
}
// PatchFinder::FindPatchCoarse (jni/PatchFinder.cc:170-235).  stats (optional): [0] += ZMSSD evaluations
bool find_patch_coarse(Finder& f, double px, double py, const OKeyFrame& kf, unsigned nRange, long* stats, int* bestSSDOut) {
  f.found = false;
  const int nLevelScale = LevelScale(f.level);
  const double ix = px / nLevelScale, iy = py / nLevelScale;
  nRange = (nRange + nLevelScale - 1) / nLevelScale;
  int nTop = iy - nRange;
  int nBottomPlusOne = iy + nRange + 1;
  int nLeft = ix - nRange;
  int nRight = ix + nRange;
  const OLevel& L = kf.lev[f.level];
  if (bestSSDOut) *bestSSDOut = f.maxSSD + 1;
  if (nTop < 0) nTop = 0;
  if (nTop >= L.im.h) return false;
  if (nBottomPlusOne <= 0) return false;
  int i = L.lut[nTop];
  const int i_end = (nBottomPlusOne >= L.im.h) ? (int)L.corners.size() : L.lut[nBottomPlusOne];
  int bestX = -1, bestY = -1, nBestSSD = f.maxSSD + 1;
  for (; i < i_end; i++) {
    const double cx = L.corners[i].x, cy = L.corners[i].y;
    if (cx < nLeft || cx > nRight) continue;
    const double dx = ix - cx, dy = iy - cy;
    double d2 = 0; d2 += dx * dx; d2 += dy * dy;
    if (d2 > nRange * nRange) continue;
    const int nSSD = zmssd_at_point(f, L.im, L.corners[i].x, L.corners[i].y);
    if (stats) stats[0]++;
    if (nSSD < nBestSSD) { bestX = L.corners[i].x; bestY = L.corners[i].y; nBestSSD = nSSD; }
  }
  if (bestSSDOut) *bestSSDOut = nBestSSD;
  if (nBestSSD < f.maxSSD) {
    f.coarsePos[0] = LevelZeroPos((double)bestX, f.level); f.coarsePos[1] = LevelZeroPos((double)bestY, f.level);
    f.found = true;
  }
  return f.found;
}
// 3x3 inverse as evaluated by the Eigen stand-in (adjugate * 1/det; oracle/shim/Eigen/Dense)
void inverse3(const double* m, double* r) {
  const double c00 = m[4] * m[8] - m[5] * m[7], c10 = m[5] * m[6] - m[3] * m[8], c20 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c00 + m[1] * c10 + m[2] * c20, invdet = 1.0 / det;
  r[0] = c00 * invdet; r[3] = c10 * invdet; r[6] = c20 * invdet;
  r[1] = (m[2] * m[7] - m[1] * m[8]) * invdet; r[4] = (m[0] * m[8] - m[2] * m[6]) * invdet; r[7] = (m[1] * m[6] - m[0] * m[7]) * invdet;
  r[2] = (m[1] * m[5] - m[2] * m[4]) * invdet; r[5] = (m[2] * m[3] - m[0] * m[5]) * invdet; r[8] = (m[0] * m[4] - m[1] * m[3]) * invdet;
}
// PatchFinder::MakeSubPixTemplate (jni/PatchFinder.cc:242-267)
void make_subpix_template(Finder& f) {
  const int P = f.P, Q = P - 2;
  f.jx.assign(Q * Q, 0); f.jy.assign(Q * Q, 0);
  double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int x = 1; x < P - 1; x++)
    for (int y = 1; y < P - 1; y++) {
      const double g[3] = {0.5 * (f.tmpl[y * P + x + 1] - f.tmpl[y * P + x - 1]), 0.5 * (f.tmpl[(y + 1) * P + x] - f.tmpl[(y - 1) * P + x]), 1.0};
      f.jx[(x - 1) * Q + (y - 1)] = g[0]; f.jy[(x - 1) * Q + (y - 1)] = g[1];
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) H[3 * i + j] += g[i] * g[j];
    }
  inverse3(H, f.hinv);
  f.subPixPos[0] = f.coarsePos[0]; f.subPixPos[1] = f.coarsePos[1];
  f.meanDiff = 0.0;
}
static int g_dbg_point = -1; static bool g_dbg = false;
// PatchFinder::IterateSubPix (jni/PatchFinder.cc:291-350)
double iterate_subpix(Finder& f, const OKeyFrame& kf) {
  const int P = f.P, Q = P - 2;
  const Image& im = kf.lev[f.level].im;
  const double c0 = LevelNPos(f.subPixPos[0], f.level), c1 = LevelNPos(f.subPixPos[1], f.level);
  const int x_border = (c0 > 0.0 ? c0 + 0.5 : c0 - 0.5), y_border = (c1 > 0.0 ? c1 + 0.5 : c1 - 0.5);
  if (!in_image_with_border(im, x_border, y_border, P / 2 + 1)) return -1.0;
  const double b0 = c0 - (double)(P / 2), b1 = c1 - (double)(P / 2);
  double acc[3] = {0, 0, 0};
  const double dX = b0 - floor(b0), dY = b1 - floor(b1);
  const float fMixTL = (1.0 - dX) * (1.0 - dY), fMixTR = (dX) * (1.0 - dY), fMixBL = (1.0 - dX) * (dY), fMixBR = (dX) * (dY);
  for (int y = 1; y < P - 1; y++) {
    const uint8_t* tl = im.row((int)b1 + y) + ((int)b0 + 1);
    const uint8_t* bl = tl + im.w;
    for (int x = 1; x < P - 1; x++) {
      float fPixel = fMixTL * tl[0] + fMixTR * tl[1] + fMixBL * bl[0] + fMixBR * bl[1];
      tl++; bl++;
      const double dDiff = fPixel - f.tmpl[y * P + x] + f.meanDiff;
      acc[0] += dDiff * f.jx[(x - 1) * Q + (y - 1)];
      acc[1] += dDiff * f.jy[(x - 1) * Q + (y - 1)];
      acc[2] += dDiff;
    }
  }
  double upd[3];
  for (int i = 0; i < 3; i++) { double s = f.hinv[3 * i] * acc[0]; s += f.hinv[3 * i + 1] * acc[1]; s += f.hinv[3 * i + 2] * acc[2]; upd[i] = s; }
  const int sc = LevelScale(f.level);
  f.subPixPos[0] -= upd[0] * sc; f.subPixPos[1] -= upd[1] * sc;
  f.meanDiff -= upd[2];
  double d = 0; d += upd[0] * upd[0]; d += upd[1] * upd[1];
  if (g_dbg) fprintf(stderr, "orc it: acc %.17g %.17g %.17g upd %.17g %.17g %.17g pos %.17g %.17g d %.17g mix %.9g %.9g %.9g %.9g\n", acc[0], acc[1], acc[2], upd[0], upd[1], upd[2], f.subPixPos[0], f.subPixPos[1], d, fMixTL, fMixTR, fMixBL, fMixBR);
  return d;
}
// PatchFinder::IterateSubPixToConvergence (jni/PatchFinder.cc:272-285)
bool iterate_subpix_to_convergence(Finder& f, const OKeyFrame& kf, int nMaxIts) {
  const double dConvLimit = 0.03;
  for (int it = 0; it < nMaxIts; it++) {
    const double d = iterate_subpix(f, kf);
    if (d < 0) return false;
    if (d < dConvLimit * dConvLimit) return true;
  }
  return false;
}

// ------------------------------------------------------------------------------------------------
// MiniPatch (jni/MiniPatch.cc)
int minipatch_ssd(const uint8_t* patch, int half, int maxSSD, const Image& im, int icol, int irow) {  // :6-27
  if (!in_image_with_border(im, icol, irow, half)) return maxSSD + 1;
  const int n = 2 * half + 1; int ssd = 0;
  for (int r = 0; r < n; r++) { const uint8_t* ip = im.row(irow - half + r) + (icol - half); for (int c = 0; c < n; c++) { const int d = ip[c] - patch[r * n + c]; ssd += d * d; } }
  return ssd;
}
bool minipatch_find(const uint8_t* patch, int half, int maxSSD, double* pos, const OLevel& L, int nRange, bool useLUT, int* bestOut) {  // :32-70
  double bestX = 0, bestY = 0; int nBestSSD = maxSSD + 1;
  const double tlx = pos[0] - nRange, tly = pos[1] - nRange, brx = pos[0] + nRange, bry = pos[1] + nRange;
  size_t i = 0;
  if (!useLUT) { for (i = 0; i < L.corners.size(); i++) if (L.corners[i].y >= tly) break; }
  else { int top = tly; if (top < 0) top = 0; if (top >= (int)L.lut.size()) top = (int)L.lut.size() - 1; i = L.lut[top]; }
  for (; i < L.corners.size(); i++) {
    if (L.corners[i].x < tlx || L.corners[i].x > brx) continue;
    if (L.corners[i].y > bry) break;
    const int s = minipatch_ssd(patch, half, maxSSD, L.im, L.corners[i].x, L.corners[i].y);
    if (s < nBestSSD) { bestX = L.corners[i].x; bestY = L.corners[i].y; nBestSSD = s; }
  }
  if (bestOut) *bestOut = nBestSSD;
  if (nBestSSD < maxSSD) { pos[0] = bestX; pos[1] = bestY; return true; }
  return false;
}

// ------------------------------------------------------------------------------------------------
// TrackerData (jni/TrackerData.h:35-136)
struct TData {
  Finder finder;
  double v3Cam[3], v2ImPlane[2], v2Image[2], derivs[4];
  bool inImage, potentiallyVisible;
  int searchLevel; bool searched, found, didSubPix;
  double v2Found[2], sqrtInvNoise, err[2], jac[12];
};
// TrackerData::Project (jni/TrackerData.h:69-86)
void td_project(TData& td, const MapPointO& p, const SE3& pose, Cam& cam) {
  td.inImage = td.potentiallyVisible = false;
  se3_apply(pose, p.world, td.v3Cam);
  if (td.v3Cam[2] < 0.001) return;
  td.v2ImPlane[0] = td.v3Cam[0] / td.v3Cam[2]; td.v2ImPlane[1] = td.v3Cam[1] / td.v3Cam[2];
  double d = 0; d += td.v2ImPlane[0] * td.v2ImPlane[0]; d += td.v2ImPlane[1] * td.v2ImPlane[1];
  if (d > cam.largestRadius * cam.largestRadius) return;
  cam_project(cam, td.v2ImPlane[0], td.v2ImPlane[1], td.v2Image);
  if (cam.invalid) return;
  if (td.v2Image[0] < 0 || td.v2Image[1] < 0 || td.v2Image[0] > cam.width || td.v2Image[1] > cam.height) return;
  td.inImage = true;
}
// TrackerData::ProjectAndDerivs (jni/TrackerData.h:98-102): derivs refreshed `if(bFound)`, from the camera's cache
void td_project_and_derivs(TData& td, const MapPointO& p, const SE3& pose, Cam& cam) { td_project(td, p, pose, cam); if (td.found) cam_derivs(cam, td.derivs); }
// TrackerData::CalcJacobian (jni/TrackerData.h:107-123) with mySE3::generator_field (jni/RT.h:285-295)
void td_calc_jacobian(TData& td) {
  const double invz = 1.0 / td.v3Cam[2];
  const double pos[4] = {td.v3Cam[0], td.v3Cam[1], td.v3Cam[2], 1.0};
  for (int m = 0; m < 6; m++) {
    double v4[4] = {0, 0, 0, 0};
    if (m < 3) v4[m] = pos[3];
    else { v4[(m + 1) % 3] = -pos[(m + 2) % 3]; v4[(m + 2) % 3] = pos[(m + 1) % 3]; }
    const double c0 = (v4[0] - td.v3Cam[0] * v4[2] * invz) * invz, c1 = (v4[1] - td.v3Cam[1] * v4[2] * invz) * invz;
    double a0 = td.derivs[0] * c0; a0 += td.derivs[1] * c1;
    double a1 = td.derivs[2] * c0; a1 += td.derivs[3] * c1;
    td.jac[m] = a0; td.jac[6 + m] = a1;
  }
}
// TrackerData::LinearUpdate (jni/TrackerData.h:126-132)
void td_linear_update(TData& td, const double* v6) {
  for (int r = 0; r < 2; r++) { double s = td.jac[6 * r] * v6[0]; for (int k = 1; k < 6; k++) s += td.jac[6 * r + k] * v6[k]; td.v2Image[r] += s; }
}

// Tukey (jni/MEstimator.h:42-77)
double tukey_sigma_squared(std::vector<double>& v) {
  std::sort(v.begin(), v.end());
  const double med = v[v.size() / 2];
  double sigma = 1.4826 * (1 + 5.0 / (v.size() * 2 - 6)) * sqrt(med);
  sigma = 4.6851 * sigma;
  return sigma * sigma;
}
double tukey_weight(double e2, double s2) { const double sq = (e2 > s2) ? 0.0 : 1.0 - (e2 / s2); return sq * sq; }

// Dynamic inverse as evaluated by the Eigen stand-in: partial-pivot LU, column by column (oracle/shim/Eigen/Dense)
void inverse_lu(const double* m, int n, double* r) {
  std::vector<double> a(m, m + n * n), x(n); std::vector<int> piv(n);
  for (int i = 0; i < n; i++) piv[i] = i;
  for (int k = 0; k < n; k++) {
    int p = k; double best = fabs(a[k * n + k]);
    for (int i = k + 1; i < n; i++) if (fabs(a[i * n + k]) > best) { best = fabs(a[i * n + k]); p = i; }
    if (p != k) { for (int j = 0; j < n; j++) std::swap(a[k * n + j], a[p * n + j]); std::swap(piv[k], piv[p]); }
    for (int i = k + 1; i < n; i++) { a[i * n + k] /= a[k * n + k]; for (int j = k + 1; j < n; j++) a[i * n + j] -= a[i * n + k] * a[k * n + j]; }
  }
  for (int c = 0; c < n; c++) {
    for (int i = 0; i < n; i++) x[i] = (piv[i] == c) ? 1.0 : 0.0;
    for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) x[i] -= a[i * n + j] * x[j];
    for (int i = n - 1; i >= 0; i--) { for (int j = i + 1; j < n; j++) x[i] -= a[i * n + j] * x[j]; x[i] /= a[i * n + i]; }
    for (int i = 0; i < n; i++) r[i * n + c] = x[i];
  }
}


// ------------------------------------------------------------------------------------------------
// SmallBlurryImage (jni/SmallBlurryImage.cc) — the f1 row of SURVEY.md §8: per-frame rotation estimate that seeds the motion model.
struct SBI {
  int w, h;
  std::vector<uint8_t> small;        // mimSmall
  std::vector<float> tmpl;           // mimTemplate (zero-mean, blurred)
  std::vector<float> jac;            // mimImageJacs, 2 floats per pixel
  bool madeJacs;
};
// cv::GaussianBlur(float, ksize 9x9, sigma, BORDER_REPLICATE) as restated by the OpenCV stand-in (oracle/shim/opencv2/core/core.hpp):
// getGaussianKernel taps in float, row pass then column pass, float accumulation in tap order.
void gaussian_blur(std::vector<float>& im, int W, int H, int n /* taps: 9 or 17 */, double sigma) {
  float k[17];
  { const double scale2x = -0.5 / (sigma * sigma); double sum = 0;
    for (int i = 0; i < n; i++) { const double x = i - (n - 1) * 0.5; const double t = std::exp(scale2x * x * x); k[i] = (float)t; sum += k[i]; }
    sum = 1. / sum; for (int i = 0; i < n; i++) k[i] = (float)(k[i] * sum); }
  std::vector<float> tmp((size_t)W * H), out((size_t)W * H);
  for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
    float s = 0; for (int i = 0; i < n; i++) { int xx = x + i - n / 2; xx = xx < 0 ? 0 : (xx >= W ? W - 1 : xx); s += k[i] * im[(size_t)y * W + xx]; }
    tmp[(size_t)y * W + x] = s; }
  for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
    float s = 0; for (int i = 0; i < n; i++) { int yy = y + i - n / 2; yy = yy < 0 ? 0 : (yy >= H ? H - 1 : yy); s += k[i] * tmp[(size_t)yy * W + x]; }
    out[(size_t)y * W + x] = s; }
  im.swap(out);
}
// SmallBlurryImage::MakeFromKF (jni/SmallBlurryImage.cc:20-55); the tracker uses dBlur 0.75 (jni/Tracker.cc:87), the relocaliser
// and MakeKeyFrame_Rest the default 2.5 (jni/SmallBlurryImage.h:19)
void sbi_make(SBI& s, const OKeyFrame& kf, double dBlur) {
  const Image& l3 = kf.lev[3].im;
  s.w = l3.w / 2; s.h = l3.h / 2; s.madeJacs = false;
  // cv::resize(level 3, mimSmall, (cols / 2, rows / 2)), INTER_LINEAR (jni/SmallBlurryImage.cc:22-30): exactly half = (a+b+c+d+2)>>2, odd level-3
  // sizes (1080p: 240 x 135 -> 120 x 67) = OpenCV's fixed-point bilinear; restated in shim/cv_resize_linear_u8.h
  s.small.assign((size_t)s.w * s.h, 0);
  cv_restated::resize_linear_u8(&l3.px[0], l3.w, l3.h, (size_t)l3.w, &s.small[0], s.w, s.h, (size_t)s.w);
  unsigned nSum = 0; for (size_t i = 0; i < s.small.size(); i++) nSum += s.small[i];
  const float fMean = ((float)nSum) / (s.h * s.w);
  s.tmpl.resize((size_t)s.w * s.h);
  for (size_t i = 0; i < s.small.size(); i++) s.tmpl[i] = s.small[i] - fMean;
  gaussian_blur(s.tmpl, s.w, s.h, dBlur <= 2.0 ? 9 : 17, dBlur);   // cv::Size(9,9) / cv::Size(17,17) (jni/SmallBlurryImage.cc:50-54)
}
// SmallBlurryImage::MakeJacs (jni/SmallBlurryImage.cc:58-79)
void sbi_make_jacs(SBI& s) {
  s.jac.assign((size_t)2 * s.w * s.h, 0.f);
  for (int x = 0; x < s.w; x++) for (int y = 0; y < s.h; y++)
    if (x >= 1 && y >= 1 && x < s.w - 1 && y < s.h - 1) {
      s.jac[2 * ((size_t)y * s.w + x)] = s.tmpl[(size_t)y * s.w + x + 1] - s.tmpl[(size_t)y * s.w + x - 1];
      s.jac[2 * ((size_t)y * s.w + x) + 1] = s.tmpl[(size_t)(y + 1) * s.w + x] - s.tmpl[(size_t)(y - 1) * s.w + x];
    }
  s.madeJacs = true;
}
// transform_image, float source and destination (jni/vision/ImageHandler.cpp:3-10,21-113)
void transform_image_f32(const std::vector<float>& in, int iw, int ih, std::vector<float>& out, int w, int h, const double* M, const double* inOrig, const double* outOrig, double def) {
  const double across[2] = {M[0], M[2]}, down[2] = {M[1], M[3]};
  double p0[2];
  { double a = M[0] * outOrig[0]; a += M[1] * outOrig[1]; double b = M[2] * outOrig[0]; b += M[3] * outOrig[1]; p0[0] = inOrig[0] - a; p0[1] = inOrig[1] - b; }
  double min_x = p0[0], min_y = p0[1], max_x = min_x, max_y = min_y;
  if (across[0] < 0) min_x += w * across[0]; else max_x += w * across[0];
  if (down[0] < 0) min_x += h * down[0]; else max_x += h * down[0];
  if (across[1] < 0) min_y += w * across[1]; else max_y += w * across[1];
  if (down[1] < 0) min_y += h * down[1]; else max_y += h * down[1];
  const double cr[2] = {down[0] - w * across[0], down[1] - w * across[1]};
  const bool inside = (min_x >= 0 && min_y >= 0 && max_x < iw - 1 && max_y < ih - 1);
  const float x_bound = iw - 1, y_bound = ih - 1;
  out.resize((size_t)w * h);
  double p[2] = {p0[0], p0[1]};
  for (int i = 0; i < h; ++i, p[0] += cr[0], p[1] += cr[1])
    for (int j = 0; j < w; ++j, p[0] += across[0], p[1] += across[1]) {
      if (inside || (0 <= p[0] && 0 <= p[1] && p[0] < x_bound && p[1] < y_bound)) {
        double x = p[0], y = p[1];
        const int lx = (int)x, ly = (int)y;
        x -= lx; y -= ly;
        const float* r0 = &in[(size_t)ly * iw]; const float* r1 = &in[(size_t)(ly + 1) * iw];
        const double v = (double)((1 - y) * ((1 - x) * r0[lx] + x * r0[lx + 1]) + y * ((1 - x) * r1[lx] + x * r1[lx + 1]));
        out[(size_t)i * w + j] = (float)v;
      } else out[(size_t)i * w + j] = (float)def;
    }
}
struct SE2 { double R[4]; double t[2]; };   // row-major rotation
SE2 se2_identity() { SE2 s; s.R[0] = s.R[3] = 1; s.R[1] = s.R[2] = 0; s.t[0] = s.t[1] = 0; return s; }
SE2 se2_mul(const SE2& a, const SE2& b) {   // jni/RT.h:512-520
  SE2 r;
  for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) { double s = a.R[2 * i] * b.R[j]; s += a.R[2 * i + 1] * b.R[2 + j]; r.R[2 * i + j] = s; }
  for (int i = 0; i < 2; i++) { double s = a.R[2 * i] * b.t[0]; s += a.R[2 * i + 1] * b.t[1]; r.t[i] = a.t[i] + s; }
  return r;
}
SE2 se2_inverse(const SE2& a) {   // jni/RT.h:502-508
  SE2 r; r.R[0] = a.R[0]; r.R[1] = a.R[2]; r.R[2] = a.R[1]; r.R[3] = a.R[3];
  for (int i = 0; i < 2; i++) { double s = r.R[2 * i] * a.t[0]; s += r.R[2 * i + 1] * a.t[1]; r.t[i] = -s; }
  return r;
}
// SmallBlurryImage::IteratePosRelToTarget (jni/SmallBlurryImage.cc:99-222)
SE2 sbi_iterate(const SBI& cur, const SBI& other, int nIterations, double* finalScore) {
  const int W = cur.w, H = cur.h;
  SE2 CtoC = se2_identity(), WfromC = se2_identity();
  const double cx = W / 2.0, cy = H / 2.0;   // irCenter = mirSize / 2
  WfromC.t[0] = cx; WfromC.t[1] = cy;
  double dMeanOffset = 0.0, dFinalScore = 0.0;
  std::vector<float> warped;
  for (int it = 0; it < nIterations; it++) {
    dFinalScore = 0.0;
    double acc[4] = {0, 0, 0, 0}, tri[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const SE2 X = se2_mul(se2_mul(WfromC, CtoC), se2_inverse(WfromC));
    const double zero[2] = {0, 0};
    transform_image_f32(cur.tmpl, W, H, warped, W, H, X.R, X.t, zero, -9e20f);
    for (int i = 0; i < W; i++)
      for (int j = 0; j < H; j++) {
        if (!(i >= 1 && j >= 1 && i < W - 1 && j < H - 1)) continue;
        const float l = warped[(size_t)j * W + i - 1], r = warped[(size_t)j * W + i + 1], u = warped[(size_t)(j - 1) * W + i], d = warped[(size_t)(j + 1) * W + i], here = warped[(size_t)j * W + i];
        if (l + r + u + d + here < -9999.9) continue;
        const double g0 = r - l, g1 = d - u;
        const double s0 = 0.25 * (g0 + other.jac[2 * ((size_t)j * W + i)]), s1 = 0.25 * (g1 + other.jac[2 * ((size_t)j * W + i) + 1]);
        const double J[4] = {s0, s1, -((double)j - cy) * s0 + ((double)i - cx) * s1, 1.0};
        const double dDiff = warped[(size_t)j * W + i] - other.tmpl[(size_t)j * W + i] + dMeanOffset;
        dFinalScore += dDiff * dDiff;
        for (int k = 0; k < 4; k++) acc[k] += dDiff * J[k];
        tri[0] += J[0] * J[0]; tri[1] += J[1] * J[0]; tri[2] += J[1] * J[1]; tri[3] += J[2] * J[0]; tri[4] += J[2] * J[1]; tri[5] += J[2] * J[2];
        tri[6] += J[0]; tri[7] += J[1]; tri[8] += J[2]; tri[9] += 1.0;
      }
    double m4[16]; int v = 0;
    for (int j = 0; j < 4; j++) for (int i = 0; i <= j; i++) { m4[4 * j + i] = m4[4 * i + j] = tri[v++]; }
    double inv[16]; inverse_lu(m4, 4, inv);
    double upd[4];
    for (int i = 0; i < 4; i++) { double s = inv[4 * i] * acc[0]; for (int k = 1; k < 4; k++) s += inv[4 * i + k] * acc[k]; upd[i] = s; }
    SE2 U; U.t[0] = -upd[0]; U.t[1] = -upd[1];
    const double ang = -upd[2];
    U.R[0] = U.R[3] = cos(ang); U.R[2] = sin(ang); U.R[1] = -U.R[2];   // mySO2::exp (jni/RT.h:461-467)
    CtoC = se2_mul(CtoC, U);
    dMeanOffset -= upd[3];
  }
  if (finalScore) *finalScore = dFinalScore;
  return CtoC;
}
// SmallBlurryImage::SE3fromSE2 (jni/SmallBlurryImage.cc:245-333): rotation-only pose whose image motion matches the SE2; returns ln() (6-vector)
SE3 se3_from_se2_pose(const SE2& se2, Cam cam /* at the SBI image size */, int W, int H) {
  double turned[2][2], orig[2][3];
  const double c[2] = {W / 2.0, H / 2.0};
  const double off[2][2] = {{5, 0}, {-5, 0}};
  for (int k = 0; k < 2; k++) {
    for (int i = 0; i < 2; i++) { double s = se2.R[2 * i] * off[k][0]; s += se2.R[2 * i + 1] * off[k][1]; turned[k][i] = c[i] + (se2.t[i] + s); }
    const double im[2] = {c[0] + off[k][0], c[1] + off[k][1]}; double u[2];
    cam_unproject(cam, im, u); orig[k][0] = u[0]; orig[k][1] = u[1]; orig[k][2] = 1.0;
  }
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int it = 0; it < 3; it++) {
    double C[9] = {10, 0, 0, 0, 10, 0, 0, 0, 10}, b[3] = {0, 0, 0};   // myWLS<3>, prior 10
    for (int k = 0; k < 2; k++) {
      double v3[3]; mat3_mul_vec(R, orig[k], v3);
      double pix[2]; cam_project(cam, v3[0] / v3[2], v3[1] / v3[2], pix);
      const double err[2] = {turned[k][0] - pix[0], turned[k][1] - pix[1]};
      double dv[4]; cam_derivs(cam, dv);
      double J[2][3];
      const double invz = 1.0 / v3[2];
      for (int m = 0; m < 3; m++) {
        double mo[3]; mo[m] = 0; mo[(m + 1) % 3] = -v3[(m + 2) % 3]; mo[(m + 2) % 3] = v3[(m + 1) % 3];   // mySO3::generator_field (jni/RT.h:71-78)
        const double c0 = (mo[0] - v3[0] * mo[2] * invz) * invz, c1 = (mo[1] - v3[1] * mo[2] * invz) * invz;
        double a0 = dv[0] * c0; a0 += dv[1] * c1; double a1 = dv[2] * c0; a1 += dv[3] * c1;
        J[0][m] = a0; J[1][m] = a1;
      }
      for (int row = 0; row < 2; row++)
        for (int r = 0; r < 3; r++) { const double Jw = 1.0 * J[row][r]; b[r] += err[row] * Jw; for (int cc = r; cc < 3; cc++) C[3 * r + cc] += Jw * J[row][cc]; }
    }
    for (int r = 1; r < 3; r++) for (int cc = 0; cc < r; cc++) C[3 * r + cc] = C[3 * cc + r];
    double Ci[9]; inverse3(C, Ci);
    double mu[3]; for (int i = 0; i < 3; i++) { double s = Ci[3 * i] * b[0]; s += Ci[3 * i + 1] * b[1]; s += Ci[3 * i + 2] * b[2]; mu[i] = s; }
    double E[9], Rn[9]; so3_exp(mu, E); mat3_mul(E, R, Rn); memcpy(R, Rn, sizeof(R));
  }
  SE3 s = se3_identity(); memcpy(s.R, R, sizeof(R));
  return s;
}
void se3_from_se2(const SE2& se2, Cam cam, int W, int H, double* v6) { se3_ln(se3_from_se2_pose(se2, cam, W, H), v6); }

// ------------------------------------------------------------------------------------------------
// Tracker (jni/Tracker.cc)
struct OTracker {
  Cam cam; int P;
  std::vector<MapPointO> pts; std::vector<TData> td; std::vector<char> hasTD;
  const OKeyFrame* srcKF;
  OKeyFrame cur;
  SE3 pose, startPose;
  double velocity[6], sbiRot[6]; bool useSBI;
  double msdScaledVel, velMag;
  double sceneDepthMean, sceneDepthSigma;
  int attempted[LEVELS], foundCnt[LEVELS];
  int quality /*0 BAD 1 DODGY 2 GOOD*/, lostFrames; bool didCoarse, justRecovered;
  bool truncateError;   // (int) cast at jni/Tracker.cc:766-767; true = reference behaviour
  GlibcRand rng;
  long zmssdEvals;
  std::vector<double> updates;  // 6-vectors of every CalcPoseUpdate of the last TrackMap, in order
  std::vector<double> sigmas;   // sigma^2 used by each of them
  // SmallBlurryImage state (Tracker::mpSBIThisFrame / mpSBILastFrame, jni/Tracker.cc:86-97)
  bool computeSBI, haveSBI; SBI sbiThis, sbiLast; Cam sbiCam; int nFrame;
  // Relocaliser (jni/Relocaliser.cc): the map keyframes' SmallBlurryImages (blur 2.5, with gradient images) and poses
  std::vector<SBI> relocSBI; std::vector<SE3> relocPose; int relocBest; double relocScore; int nRecoveries;
  // keyframe hand-off (jni/Tracker.cc:127-132, :866-872): MapMaker's heuristics over the map keyframes (= relocPose), off by default
  bool kfPolicy = false; double kfWiggle = 0.1, kfWiggleDN = 0.1, kfMult = 0.2; int kfMinFrames = 20, lastKeyFrameDropped = -20, kfAddedThisFrame = 0;
};

TData& ensure_td(OTracker& t, int i) {
  if (!t.hasTD[i]) { t.td[i].finder.init(t.P); t.td[i].found = t.td[i].searched = t.td[i].didSubPix = false; t.td[i].inImage = false; t.td[i].searchLevel = -1;
    memset(t.td[i].jac, 0, sizeof(t.td[i].jac)); memset(t.td[i].derivs, 0, sizeof(t.td[i].derivs)); memset(t.td[i].v2Image, 0, 16); memset(t.td[i].v2Found, 0, 16);
    memset(t.td[i].v3Cam, 0, 24); memset(t.td[i].err, 0, 16); t.td[i].sqrtInvNoise = 0; t.hasTD[i] = 1; }
  return t.td[i];
}

// Tracker::SearchForPoints (jni/Tracker.cc:629-674)
int search_for_points(OTracker& t, const std::vector<int>& v, int nRange, int nSubPixIts) {
  int nFound = 0;
  for (size_t k = 0; k < v.size(); k++) {
    TData& TD = t.td[v[k]]; Finder& F = TD.finder;
    g_dbg = (v[k] == g_dbg_point);
    make_template_coarse_cont(F, t.pts[v[k]], *t.pts[v[k]].srcKF);
    if (F.templateBad) { TD.inImage = TD.potentiallyVisible = TD.found = false; continue; }
    t.attempted[F.level]++;
    const bool bFound = find_patch_coarse(F, TD.v2Image[0], TD.v2Image[1], t.cur, nRange, &t.zmssdEvals, 0);
    TD.searched = true;
    if (!bFound) { TD.found = false; continue; }
    TD.found = true;
    TD.sqrtInvNoise = (1.0 / LevelScale(F.level));
    nFound++; t.foundCnt[F.level]++;
    if (nSubPixIts > 0) {
      TD.didSubPix = true;
      make_subpix_template(F);
      if (!iterate_subpix_to_convergence(F, t.cur, nSubPixIts)) { TD.found = false; nFound--; t.foundCnt[F.level]--; continue; }
      TD.v2Found[0] = F.subPixPos[0]; TD.v2Found[1] = F.subPixPos[1];
    } else { TD.v2Found[0] = F.coarsePos[0]; TD.v2Found[1] = F.coarsePos[1]; TD.didSubPix = false; }
  }
  return nFound;
}

// Tracker::CalcPoseUpdate (jni/Tracker.cc:683-774) with myWLS<6> (jni/myWLS.h:29-62)
void calc_pose_update(OTracker& t, const std::vector<int>& v, double dOverrideSigma, bool bMarkOutliers, double* out6, double* sigmaOut) {
  std::vector<double> errSq;
  for (size_t k = 0; k < v.size(); k++) {
    TData& TD = t.td[v[k]];
    if (!TD.found) continue;
    TD.err[0] = (TD.v2Found[0] - TD.v2Image[0]) * TD.sqrtInvNoise; TD.err[1] = (TD.v2Found[1] - TD.v2Image[1]) * TD.sqrtInvNoise;
    double e = 0; e += TD.err[0] * TD.err[0]; e += TD.err[1] * TD.err[1];
    errSq.push_back(e);
  }
  for (int i = 0; i < 6; i++) out6[i] = 0;
  if (sigmaOut) *sigmaOut = 0;
  if (errSq.size() == 0) return;
  double dSigmaSquared = (dOverrideSigma > 0) ? dOverrideSigma : tukey_sigma_squared(errSq);
  if (sigmaOut) *sigmaOut = dSigmaSquared;
  double C[36], b[6];
  memset(C, 0, sizeof(C)); memset(b, 0, sizeof(b));
  for (int i = 0; i < 6; i++) C[7 * i] += 100.0;
  for (size_t k = 0; k < v.size(); k++) {
    TData& TD = t.td[v[k]];
    if (!TD.found) continue;
    double e2 = 0; e2 += TD.err[0] * TD.err[0]; e2 += TD.err[1] * TD.err[1];
    const double w = tukey_weight(e2, dSigmaSquared);
    if (w == 0.0) { if (bMarkOutliers) t.pts[v[k]].outlierCount++; continue; }
    else if (bMarkOutliers) t.pts[v[k]].inlierCount++;
    for (int row = 0; row < 2; row++) {
      double J[6]; for (int c = 0; c < 6; c++) J[c] = TD.sqrtInvNoise * TD.jac[6 * row + c];
      const double m = t.truncateError ? (double)(int)TD.err[row] : TD.err[row];
      for (int r = 0; r < 6; r++) { const double Jw = w * J[r]; b[r] += m * Jw; for (int c = r; c < 6; c++) C[6 * r + c] += Jw * J[c]; }
    }
  }
  for (int r = 1; r < 6; r++) for (int c = 0; c < r; c++) C[6 * r + c] = C[6 * c + r];
  double Ci[36]; inverse_lu(C, 6, Ci);
  for (int i = 0; i < 6; i++) { double s = Ci[6 * i] * b[0]; for (int j = 1; j < 6; j++) s += Ci[6 * i + j] * b[j]; out6[i] = s; }
}

// std::random_shuffle(first,last) of this libstdc++ (bits/stl_algo.h:4581-4597) drawing from rand()
void random_shuffle(std::vector<int>& v, GlibcRand& rng) {
  for (size_t i = 1; i < v.size(); i++) { const size_t j = (size_t)(rng.next() % (int)(i + 1)); if (i != j) std::swap(v[i], v[j]); }
}

void pose_step(OTracker& t, const std::vector<int>& set, double sigma, bool mark, double* upd) {
  double s2; calc_pose_update(t, set, sigma, mark, upd, &s2);
  t.pose = se3_mul(se3_exp(upd), t.pose);
  for (int i = 0; i < 6; i++) t.updates.push_back(upd[i]);
  t.sigmas.push_back(s2);
}

// Tracker::TrackMap (jni/Tracker.cc:358-626)
void track_map(OTracker& t) {
  for (int i = 0; i < LEVELS; i++) t.attempted[i] = t.foundCnt[i] = 0;
  t.updates.clear(); t.sigmas.clear();
  std::vector<int> avPVS[LEVELS];
  for (size_t i = 0; i < t.pts.size(); i++) {  // :369-392
    TData& TD = ensure_td(t, (int)i);
    td_project(TD, t.pts[i], t.pose, t.cam);
    if (!TD.inImage) continue;
    cam_derivs(t.cam, TD.derivs);
    TD.searchLevel = calc_search_level_and_warp(TD.finder, t.pts[i], t.pose, TD.derivs);
    if (TD.searchLevel == -1) continue;
    TD.searched = false; TD.found = false;
    avPVS[TD.searchLevel].push_back((int)i);
  }
  for (int i = 0; i < LEVELS; i++) random_shuffle(avPVS[i], t.rng);  // :396-397
  std::vector<int> vNext, vIter;
  const unsigned gvnCoarseMin = 20, gvnCoarseMax = 60, gvnCoarseRange = 30; const int gvnCoarseSubPixIts = 8; const double gvdCoarseMinVel = 0.006;  // :405-410
  unsigned nCoarseMax = gvnCoarseMax, nCoarseRange = gvnCoarseRange;
  t.didCoarse = false;
  bool bTryCoarse = true;
  if (t.msdScaledVel < gvdCoarseMinVel || nCoarseMax == 0) bTryCoarse = false;
  if (t.justRecovered) { bTryCoarse = true; nCoarseMax *= 2; nCoarseRange *= 2; t.justRecovered = false; }
  if (bTryCoarse && avPVS[LEVELS - 1].size() + avPVS[LEVELS - 2].size() > gvnCoarseMin) {  // :439-490
    if (avPVS[LEVELS - 1].size() <= nCoarseMax) { vNext = avPVS[LEVELS - 1]; avPVS[LEVELS - 1].clear(); }
    else { for (unsigned i = 0; i < nCoarseMax; i++) vNext.push_back(avPVS[LEVELS - 1][i]); avPVS[LEVELS - 1].erase(avPVS[LEVELS - 1].begin(), avPVS[LEVELS - 1].begin() + nCoarseMax); }
    if (vNext.size() < nCoarseMax) {
      const unsigned more = nCoarseMax - (unsigned)vNext.size();
      if (avPVS[LEVELS - 2].size() <= more) { vNext = avPVS[LEVELS - 2]; avPVS[LEVELS - 2].clear(); }  // overwrite quirk (:454-456)
      else { for (unsigned i = 0; i < more; i++) vNext.push_back(avPVS[LEVELS - 2][i]); avPVS[LEVELS - 2].erase(avPVS[LEVELS - 2].begin(), avPVS[LEVELS - 2].begin() + more); }
    }
    const unsigned nFound = search_for_points(t, vNext, nCoarseRange, gvnCoarseSubPixIts);
    vIter = vNext;
    if (nFound >= gvnCoarseMin) {
      t.didCoarse = true;
      for (int iter = 0; iter < 10; iter++) {
        if (iter != 0) for (size_t i = 0; i < vIter.size(); i++) if (t.td[vIter[i]].found) td_project_and_derivs(t.td[vIter[i]], t.pts[vIter[i]], t.pose, t.cam);
        for (size_t i = 0; i < vIter.size(); i++) if (t.td[vIter[i]].found) td_calc_jacobian(t.td[vIter[i]]);
        double upd[6]; pose_step(t, vIter, iter > 5 ? 1.0 : 0.0, false, upd);
      }
    }
  }
  const int nFineRange = t.didCoarse ? 5 : 10;  // :495-497
  {
    const int l = LEVELS - 1;
    for (size_t i = 0; i < avPVS[l].size(); i++) td_project_and_derivs(t.td[avPVS[l][i]], t.pts[avPVS[l][i]], t.pose, t.cam);
    search_for_points(t, avPVS[l], nFineRange, 8);
    for (size_t i = 0; i < avPVS[l].size(); i++) vIter.push_back(avPVS[l][i]);
  }
  vNext.clear();
  for (int l = LEVELS - 2; l >= 0; l--) for (size_t i = 0; i < avPVS[l].size(); i++) vNext.push_back(avPVS[l][i]);
  int nFinePatchesToUse = 1000 - (int)vIter.size();  // :518-527
  if (nFinePatchesToUse < 0) nFinePatchesToUse = 0;
  if ((int)vNext.size() > nFinePatchesToUse) { random_shuffle(vNext, t.rng); vNext.resize(nFinePatchesToUse); }
  if (t.didCoarse) for (size_t i = 0; i < vNext.size(); i++) td_project_and_derivs(t.td[vNext[i]], t.pts[vNext[i]], t.pose, t.cam);
  search_for_points(t, vNext, nFineRange, 0);
  for (size_t i = 0; i < vNext.size(); i++) vIter.push_back(vNext[i]);
  double last[6] = {0, 0, 0, 0, 0, 0};
  for (int iter = 0; iter < 10; iter++) {  // :543-577
    const bool nonlinear = (iter == 0 || iter == 4 || iter == 9);
    if (iter != 0) {
      if (nonlinear) { for (size_t i = 0; i < vIter.size(); i++) if (t.td[vIter[i]].found) td_project_and_derivs(t.td[vIter[i]], t.pts[vIter[i]], t.pose, t.cam); }
      else { for (size_t i = 0; i < vIter.size(); i++) if (t.td[vIter[i]].found) td_linear_update(t.td[vIter[i]], last); }
    }
    if (nonlinear) for (size_t i = 0; i < vIter.size(); i++) if (t.td[vIter[i]].found) td_calc_jacobian(t.td[vIter[i]]);
    double upd[6]; pose_step(t, vIter, iter > 5 ? 16.0 : 0.0, iter == 9, upd);
    memcpy(last, upd, sizeof(last));
  }
  {  // scene depth (:610-625)
    double dSum = 0, dSumSq = 0; int nNum = 0;
    for (size_t i = 0; i < vIter.size(); i++) if (t.td[vIter[i]].found) { const double z = t.td[vIter[i]].v3Cam[2]; dSum += z; dSumSq += z * z; nNum++; }
    if (nNum > 20) { t.sceneDepthMean = dSum / nNum; t.sceneDepthSigma = sqrt((dSumSq / nNum) - (t.sceneDepthMean) * (t.sceneDepthMean)); }
  }
}
// Tracker::ApplyMotionModel (jni/Tracker.cc:781-798)
void apply_motion_model(OTracker& t) {
  double v[6]; memcpy(v, t.velocity, sizeof(v));
  t.startPose = t.pose;
  if (t.useSBI) { v[0] = 0.0; v[1] = 0.0; v[2] = t.velocity[2]; v[3] = t.sbiRot[3]; v[4] = t.sbiRot[4]; v[5] = t.sbiRot[5]; }
  t.pose = se3_mul(se3_exp(v), t.startPose);
}
// Tracker::UpdateMotionModel (jni/Tracker.cc:802-820)
void update_motion_model(OTracker& t) {
  const SE3 nfo = se3_mul(t.pose, se3_inverse(t.startPose));
  double m[6]; se3_ln(nfo, m);
  for (int i = 0; i < 6; i++) t.velocity[i] = 0.9 * (0.5 * m[i] + 0.5 * t.velocity[i]);
  double s = 0; for (int i = 0; i < 6; i++) s += t.velocity[i] * t.velocity[i];
  t.velMag = sqrt(s);
  double v[6]; memcpy(v, t.velocity, sizeof(v));
  for (int i = 0; i < 3; i++) v[i] *= 1.0 / t.sceneDepthMean;
  s = 0; for (int i = 0; i < 6; i++) s += v[i] * v[i];
  t.msdScaledVel = sqrt(s);
}
// Tracker::AssessTrackingQuality (jni/Tracker.cc:832-878); MapMaker::IsDistanceToNearestKeyFrameExcessive is out of scope (false)
// MapMaker::KeyFrameLinearDist (jni/MapMaker.cc:705-712) of the current keyframe (pose = the tracker's, jni/Tracker.cc:594) to each
// map keyframe, minimum as MapMaker::ClosestKeyFrame takes it (jni/MapMaker.cc:736-754: strict <, first wins).
double dist_to_nearest_keyframe(OTracker& t, int* closest) {
  double dClosestDist = 9999999999.9; int nClosest = -1;
  const SE3 cur = se3_inverse(t.pose);
  for (size_t i = 0; i < t.relocPose.size(); i++) {
    const SE3 kf = se3_inverse(t.relocPose[i]);
    const double d[3] = {kf.t[0] - cur.t[0], kf.t[1] - cur.t[1], kf.t[2] - cur.t[2]};
    double dd = d[0] * d[0]; dd += d[1] * d[1]; dd += d[2] * d[2];
    const double dDist = sqrt(dd);
    if (dDist < dClosestDist) { dClosestDist = dDist; nClosest = (int)i; }
  }
  if (closest) *closest = nClosest;
  return dClosestDist;
}

void assess_tracking_quality(OTracker& t) {
  int nTotalAttempted = 0, nTotalFound = 0, nLargeAttempted = 0, nLargeFound = 0;
  for (int i = 0; i < LEVELS; i++) { nTotalAttempted += t.attempted[i]; nTotalFound += t.foundCnt[i]; if (i >= 2) { nLargeAttempted += t.attempted[i]; nLargeFound += t.foundCnt[i]; } }
  if (nTotalFound == 0 || nTotalAttempted == 0) t.quality = 0;
  else {
    const double dTotalFracFound = (double)nTotalFound / nTotalAttempted;
    const double dLargeFracFound = (nLargeAttempted > 10) ? (double)nLargeFound / nLargeAttempted : dTotalFracFound;
    if (dTotalFracFound > 0.3) t.quality = 2; else if (dLargeFracFound < 0.13) t.quality = 0; else t.quality = 1;
  }
  // jni/Tracker.cc:866-872: a DODGY tracker whose pose ran far away from every keyframe is BAD
  // (MapMaker::IsDistanceToNearestKeyFrameExcessive, jni/MapMaker.cc:1098-1101: DistToNearestKeyFrame > mdWiggleScale * 10)
  if (t.quality == 1 && t.kfPolicy && !t.relocPose.empty()) { if (dist_to_nearest_keyframe(t, 0) > t.kfWiggle * 10.0) t.quality = 0; }
  if (t.quality == 0) t.lostFrames++; else t.lostFrames = 0;
}

Cam cam_from13(const double* s) {
  Cam c; memset(&c, 0, sizeof(c));
  c.fx = s[0]; c.fy = s[1]; c.cx = s[2]; c.cy = s[3]; c.W = s[4]; c.Winv = s[5]; c.twoTan = s[6]; c.oneOver2Tan = s[7]; c.distEnabled = s[8];
  c.largestRadius = s[9]; c.maxR = s[10]; c.width = s[11]; c.height = s[12];
  return c;
}

}  // namespace

// =================================================================================================
// SmallBlurryImage::ZMSSD (jni/SmallBlurryImage.cc:82-94): plain SSD of the two zero-mean templates, x outer / y inner, float difference
double sbi_zmssd(const SBI& a, const SBI& b) {
  double dSSD = 0.0;
  for (int x = 0; x < a.w; x++) for (int y = 0; y < a.h; y++) { const double dDiff = a.tmpl[(size_t)y * a.w + x] - b.tmpl[(size_t)y * a.w + x]; dSSD += dDiff * dDiff; }
  return dSSD;
}
// Relocaliser::AttemptRecovery + ScoreKFs (jni/Relocaliser.cc:17-58) and Tracker::AttemptRecovery (jni/Tracker.cc:167-180)
bool attempt_recovery(OTracker& t) {
  SBI cur; sbi_make(cur, t.cur, 2.5);
  double best = 99999999999999.9; int nBest = -1;
  for (size_t i = 0; i < t.relocSBI.size(); i++) { const double d = sbi_zmssd(cur, t.relocSBI[i]); if (d < best) { best = d; nBest = (int)i; } }
  double dScore = 0;
  const SE2 se2 = sbi_iterate(cur, t.relocSBI[nBest], 6, &dScore);
  const SE3 se3Best = se3_mul(se3_from_se2_pose(se2, t.sbiCam, cur.w, cur.h), t.relocPose[nBest]);
  t.relocBest = nBest; t.relocScore = dScore;
  if (!(dScore < 9e6)) return false;
  t.pose = t.startPose = se3Best;
  memset(t.velocity, 0, sizeof(t.velocity));
  t.justRecovered = true;
  t.nRecoveries++;
  return true;
}

extern "C" {

void orc_debug_point(int i) { g_dbg_point = i; }
// ---- RNG
void* orc_rand_create(unsigned seed) { GlibcRand* r = new GlibcRand(); r->seed(seed); return r; }
int orc_rand_next(void* r) { return ((GlibcRand*)r)->next(); }
void orc_rand_destroy(void* r) { delete (GlibcRand*)r; }

// ---- KeyFrame
void* orc_kf_create() { return new OKeyFrame(); }
void orc_kf_destroy(void* k) { delete (OKeyFrame*)k; }
void orc_kf_make_lite(void* k, const uint8_t* gray, int w, int h, int stride) { make_keyframe_lite(*(OKeyFrame*)k, gray, w, h, stride); }
void orc_kf_make_rest(void* k) { make_keyframe_rest(*(OKeyFrame*)k); }
void orc_kf_level_dims(void* k, int l, int* w, int* h) { const Image& im = ((OKeyFrame*)k)->lev[l].im; *w = im.w; *h = im.h; }
void orc_kf_level_pixels(void* k, int l, uint8_t* out) { const Image& im = ((OKeyFrame*)k)->lev[l].im; memcpy(out, &im.px[0], im.px.size()); }
int orc_kf_num_corners(void* k, int l) { return (int)((OKeyFrame*)k)->lev[l].corners.size(); }
void orc_kf_corners(void* k, int l, int32_t* xy) { const std::vector<Corner>& c = ((OKeyFrame*)k)->lev[l].corners; for (size_t i = 0; i < c.size(); i++) { xy[2 * i] = c[i].x; xy[2 * i + 1] = c[i].y; } }
int orc_kf_row_lut(void* k, int l, int32_t* out) { const std::vector<int>& v = ((OKeyFrame*)k)->lev[l].lut; for (size_t i = 0; i < v.size(); i++) out[i] = v[i]; return (int)v.size(); }
int orc_kf_num_max_corners(void* k, int l) { return (int)((OKeyFrame*)k)->lev[l].maxCorners.size(); }
void orc_kf_max_corners(void* k, int l, int32_t* xy) { const std::vector<Corner>& c = ((OKeyFrame*)k)->lev[l].maxCorners; for (size_t i = 0; i < c.size(); i++) { xy[2 * i] = c[i].x; xy[2 * i + 1] = c[i].y; } }
int orc_kf_num_candidates(void* k, int l) { return (int)((OKeyFrame*)k)->lev[l].candidates.size(); }
void orc_kf_candidates(void* k, int l, int32_t* xy, double* sc) {
  const OLevel& L = ((OKeyFrame*)k)->lev[l];
  for (size_t i = 0; i < L.candidates.size(); i++) { xy[2 * i] = L.candidates[i].x; xy[2 * i + 1] = L.candidates[i].y; sc[i] = L.candScores[i]; }
}
void orc_kf_fast_scores(void* k, int l, int barrier, int32_t* out) {
  const OLevel& L = ((OKeyFrame*)k)->lev[l];
  for (size_t i = 0; i < L.corners.size(); i++) out[i] = fast_score(L.im, L.corners[i].x, L.corners[i].y, barrier);
}
double orc_shi_tomasi(void* k, int l, int nsize, int px, int py) { return shi_tomasi(((OKeyFrame*)k)->lev[l].im, nsize, px, py); }

// ---- camera (13 scalars in, see synth.Camera.scalars)
void orc_cam_project(const double* cam13, const double* cam2, double* im2, int* invalid, double* derivs4) {
  Cam c = cam_from13(cam13); cam_project(c, cam2[0], cam2[1], im2); if (invalid) *invalid = c.invalid; if (derivs4) cam_derivs(c, derivs4);
}
void orc_cam_unproject(const double* cam13, const double* im2, double* cam2) { Cam c = cam_from13(cam13); cam_unproject(c, im2, cam2); }

// Patch-source fields of a new map point (tail of MapMaker::AddPointEpipolar, jni/MapMaker.cc:655-684) + MapPoint::RefreshPixelVectors
// (jni/MapPoint.cc:4-29) with v3Normal_NC = (0,0,-1).  out15 = Center_NC, OneRightFromCenter_NC, OneDownFromCenter_NC, PixelRight_W, PixelDown_W.
void orc_epipolar_point_fields(const double* cam13, const double* src12, int level, int cx, int cy, const double* world3, double* out15) {
  Cam cam = cam_from13(cam13);
  const SE3 S = se3_from12(src12);
  const int nLevelScale = 1 << level;
  const double root[2] = {(cx + 0.5) * nLevelScale - 0.5, (cy + 0.5) * nLevelScale - 0.5};
  const double at[3][2] = {{root[0], root[1]}, {root[0] + nLevelScale, root[1]}, {root[0], root[1] + nLevelScale}};
  double ray[3][3];
  for (int k = 0; k < 3; k++) {
    double u[2]; cam_unproject(cam, at[k], u);
    ray[k][0] = u[0]; ray[k][1] = u[1]; ray[k][2] = 1.0;
    double nn = ray[k][0] * ray[k][0]; nn += ray[k][1] * ray[k][1]; nn += ray[k][2] * ray[k][2];
    const double nrm = sqrt(nn); for (int q = 0; q < 3; q++) ray[k][q] /= nrm;
  }
  double pc[3]; se3_apply(S, world3, pc);
  const double nrmz[3] = {0, 0, -1};
  auto dot3 = [](const double* a, const double* b) { double s = a[0] * b[0]; s += a[1] * b[1]; s += a[2] * b[2]; return s; };
  const double dCamHeight = fabs(dot3(pc, nrmz));
  double on[3][3];
  for (int k = 0; k < 3; k++) { const double rate = fabs(dot3(ray[k], nrmz)); for (int q = 0; q < 3; q++) on[k][q] = ray[k][q] * dCamHeight / rate; }
  const SE3 Sinv = se3_inverse(S);
  for (int k = 1; k < 3; k++) { double d[3]; for (int q = 0; q < 3; q++) d[q] = on[k][q] - on[0][q]; mat3_mul_vec(Sinv.R, d, out15 + 6 + 3 * k); }
  for (int k = 0; k < 3; k++) for (int q = 0; q < 3; q++) out15[3 * k + q] = ray[k][q];
}

// ---- SE3
void orc_se3_exp(const double* mu6, double* pose12) { se3_to12(se3_exp(mu6), pose12); }
void orc_se3_ln(const double* pose12, double* mu6) { se3_ln(se3_from12(pose12), mu6); }
void orc_se3_mul(const double* a, const double* b, double* o) { se3_to12(se3_mul(se3_from12(a), se3_from12(b)), o); }
void orc_se3_inverse(const double* a, double* o) { se3_to12(se3_inverse(se3_from12(a)), o); }

// ---- stand-alone PatchFinder pieces
// template from an explicit m2 (= inverse(warp) * levelscale), row-major; returns samples outside
int orc_make_template(void* srckf, int srcLevel, const int32_t* irCenter2, int P, const double* m2, uint8_t* tmpl, int* sum, int* sumsq) {
  const double inOrig[2] = {(double)irCenter2[0], (double)irCenter2[1]}, outOrig[2] = {(double)(P / 2), (double)(P / 2)};
  const int n = transform_image_u8(((OKeyFrame*)srckf)->lev[srcLevel].im, tmpl, P, m2, inOrig, outOrig);
  int s = 0, q = 0; for (int i = 0; i < P * P; i++) { s += tmpl[i]; q += tmpl[i] * tmpl[i]; } *sum = s; *sumsq = q;
  return n;
}
int orc_zmssd(void* kf, int level, const uint8_t* tmpl, int P, int x, int y) {
  Finder f; f.init(P); memcpy(&f.tmpl[0], tmpl, P * P); make_template_sums(f);
  return zmssd_at_point(f, ((OKeyFrame*)kf)->lev[level].im, x, y);
}
// returns found; pos2 = coarse position (L0); best = best ZMSSD; evals = number of ZMSSD evaluations
int orc_find_patch_coarse(void* kf, int level, const uint8_t* tmpl, int P, double x, double y, unsigned range, double* pos2, int* best, long* evals) {
  Finder f; f.init(P); memcpy(&f.tmpl[0], tmpl, P * P); make_template_sums(f); f.level = level;
  long st = 0; const bool ok = find_patch_coarse(f, x, y, *(OKeyFrame*)kf, range, &st, best);
  if (ok) { pos2[0] = f.coarsePos[0]; pos2[1] = f.coarsePos[1]; }
  if (evals) *evals = st;
  return ok;
}
int orc_subpix(void* kf, int level, const uint8_t* tmpl, int P, const double* coarse2, int max_its, double* pos2, double* hinv9) {
  Finder f; f.init(P); memcpy(&f.tmpl[0], tmpl, P * P); make_template_sums(f); f.level = level;
  f.coarsePos[0] = coarse2[0]; f.coarsePos[1] = coarse2[1];
  make_subpix_template(f);
  if (hinv9) memcpy(hinv9, f.hinv, sizeof(f.hinv));
  const bool ok = iterate_subpix_to_convergence(f, *(OKeyFrame*)kf, max_its);
  pos2[0] = f.subPixPos[0]; pos2[1] = f.subPixPos[1];
  return ok;
}
// MiniPatch: patch sampled at (sx,sy) of `src` level 0 (MiniPatch::SampleFromImage, jni/MiniPatch.cc:73-83), searched in `kf` level 0
int orc_minipatch_find(void* src, int sx, int sy, void* kf, double* pos2, int range, int use_lut, int max_ssd, int* best) {
  const int half = 4, n = 9; uint8_t patch[81];
  const Image& s = ((OKeyFrame*)src)->lev[0].im;
  for (int r = 0; r < n; r++) memcpy(patch + r * n, s.row(sy - half + r) + (sx - half), n);
  return minipatch_find(patch, half, max_ssd, pos2, ((OKeyFrame*)kf)->lev[0], range, use_lut != 0, best);
}

// ---- Trail tracking for the initial map (Tracker::TrailTracking_Start / _Advance, jni/Tracker.cc:264-346)
struct OTrail { uint8_t patch[81]; double cur[2], init[2]; };
struct OTrails { std::list<OTrail> trails; OKeyFrame prev; };
// jni/Tracker.h:47-52: `lhs.first > rhs.first` on first = -dSTScore, i.e. the LOWEST Shi-Tomasi scores come first (the comment at
// jni/Tracker.cc:275 says the opposite; reproduced as shipped — it only matters when there are more than 1000 candidates)
struct CompareFirstO { bool operator()(const std::pair<double, Corner>& a, const std::pair<double, Corner>& b) const { return a.first > b.first; } };
void* orc_trails_create() { return new OTrails(); }
void orc_trails_destroy(void* t) { delete (OTrails*)t; }
// jni/Tracker.cc:264-292: MakeKeyFrame_Rest, level-0 candidates inside the MiniPatch border sorted by -dSTScore (std::sort, the
// reference's comparator), at most 1000 trails, patches sampled from the frame, previous frame = this frame.
int orc_trails_start(void* t_, void* kf_) {
  OTrails* t = (OTrails*)t_; OKeyFrame* kf = (OKeyFrame*)kf_;
  make_keyframe_rest(*kf);
  const OLevel& L = kf->lev[0];
  const int half = 4;
  std::vector<std::pair<double, Corner> > v;
  for (size_t i = 0; i < L.candidates.size(); i++) {
    const Corner c = L.candidates[i];
    if (!(c.x >= half && c.y >= half && c.x < L.im.w - half && c.y < L.im.h - half)) continue;
    v.push_back(std::make_pair(-1.0 * L.candScores[i], c));
  }
  std::sort(v.begin(), v.end(), CompareFirstO());
  int nToAdd = 1000;
  t->trails.clear();
  for (size_t i = 0; i < v.size() && nToAdd > 0; i++) {
    const Corner c = v[i].second;
    if (!(c.x >= half && c.y >= half && c.x < L.im.w - half && c.y < L.im.h - half)) continue;
    OTrail tr;
    for (int r = 0; r < 9; r++) memcpy(tr.patch + r * 9, L.im.row(c.y - half + r) + (c.x - half), 9);
    tr.init[0] = tr.cur[0] = c.x; tr.init[1] = tr.cur[1] = c.y;
    t->trails.push_back(tr);
    nToAdd--;
  }
  t->prev = *kf;
  return (int)t->trails.size();
}
// jni/Tracker.cc:294-346: forward FindPatch (range 10, no LUT argument => linear corner scan), married-match check backwards into
// the previous frame, erase trails that fail; nGoodTrails counts the forward hits.
int orc_trails_advance(void* t_, void* kf_, int max_ssd) {
  OTrails* t = (OTrails*)t_; OKeyFrame* kf = (OKeyFrame*)kf_;
  const OLevel& cur = kf->lev[0]; const OLevel& prev = t->prev.lev[0];
  const int half = 4;
  int nGood = 0;
  for (std::list<OTrail>::iterator i = t->trails.begin(); i != t->trails.end();) {
    std::list<OTrail>::iterator next = i; ++next;
    OTrail& tr = *i;
    const double start[2] = {tr.cur[0], tr.cur[1]};
    double end[2] = {start[0], start[1]};
    bool bFound = minipatch_find(tr.patch, half, max_ssd, end, cur, 10, false, 0);
    if (bFound) {
      uint8_t back[81];
      const int ex = (int)end[0], ey = (int)end[1];
      for (int r = 0; r < 9; r++) memcpy(back + r * 9, cur.im.row(ey - half + r) + (ex - half), 9);
      double bw[2] = {end[0], end[1]};
      bFound = minipatch_find(back, half, max_ssd, bw, prev, 10, false, 0);
      const double dx = bw[0] - start[0], dy = bw[1] - start[1];
      if (dx * dx + dy * dy > 2) bFound = false;
      tr.cur[0] = end[0]; tr.cur[1] = end[1];
      nGood++;
    }
    if (!bFound) t->trails.erase(i);
    i = next;
  }
  t->prev = *kf;
  return nGood;
}
int orc_trails_count(void* t) { return (int)((OTrails*)t)->trails.size(); }
void orc_trails_get(void* t, double* init_cur4) {
  int k = 0;
  for (std::list<OTrail>::iterator i = ((OTrails*)t)->trails.begin(); i != ((OTrails*)t)->trails.end(); ++i, ++k) {
    init_cur4[4 * k] = i->init[0]; init_cur4[4 * k + 1] = i->init[1]; init_cur4[4 * k + 2] = i->cur[0]; init_cur4[4 * k + 3] = i->cur[1];
  }
}

// ---- Tracker
void* orc_tracker_create(const double* cam13, int P) {
  OTracker* t = new OTracker();
  t->cam = cam_from13(cam13); t->P = P; t->srcKF = 0; t->pose = se3_identity(); t->startPose = se3_identity();
  memset(t->velocity, 0, sizeof(t->velocity)); memset(t->sbiRot, 0, sizeof(t->sbiRot)); t->useSBI = true;
  t->msdScaledVel = 0; t->velMag = 0; t->sceneDepthMean = 1.0; t->sceneDepthSigma = 1.0;
  for (int i = 0; i < LEVELS; i++) t->attempted[i] = t->foundCnt[i] = 0;
  t->quality = 2; t->lostFrames = 0; t->didCoarse = false; t->justRecovered = false; t->truncateError = true; t->zmssdEvals = 0;
  t->rng.seed(1);
  t->computeSBI = false; t->haveSBI = false; t->nFrame = 0; memset(&t->sbiCam, 0, sizeof(t->sbiCam));
  t->relocBest = -1; t->relocScore = 0; t->nRecoveries = 0;
  return t;
}
void orc_tracker_destroy(void* t) { delete (OTracker*)t; }
void orc_tracker_seed(void* t, unsigned s) { ((OTracker*)t)->rng.seed(s); }
void orc_tracker_set_truncate(void* t, int on) { ((OTracker*)t)->truncateError = on != 0; }
void orc_tracker_set_map(void* t_, void* srckf, int n, const double* world, const double* right, const double* down, const int32_t* irCenter, const int32_t* srcLevel) {
  OTracker* t = (OTracker*)t_;
  t->srcKF = (OKeyFrame*)srckf; t->pts.resize(n); t->td.clear(); t->td.resize(n); t->hasTD.assign(n, 0);
  for (int i = 0; i < n; i++) {
    MapPointO& p = t->pts[i];
    for (int k = 0; k < 3; k++) { p.world[k] = world[3 * i + k]; p.right[k] = right[3 * i + k]; p.down[k] = down[3 * i + k]; }
    p.irCenter[0] = irCenter[2 * i]; p.irCenter[1] = irCenter[2 * i + 1]; p.srcLevel = srcLevel[i]; p.outlierCount = p.inlierCount = 0;
    p.srcKF = t->srcKF;
  }
}
// Map::vpPoints.push_back of new points while the tracker runs (MapMaker::AddPointEpipolar, jni/MapMaker.cc:685): existing points keep
// their TrackerData, the new ones get theirs on first use (jni/Tracker.cc:372)
void orc_tracker_append_points(void* t_, void* srckf, int n, const double* world, const double* right, const double* down, const int32_t* irCenter, const int32_t* srcLevel) {
  OTracker* t = (OTracker*)t_;
  const size_t base = t->pts.size();
  t->pts.resize(base + n); t->td.resize(base + n); t->hasTD.resize(base + n, 0);
  for (int i = 0; i < n; i++) {
    MapPointO& p = t->pts[base + i];
    for (int k = 0; k < 3; k++) { p.world[k] = world[3 * i + k]; p.right[k] = right[3 * i + k]; p.down[k] = down[3 * i + k]; }
    p.irCenter[0] = irCenter[2 * i]; p.irCenter[1] = irCenter[2 * i + 1]; p.srcLevel = srcLevel[i]; p.outlierCount = p.inlierCount = 0;
    p.srcKF = (OKeyFrame*)srckf;
  }
}
// a map point whose patch comes from another keyframe of the map (MapPoint::pPatchSourceKF)
void orc_tracker_set_point_source_kf(void* t_, int i, void* kf) { ((OTracker*)t_)->pts[i].srcKF = (OKeyFrame*)kf; }
void orc_tracker_set_pose(void* t, const double* p12) { ((OTracker*)t)->pose = se3_from12(p12); }
void orc_tracker_get_pose(void* t, double* p12) { se3_to12(((OTracker*)t)->pose, p12); }
void orc_tracker_set_velocity(void* t_, const double* v6, double msd) { OTracker* t = (OTracker*)t_; memcpy(t->velocity, v6, sizeof(t->velocity)); t->msdScaledVel = msd; }
void orc_tracker_get_velocity(void* t_, double* v6, double* msd) { OTracker* t = (OTracker*)t_; memcpy(v6, t->velocity, sizeof(t->velocity)); *msd = t->msdScaledVel; }
void orc_tracker_set_scene_depth(void* t_, double m, double s) { OTracker* t = (OTracker*)t_; t->sceneDepthMean = m; t->sceneDepthSigma = s; }
void orc_tracker_get_scene_depth(void* t_, double* m, double* s) { OTracker* t = (OTracker*)t_; *m = t->sceneDepthMean; *s = t->sceneDepthSigma; }
void orc_tracker_set_sbi_rot(void* t_, const double* v6, int use) { OTracker* t = (OTracker*)t_; memcpy(t->sbiRot, v6, sizeof(t->sbiRot)); t->useSBI = use != 0; }
void* orc_tracker_current_kf(void* t) { return &((OTracker*)t)->cur; }
void orc_tracker_make_current_kf(void* t, const uint8_t* gray, int w, int h, int stride) { make_keyframe_lite(((OTracker*)t)->cur, gray, w, h, stride); }
void orc_tracker_track_map(void* t) { track_map(*(OTracker*)t); }
void orc_tracker_motion_model(void* t, int apply_not_update) { if (apply_not_update) apply_motion_model(*(OTracker*)t); else update_motion_model(*(OTracker*)t); }
void orc_tracker_assess_quality(void* t) { assess_tracking_quality(*(OTracker*)t); }
// Tracker::TrackFrame for a good map (jni/Tracker.cc:76-112) with the SBI rotation supplied by the caller
// (SmallBlurryImage is the f1 "next" row of SURVEY.md §8): lost (>= 3 bad frames) streams are left alone.
void orc_tracker_track_frame(void* t_, const uint8_t* gray, int w, int h, int stride) {
  OTracker* t = (OTracker*)t_;
  make_keyframe_lite(t->cur, gray, w, h, stride);
  if (t->computeSBI) {   // jni/Tracker.cc:86-97: rotate the two SmallBlurryImages; on the first frame both come from the same keyframe
    if (!t->haveSBI) { sbi_make(t->sbiThis, t->cur, 0.75); t->sbiLast = t->sbiThis; t->haveSBI = true; }
    else { t->sbiLast = t->sbiThis; sbi_make(t->sbiThis, t->cur, 0.75); }
  }
  t->nFrame++;
  if (t->lostFrames < 3) {
    if (t->computeSBI && t->useSBI) {   // Tracker::CalcSBIRotation (jni/Tracker.cc:885-893)
      sbi_make_jacs(t->sbiLast);
      const SE2 se2 = sbi_iterate(t->sbiThis, t->sbiLast, 6, 0);
      se3_from_se2(se2, t->sbiCam, t->sbiThis.w, t->sbiThis.h, t->sbiRot);
    }
    apply_motion_model(*t); track_map(*t); update_motion_model(*t); assess_tracking_quality(*t);
    // jni/Tracker.cc:127-132: heuristics to add a keyframe (MapMaker::NeedNewKeyFrame, jni/MapMaker.cc:763-773; the queue is always
    // empty here: the keyframe joins the map at once)
    t->kfAddedThisFrame = 0;
    if (t->kfPolicy && !t->relocPose.empty() && t->quality == 2) {
      double dDist = dist_to_nearest_keyframe(*t, 0);
      dDist *= (1.0 / t->sceneDepthMean);
      if (dDist > t->kfMult * t->kfWiggleDN && t->nFrame - t->lastKeyFrameDropped > t->kfMinFrames) {
        // Tracker::AddNewKeyFrame (jni/Tracker.cc:823-827) -> MapMaker::AddKeyFrame copies mCurrentKF
        SBI sb; sbi_make(sb, t->cur, 2.5); sbi_make_jacs(sb);
        t->relocSBI.push_back(sb); t->relocPose.push_back(t->pose);
        t->lastKeyFrameDropped = t->nFrame; t->kfAddedThisFrame = 1;
      }
    }
  } else if (!t->relocSBI.empty()) {   // jni/Tracker.cc:134-140: tracking lost -> relocalise against the map keyframes
    if (attempt_recovery(*t)) { track_map(*t); assess_tracking_quality(*t); }
  }
}
// Relocaliser keyframe: SmallBlurryImage(kf) with the default blur 2.5 (KeyFrame::MakeKeyFrame_Rest, jni/KeyFrame.cc:98) + MakeJacs
void orc_tracker_add_reloc_keyframe(void* t_, void* kf, const double* pose12) {
  OTracker* t = (OTracker*)t_;
  SBI s; sbi_make(s, *(OKeyFrame*)kf, 2.5); sbi_make_jacs(s);
  t->relocSBI.push_back(s); t->relocPose.push_back(se3_from12(pose12));
}
void orc_tracker_set_keyframe_policy(void* t_, int enable, double wiggle, double wiggle_dn, double mult, int min_frames) {
  OTracker* t = (OTracker*)t_; t->kfPolicy = enable != 0; t->kfWiggle = wiggle; t->kfWiggleDN = wiggle_dn; t->kfMult = mult; t->kfMinFrames = min_frames;
}
void orc_tracker_keyframe_info(void* t_, int* n_keyframes, int* added_this_frame, int* n_frame, int* last_dropped) {
  OTracker* t = (OTracker*)t_; *n_keyframes = (int)t->relocPose.size(); *added_this_frame = t->kfAddedThisFrame; *n_frame = t->nFrame; *last_dropped = t->lastKeyFrameDropped;
}
void orc_tracker_reloc_info(void* t_, int* best, double* score, int* n_recoveries) { OTracker* t = (OTracker*)t_; *best = t->relocBest; *score = t->relocScore; *n_recoveries = t->nRecoveries; }
void orc_tracker_set_lost(void* t_, int lost_frames, int quality) { OTracker* t = (OTracker*)t_; t->lostFrames = lost_frames; t->quality = quality; }
// Turn the on-board SmallBlurryImage path on: cam13 = camera scalars at the SBI image size (level 3 halved)
void orc_tracker_enable_sbi(void* t_, const double* sbi_cam13) { OTracker* t = (OTracker*)t_; t->computeSBI = true; t->useSBI = true; t->sbiCam = cam_from13(sbi_cam13); }
void orc_tracker_get_sbi_rot(void* t_, double* v6) { memcpy(v6, ((OTracker*)t_)->sbiRot, sizeof(double) * 6); }
// stand-alone SBI pieces for the tests
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) { cv_restated::resize_linear_u8(src, sw, sh, (size_t)sw, dst, dw, dh, (size_t)dw); }
void* orc_sbi_create(void* kf, double blur) { SBI* s = new SBI(); sbi_make(*s, *(OKeyFrame*)kf, blur); return s; }
void orc_sbi_destroy(void* s) { delete (SBI*)s; }
void orc_sbi_dims(void* s, int* w, int* h) { *w = ((SBI*)s)->w; *h = ((SBI*)s)->h; }
void orc_sbi_template(void* s_, float* out) { SBI* s = (SBI*)s_; memcpy(out, s->tmpl.data(), s->tmpl.size() * sizeof(float)); }
void orc_sbi_small(void* s_, uint8_t* out) { SBI* s = (SBI*)s_; memcpy(out, s->small.data(), s->small.size()); }
double orc_sbi_rotation(void* this_, void* other_, const double* sbi_cam13, int its, double* se2_3, double* v6) {
  SBI* a = (SBI*)this_; SBI* b = (SBI*)other_;
  sbi_make_jacs(*b);
  double score; const SE2 r = sbi_iterate(*a, *b, its, &score);
  if (se2_3) { se2_3[0] = r.t[0]; se2_3[1] = r.t[1]; se2_3[2] = atan2(r.R[2], r.R[0]); }
  se3_from_se2(r, cam_from13(sbi_cam13), a->w, a->h, v6);
  return score;
}
void orc_tracker_counters(void* t_, int32_t* attempted4, int32_t* found4, int* quality, int* lost, int* did_coarse) {
  OTracker* t = (OTracker*)t_;
  for (int i = 0; i < LEVELS; i++) { attempted4[i] = t->attempted[i]; found4[i] = t->foundCnt[i]; }
  *quality = t->quality; *lost = t->lostFrames; *did_coarse = t->didCoarse;
}
long orc_tracker_zmssd_evals(void* t) { return ((OTracker*)t)->zmssdEvals; }
int orc_tracker_num_updates(void* t) { return (int)((OTracker*)t)->sigmas.size(); }
void orc_tracker_updates(void* t_, double* upd6n, double* sigmas) {
  OTracker* t = (OTracker*)t_;
  memcpy(upd6n, t->updates.data(), t->updates.size() * sizeof(double)); memcpy(sigmas, t->sigmas.data(), t->sigmas.size() * sizeof(double));
}
void orc_tracker_project_all(void* t_) {  // first loop of TrackMap on its own (jni/Tracker.cc:369-392)
  OTracker* t = (OTracker*)t_;
  for (size_t i = 0; i < t->pts.size(); i++) {
    TData& TD = ensure_td(*t, (int)i);
    TD.searchLevel = -1; TD.searched = false; TD.found = false; TD.didSubPix = false;
    td_project(TD, t->pts[i], t->pose, t->cam);
    if (!TD.inImage) continue;
    cam_derivs(t->cam, TD.derivs);
    TD.searchLevel = calc_search_level_and_warp(TD.finder, t->pts[i], t->pose, TD.derivs);
  }
}
// same layout as ref_tracker_point_state (oracle/ref_harness.cc)
void orc_tracker_point_state(void* t_, int i, int32_t* ints, double* dbl) {
  OTracker* t = (OTracker*)t_;
  memset(ints, 0, 8 * sizeof(int32_t)); memset(dbl, 0, 32 * sizeof(double));
  if (!t->hasTD[i]) return;
  const TData& TD = t->td[i];
  ints[0] = TD.inImage; ints[1] = TD.searchLevel; ints[2] = TD.searched; ints[3] = TD.found; ints[4] = TD.didSubPix; ints[5] = TD.finder.templateBad; ints[6] = 1;
  dbl[0] = TD.v2Image[0]; dbl[1] = TD.v2Image[1]; dbl[2] = TD.v2Found[0]; dbl[3] = TD.v2Found[1];
  for (int k = 0; k < 4; k++) dbl[4 + k] = TD.derivs[k];
  for (int k = 0; k < 3; k++) dbl[8 + k] = TD.v3Cam[k];
  for (int k = 0; k < 4; k++) dbl[11 + k] = TD.finder.warpInv[k];
  dbl[15] = TD.sqrtInvNoise;
  for (int k = 0; k < 12; k++) dbl[16 + k] = TD.jac[k];
  dbl[28] = TD.err[0]; dbl[29] = TD.err[1]; dbl[30] = TD.finder.coarsePos[0]; dbl[31] = TD.finder.coarsePos[1];
}
void orc_tracker_point_template(void* t_, int i, uint8_t* tmpl, int* sum, int* sumsq) {
  const Finder& F = ((OTracker*)t_)->td[i].finder; memcpy(tmpl, &F.tmpl[0], F.P * F.P); *sum = F.tsum; *sumsq = F.tsumsq;
}
void orc_tracker_point_counts(void* t_, int i, int* outlier, int* inlier) { const MapPointO& p = ((OTracker*)t_)->pts[i]; *outlier = p.outlierCount; *inlier = p.inlierCount; }
int orc_tracker_search_for_points(void* t_, const int32_t* idx, int n, int range, int subpix) {
  std::vector<int> v(idx, idx + n); return search_for_points(*(OTracker*)t_, v, range, subpix);
}
// MapMaker::ReFind_Common (jni/MapMaker.cc:967-1036) for the listed points in the tracker's current keyframe at the tracker's pose
// (k.se3CfromW), without the Measurement / sNeverRetryKFs bookkeeping.  ONE PatchFinder for the whole list, like the function's
// `static PatchFinder Finder` (so the template reuse rule sees consecutive calls); cold != 0 forgets it between points.
// out3[k] = {found, Finder.GetLevel(), bSubPix}, pos2[k] = m.v2RootPos.
void orc_tracker_refind(void* t_, const int32_t* idx, int n, int range, int subpix_its, int cold, int32_t* out3, double* pos2) {
  OTracker* t = (OTracker*)t_;
  Finder F; F.init(t->P);
  int last = -1;
  for (int k = 0; k < n; k++) {
    const MapPointO& p = t->pts[idx[k]];
    out3[3 * k] = 0; out3[3 * k + 1] = -1; out3[3 * k + 2] = 0; pos2[2 * k] = pos2[2 * k + 1] = 0.0;
    double v3Cam[3]; se3_apply(t->pose, p.world, v3Cam);
    if (v3Cam[2] < 0.001) continue;
    const double ipx = v3Cam[0] / v3Cam[2], ipy = v3Cam[1] / v3Cam[2];
    double d = 0; d += ipx * ipx; d += ipy * ipy;
    if (d > t->cam.largestRadius * t->cam.largestRadius) continue;
    double v2Image[2]; cam_project(t->cam, ipx, ipy, v2Image);
    if (t->cam.invalid) continue;
    if (v2Image[0] < 0 || v2Image[1] < 0 || v2Image[0] > t->cur.lev[0].im.w || v2Image[1] > t->cur.lev[0].im.h) continue;
    double derivs[4]; cam_derivs(t->cam, derivs);
    // Finder.MakeTemplateCoarse(p, k.se3CfromW, m2CamDerivs)
    if (cold || last != idx[k]) F.haveLast = false;          // &p != mpLastTemplateMapPoint
    calc_search_level_and_warp(F, p, t->pose, derivs);       // may set templateBad; MakeTemplateCoarseCont runs regardless
    make_template_coarse_cont(F, p, *p.srcKF);
    last = idx[k];
    out3[3 * k + 1] = F.level;
    if (F.templateBad) continue;
    if (!find_patch_coarse(F, v2Image[0], v2Image[1], t->cur, (unsigned)range, &t->zmssdEvals, 0)) continue;
    out3[3 * k] = 1;
    if (F.level > 0) {
      make_subpix_template(F);
      iterate_subpix_to_convergence(F, t->cur, subpix_its);
      pos2[2 * k] = F.subPixPos[0]; pos2[2 * k + 1] = F.subPixPos[1]; out3[3 * k + 2] = 1;
    } else { pos2[2 * k] = F.coarsePos[0]; pos2[2 * k + 1] = F.coarsePos[1]; }
  }
}
// The search of MapMaker::AddPointEpipolar (jni/MapMaker.cc:525-640) for candidate `cand` (level pixels) of level `level` of kSrc in
// kTarget; the tracker only lends its camera and patch size.  out3 = {converged match, best corner index (-1 none), its ZMSSD};
// pos2 = Finder.GetSubPixPos().  geom8 (may be NULL): v2Normal, v2Along, dNormDist, dMinLen, dMaxLen, dMaxDistSq.
void orc_epipolar_search(void* t_, void* ksrc, void* ktgt, const double* src12, const double* tgt12, double depth_mean, double depth_sigma, double wiggle,
                         int level, int cx_, int cy_, int32_t* out3, double* pos2, double* geom8) {
  OTracker* t = (OTracker*)t_;
  Cam cam = t->cam;
  const OKeyFrame& kSrc = *(OKeyFrame*)ksrc; const OKeyFrame& kTarget = *(OKeyFrame*)ktgt;
  const SE3 S = se3_from12(src12), T = se3_from12(tgt12);
  out3[0] = 0; out3[1] = -1; out3[2] = 0; pos2[0] = pos2[1] = 0.0;
  const int nLevelScale = LevelScale(level);
  const double root[2] = {LevelZeroPos(cx_, level), LevelZeroPos(cy_, level)};
  double un[2]; cam_unproject(cam, root, un);
  double ray[3] = {un[0], un[1], 1.0};
  { double nn = 0; nn += ray[0] * ray[0]; nn += ray[1] * ray[1]; nn += ray[2] * ray[2]; const double n = sqrt(nn); for (int i = 0; i < 3; i++) ray[i] /= n; }
  // v3LineDirn_TC = R_T * (R_S^-1 * ray)
  double St[9]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) St[3 * i + j] = S.R[3 * j + i];
  double tmp[3], dirn[3]; mat3_mul_vec(St, ray, tmp); mat3_mul_vec(T.R, tmp, dirn);
  const double dStart = std::max(wiggle, depth_mean - depth_sigma), dEnd = std::min(40 * wiggle, depth_mean + depth_sigma);
  const SE3 Sinv = se3_inverse(S);
  double center[3]; se3_apply(T, Sinv.t, center);
  double start[3], end[3];
  for (int i = 0; i < 3; i++) { start[i] = center[i] + dStart * dirn[i]; end[i] = center[i] + dEnd * dirn[i]; }
  if (end[2] <= start[2]) return;
  if (end[2] <= 0.0) return;
  if (start[2] <= 0.0) { const double f = 0.001 - start[2] / dirn[2]; for (int i = 0; i < 3; i++) start[i] += dirn[i] * f; }
  const double A[2] = {start[0] / start[2], start[1] / start[2]}, B[2] = {end[0] / end[2], end[1] / end[2]};
  double al[2] = {A[0] - B[0], A[1] - B[1]};
  double aa = 0; aa += al[0] * al[0]; aa += al[1] * al[1];
  if (aa < 0.00000001) return;
  { const double n = sqrt(aa); al[0] /= n; al[1] /= n; }
  const double nr[2] = {al[1], -al[0]};
  double dNormDist = 0; dNormDist += A[0] * nr[0]; dNormDist += A[1] * nr[1];
  if (fabs(dNormDist) > cam.largestRadius) return;
  double aA = 0; aA += al[0] * A[0]; aA += al[1] * A[1];
  double aB = 0; aB += al[0] * B[0]; aB += al[1] * B[1];
  double dMinLen = std::min(aA, aB) - 0.05, dMaxLen = std::max(aA, aB) + 0.05;
  if (dMinLen < -2.0) dMinLen = -2.0;
  if (dMaxLen < -2.0) dMaxLen = -2.0;
  if (dMinLen > 2.0) dMinLen = 2.0;
  if (dMaxLen > 2.0) dMaxLen = 2.0;
  // mdOnePixelDist (jni/ATANCamera.cc:86-91)
  double uc[2], ur[2]; const double c0[2] = {cam.width / 2, cam.height / 2}, c1[2] = {cam.width / 2 + 1.0, cam.height / 2 + 1.0};
  cam_unproject(cam, c0, uc); cam_unproject(cam, c1, ur);
  double dd = 0; dd += (uc[0] - ur[0]) * (uc[0] - ur[0]); dd += (uc[1] - ur[1]) * (uc[1] - ur[1]);
  const double onePixelDist = sqrt(dd) / sqrt(2.0);
  const double dMaxDistDiff = onePixelDist * (4.0 + 1.0 * nLevelScale), dMaxDistSq = dMaxDistDiff * dMaxDistDiff;
  if (geom8) { geom8[0] = nr[0]; geom8[1] = nr[1]; geom8[2] = al[0]; geom8[3] = al[1]; geom8[4] = dNormDist; geom8[5] = dMinLen; geom8[6] = dMaxLen; geom8[7] = dMaxDistSq; }
  // Finder.MakeTemplateCoarseNoWarp(kSrc, nLevel, a, b) (jni/PatchFinder.cc:130-143)
  Finder F; F.init(t->P); F.level = level;
  const Image& sim = kSrc.lev[level].im;
  if (!in_image_with_border(sim, cx_, cy_, F.P / 2 + 1)) return;
  for (int r = 0; r < F.P; r++) memcpy(&F.tmpl[r * F.P], sim.row(cy_ - F.P / 2 + r) + (cx_ - F.P / 2), F.P);
  make_template_sums(F);
  const OLevel& TL = kTarget.lev[level];
  const int W0 = kTarget.lev[0].im.w;
  int nBest = -1, nBestZMSSD = F.maxSSD + 1;
  for (size_t i = 0; i < TL.corners.size(); i++) {
    // vv2Corners[i] = imUnProj(zpos(1), zpos(0)): the table holds UnProject of integer pixels and is indexed with the truncated position
    const int zx = (int)LevelZeroPos(TL.corners[i].x, level), zy = (int)LevelZeroPos(TL.corners[i].y, level);
    (void)W0;
    const double px[2] = {(double)zx, (double)zy}; double v2Im[2]; cam_unproject(cam, px, v2Im);
    double dn = 0; dn += v2Im[0] * nr[0]; dn += v2Im[1] * nr[1];
    const double dDistDiff = dNormDist - dn;
    if (dDistDiff * dDistDiff > dMaxDistSq) continue;
    double da = 0; da += v2Im[0] * al[0]; da += v2Im[1] * al[1];
    if (da < dMinLen) continue;
    if (da > dMaxLen) continue;
    const int z = zmssd_at_point(F, TL.im, TL.corners[i].x, TL.corners[i].y);
    if (z < nBestZMSSD) { nBest = (int)i; nBestZMSSD = z; }
  }
  out3[1] = nBest; out3[2] = nBestZMSSD;
  if (nBest == -1) return;
  make_subpix_template(F);
  F.subPixPos[0] = LevelZeroPos(TL.corners[nBest].x, level); F.subPixPos[1] = LevelZeroPos(TL.corners[nBest].y, level);
  const bool ok = iterate_subpix_to_convergence(F, kTarget, 10);
  pos2[0] = F.subPixPos[0]; pos2[1] = F.subPixPos[1];
  out3[0] = ok ? 1 : 0;
}
void orc_tracker_clear_counters(void* t_) { OTracker* t = (OTracker*)t_; for (int i = 0; i < LEVELS; i++) t->attempted[i] = t->foundCnt[i] = 0; }
void orc_tracker_calc_jacobians(void* t_, const int32_t* idx, int n) { OTracker* t = (OTracker*)t_; for (int i = 0; i < n; i++) if (t->td[idx[i]].found) td_calc_jacobian(t->td[idx[i]]); }
void orc_tracker_project_and_derivs(void* t_, const int32_t* idx, int n, int only_found) {
  OTracker* t = (OTracker*)t_;
  for (int i = 0; i < n; i++) if (!only_found || t->td[idx[i]].found) td_project_and_derivs(t->td[idx[i]], t->pts[idx[i]], t->pose, t->cam);
}
void orc_tracker_linear_update(void* t_, const int32_t* idx, int n, const double* v6) { OTracker* t = (OTracker*)t_; for (int i = 0; i < n; i++) if (t->td[idx[i]].found) td_linear_update(t->td[idx[i]], v6); }
void orc_tracker_calc_pose_update(void* t_, const int32_t* idx, int n, double sigma, int mark, int apply, double* out6) {
  OTracker* t = (OTracker*)t_; std::vector<int> v(idx, idx + n);
  calc_pose_update(*t, v, sigma, mark != 0, out6, 0);
  if (apply) t->pose = se3_mul(se3_exp(out6), t->pose);
}
double orc_tukey_sigma_squared(const double* e, int n) { std::vector<double> v(e, e + n); return tukey_sigma_squared(v); }

}  // extern "C"

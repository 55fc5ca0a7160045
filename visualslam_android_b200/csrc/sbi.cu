// SmallBlurryImage on the GPU (reference: jni/SmallBlurryImage.cc; the f1 row of SURVEY.md §8).
// One CTA per stream, once per frame, between the pyramid kernels and the projection kernel:
//   MakeFromKF   level 3 halved ((a+b+c+d+2)>>2), mean removed in float, 9x9 float Gaussian (sigma 0.75, BORDER_REPLICATE)
//   MakeJacs     central differences of the previous frame's template
//   IteratePosRelToTarget   6 ESM iterations aligning this frame's template to the previous one (SE2 + mean offset)
//   SE3fromSE2   3 Gauss-Newton steps for the camera rotation that reproduces the SE2; ln() of it seeds Tracker::ApplyMotionModel
// Deviations from the serial reference, both below 1e-12 relative on the result: the warp positions are evaluated in closed
// form (p0 + i*down + j*across) instead of by running sums, and the 15 ESM sums are reduced in a fixed tree, not serially.
#include "geometry.cuh"
#include <cstdio>

namespace {

constexpr int kT = 256;

// cv::resize(level 3 -> SmallBlurryImage size), INTER_LINEAR on u8 (OpenCV's fixed-point bilinear; the third-party algorithm is restated
// and pinned against cv2 in oracle/shim/cv_resize_linear_u8.h): exactly half in both directions = OpenCV's fast area path (a+b+c+d+2)>>2,
// otherwise per output column / row a source index and two 11-bit weights, built on the host by vslam_enable_sbi
struct ResizeTab { int exact_half; int l3w; const int* tab; };   // tab: xofs[w], alpha[2 w], yofs[h], beta[2 h]

struct SbiDev {
  ResizeTab rs;
  int w, h;                 // SBI size = level 3 / 2
  int l3w, l3h, l3pitch; const uint8_t* l3;   // [S][l3h][l3pitch]
  float taps[9];            // getGaussianKernel(9, 0.75) in float, computed on the host
  CamDev cam;               // camera scalars at the SBI image size
  double orig[2][3];        // un-projected (w/2 +- 5, h/2) points (host: tan)
  float* tmpl;              // [S][2][n]  this / last template (ping-pong by frame parity)
  float* scratch;           // [S][3][n]  tmp, warped, (spare)
  float* jac;               // [S][2n]    gradient image of the last template
  uint8_t* small;           // [S][n]
  StreamState* ss; int* have; int* parity;   // per stream
  double* rot_out;          // [S][6] Tracker::mv6SBIRot of this frame (the frame set's slot of ctx->sbi_rot_buf); k_project_lists hands it to StreamState::sbi_rot
  float* reloc_scratch; uint8_t* reloc_small;   // k_relocalise's own scratch [S][3n] / [S][n]
  int use_sbi, s0;
  int smem_floats;          // k_sbi: 3n when this frame's template and the two scratch images fit in dynamic shared memory, else 0 (they stay in global memory)
};

struct RelocDev {
  ResizeTab rs;
  int n_kf, w, h, l3h, l3pitch;
  const uint8_t* src_l3;    // level 3 of the source keyframes [K_src][l3h][l3pitch]
  float taps17[17];         // getGaussianKernel(17, 2.5) in float
  float* kf_tmpl; float* kf_jac; float* kf_tmp; uint8_t* kf_small;   // [n_kf][n], [n_kf][2n], scratch
  const double* kf_pose;    // [n_kf][12]
  double* scores;           // [S][n_kf]
};

// i / W for 0 <= i < W * H (a few thousand) with the 32-bit reciprocal rw = 0xffffffff / W + 1: one IMAD.HI instead of an integer division
__device__ __forceinline__ int div_w(int i, uint32_t rw) { return (int)__umulhi((uint32_t)i, rw); }

__device__ inline double block_sum(double v, double* red) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0;
  for (int k = 0; k < kT / 32; k++) r += red[k];
  return r;
}

// 4x4 inverse by partial-pivot LU (same routine as the oracle's inverse_lu), then inv * b.  Every loop is unrolled and the row swap is a chain
// of selects, so that the matrix stays in registers (a run-time row index would put it in local memory); the arithmetic and its order are those
// of the plain loops.
__device__ inline void solve4(const double* m, const double* b, double* x4) {
  double a[16], inv[16]; int piv[4];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = m[i];
#pragma unroll
  for (int i = 0; i < 4; i++) piv[i] = i;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    int p = k; double best = fabs(a[k * 4 + k]);
#pragma unroll
    for (int i = k + 1; i < 4; i++) if (fabs(a[i * 4 + k]) > best) { best = fabs(a[i * 4 + k]); p = i; }
#pragma unroll
    for (int i = k + 1; i < 4; i++) {
      const bool sw = p == i;
#pragma unroll
      for (int j = 0; j < 4; j++) { const double u = a[k * 4 + j], v = a[i * 4 + j]; a[k * 4 + j] = sw ? v : u; a[i * 4 + j] = sw ? u : v; }
      const int u = piv[k], v = piv[i]; piv[k] = sw ? v : u; piv[i] = sw ? u : v;
    }
#pragma unroll
    for (int i = k + 1; i < 4; i++) {
      a[i * 4 + k] /= a[k * 4 + k];
#pragma unroll
      for (int j = k + 1; j < 4; j++) a[i * 4 + j] -= a[i * 4 + k] * a[k * 4 + j];
    }
  }
#pragma unroll
  for (int c = 0; c < 4; c++) {
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = (piv[i] == c) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < i; j++) x[i] -= a[i * 4 + j] * x[j];
#pragma unroll
    for (int i = 3; i >= 0; i--) {
#pragma unroll
      for (int j = i + 1; j < 4; j++) x[i] -= a[i * 4 + j] * x[j];
      x[i] /= a[i * 4 + i];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) inv[i * 4 + c] = x[i];
  }
#pragma unroll
  for (int i = 0; i < 4; i++) { double sacc = inv[4 * i] * b[0]; for (int k = 1; k < 4; k++) sacc += inv[4 * i + k] * b[k]; x4[i] = sacc; }
}

__device__ inline void inverse3(const double* m, double* r) {
  const double c00 = m[4] * m[8] - m[5] * m[7], c10 = m[5] * m[6] - m[3] * m[8], c20 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c00 + m[1] * c10 + m[2] * c20, invdet = 1.0 / det;
  r[0] = c00 * invdet; r[3] = c10 * invdet; r[6] = c20 * invdet;
  r[1] = (m[2] * m[7] - m[1] * m[8]) * invdet; r[4] = (m[0] * m[8] - m[2] * m[6]) * invdet; r[7] = (m[1] * m[6] - m[0] * m[7]) * invdet;
  r[2] = (m[1] * m[5] - m[2] * m[4]) * invdet; r[5] = (m[2] * m[3] - m[0] * m[5]) * invdet; r[8] = (m[0] * m[4] - m[1] * m[3]) * invdet;
}

// ---- building blocks (called by every thread of a kT-thread CTA) -----------------------------------------------------------------
#ifdef VS_SBI_TIMING   // instrumented build (scratch experiments): cycles per phase of one CTA, printed by stream 7
#define SBI_MARK(k) do { __syncthreads(); if (threadIdx.x == 0) { const long long t_ = clock64(); sh.tacc[k] += t_ - sh.tlast; sh.tlast = t_; } } while (0)
#else
#define SBI_MARK(k) do { } while (0)
#endif

struct SbiShared {
#ifdef VS_SBI_TIMING
  long long tacc[12], tlast;
#endif
  double red[kT / 32];
  double red15[kT / 32][15];
  double X[6];        // current warp: R (4, row-major) and t (2)
  double CtoC[6];     // accumulated SE2
  double mean, score;
  int best;
};

// SmallBlurryImage::MakeFromKF (jni/SmallBlurryImage.cc:20-55): level 3 halved, mean removed, separable float Gaussian (ntaps = 9 for
// dBlur <= 2, else 17), BORDER_REPLICATE.  `out` receives the template; `small` the u8 thumbnail; `tmp` is scratch.
__device__ void sbi_make(const uint8_t* __restrict__ l3, int l3pitch, int l3h, const ResizeTab& rs, int W, int H, const float* taps, int ntaps, uint8_t* small, float* tmp, float* out, SbiShared& sh) {
  const int tid = threadIdx.x, n = W * H, half = ntaps / 2;
  const uint32_t rw = 0xffffffffu / (uint32_t)W + 1u;      // W >= 2
  double isum = 0;
  if (rs.exact_half) {
    for (int i = tid; i < n; i += kT) {
      const int y = div_w(i, rw), x = i - y * W;
      const uint8_t* a = l3 + (size_t)(2 * y) * l3pitch + 2 * x;
      const int v = (a[0] + a[1] + a[l3pitch] + a[l3pitch + 1] + 2) >> 2;
      small[i] = (uint8_t)v; isum += v;
    }
  } else {
    const int* xofs = rs.tab; const int* alpha = xofs + W; const int* yofs = alpha + 2 * W; const int* beta = yofs + H;
    for (int i = tid; i < n; i += kT) {
      const int y = div_w(i, rw), x = i - y * W;
      const int sy = yofs[y], sy0 = sy < 0 ? 0 : (sy < l3h ? sy : l3h - 1), sy1 = sy + 1 < 0 ? 0 : (sy + 1 < l3h ? sy + 1 : l3h - 1);
      const int sx = xofs[x], sx1 = sx + 1 < rs.l3w ? sx + 1 : rs.l3w - 1, a0 = alpha[2 * x], a1 = alpha[2 * x + 1], b0 = beta[2 * y], b1 = beta[2 * y + 1];
      const uint8_t* S0 = l3 + (size_t)sy0 * l3pitch; const uint8_t* S1 = l3 + (size_t)sy1 * l3pitch;
      const int r0 = S0[sx] * a0 + S0[sx1] * a1, r1 = S1[sx] * a0 + S1[sx1] * a1;
      const int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
      small[i] = (uint8_t)v; isum += v;
    }
  }
  isum = block_sum(isum, sh.red);            // integer valued: exact
  const float fMean = ((float)(unsigned)isum) / (H * W);
  for (int i = tid; i < n; i += kT) out[i] = (float)small[i] - fMean;
  __syncthreads();
  for (int i = tid; i < n; i += kT) {    // row pass
    const int y = div_w(i, rw), x = i - y * W; float acc = 0;
    for (int k = 0; k < ntaps; k++) { int xx = x + k - half; xx = xx < 0 ? 0 : (xx >= W ? W - 1 : xx); acc += taps[k] * out[y * W + xx]; }
    tmp[i] = acc;
  }
  __syncthreads();
  for (int i = tid; i < n; i += kT) {    // column pass
    const int y = div_w(i, rw), x = i - y * W; float acc = 0;
    for (int k = 0; k < ntaps; k++) { int yy = y + k - half; yy = yy < 0 ? 0 : (yy >= H ? H - 1 : yy); acc += taps[k] * tmp[yy * W + x]; }
    out[i] = acc;
  }
  __syncthreads();
}

// SmallBlurryImage::MakeJacs (jni/SmallBlurryImage.cc:58-79)
__device__ void sbi_make_jacs(const float* __restrict__ t, float* jac, int W, int H) {
  const int n = W * H;
  const uint32_t rw = 0xffffffffu / (uint32_t)W + 1u;
  for (int i = threadIdx.x; i < n; i += kT) {
    const int y = div_w(i, rw), x = i - y * W;
    float gx = 0.f, gy = 0.f;
    if (x >= 1 && y >= 1 && x < W - 1 && y < H - 1) { gx = t[i + 1] - t[i - 1]; gy = t[i + W] - t[i - W]; }
    jac[2 * i] = gx; jac[2 * i + 1] = gy;
  }
}

// X = WfromC * CtoC * WfromC^-1 with WfromC = (I, centre): the image-space transform of the current SE2 estimate (one thread)
__device__ __forceinline__ void sbi_esm_transform(SbiShared& sh, double cx, double cy) {
  const double* R = sh.CtoC; const double t0 = sh.CtoC[4], t1 = sh.CtoC[5];
  // A = WfromC * CtoC : rotation R (I*R evaluated like the reference: sums with exact zeros), translation c + (1*t0 + 0*t1, ...)
  double AR[4]; for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) { double sacc = (i == 0 ? 1.0 : 0.0) * R[j]; sacc += (i == 1 ? 1.0 : 0.0) * R[2 + j]; AR[2 * i + j] = sacc; }
  double At[2]; { double a = 1.0 * t0; a += 0.0 * t1; At[0] = cx + a; double b = 0.0 * t0; b += 1.0 * t1; At[1] = cy + b; }
  // inverse of WfromC: rotation I, translation -(I * c)
  double it0, it1; { double a = 1.0 * cx; a += 0.0 * cy; it0 = -a; double b = 0.0 * cx; b += 1.0 * cy; it1 = -b; }
  for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) { double sacc = AR[2 * i] * (j == 0 ? 1.0 : 0.0); sacc += AR[2 * i + 1] * (j == 1 ? 1.0 : 0.0); sh.X[2 * i + j] = sacc; }
  for (int i = 0; i < 2; i++) { double sacc = AR[2 * i] * it0; sacc += AR[2 * i + 1] * it1; sh.X[4 + i] = At[i] + sacc; }
}

// SmallBlurryImage::IteratePosRelToTarget (jni/SmallBlurryImage.cc:99-222): `its` ESM iterations aligning `cur` to the target `last`
// (gradient image `jac`).  Leaves the SE2 in sh.CtoC and the final score (sum of squared differences of the last iteration) in sh.score.
__device__ void sbi_esm(const float* __restrict__ cur, const float* __restrict__ last, const float* __restrict__ jac, float* warped, int W, int H, int its, SbiShared& sh) {
  const int tid = threadIdx.x, n = W * H;
  const uint32_t rw = 0xffffffffu / (uint32_t)W + 1u;
  const double cx = W / 2.0, cy = H / 2.0;
  __syncthreads();
  if (tid == 0) { sh.CtoC[0] = 1; sh.CtoC[1] = 0; sh.CtoC[2] = 0; sh.CtoC[3] = 1; sh.CtoC[4] = 0; sh.CtoC[5] = 0; sh.mean = 0.0; sh.score = 0.0; sbi_esm_transform(sh, cx, cy); }
  __syncthreads();
  for (int it = 0; it < its; it++) {
    {   // transform_image (float): out(i,j) = bilinear(cur, p0 + i*down + j*across), default -9e20f outside
      const double a0 = sh.X[0], a1 = sh.X[2], d0 = sh.X[1], d1 = sh.X[3];
      const double p00 = sh.X[4], p01 = sh.X[5];   // outOrig = 0  =>  p0 = inOrig
      const float xb = W - 1, yb = H - 1;
      for (int i = tid; i < n; i += kT) {
        const int r = div_w(i, rw), c = i - r * W;
        double x = p00 + r * d0 + c * a0, y = p01 + r * d1 + c * a1;
        float v = -9e20f;
        if (0 <= x && 0 <= y && x < xb && y < yb) {
          const int lx = (int)x, ly = (int)y;
          x -= lx; y -= ly;
          const float* r0 = cur + ly * W + lx; const float* r1 = r0 + W;
          v = (float)((1 - y) * ((1 - x) * r0[0] + x * r0[1]) + y * ((1 - x) * r1[0] + x * r1[1]));
        }
        warped[i] = v;
      }
    }
    __syncthreads();
    SBI_MARK(3);
    double acc[15];
#pragma unroll
    for (int k = 0; k < 15; k++) acc[k] = 0;
    const double mean = sh.mean;
    for (int i = tid; i < n; i += kT) {
      const int y = div_w(i, rw), x = i - y * W;
      if (!(x >= 1 && y >= 1 && x < W - 1 && y < H - 1)) continue;
      const float l = warped[i - 1], r = warped[i + 1], u = warped[i - W], d = warped[i + W], here = warped[i];
      if (l + r + u + d + here < -9999.9) continue;
      const double g0 = r - l, g1 = d - u;
      const double s0 = 0.25 * (g0 + jac[2 * i]), s1 = 0.25 * (g1 + jac[2 * i + 1]);
      const double J0 = s0, J1 = s1, J2 = -((double)y - cy) * s0 + ((double)x - cx) * s1;
      const double dd = here - last[i] + mean;
      acc[0] += dd * J0; acc[1] += dd * J1; acc[2] += dd * J2; acc[3] += dd * 1.0;
      acc[4] += J0 * J0; acc[5] += J1 * J0; acc[6] += J1 * J1; acc[7] += J2 * J0; acc[8] += J2 * J1; acc[9] += J2 * J2;
      acc[10] += J0; acc[11] += J1; acc[12] += J2; acc[13] += 1.0; acc[14] += dd * dd;
    }
    // one fixed-shape reduction for all 15 sums: a transposing butterfly (each step halves what a lane holds: 8 + 4 + 2 + 1 + 1 = 16 shuffles of
    // doubles instead of 15 x 5; every sum still pairs lanes at distance 16, 8, 4, 2, 1 in that order, so the values are those of the plain
    // butterfly), then 8 partials per sum through shared memory
    {
      const int lane = tid & 31;
      const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
      double v8[8], v4[4], v2[2];
#pragma unroll
      for (int i = 0; i < 8; i++) { const double hi = (8 + i < 15) ? acc[(8 + i < 15) ? 8 + i : 0] : 0.0; const double keep = b4 ? hi : acc[i], give = b4 ? acc[i] : hi; v8[i] = keep + __shfl_xor_sync(0xffffffffu, give, 16); }
#pragma unroll
      for (int i = 0; i < 4; i++) { const double keep = b3 ? v8[4 + i] : v8[i], give = b3 ? v8[i] : v8[4 + i]; v4[i] = keep + __shfl_xor_sync(0xffffffffu, give, 8); }
#pragma unroll
      for (int i = 0; i < 2; i++) { const double keep = b2 ? v4[2 + i] : v4[i], give = b2 ? v4[i] : v4[2 + i]; v2[i] = keep + __shfl_xor_sync(0xffffffffu, give, 4); }
      const double keep1 = b1 ? v2[1] : v2[0], give1 = b1 ? v2[0] : v2[1];
      const double v1 = keep1 + __shfl_xor_sync(0xffffffffu, give1, 2);
      const double r = v1 + __shfl_xor_sync(0xffffffffu, v1, 1);         // both lanes of a pair end with the sum
      const int q = (b4 ? 8 : 0) + (b3 ? 4 : 0) + (b2 ? 2 : 0) + (b1 ? 1 : 0);   // the sum this lane pair holds (15 = the padding)
      __syncthreads();
      if (!b0 && q < 15) sh.red15[tid >> 5][q] = r;
    }
    __syncthreads();
    if (tid < 15) {   // lane k adds the eight partials of sum k (same order as a serial loop), lane 0 collects them
      double r = 0; for (int w = 0; w < kT / 32; w++) r += sh.red15[w][tid];
      sh.red15[0][tid] = r;
    }
    if (tid < 32) __syncwarp();
    SBI_MARK(4);
    if (tid == 0) {
      for (int k = 0; k < 15; k++) acc[k] = sh.red15[0][k];
      sh.score = acc[14];
      double m4[16]; int v = 0;
      for (int j = 0; j < 4; j++) for (int i = 0; i <= j; i++) { m4[4 * j + i] = m4[4 * i + j] = acc[4 + v]; v++; }
      double upd[4]; solve4(m4, acc, upd);
      const double ang = -upd[2];
      const double c = cos(ang), sn = sin(ang);
      const double U[6] = {c, -sn, sn, c, -upd[0], -upd[1]};   // mySO2::exp (jni/RT.h:461-467), translation -update
      double R[4], t[2];
      for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) { double sacc = sh.CtoC[2 * i] * U[j]; sacc += sh.CtoC[2 * i + 1] * U[2 + j]; R[2 * i + j] = sacc; }
      for (int i = 0; i < 2; i++) { double sacc = sh.CtoC[2 * i] * U[4]; sacc += sh.CtoC[2 * i + 1] * U[5]; t[i] = sh.CtoC[4 + i] + sacc; }
      for (int k = 0; k < 4; k++) sh.CtoC[k] = R[k];
      sh.CtoC[4] = t[0]; sh.CtoC[5] = t[1];
      sh.mean -= upd[3];
      if (it + 1 < its) sbi_esm_transform(sh, cx, cy);   // for the next iteration, in the same single-thread section (one barrier less per iteration)
    }
    __syncthreads();
    SBI_MARK(5);
  }
}

// SmallBlurryImage::SE3fromSE2 (jni/SmallBlurryImage.cc:245-333), one thread: the rotation (row-major 3x3) that reproduces the SE2
__device__ void sbi_se3_from_se2(const double* CtoC, const CamDev& cam, const double (*orig)[3], int W, int H, double* R) {
  const double c2[2] = {W / 2.0, H / 2.0};
  const double off[2][2] = {{5, 0}, {-5, 0}};
  double turned[2][2];
  for (int k = 0; k < 2; k++) for (int i = 0; i < 2; i++) { double sacc = CtoC[2 * i] * off[k][0]; sacc += CtoC[2 * i + 1] * off[k][1]; turned[k][i] = c2[i] + (CtoC[4 + i] + sacc); }
  for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int it = 0; it < 3; it++) {
    double C[9] = {10, 0, 0, 0, 10, 0, 0, 0, 10}, b[3] = {0, 0, 0};
    for (int k = 0; k < 2; k++) {
      double v3[3];
      for (int i = 0; i < 3; i++) { double sacc = R[3 * i] * orig[k][0]; sacc += R[3 * i + 1] * orig[k][1]; sacc += R[3 * i + 2] * orig[k][2]; v3[i] = sacc; }
      double pix[2]; CamCache cc; cam_project(cam, v3[0] / v3[2], v3[1] / v3[2], pix, cc);
      const double err[2] = {turned[k][0] - pix[0], turned[k][1] - pix[1]};
      double dv[4]; cam_derivs(cam, cc, dv);
      double J[2][3];
      const double invz = 1.0 / v3[2];
      for (int m = 0; m < 3; m++) {
        double mo[3]; mo[m] = 0; mo[(m + 1) % 3] = -v3[(m + 2) % 3]; mo[(m + 2) % 3] = v3[(m + 1) % 3];
        const double c0 = (mo[0] - v3[0] * mo[2] * invz) * invz, c1 = (mo[1] - v3[1] * mo[2] * invz) * invz;
        double a0 = dv[0] * c0; a0 += dv[1] * c1; double a1 = dv[2] * c0; a1 += dv[3] * c1;
        J[0][m] = a0; J[1][m] = a1;
      }
      for (int row = 0; row < 2; row++)
        for (int r = 0; r < 3; r++) { const double Jw = 1.0 * J[row][r]; b[r] += err[row] * Jw; for (int q = r; q < 3; q++) C[3 * r + q] += Jw * J[row][q]; }
    }
    for (int r = 1; r < 3; r++) for (int q = 0; q < r; q++) C[3 * r + q] = C[3 * q + r];
    double Ci[9]; inverse3(C, Ci);
    double mu[3]; for (int i = 0; i < 3; i++) { double sacc = Ci[3 * i] * b[0]; sacc += Ci[3 * i + 1] * b[1]; sacc += Ci[3 * i + 2] * b[2]; mu[i] = sacc; }
    double E[9], Rn[9]; so3_exp(mu, E);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double sacc = E[3 * i] * R[j]; sacc += E[3 * i + 1] * R[3 + j]; sacc += E[3 * i + 2] * R[6 + j]; Rn[3 * i + j] = sacc; }
    for (int i = 0; i < 9; i++) R[i] = Rn[i];
  }
}

// Per frame and stream: the tracker's SmallBlurryImage pair (blur 0.75) and Tracker::CalcSBIRotation (jni/Tracker.cc:86-97,885-893)
__global__ void __launch_bounds__(kT) k_sbi(SbiDev D) {
  cudaGridDependencySynchronize(); cudaTriggerProgrammaticLaunchCompletion();   // programmatic dependent launch (vs_launch_pdl); no-ops otherwise
  __shared__ SbiShared sh;
  const int s = blockIdx.x + D.s0, tid = threadIdx.x;
  const int W = D.w, H = D.h, n = W * H;
  const int par = D.parity[s];
  float* cur_g = D.tmpl + ((size_t)s * 2 + par) * n;
  float* last = D.tmpl + ((size_t)s * 2 + (par ^ 1)) * n;
  // This frame's template and the two scratch images (blur row pass / warped image) are written and re-read by the whole CTA between barriers
  // a dozen times: in shared memory when they fit (VGA 14 KB, 1080p 94 KB), so that those hand-overs do not go through L2
  extern __shared__ float sbi_dyn[];
  const bool in_smem = D.smem_floats > 0;
  float* tmp = in_smem ? sbi_dyn : D.scratch + (size_t)s * 3 * n;
  float* warped = tmp + n;
  float* cur = in_smem ? sbi_dyn + 2 * n : cur_g;
#ifdef VS_SBI_TIMING
  if (tid < 12) sh.tacc[tid] = 0;
  if (tid == 0) sh.tlast = clock64();
  __syncthreads();
#endif
  sbi_make(D.l3 + (size_t)s * D.l3h * D.l3pitch, D.l3pitch, D.l3h, D.rs, W, H, D.taps, 9, D.small + (size_t)s * n, tmp, cur, sh);
  SBI_MARK(0);
  const bool first = !D.have[s];
  if (in_smem) { for (int i = tid; i < n; i += kT) cur_g[i] = cur[i]; }   // the next frame's `last`
  if (first) { for (int i = tid; i < n; i += kT) last[i] = cur[i]; }   // first frame: both SBIs come from the same keyframe (jni/Tracker.cc:90-93)
  __syncthreads();
  if (tid == 0) { D.have[s] = 1; D.parity[s] = par ^ 1; }               // next frame: `cur` becomes `last`
  // (The kernel reads no tracker state: it belongs to the pose-independent front end of a frame and may run before the previous frame's pose
  // is known.  A lost stream's rotation estimate is computed and never used: k_project_lists returns before the motion model.)
  if (!D.use_sbi) return;
  float* jac = D.jac + (size_t)s * 2 * n;
  sbi_make_jacs(last, jac, W, H);
  SBI_MARK(1);
  sbi_esm(cur, last, jac, warped, W, H, 6, sh);
  SBI_MARK(2);
  if (tid == 0) {   // SE3fromSE2 and ln() -> Tracker::mv6SBIRot
    double R[9]; sbi_se3_from_se2(sh.CtoC, D.cam, D.orig, W, H, R);
    double P[12]; for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) P[4 * i + j] = R[3 * i + j]; P[4 * i + 3] = 0.0; }
    double v6[6]; se3_ln(P, v6);
    for (int k = 0; k < 6; k++) D.rot_out[6 * (size_t)s + k] = v6[k];
  }
#ifdef VS_SBI_TIMING
  SBI_MARK(6);
  if (tid == 0 && s == 7) printf("k_sbi cycles: make %lld jacs %lld esm_rest %lld | esm transform %lld accumulate %lld solve %lld | se3fromse2 %lld\n", sh.tacc[0], sh.tacc[1], sh.tacc[2], sh.tacc[3], sh.tacc[4], sh.tacc[5], sh.tacc[6]);
#endif
}

// Relocaliser keyframe k: SmallBlurryImage(kf) with the default blur 2.5 (jni/KeyFrame.cc:98) + MakeJacs, from the level-3 image of a
// source keyframe.  One CTA per keyframe.
__global__ void __launch_bounds__(kT) k_reloc_make(RelocDev Rd, const int* __restrict__ src_ids) {
  __shared__ SbiShared sh;
  const int k = blockIdx.x, n = Rd.w * Rd.h;
  const uint8_t* l3 = Rd.src_l3 + (size_t)src_ids[k] * Rd.l3h * Rd.l3pitch;
  float* out = Rd.kf_tmpl + (size_t)k * n;
  sbi_make(l3, Rd.l3pitch, Rd.l3h, Rd.rs, Rd.w, Rd.h, Rd.taps17, 17, Rd.kf_small + (size_t)k * n, Rd.kf_tmp + (size_t)k * n, out, sh);
  sbi_make_jacs(out, Rd.kf_jac + (size_t)k * 2 * n, Rd.w, Rd.h);
}

// The lost branch of Tracker::TrackFrame (jni/Tracker.cc:134-140), one CTA per lost stream: Relocaliser::AttemptRecovery
// (jni/Relocaliser.cc:17-58: SmallBlurryImage of the frame with blur 2.5, ScoreKFs = SSD against every map keyframe's, six ESM
// iterations against the best, SE3fromSE2 * keyframe pose, accepted if the final score < 9e6) and Tracker::AttemptRecovery
// (jni/Tracker.cc:167-180: pose = start pose = best, velocity zero, coarse stage forced).  TrackMap + AssessTrackingQuality follow in
// the usual kernels, which treat a stream with `recovered` set as alive (no motion model before, no UpdateMotionModel after).
__global__ void __launch_bounds__(kT) k_relocalise(SbiDev D, RelocDev Rd) {
  cudaGridDependencySynchronize(); cudaTriggerProgrammaticLaunchCompletion();
  __shared__ SbiShared sh;
  const int s = blockIdx.x + D.s0, tid = threadIdx.x;
  StreamState* st = D.ss + s;
  if (tid == 0) st->recovered = 0;      // set again below if this frame relocalises the stream
  if (st->lost_frames < 3 || Rd.n_kf <= 0) return;
  const int W = D.w, H = D.h, n = W * H;
  float* tmp = D.reloc_scratch + (size_t)s * 3 * n;      // (not k_sbi's scratch: with frame look-ahead the next frame's k_sbi may be running)
  float* warped = tmp + n;
  float* cur = tmp + 2 * n;
  sbi_make(D.l3 + (size_t)s * D.l3h * D.l3pitch, D.l3pitch, D.l3h, D.rs, W, H, Rd.taps17, 17, D.reloc_small + (size_t)s * n, tmp, cur, sh);
  // ScoreKFs: SmallBlurryImage::ZMSSD (jni/SmallBlurryImage.cc:82-94), serial sum in the reference's order (x outer, y inner), one thread per keyframe
  double* scores = Rd.scores + (size_t)s * Rd.n_kf;
  for (int k = tid; k < Rd.n_kf; k += kT) {
    const float* o = Rd.kf_tmpl + (size_t)k * n;
    double d = 0.0;
    for (int x = 0; x < W; x++) for (int y = 0; y < H; y++) { const double df = cur[y * W + x] - o[y * W + x]; d += df * df; }
    scores[k] = d;
  }
  __syncthreads();
  if (tid == 0) {
    double best = 99999999999999.9; int nb = -1;
    for (int k = 0; k < Rd.n_kf; k++) if (scores[k] < best) { best = scores[k]; nb = k; }
    sh.best = nb;
  }
  __syncthreads();
  const int nb = sh.best;
  if (nb < 0) return;
  sbi_esm(cur, Rd.kf_tmpl + (size_t)nb * n, Rd.kf_jac + (size_t)nb * 2 * n, warped, W, H, 6, sh);
  if (tid == 0) {
    double R[9]; sbi_se3_from_se2(sh.CtoC, D.cam, D.orig, W, H, R);
    double P[12]; for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) P[4 * i + j] = R[3 * i + j]; P[4 * i + 3] = 0.0; }
    double bestp[12]; se3_mul(P, Rd.kf_pose + 12 * (size_t)nb, bestp);
    st->reloc_best = nb; st->reloc_score = sh.score;
    if (sh.score < 9e6) {
      for (int k = 0; k < 12; k++) { st->pose[k] = bestp[k]; st->start_pose[k] = bestp[k]; }
      for (int k = 0; k < 6; k++) st->velocity[k] = 0.0;
      st->just_recovered = 1; st->recovered = 1; st->n_recoveries++;
    }
  }
}

}  // namespace

static SbiDev make_sbi_dev(vslam_ctx* ctx) {
  SbiDev D;
  const LevelDesc& L3 = ctx->lev[3];
  D.w = L3.w / 2; D.h = L3.h / 2; D.l3w = L3.w; D.l3h = L3.h; D.l3pitch = L3.pitch; D.l3 = L3.img;
  D.rs.exact_half = ctx->sbi_exact_half; D.rs.l3w = L3.w; D.rs.tab = ctx->sbi_resize_tab;
  for (int k = 0; k < 9; k++) D.taps[k] = ctx->sbi_taps[k];
  D.cam = ctx->sbi_cam; memcpy(D.orig, ctx->sbi_orig, sizeof(D.orig));
  D.tmpl = ctx->sbi_tmpl; D.scratch = ctx->sbi_scratch; D.jac = ctx->sbi_jac; D.small = ctx->sbi_small; D.ss = ctx->ss; D.have = ctx->sbi_have; D.parity = ctx->sbi_have + ctx->S;
  D.use_sbi = ctx->params.use_sbi; D.s0 = ctx->cur_s0;
  { const size_t fl = 3 * (size_t)D.w * D.h; D.smem_floats = fl * sizeof(float) <= 200 * 1024 ? (int)fl : 0; }
  D.rot_out = ctx->sbi_rot_buf + (size_t)ctx->cur_set * ctx->S * 6; D.reloc_scratch = ctx->reloc_frame_scratch; D.reloc_small = ctx->reloc_frame_small;
  return D;
}
static RelocDev make_reloc_dev(vslam_ctx* ctx) {
  RelocDev R;
  const LevelDesc& L3 = ctx->lev[3];
  R.n_kf = ctx->reloc_n; R.w = L3.w / 2; R.h = L3.h / 2; R.l3h = ctx->src.h[3]; R.l3pitch = ctx->src.pitch[3]; R.src_l3 = ctx->src.img[3];
  R.rs.exact_half = ctx->sbi_exact_half; R.rs.l3w = L3.w; R.rs.tab = ctx->sbi_resize_tab;
  for (int k = 0; k < 17; k++) R.taps17[k] = ctx->reloc_taps[k];
  R.kf_tmpl = ctx->reloc_tmpl; R.kf_jac = ctx->reloc_jac; R.kf_tmp = ctx->reloc_tmp; R.kf_small = ctx->reloc_small; R.kf_pose = ctx->reloc_pose; R.scores = ctx->reloc_scores;
  return R;
}

int vs_launch_sbi(vslam_ctx* ctx) {
  if (!ctx->sbi_on) return VSLAM_OK;
  const SbiDev D = make_sbi_dev(ctx);
  vs_time_begin(ctx, VS_ST_OTHER);
  const size_t smem = (size_t)D.smem_floats * sizeof(float);
  static size_t smem_opted = 48 * 1024;
  if (smem > smem_opted) { VS_CUDA(cudaFuncSetAttribute(k_sbi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); smem_opted = smem; }
  VS_CUDA(vs_launch_pdl(k_sbi, dim3(ctx->cur_cnt), dim3(kT), smem, ctx->stream, ctx->pdl && !ctx->timing, D));
  vs_time_end(ctx);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  return VSLAM_OK;
}

// The lost branch of Tracker::TrackFrame: needs the lost state the previous frame left, i.e. it opens the back end of a frame
int vs_launch_relocalise(vslam_ctx* ctx) {
  if (!ctx->sbi_on) return VSLAM_OK;
  const SbiDev D = make_sbi_dev(ctx);
  if (ctx->reloc_n > 0) {   // lost streams try to relocalise; CTAs of streams that are not lost return at once
    vs_time_begin(ctx, VS_ST_OTHER);
    VS_CUDA(vs_launch_pdl(k_relocalise, dim3(ctx->cur_cnt), dim3(kT), 0, ctx->stream, ctx->pdl && !ctx->timing, D, make_reloc_dev(ctx)));
    vs_time_end(ctx);
    VS_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  return VSLAM_OK;
}

// SmallBlurryImages (blur 2.5) + gradient images of ctx->reloc_n relocaliser keyframes from the source-keyframe pyramids
int vs_launch_reloc_make(vslam_ctx* ctx, const int* src_ids_dev) {
  k_reloc_make<<<ctx->reloc_n, kT, 0, ctx->stream>>>(make_reloc_dev(ctx), src_ids_dev);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  return VSLAM_OK;
}

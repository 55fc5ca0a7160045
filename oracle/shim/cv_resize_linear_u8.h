// cv::resize(src, dst, dsize) with the default INTER_LINEAR on CV_8UC1 -- restatement of OpenCV's published algorithm (imgproc/src/imgwarp.cpp /
// resize.cpp, stable from 2.4 to 4.x: resizeGeneric_ + HResizeLinear / VResizeLinear<uchar, int, short>), because OpenCV itself is absent from
// this image's C++ toolchain.  TEST INFRASTRUCTURE (oracle/ and the stand-in OpenCV of the reference build); pinned against cv2.resize in
// tests/test_oracle_golden.py and tests/golden/resize_linear.npz.
//   * both scale factors exactly 2: OpenCV switches to its fast area path, (a + b + c + d + 2) >> 2;
//   * otherwise fixed-point bilinear: coefficients in 11 bits (cvRound of the float weights), horizontal pass without shift, vertical pass
//       dst = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2,
//     sample positions (d + 0.5) * scale - 0.5 evaluated in float, scale = 1 / (dsize / ssize) in double, borders clamped.
#pragma once
#include <cfloat>
#include <cmath>
#include <vector>

namespace cv_restated {

struct ResizeTables { std::vector<int> xofs, yofs; std::vector<short> ialpha, ibeta; bool area_fast; };

inline short sat_short_round(float v) { long r = lrintf(v); return (short)(r < -32768 ? -32768 : (r > 32767 ? 32767 : r)); }   // saturate_cast<short>(float): cvRound, ties to even

inline ResizeTables resize_linear_tables(int sw, int sh, int dw, int dh) {
  ResizeTables T;
  const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
  const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
  const int iscale_x = (int)lrint(scale_x), iscale_y = (int)lrint(scale_y);
  T.area_fast = std::fabs(scale_x - iscale_x) < DBL_EPSILON && std::fabs(scale_y - iscale_y) < DBL_EPSILON && iscale_x == 2 && iscale_y == 2;
  T.xofs.resize(dw); T.ialpha.resize(2 * dw); T.yofs.resize(dh); T.ibeta.resize(2 * dh);
  for (int dx = 0; dx < dw; dx++) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)std::floor(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    T.xofs[dx] = sx;
    T.ialpha[2 * dx] = sat_short_round((1.f - fx) * 2048); T.ialpha[2 * dx + 1] = sat_short_round(fx * 2048);
  }
  for (int dy = 0; dy < dh; dy++) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    const int sy = (int)std::floor(fy);
    fy -= sy;
    T.yofs[dy] = sy;
    T.ibeta[2 * dy] = sat_short_round((1.f - fy) * 2048); T.ibeta[2 * dy + 1] = sat_short_round(fy * 2048);
  }
  return T;
}

inline int clip(int x, int a, int b) { return x >= a ? (x < b ? x : b - 1) : a; }

inline void resize_linear_u8(const unsigned char* src, int sw, int sh, size_t sstep, unsigned char* dst, int dw, int dh, size_t dstep) {
  const ResizeTables T = resize_linear_tables(sw, sh, dw, dh);
  if (T.area_fast) {
    for (int y = 0; y < dh; y++) {
      const unsigned char* a = src + (size_t)(2 * y) * sstep; const unsigned char* b = a + sstep;
      for (int x = 0; x < dw; x++) dst[(size_t)y * dstep + x] = (unsigned char)((a[2 * x] + a[2 * x + 1] + b[2 * x] + b[2 * x + 1] + 2) >> 2);
    }
    return;
  }
  for (int dy = 0; dy < dh; dy++) {
    const unsigned char* S0 = src + (size_t)clip(T.yofs[dy], 0, sh) * sstep; const unsigned char* S1 = src + (size_t)clip(T.yofs[dy] + 1, 0, sh) * sstep;
    const int b0 = T.ibeta[2 * dy], b1 = T.ibeta[2 * dy + 1];
    for (int dx = 0; dx < dw; dx++) {
      const int sx = T.xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1, a0 = T.ialpha[2 * dx], a1 = T.ialpha[2 * dx + 1];
      const int r0 = S0[sx] * a0 + S0[sx1] * a1, r1 = S1[sx] * a0 + S1[sx1] * a1;
      dst[(size_t)dy * dstep + dx] = (unsigned char)((((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2);
    }
  }
}

}  // namespace cv_restated

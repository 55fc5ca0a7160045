// -*- c++ -*-
// Minimal stand-in for the slice of OpenCV 2.4 (cv::Mat + 5 functions) used by
// the reference's tracking front-end.  TEST INFRASTRUCTURE
// ONLY — lets the unmodified reference sources compile into oracle/_ref/ in an
// image without OpenCV headers.  Third-party arithmetic restated here:
//   cv::resize      INTER_LINEAR on CV_8UC1, restated in cv_resize_linear_u8.h; exact 2:1: (a+b+c+d+2)>>2  (what OpenCV's INTER_LINEAR
//                   gives for an exact half-size u8 image); other sizes: OpenCV's fixed-point bilinear.  Pinned against
//                   cv2 4.13 in tests/test_oracle_golden.py.
//   cv::cvtColor    CV_RGB2BGR on 3/4-channel u8 -> 3-channel.
//   cv::GaussianBlur float, separable, BORDER_REPLICATE, OpenCV's
//                   getGaussianKernel weights (only the SmallBlurryImage path).
#ifndef VSLAM_ORACLE_CV_SHIM_CORE
#define VSLAM_ORACLE_CV_SHIM_CORE

#include "../../cv_resize_linear_u8.h"
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8UC1 0
#define CV_8UC3 16
#define CV_8UC4 24
#define CV_32FC1 5
#define CV_32FC2 13
#define CV_RGB2BGR 4

#define CVAPI(rettype) rettype
#define CV_FUNCNAME(Name) static const char cvFuncName[] = Name; (void)cvFuncName
#define __CV_BEGIN__ {
#define __CV_END__ goto exit; exit: ; }
#define CV_StsNullPtr (-27)
#define CV_StsOutOfRange (-211)
#define CV_StsUnsupportedFormat (-210)
#define CV_ERROR(Code, Msg) do { fprintf(stderr, "cv shim error %d: %s\n", (Code), (Msg)); goto exit; } while (0)

struct CvPoint { int x, y; };
typedef unsigned char uchar;

namespace cv {

enum { BORDER_REPLICATE = 1 };

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Point { int x, y; Point() : x(0), y(0) {} template <class A, class B> Point(A a, B b) : x((int)a), y((int)b) {} };
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; } };
struct Vec2f { float v[2]; float& operator[](int i) { return v[i]; } const float& operator[](int i) const { return v[i]; } };

inline int shim_elem_size(int type) {
  switch (type) {
    case CV_8UC1: return 1; case CV_8UC3: return 3; case CV_8UC4: return 4;
    case CV_32FC1: return 4; case CV_32FC2: return 8;
  }
  fprintf(stderr, "cv shim: unsupported Mat type %d\n", type); abort();
}
inline int shim_channels(int type) {
  switch (type) { case CV_8UC3: return 3; case CV_8UC4: return 4; case CV_32FC2: return 2; default: return 1; }
}

class Mat {
 public:
  int rows, cols;
  unsigned char* data;
  size_t step;

  Mat() : rows(0), cols(0), data(0), step(0), type_(CV_8UC1) {}
  Mat(int r, int c, int type) : rows(0), cols(0), data(0), step(0), type_(CV_8UC1) { create(r, c, type); }
  // Wrap caller-owned memory (no copy), like cv::Mat(rows, cols, type, void*, step).
  Mat(int r, int c, int type, void* ext, size_t stp = 0)
      : rows(r), cols(c), data((unsigned char*)ext), step(stp ? stp : (size_t)c * shim_elem_size(type)), type_(type) {}

  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_ && step == (size_t)c * shim_elem_size(type)) return;
    rows = r; cols = c; type_ = type; step = (size_t)c * shim_elem_size(type);
    buf_.reset(new std::vector<unsigned char>(step * (size_t)r + 16));
    data = buf_->empty() ? 0 : &(*buf_)[0];
  }
  void create(Size s, int type) { create(s.height, s.width, type); }
  int type() const { return type_; }
  int channels() const { return shim_channels(type_); }
  Size size() const { return Size(cols, rows); }
  bool empty() const { return data == 0 || rows == 0 || cols == 0; }

  template <class T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <class T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
  template <class T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
  template <class T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }

  void copyTo(Mat& dst) const {
    const size_t rowbytes = (size_t)cols * shim_elem_size(type_);
    if (dst.data == data && dst.rows == rows && dst.cols == cols) return;
    dst.create(rows, cols, type_);
    for (int r = 0; r < rows; r++) memcpy(dst.data + (size_t)r * dst.step, data + (size_t)r * step, rowbytes);
  }
  Mat clone() const { Mat m; copyTo(m); return m; }

  Mat operator()(const Rect& roi) const {
    assert(roi.x >= 0 && roi.y >= 0 && roi.x + roi.width <= cols && roi.y + roi.height <= rows);
    Mat m(*this);
    m.rows = roi.height; m.cols = roi.width;
    m.data = data + (size_t)roi.y * step + (size_t)roi.x * shim_elem_size(type_);
    return m;
  }

 private:
  int type_;
  std::shared_ptr<std::vector<unsigned char> > buf_;
};

inline void resize(const Mat& src, Mat& dst, Size dsize) {   // INTER_LINEAR, CV_8UC1: cv_resize_linear_u8.h
  if (src.type() != CV_8UC1) { fprintf(stderr, "cv shim: resize supports CV_8UC1 only\n"); abort(); }
  Mat out; out.create(dsize.height, dsize.width, CV_8UC1);
  cv_restated::resize_linear_u8(src.data, src.cols, src.rows, src.step, out.data, dsize.width, dsize.height, out.step);
  dst = out;
}

inline void cvtColor(const Mat& src, Mat& dst, int code) {
  assert(code == CV_RGB2BGR); (void)code;
  const int scn = src.channels();
  if (scn != 3 && scn != 4) { fprintf(stderr, "cv shim: cvtColor needs 3/4 channels\n"); abort(); }
  Mat out(src.rows, src.cols, CV_8UC3);
  for (int y = 0; y < src.rows; y++) {
    const unsigned char* s = src.data + (size_t)y * src.step;
    unsigned char* d = out.data + (size_t)y * out.step;
    for (int x = 0; x < src.cols; x++) { d[3 * x] = s[scn * x + 2]; d[3 * x + 1] = s[scn * x + 1]; d[3 * x + 2] = s[scn * x]; }
  }
  dst = out;
}

inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
  assert(src.type() == CV_32FC1 && borderType == BORDER_REPLICATE); (void)borderType;
  // OpenCV getGaussianKernel(n, sigma, CV_32F): exp(-x^2/(2 sigma^2)) normalised, float taps.
  std::vector<float> kx(ksize.width), ky(ksize.height);
  for (int pass = 0; pass < 2; pass++) {
    std::vector<float>& k = pass ? ky : kx; const int n = (int)k.size(); const double sg = pass ? sigmaY : sigmaX;
    const double scale2x = -0.5 / (sg * sg); double sum = 0;
    for (int i = 0; i < n; i++) { double x = i - (n - 1) * 0.5; double t = std::exp(scale2x * x * x); k[i] = (float)t; sum += k[i]; }
    sum = 1. / sum;
    for (int i = 0; i < n; i++) k[i] = (float)(k[i] * sum);
  }
  const int W = src.cols, H = src.rows;
  std::vector<float> tmp((size_t)W * H);
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      float s = 0;
      for (int i = 0; i < ksize.width; i++) { int xx = x + i - ksize.width / 2; xx = xx < 0 ? 0 : (xx >= W ? W - 1 : xx); s += kx[i] * src.at<float>(y, xx); }
      tmp[(size_t)y * W + x] = s;
    }
  Mat out(H, W, CV_32FC1);
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      float s = 0;
      for (int i = 0; i < ksize.height; i++) { int yy = y + i - ksize.height / 2; yy = yy < 0 ? 0 : (yy >= H ? H - 1 : yy); s += ky[i] * tmp[(size_t)yy * W + x]; }
      out.at<float>(y, x) = s;
    }
  dst = out;
}

inline void line(Mat&, Point, Point, const Scalar&, int = 1) {}
inline void circle(Mat&, Point, int, const Scalar&, int = 1) {}

}  // namespace cv

#endif

// vslam_b200_shell.hpp — host C++ shell that keeps the reference's KeyFrame / Tracker API on top of the C-ABI
// (include/vslam_b200.h).  Header-only; compiles against the reference's own types (cv::Mat from OpenCV 2.4, Eigen) or
// against the stand-ins in oracle/shim (tests/test_abi_cpu.py compiles it that way).
//
//   reference                                             here
//   KeyFrame::MakeKeyFrame_Lite(cv::Mat&, cv::Mat&)       vslam_b200::KeyFrame::MakeKeyFrame_Lite   (jni/KeyFrame.h:89)
//   Level::{im, vCorners, vCornerRowLUT}                  filled from vslam_get_level / _corners / _row_lut (jni/KeyFrame.h:45-62)
//   Tracker::TrackFrame(cv::Mat&, cv::Mat&, bool)         vslam_b200::Tracker::TrackFrame            (jni/Tracker.h:55)
//   Tracker::GetCurrentPose()                             vslam_b200::Tracker::GetCurrentPose        (jni/Tracker.h:58)
//   Tracker::GetMessageForUser()                          same text format                           (jni/Tracker.cc:113-125)
// The colour image argument is accepted and ignored (it is only used for drawing / map-point colouring, off the hot path).
#ifndef VSLAM_B200_SHELL_HPP
#define VSLAM_B200_SHELL_HPP

#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <Eigen/Dense>
#include <opencv2/core/core.hpp>

#include "vslam_b200.h"

namespace vslam_b200 {

#ifndef LEVELS
#define LEVELS VSLAM_LEVELS
#endif

inline void check(vslam_ctx* c, int rc) { if (rc != VSLAM_OK) throw std::runtime_error(std::string("vslam_b200: ") + vslam_last_error(c)); }

struct Level {
  cv::Mat im;
  std::vector<Eigen::Vector2d> vCorners;   // all FAST corners on this level, raster order
  std::vector<int> vCornerRowLUT;          // row index into vCorners
};

// One camera stream of a context.  Several KeyFrame/Tracker objects may share one context (one per GPU).
class Context {
 public:
  Context(int width, int height, int n_streams, int max_points, int patch_size = 11, int device = 0) {
    vslam_config cfg; vslam_default_config(&cfg);
    cfg.width = width; cfg.height = height; cfg.n_streams = n_streams; cfg.max_points = max_points; cfg.patch_size = patch_size; cfg.device = device;
    if (vslam_create(&cfg, &c_) != VSLAM_OK) throw std::runtime_error(std::string("vslam_b200: ") + vslam_last_error(0));
  }
  ~Context() { vslam_destroy(c_); }
  vslam_ctx* get() const { return c_; }
 private:
  Context(const Context&); Context& operator=(const Context&);
  vslam_ctx* c_;
};

struct KeyFrame {
  KeyFrame(Context& ctx, int stream) : ctx_(ctx), stream_(stream) {}
  Level aLevels[LEVELS];

  // jni/KeyFrame.cc:5-51 — pyramid + FAST-10 + row LUT on the GPU; results are copied back into the reference's containers.
  void MakeKeyFrame_Lite(cv::Mat& im, cv::Mat& /*imColor*/) {
    vslam_ctx* c = ctx_.get();
    check(c, vslam_make_keyframe_lite(c, stream_, 1, im.data, (int)im.step, 0));
    Fetch();
  }
  void Fetch() {
    vslam_ctx* c = ctx_.get();
    for (int l = 0; l < LEVELS; l++) {
      int w, h, n; check(c, vslam_level_dims(c, l, &w, &h));
      aLevels[l].im.create(h, w, CV_8UC1);
      check(c, vslam_get_level(c, stream_, l, aLevels[l].im.data, (int)aLevels[l].im.step));
      check(c, vslam_get_num_corners(c, stream_, l, &n));
      std::vector<int32_t> xy(2 * (size_t)n + 2);
      check(c, vslam_get_corners(c, stream_, l, &xy[0], n));
      aLevels[l].vCorners.resize(n);
      for (int i = 0; i < n; i++) aLevels[l].vCorners[i] = Eigen::Vector2d(xy[2 * i], xy[2 * i + 1]);
      std::vector<int32_t> lut(h);
      check(c, vslam_get_row_lut(c, stream_, l, &lut[0]));
      aLevels[l].vCornerRowLUT.assign(lut.begin(), lut.end());
    }
  }
 private:
  Context& ctx_; int stream_;
};

// Pose as the reference's mySE3 stores it: rotation matrix + translation (camera-from-world).
struct SE3 { Eigen::Matrix3d R; Eigen::Vector3d t; };

class Tracker {
 public:
  // cam13: see vslam_set_camera; the map is set once with SetMap (the reference reads Map::vpPoints directly).
  Tracker(Context& ctx, int stream, const double* cam13) : ctx_(ctx), stream_(stream) { check(ctx_.get(), vslam_set_camera(ctx_.get(), cam13)); }

  void SetMap(int n, const double* world3, const double* right3, const double* down3, const int32_t* irCenter2, const int32_t* srcLevel, const int32_t* srcKF) {
    check(ctx_.get(), vslam_set_map(ctx_.get(), n, world3, right3, down3, irCenter2, srcLevel, srcKF)); n_points_ = n;
  }
  void SetSourceKeyFrame(int id, cv::Mat& gray) { check(ctx_.get(), vslam_upload_source_keyframe(ctx_.get(), id, gray.data, (int)gray.step)); }

  // jni/Tracker.cc:76-146, good-map branch.  With several streams per context use vslam_track_frame directly (one call tracks all streams).
  void TrackFrame(cv::Mat& imFrame, cv::Mat& /*imageColor*/, bool /*bDraw*/) {
    vslam_ctx* c = ctx_.get();
    check(c, vslam_track_frame(c, imFrame.data, (int)imFrame.step, 0));
    int32_t att[LEVELS], fnd[LEVELS]; int q, lost, coarse;
    check(c, vslam_get_counters(c, stream_, att, fnd, &q, &lost, &coarse));
    msg_.str("");
    msg_ << "Tracking Map, quality " << (q == 2 ? "good." : (q == 1 ? "poor." : "bad.")) << " Found:";
    for (int l = 0; l < LEVELS; l++) msg_ << " " << fnd[l] << "/" << att[l];
    msg_ << " Map: " << n_points_ << "P";
  }
  SE3 GetCurrentPose() {
    double p[12]; check(ctx_.get(), vslam_get_pose(ctx_.get(), stream_, p));
    SE3 s;
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) s.R(i, j) = p[4 * i + j]; s.t(i) = p[4 * i + 3]; }
    return s;
  }
  void SetCurrentPose(const SE3& s) {
    double p[12];
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) p[4 * i + j] = s.R(i, j); p[4 * i + 3] = s.t(i); }
    check(ctx_.get(), vslam_set_pose(ctx_.get(), stream_, p));
  }
  std::string GetMessageForUser() { return msg_.str(); }
 private:
  Context& ctx_; int stream_; int n_points_ = 0; std::ostringstream msg_;
};

}  // namespace vslam_b200
#endif

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
//   Tracker::Reset()                                      vslam_reset_stream                         (jni/Tracker.cc:45-60)
//   Tracker::TrackMap()                                   vslam_track_map                            (jni/Tracker.cc:358-626)
//   Tracker::SearchForPoints(vTD, nRange, nSubPixIts)     vslam_set_lists + vslam_search_for_points  (jni/Tracker.cc:629-674)
//   Tracker::CalcPoseUpdate(vTD, dOverrideSigma, bMark)   vslam_set_lists + vslam_calc_pose_update   (jni/Tracker.cc:683-774)
//   Tracker::TrackForInitialMap / TrailTracking_Start /   vslam_make_keyframe_rest, vslam_minipatch_*, vslam_snapshot_keyframe
//            TrailTracking_Advance                        (jni/Tracker.cc:203-346); InitFromStereo is MapMaker's: a hook
//   KeyFrame::MakeKeyFrame_Rest()                         vslam_make_keyframe_rest                   (jni/KeyFrame.cc:53-95)
//   MiniPatch::SampleFromImage / FindPatch                vslam_minipatch_sample / _find, one patch   (jni/MiniPatch.cc:6-83)
//   PatchFinder (per-object, slow path)                   stage calls on a one-entry list + vslam_pf_* (jni/PatchFinder.h:45-121): CalcSearchLevelAndWarpMatrix,
//                                                         MakeTemplateCoarse / Cont / NoWarp, FindPatchCoarse, ZMSSDAtPoint, MakeSubPixTemplate,
//                                                         IterateSubPix, IterateSubPixToConvergence, Get / SetSubPixPos
//   Tracker(width, height, camera, map, mapmaker)         vslam_b200::TrackerOnReferenceTypes         (jni/Tracker.h:43, jni/jni_part.cpp:27-46)
//   SystemPTAM::onTouchScreen                             Tracker::mbUserPressedSpacebar / vslam_user_event (jni/jni_part.cpp:49-51)
//   MapMaker::ReFindInSingleKeyFrame / ReFind_Common,     vslam_b200::MapSearch (vslam_refind, vslam_epipolar_search)
//            the search of AddPointEpipolar               (jni/MapMaker.cc:525-640, 967-1056)
//   Relocaliser (inside Tracker::TrackFrame when lost)    Tracker::SetRelocKeyFrames                 (jni/Relocaliser.cc:17-58)
// Map points are addressed by their index in the arrays given to SetMap (the reference passes MapPoint& / TrackerData*).
// The colour image argument is accepted and ignored (it is only used for drawing / map-point colouring, off the hot path).
#ifndef VSLAM_B200_SHELL_HPP
#define VSLAM_B200_SHELL_HPP

#include <algorithm>
#include <list>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include <Eigen/Dense>
#include <opencv2/core/core.hpp>

#include "vslam_b200.h"

namespace vslam_b200 {

#ifndef LEVELS
#define LEVELS VSLAM_LEVELS
#endif

inline void check(vslam_ctx* c, int rc) { if (rc != VSLAM_OK) throw std::runtime_error(std::string("vslam_b200: ") + vslam_last_error(c)); }

struct Candidate { Eigen::Vector2d irLevelPos; double dSTScore; };   // jni/KeyFrame.h:33-38
struct Level {
  cv::Mat im;
  std::vector<Eigen::Vector2d> vCorners;     // all FAST corners on this level, raster order
  std::vector<int> vCornerRowLUT;            // row index into vCorners
  std::vector<Eigen::Vector2d> vMaxCorners;  // maximal FAST corners (MakeKeyFrame_Rest)
  std::vector<Candidate> vCandidates;        // Shi-Tomasi thresholded maximal corners (MakeKeyFrame_Rest)
};

// One camera stream of a context.  Several KeyFrame/Tracker objects may share one context (one per GPU).
class Context {
 public:
  Context(int width, int height, int n_streams, int max_points, int patch_size = 11, int device = 0, int max_source_keyframes = 0 /* 0: library default */) {
    vslam_config cfg; vslam_default_config(&cfg);
    cfg.width = width; cfg.height = height; cfg.n_streams = n_streams; cfg.max_points = max_points; cfg.patch_size = patch_size; cfg.device = device;
    if (max_source_keyframes > 0) cfg.max_source_keyframes = max_source_keyframes;
    if (vslam_create(&cfg, &c_) != VSLAM_OK) throw std::runtime_error(std::string("vslam_b200: ") + vslam_last_error(0));
    n_streams_ = n_streams;
  }
  ~Context() { vslam_destroy(c_); }
  vslam_ctx* get() const { return c_; }
  int Streams() const { return n_streams_; }
  int MapSize() const { return n_points_; }
  void SetMapSize(int n) { n_points_ = n; }
  // Map files (no reference counterpart; layout in csrc/mapfile.cu): what this context holds of the map -> disk and back.
  void SaveMap(const std::string& path) { if (vslam_save_map_file(c_, path.c_str()) != VSLAM_OK) throw std::runtime_error(std::string("vslam_b200: ") + vslam_last_error(c_)); }
  void LoadMap(const std::string& path, int flags = VSLAM_MAP_LOAD_CAMERA) {
    if (vslam_load_map_file(c_, path.c_str(), flags) != VSLAM_OK) throw std::runtime_error(std::string("vslam_b200: ") + vslam_last_error(c_));
    vslam_map_file_info_t info; vslam_map_file_info(path.c_str(), &info); n_points_ = info.n_points;
  }
  // MapMaker::GUICommandHandler("SaveMap") (jni/MapMaker.cc:1254-1297): map.dump + keyframes/<i>.info under `dir`
  void SaveMapText(const std::string& dir) { if (vslam_export_map_text(c_, dir.c_str()) != VSLAM_OK) throw std::runtime_error(std::string("vslam_b200: ") + vslam_last_error(c_)); }
 private:
  int n_points_ = 0, n_streams_ = 0;
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
  // jni/KeyFrame.cc:53-95 — FAST score + non-max suppression + Shi-Tomasi candidates of the stream's current keyframe.
  void MakeKeyFrame_Rest() {
    vslam_ctx* c = ctx_.get();
    check(c, vslam_make_keyframe_rest(c, stream_));
    for (int l = 0; l < LEVELS; l++) {
      int n = 0; check(c, vslam_get_max_corners(c, stream_, l, 0, 0, &n));
      std::vector<int32_t> xy(2 * (size_t)n + 2); std::vector<double> sc((size_t)n + 1);
      check(c, vslam_get_max_corners(c, stream_, l, &xy[0], n, &n));
      aLevels[l].vMaxCorners.resize(n);
      for (int i = 0; i < n; i++) aLevels[l].vMaxCorners[i] = Eigen::Vector2d(xy[2 * i], xy[2 * i + 1]);
      int m = 0; check(c, vslam_get_candidates(c, stream_, l, &xy[0], &sc[0], n, &m));
      aLevels[l].vCandidates.resize(m);
      for (int i = 0; i < m; i++) { aLevels[l].vCandidates[i].irLevelPos = Eigen::Vector2d(xy[2 * i], xy[2 * i + 1]); aLevels[l].vCandidates[i].dSTScore = sc[i]; }
    }
  }
  int stream() const { return stream_; }
  Context& context() const { return ctx_; }
 private:
  Context& ctx_; int stream_;
};

// jni/MiniPatch.h:16-26.  The searched image is a device-resident level-0 image of the stream: which = 0 the current keyframe,
// 1 its snapshot (vslam_snapshot_keyframe; Tracker::mPreviousFrameKF).  One patch per call: the slow path; the trail tracker
// below batches all trails into one call.
struct MiniPatch {
  static int& mnHalfPatchSize() { static int v = 4; return v; }       // jni/MiniPatch.cc:86-88
  static int& mnRange() { static int v = 10; return v; }
  static int& mnMaxSSD() { static int v = 9999; return v; }
  unsigned char im[81];
  void SampleFromImage(const Eigen::Vector2d& irPos, KeyFrame& kf, int which = 0) {
    const int32_t xy[2] = {(int32_t)irPos(0), (int32_t)irPos(1)};
    check(kf.context().get(), vslam_minipatch_sample(kf.context().get(), kf.stream(), which, 1, xy, im));
  }
  bool FindPatch(Eigen::Vector2d& irPos, KeyFrame& kf, int nRange, int which = 0) {
    double pos[2] = {irPos(0), irPos(1)}; int32_t found = 0;
    check(kf.context().get(), vslam_minipatch_find(kf.context().get(), kf.stream(), which, 1, im, pos, &found, 0, nRange, mnMaxSSD()));
    if (found) irPos = Eigen::Vector2d(pos[0], pos[1]);
    return found != 0;
  }
};

// Pose as the reference's mySE3 stores it: rotation matrix + translation (camera-from-world).
struct SE3 { Eigen::Matrix3d R; Eigen::Vector3d t; };

// Initial-map trail (jni/Tracker.h:37-41)
struct Trail { MiniPatch mPatch; Eigen::Vector2d irCurrentPos, irInitialPos; };

class Tracker {
 public:
  // cam13: see vslam_set_camera; the map is set once with SetMap (the reference reads Map::vpPoints directly).
  Tracker(Context& ctx, int stream, const double* cam13) : mCurrentKF(ctx, stream), ctx_(ctx), stream_(stream) {
    check(ctx_.get(), vslam_set_camera(ctx_.get(), cam13)); Reset();
  }

  void SetMap(int n, const double* world3, const double* right3, const double* down3, const int32_t* irCenter2, const int32_t* srcLevel, const int32_t* srcKF) {
    check(ctx_.get(), vslam_set_map(ctx_.get(), n, world3, right3, down3, irCenter2, srcLevel, srcKF)); ctx_.SetMapSize(n); mbMapGood = n > 0;
  }
  // SmallBlurryImage rotation estimate inside TrackFrame, as the reference always does (jni/Tracker.cc:87-97); cam13_small = the
  // camera scalars for the thumbnail size (width/16 x height/16), see vslam_camera_from_params
  void EnableSBI(const double* cam13_small) { check(ctx_.get(), vslam_enable_sbi(ctx_.get(), cam13_small)); }
  void SetSourceKeyFrame(int id, cv::Mat& gray) { check(ctx_.get(), vslam_upload_source_keyframe(ctx_.get(), id, gray.data, (int)gray.step)); }

  // jni/Tracker.cc:45-60 (the tracker's part; MapMaker::RequestReset is the caller's)
  void Reset() {
    check(ctx_.get(), vslam_reset_stream(ctx_.get(), stream_));
    mbUserPressedSpacebar = false; mnInitialStage = TRAIL_TRACKING_NOT_STARTED; mlTrails.clear(); mnFrame = 0; mnLastKeyFrameDropped = -20; mbMapGood = false;
  }
  // jni/Tracker.cc:349-353 (the reference's GUI handler sets the flag; SystemPTAM::onTouchScreen writes the member directly, jni/jni_part.cpp:49-51)
  void PressSpacebar() { mbUserPressedSpacebar = true; }
  bool mbUserPressedSpacebar = false;

  // jni/Tracker.cc:76-146.  Good map: MakeKeyFrame_Lite, SmallBlurryImage, motion model, TrackMap, quality — one call, all on the
  // device.  No map yet: MakeKeyFrame_Lite + TrackForInitialMap.  With several streams per context use vslam_track_frame
  // directly (one call tracks all streams).
  void TrackFrame(cv::Mat& imFrame, cv::Mat& /*imageColor*/, bool /*bDraw*/) {
    vslam_ctx* c = ctx_.get();
    msg_.str("");
    mnFrame++;
    { int ev = 0; check(c, vslam_take_user_event(c, stream_, &ev)); if (ev & VSLAM_EVENT_SPACEBAR) mbUserPressedSpacebar = true; }   // events posted through the C-ABI
    if (!mbMapGood) {
      check(c, vslam_make_keyframe_lite(c, stream_, 1, imFrame.data, (int)imFrame.step, 0));
      TrackForInitialMap();
      return;
    }
    int32_t att[LEVELS], fnd[LEVELS]; int q, lost, coarse;
    check(c, vslam_get_counters(c, stream_, att, fnd, &q, &lost, &coarse));
    const bool was_lost = lost >= 3;
    check(c, vslam_track_frame(c, imFrame.data, (int)imFrame.step, 0));
    if (was_lost) { msg_ << "** Attempting recovery **."; return; }   // jni/Tracker.cc:134-140 (the relocaliser runs inside vslam_track_frame)
    check(c, vslam_get_counters(c, stream_, att, fnd, &q, &lost, &coarse));
    msg_ << "Tracking Map, quality " << (q == 2 ? "good." : (q == 1 ? "poor." : "bad.")) << " Found:";
    for (int l = 0; l < LEVELS; l++) msg_ << " " << fnd[l] << "/" << att[l];
    msg_ << " Map: " << ctx_.MapSize() << "P";
    if (mnKeyFrames >= 0) msg_ << ", " << mnKeyFrames << "KF";
    // jni/Tracker.cc:127-132: the first three terms are evaluated on the device (vslam_set_keyframe_policy); the queue length is ours
    if (mnKeyFrames >= 0) {
      std::vector<int32_t> req(ctx_.Streams());
      check(c, vslam_get_keyframe_requests(c, &req[0], 0, 0));
      if (req[stream_] && mnQueueSize < 3) { msg_ << " Adding key-frame."; AddNewKeyFrame(); }
    }
  }
  // The MapMaker heuristics TrackFrame consults (MapMaker::NeedNewKeyFrame, IsDistanceToNearestKeyFrameExcessive).  nKeyFrames = how
  // many source-keyframe slots the map already uses: keyframes added by this tracker take the following ids.
  void SetKeyFramePolicy(int nKeyFrames, double dWiggleScale, double dWiggleScaleDepthNormalized, double dMaxKFDistWiggleMult = 0.2, int nMinFrames = 20) {
    check(ctx_.get(), vslam_set_keyframe_policy(ctx_.get(), 1, dWiggleScale, dWiggleScaleDepthNormalized, dMaxKFDistWiggleMult, nMinFrames));
    mnKeyFrames = nKeyFrames;
  }
  // Tracker::AddNewKeyFrame (jni/Tracker.cc:823-827) + MapMaker::AddKeyFrame: device-to-device copy of mCurrentKF into the next slot
  void AddNewKeyFrame() {
    check(ctx_.get(), vslam_add_keyframe_from_stream(ctx_.get(), stream_, mnKeyFrames));
    mnLastKeyFrameDropped = mnFrame; mnKeyFrames++;
  }
  int mnQueueSize = 0;           // MapMaker::QueueSize(): set by the host map maker while it digests keyframes
  int mnKeyFrames = -1;          // -1: keyframe policy off
  int mnLastKeyFrameDropped = -20;
  // Relocaliser keyframes (Map::vpKeyFrames with their SmallBlurryImages, jni/Relocaliser.cc): ids of uploaded source keyframes + poses (n x 12)
  void SetRelocKeyFrames(int n, const int32_t* src_kf_ids, const double* poses12) { check(ctx_.get(), vslam_set_reloc_keyframes(ctx_.get(), n, src_kf_ids, poses12)); }
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

  // ---- the reference's protected stage functions, usable one at a time on this stream (other streams' lists are left empty)
  void TrackMap() { check(ctx_.get(), vslam_track_map(ctx_.get())); }
  // vTD: indices of the map points to search (the reference passes vector<TrackerData*>); the points must have been projected
  // (TrackMap / vslam_project_all).  Returns the number found.
  int SearchForPoints(const std::vector<int>& vTD, int nRange, int nSubPixIts) {
    vslam_ctx* c = ctx_.get();
    SetList(vTD);
    check(c, vslam_clear_counters(c));
    check(c, vslam_search_for_points(c, nRange, nSubPixIts));
    int32_t att[LEVELS], fnd[LEVELS]; int q, lost, coarse, n = 0;
    check(c, vslam_get_counters(c, stream_, att, fnd, &q, &lost, &coarse));
    for (int l = 0; l < LEVELS; l++) n += fnd[l];
    return n;
  }
  // TrackerData::CalcJacobian for the found points of the list (jni/TrackerData.h:107-123)
  void CalcJacobians(const std::vector<int>& vTD) { SetList(vTD); check(ctx_.get(), vslam_calc_jacobians(ctx_.get())); }
  // Jacobians of the listed points must be current (CalcJacobians, or TrackMap's own iterations).
  Eigen::VectorXd CalcPoseUpdate(const std::vector<int>& vTD, double dOverrideSigma = 0.0, bool bMarkOutliers = false) {
    vslam_ctx* c = ctx_.get();
    SetList(vTD);
    std::vector<double> upd(6 * (size_t)ctx_.Streams());
    check(c, vslam_calc_pose_update(c, dOverrideSigma, bMarkOutliers ? 1 : 0, 0, &upd[0]));
    Eigen::VectorXd v(6);
    for (int k = 0; k < 6; k++) v(k) = upd[6 * (size_t)stream_ + k];
    return v;
  }

  // ---- initial map: trails between the first two keyframes (jni/Tracker.cc:203-346) -------------------------------------------
  enum { TRAIL_TRACKING_NOT_STARTED, TRAIL_TRACKING_STARTED, TRAIL_TRACKING_COMPLETE };
  int mnInitialStage;
  std::list<Trail> mlTrails;
  KeyFrame mCurrentKF;
  // Matches handed to MapMaker::InitFromStereo at the second spacebar press (MapMaker is outside this library)
  std::vector<std::pair<Eigen::Vector2d, Eigen::Vector2d> > vInitMatches;

  void TrackForInitialMap() {
    MiniPatch::mnMaxSSD() = 100000;   // "Tracker.MiniPatchMaxSSD" (jni/Tracker.cc:226-227)
    if (mnInitialStage == TRAIL_TRACKING_NOT_STARTED) {
      if (mbUserPressedSpacebar) { mbUserPressedSpacebar = false; TrailTracking_Start(); mnInitialStage = TRAIL_TRACKING_STARTED; }
      else msg_ << "Point camera at planar scene and press spacebar to start tracking for initial map." << std::endl;
      return;
    }
    if (mnInitialStage == TRAIL_TRACKING_STARTED) {
      const int nGoodTrails = TrailTracking_Advance();
      if (nGoodTrails < 10) { Reset(); return; }
      if (mbUserPressedSpacebar) {
        mbUserPressedSpacebar = false;
        vInitMatches.clear();
        for (std::list<Trail>::iterator i = mlTrails.begin(); i != mlTrails.end(); ++i) vInitMatches.push_back(std::make_pair(i->irInitialPos, i->irCurrentPos));
        mnInitialStage = TRAIL_TRACKING_COMPLETE;   // the caller runs InitFromStereo on vInitMatches and then SetMap
      } else msg_ << "Translate the camera slowly sideways, and press spacebar again to perform stereo init." << std::endl;
    }
  }

  // The current frame is to be the first keyframe (jni/Tracker.cc:264-292)
  void TrailTracking_Start() {
    vslam_ctx* c = ctx_.get();
    mCurrentKF.MakeKeyFrame_Rest();
    int w, h; check(c, vslam_level_dims(c, 0, &w, &h));
    const int b = MiniPatch::mnHalfPatchSize();
    std::vector<std::pair<double, Eigen::Vector2d> > v;
    const std::vector<Candidate>& cand = mCurrentKF.aLevels[0].vCandidates;
    for (size_t i = 0; i < cand.size(); i++) {
      const int x = (int)cand[i].irLevelPos(0), y = (int)cand[i].irLevelPos(1);
      if (!(x >= b && y >= b && x < w - b && y < h - b)) continue;      // in_image_with_border
      v.push_back(std::make_pair(-1.0 * cand[i].dSTScore, cand[i].irLevelPos));
    }
    std::sort(v.begin(), v.end(), CompareFirst());   // the reference's comparator and std::sort: as shipped, the LOWEST scores come first
    int nToAdd = 1000;                               // "MaxInitialTrails"
    std::vector<int32_t> xy;
    for (size_t i = 0; i < v.size() && nToAdd > 0; i++, nToAdd--) { xy.push_back((int32_t)v[i].second(0)); xy.push_back((int32_t)v[i].second(1)); }
    const int n = (int)(xy.size() / 2);
    std::vector<unsigned char> patches(81 * (size_t)n + 1);
    if (n) check(c, vslam_minipatch_sample(c, stream_, 0, n, &xy[0], &patches[0]));    // all trails in one call
    mlTrails.clear();
    for (int i = 0; i < n; i++) {
      Trail t;
      std::copy(patches.begin() + 81 * (size_t)i, patches.begin() + 81 * (size_t)(i + 1), t.mPatch.im);
      t.irInitialPos = Eigen::Vector2d(xy[2 * i], xy[2 * i + 1]); t.irCurrentPos = t.irInitialPos;
      mlTrails.push_back(t);
    }
    check(c, vslam_snapshot_keyframe(c, stream_));   // mPreviousFrameKF = mFirstKF
  }

  // Steady-state trail tracking: advance from the previous frame, remove duds (jni/Tracker.cc:294-346).  Forward search of every
  // trail in the current frame, patches re-sampled at the hits, backward search in the previous frame: three batched calls.
  int TrailTracking_Advance() {
    vslam_ctx* c = ctx_.get();
    const int n = (int)mlTrails.size();
    int nGoodTrails = 0;
    if (n) {
      std::vector<unsigned char> patches(81 * (size_t)n);
      std::vector<double> pos(2 * (size_t)n), bpos;
      std::vector<int32_t> found(n), bfound, fxy;
      int k = 0;
      for (std::list<Trail>::iterator i = mlTrails.begin(); i != mlTrails.end(); ++i, ++k) {
        std::copy(i->mPatch.im, i->mPatch.im + 81, patches.begin() + 81 * (size_t)k);
        pos[2 * k] = i->irCurrentPos(0); pos[2 * k + 1] = i->irCurrentPos(1);
      }
      const std::vector<double> start(pos);
      check(c, vslam_minipatch_find(c, stream_, 0, n, &patches[0], &pos[0], &found[0], 0, 10, MiniPatch::mnMaxSSD()));
      std::vector<int> idx;
      for (k = 0; k < n; k++) if (found[k]) { idx.push_back(k); fxy.push_back((int32_t)pos[2 * k]); fxy.push_back((int32_t)pos[2 * k + 1]); bpos.push_back(pos[2 * k]); bpos.push_back(pos[2 * k + 1]); }
      const int m = (int)idx.size();
      std::vector<unsigned char> back(81 * (size_t)m + 1);
      bfound.assign(m + 1, 0);
      if (m) {
        check(c, vslam_minipatch_sample(c, stream_, 0, m, &fxy[0], &back[0]));                                            // BackwardsPatch.SampleFromImage
        check(c, vslam_minipatch_find(c, stream_, 1, m, &back[0], &bpos[0], &bfound[0], 0, 10, MiniPatch::mnMaxSSD()));   // married-match check
      }
      std::vector<char> keep(n, 0);
      for (int j = 0; j < m; j++) {
        k = idx[j];
        const double dx = bpos[2 * j] - start[2 * k], dy = bpos[2 * j + 1] - start[2 * k + 1];
        bool bFound = bfound[j] != 0;
        if (dx * dx + dy * dy > 2) bFound = false;
        keep[k] = bFound ? 1 : 0;
        nGoodTrails++;                                  // counted before the married-match verdict, as the reference does (:317-318)
      }
      k = 0;
      for (std::list<Trail>::iterator i = mlTrails.begin(); i != mlTrails.end(); ++k) {
        if (found[k]) i->irCurrentPos = Eigen::Vector2d(pos[2 * k], pos[2 * k + 1]);
        if (!keep[k]) i = mlTrails.erase(i); else ++i;
      }
    }
    check(c, vslam_snapshot_keyframe(c, stream_));   // mPreviousFrameKF = mCurrentKF
    return nGoodTrails;
  }

 private:
  struct CompareFirst { bool operator()(const std::pair<double, Eigen::Vector2d>& a, const std::pair<double, Eigen::Vector2d>& b) const { return a.first > b.first; } };   // jni/Tracker.h:47-52 (`>` on -dSTScore)
  void SetList(const std::vector<int>& vTD) {
    vslam_ctx* c = ctx_.get();
    const int stride = (int)vTD.size() > 0 ? (int)vTD.size() : 1;
    std::vector<int32_t> idx((size_t)ctx_.Streams() * stride, 0), n(ctx_.Streams(), 0);
    for (size_t k = 0; k < vTD.size(); k++) idx[(size_t)stream_ * stride + k] = vTD[k];
    n[stream_] = (int32_t)vTD.size();
    check(c, vslam_set_lists(c, &idx[0], &n[0], stride));
  }
 protected:
  Context& ctx_; int stream_; int mnFrame = 0; bool mbMapGood = false; std::ostringstream msg_;
};

// jni/PatchFinder.h:45-121, per object, for one map point of one stream at a time: the slow path (every call moves the whole
// stream's point state across PCIe).  The batched fast path is Tracker::SearchForPoints / vslam_track_frame.
class PatchFinder {
 public:
  PatchFinder(Context& ctx, int stream, int nPatchSize = 11) : mnMaxSSD(nPatchSize * nPatchSize * 500), ctx_(ctx), stream_(stream), point_(-1), mnSearchLevel(-1), mbFound(false), mbTemplateBad(false) {}
  // Projects the map with se3CFromW as the stream's pose and returns the search level of `point` (negative: inappropriate warp)
  int CalcSearchLevelAndWarpMatrix(int point, const SE3& se3CFromW) {
    vslam_ctx* c = ctx_.get();
    double p[12];
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) p[4 * i + j] = se3CFromW.R(i, j); p[4 * i + 3] = se3CFromW.t(i); }
    check(c, vslam_set_pose(c, stream_, p));
    check(c, vslam_project_all(c));
    point_ = point;
    Pull();
    mnSearchLevel = ints_[8 * (size_t)point + 1];
    for (int r = 0; r < 2; r++) for (int q = 0; q < 2; q++) mm2WarpInverse(r, q) = dbl_[32 * (size_t)point + 11 + 2 * r + q];
    mbTemplateBad = mnSearchLevel < 0;
    return mnSearchLevel;
  }
  int GetLevel() { return mnSearchLevel; }
  int GetLevelScale() { return 1 << mnSearchLevel; }
  Eigen::Matrix2d GetWarpInverse() { return mm2WarpInverse; }
  // TrackerData::v2Image of the point at that pose (what the reference's callers pass to FindPatchCoarse), and any point's level
  Eigen::Vector2d GetProjection() { return Eigen::Vector2d(dbl_[32 * (size_t)point_], dbl_[32 * (size_t)point_ + 1]); }
  int LevelOf(int point) { return ints_[8 * (size_t)point + 1]; }
  // jni/PatchFinder.cc:79-125, for `point` with the warp CalcSearchLevelAndWarpMatrix left on the device; the re-use rule (:91-102) applies
  void MakeTemplateCoarseCont(int point) {
    point_ = point;
    int bad = 0; check(ctx_.get(), vslam_pf_make_template(ctx_.get(), stream_, point_, &bad));
    mbTemplateBad = bad != 0;
  }
  // jni/PatchFinder.cc:72-76.  (The reference also takes the camera derivatives at the projection; the device recomputes them.)
  void MakeTemplateCoarse(int point, const SE3& se3CFromW) { if (CalcSearchLevelAndWarpMatrix(point, se3CFromW) >= 0) MakeTemplateCoarseCont(point); }
  // jni/PatchFinder.cc:146-149: un-warped pixels of the point's own source keyframe, level and irCenter
  void MakeTemplateCoarseNoWarp(int point) { NoWarp(point, -1, 0, 0, 0); }
  // jni/PatchFinder.cc:130-143 with the keyframe named by its source-keyframe slot (Tracker::SetSourceKeyFrame)
  void MakeTemplateCoarseNoWarp(int point, int src_kf_id, int nLevel, int irLevelPos0, int irLevelPos1) { NoWarp(point, src_kf_id, nLevel, irLevelPos0, irLevelPos1); }
  bool TemplateBad() { return mbTemplateBad; }
  // jni/PatchFinder.cc:352-380.  img must be a level image of the stream's current keyframe (kf.aLevels[l].im): it names the level, the
  // pixels that are compared are the device's copy of that level.
  int ZMSSDAtPoint(cv::Mat& img, int icol, int irow) {
    vslam_ctx* c = ctx_.get();
    int level = -1;
    for (int l = 0; l < LEVELS && level < 0; l++) { int w, h; check(c, vslam_level_dims(c, l, &w, &h)); if (w == img.cols && h == img.rows) level = l; }
    if (level < 0) throw std::runtime_error("vslam_b200: ZMSSDAtPoint: the image is not a level of the stream's keyframe");
    const int32_t xy[2] = {icol, irow}; int32_t ssd = 0;
    check(c, vslam_pf_zmssd_at(c, stream_, point_, level, 1, xy, &ssd));
    return ssd;
  }
  // Search around v2Pos (level-0 pixels) in the stream's current keyframe (kf must be that keyframe)
  bool FindPatchCoarse(const Eigen::Vector2d& v2Pos, KeyFrame& /*kf*/, unsigned int nRange) { return Search(v2Pos, (int)nRange, 0); }
  Eigen::Vector2d GetCoarsePosAsVector() { return mv2CoarsePos; }
  // jni/PatchFinder.cc:242-267: start of the refinement (the inverse of JtJ is rebuilt from the template by every device call)
  void MakeSubPixTemplate() { mv2SubPixPos = mv2CoarsePos; mdMeanDiff = 0.0; }
  // jni/PatchFinder.cc:272-285
  bool IterateSubPixToConvergence(KeyFrame& /*kf*/, int nMaxIts) {
    double pos[2] = {mv2SubPixPos(0), mv2SubPixPos(1)}; int conv = 0;
    check(ctx_.get(), vslam_pf_subpix(ctx_.get(), stream_, point_, nMaxIts, pos, &mdMeanDiff, &conv, 0));
    mv2SubPixPos = Eigen::Vector2d(pos[0], pos[1]);
    return conv != 0;
  }
  // jni/PatchFinder.cc:290-350: one iteration; returns the squared pixel update, negative when the patch left the image
  double IterateSubPix(KeyFrame& /*kf*/) {
    double pos[2] = {mv2SubPixPos(0), mv2SubPixPos(1)}, upd = -1.0;
    check(ctx_.get(), vslam_pf_subpix(ctx_.get(), stream_, point_, 1, pos, &mdMeanDiff, 0, &upd));
    mv2SubPixPos = Eigen::Vector2d(pos[0], pos[1]);
    return upd;
  }
  void SetSubPixPos(const Eigen::Vector2d& v2) { mv2SubPixPos = v2; }
  // Coarse search + inverse-compositional refinement in one device call (the reference splits them; the result is the same)
  bool FindPatchCoarseAndSubPix(const Eigen::Vector2d& v2Pos, KeyFrame& /*kf*/, unsigned int nRange, int nMaxIts) { return Search(v2Pos, (int)nRange, nMaxIts); }
  Eigen::Vector2d GetSubPixPos() { return mv2SubPixPos; }
  int mnMaxSSD;

 private:
  void Pull() {
    vslam_ctx* c = ctx_.get();
    const int n = ctx_.MapSize();
    ints_.resize(8 * (size_t)n); dbl_.resize(32 * (size_t)n);
    check(c, vslam_get_point_states(c, stream_, &ints_[0], &dbl_[0]));
  }
  void NoWarp(int point, int kf, int level, int x, int y) {
    point_ = point;
    int bad = 0; check(ctx_.get(), vslam_pf_make_template_nowarp(ctx_.get(), stream_, point_, kf, level, x, y, &bad));
    mbTemplateBad = bad != 0;
    Pull(); mnSearchLevel = ints_[8 * (size_t)point_ + 1];
  }
  bool Search(const Eigen::Vector2d& v2Pos, int nRange, int nSubPix) {
    vslam_ctx* c = ctx_.get();
    const int n = ctx_.MapSize();
    Pull();
    std::vector<double> v2(2 * (size_t)n), warp(4 * (size_t)n); std::vector<int32_t> lev(n);
    for (int i = 0; i < n; i++) { v2[2 * i] = dbl_[32 * (size_t)i]; v2[2 * i + 1] = dbl_[32 * (size_t)i + 1]; for (int q = 0; q < 4; q++) warp[4 * i + q] = dbl_[32 * (size_t)i + 11 + q]; lev[i] = ints_[8 * (size_t)i + 1]; }
    v2[2 * point_] = v2Pos(0); v2[2 * point_ + 1] = v2Pos(1);
    check(c, vslam_set_point_projection(c, stream_, &v2[0], &warp[0], &lev[0]));
    const int ns = ctx_.Streams();
    std::vector<int32_t> idx(ns, 0), cnt(ns, 0);
    idx[stream_] = point_; cnt[stream_] = 1;
    check(c, vslam_set_lists(c, &idx[0], &cnt[0], 1));
    check(c, vslam_search_for_points(c, nRange, nSubPix));
    Pull();
    const int32_t* I = &ints_[8 * (size_t)point_]; const double* Dd = &dbl_[32 * (size_t)point_];
    mbTemplateBad = I[5] != 0;
    mbFound = I[3] != 0;
    mv2CoarsePos = Eigen::Vector2d(Dd[30], Dd[31]);
    mv2SubPixPos = Eigen::Vector2d(Dd[2], Dd[3]);
    return mbFound;
  }
  Context& ctx_; int stream_, point_, mnSearchLevel; bool mbFound, mbTemplateBad;
  Eigen::Matrix2d mm2WarpInverse; Eigen::Vector2d mv2CoarsePos, mv2SubPixPos; double mdMeanDiff = 0.0;
  std::vector<int32_t> ints_; std::vector<double> dbl_;
};

// The reference's own constructor shape (jni/Tracker.h:43; jni/jni_part.cpp:27-46 builds `new Tracker(800, 480, *mpCamera, *mpMap,
// *mpMapMaker)`), on the reference's own types -- or anything shaped like them:
//   ATANCameraT  mvDefaultParams: the five camera parameters (jni/ATANCamera.h:97; the port reads no others)
//   MapT         vpPoints (MapPoint*: v3WorldPos, v3PixelRight_W, v3PixelDown_W, irCenter, nSourceLevel, pPatchSourceKF), vpKeyFrames
//                (KeyFrame*: aLevels[0].im), IsGood()                                             (jni/Map.h:29-45, jni/MapPoint.h:33-54)
//   MapMakerT    kept by reference only (its thread is switched off in the reference)
// The tracker owns a one-stream context.  Every TrackFrame first brings the device's flat copy of the map up to date: new keyframes are
// uploaded to source slots (slot = index in vpKeyFrames), new points appended (the reference reads Map::vpPoints live, jni/Tracker.cc:372).
struct OwnedContext { explicit OwnedContext(Context* c) : own_(c) {} ~OwnedContext() { delete own_; } Context* own_; };
template <class ATANCameraT, class MapT, class MapMakerT>
class TrackerOnReferenceTypes : private OwnedContext, public Tracker {
 public:
  // bAsShippedRadius: reproduce the int-temporary bug of jni/ATANCamera.cc:70-82 (largest radius 0: no map point is ever in view)
  TrackerOnReferenceTypes(int width, int height, const ATANCameraT& c, MapT& m, MapMakerT& mm, int max_points = 8192, int max_keyframes = 16, int device = 0, bool bAsShippedRadius = false)
      : OwnedContext(new Context(width, height, 1, max_points, 11, device, max_keyframes)), Tracker(*own_, 0, Cam13(c, width, height, bAsShippedRadius).v), mMap(m), mMapMaker(mm),
        mnPointsOnDevice(0), mnKeyFramesOnDevice(0) {
    if (height % 16 == 0) EnableSBI(Cam13(c, width / 16, height / 16, bAsShippedRadius).v);   // SmallBlurryImage, as the reference's TrackFrame always does
  }
  void TrackFrame(cv::Mat& imFrame, cv::Mat& imageColor, bool bDraw) { SyncMap(); Tracker::TrackFrame(imFrame, imageColor, bDraw); }
  // Flatten what is new in Map::vpKeyFrames / Map::vpPoints onto the device
  void SyncMap() {
    vslam_ctx* c = ctx_.get();
    for (; mnKeyFramesOnDevice < (int)mMap.vpKeyFrames.size(); mnKeyFramesOnDevice++) {
      cv::Mat& im = mMap.vpKeyFrames[mnKeyFramesOnDevice]->aLevels[0].im;
      check(c, vslam_upload_source_keyframe(c, mnKeyFramesOnDevice, im.data, (int)im.step));
    }
    const int n = (int)mMap.vpPoints.size();
    if (n > mnPointsOnDevice) {
      const int k0 = mnPointsOnDevice, m = n - k0;
      std::vector<double> w(3 * (size_t)m), r(3 * (size_t)m), d(3 * (size_t)m); std::vector<int32_t> irc(2 * (size_t)m), lvl(m), kf(m);
      for (int k = 0; k < m; k++) {
        const auto& p = *mMap.vpPoints[k0 + k];
        for (int q = 0; q < 3; q++) { w[3 * k + q] = p.v3WorldPos(q); r[3 * k + q] = p.v3PixelRight_W(q); d[3 * k + q] = p.v3PixelDown_W(q); }
        irc[2 * k] = (int32_t)p.irCenter(0); irc[2 * k + 1] = (int32_t)p.irCenter(1); lvl[k] = p.nSourceLevel;
        int id = 0; while (id < (int)mMap.vpKeyFrames.size() && mMap.vpKeyFrames[id] != p.pPatchSourceKF) id++;
        if (id == (int)mMap.vpKeyFrames.size()) throw std::runtime_error("vslam_b200: a map point's source keyframe is not in Map::vpKeyFrames");
        kf[k] = id;
      }
      if (k0 == 0) check(c, vslam_set_map(c, m, &w[0], &r[0], &d[0], &irc[0], &lvl[0], &kf[0]));
      else check(c, vslam_append_map_points(c, m, &w[0], &r[0], &d[0], &irc[0], &lvl[0], &kf[0]));
      mnPointsOnDevice = n; ctx_.SetMapSize(n);
    }
    mbMapGood = mMap.IsGood() && mnPointsOnDevice > 0;
  }
  MapT& mMap; MapMakerT& mMapMaker;
 private:
  struct Cam13 { double v[13]; Cam13(const ATANCameraT& c, int w, int h, bool shipped) { double p5[5]; for (int k = 0; k < 5; k++) p5[k] = c.mvDefaultParams(k); vslam_camera_from_params(p5, w, h, shipped ? 1 : 0, v); } };
  int mnPointsOnDevice, mnKeyFramesOnDevice;
};

// jni/KeyFrame.h:45-50
struct Measurement {
  int nLevel; bool bSubPix; Eigen::Vector2d v2RootPos;
  enum { SRC_TRACKER, SRC_REFIND, SRC_ROOT, SRC_TRAIL, SRC_EPIPOLAR } Source;
};

// The two searches MapMaker runs on keyframes (jni/MapMaker.cc), over the device kernels of the tracker.  A keyframe is a stream of
// the context: its image must be that stream's current keyframe (KeyFrame::MakeKeyFrame_Lite) and `se3CfromW` its pose.  MapMaker's
// bookkeeping (sMeasurementKFs, sNeverRetryKFs, the new MapPoint and its triangulation) stays with the caller.
class MapSearch {
 public:
  explicit MapSearch(Context& ctx) : ctx_(ctx), mdWiggleScale(0.1) {}
  double mdWiggleScale;   // MapMaker::mdWiggleScale (jni/MapMaker.cc:206)

  // MapMaker::ReFindInSingleKeyFrame / ReFind_Common (jni/MapMaker.cc:967-1056) for the listed map points in keyframe `k`.
  // Returns the number found; out[i] is valid where found[i] != 0.
  int ReFindInSingleKeyFrame(KeyFrame& k, const SE3& se3CfromW, const std::vector<int>& points, std::vector<Measurement>& out, std::vector<char>& found) {
    vslam_ctx* c = ctx_.get();
    double p[12];
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) p[4 * i + j] = se3CfromW.R(i, j); p[4 * i + 3] = se3CfromW.t(i); }
    check(c, vslam_set_pose(c, k.stream(), p));
    const int ns = ctx_.Streams(), n = (int)points.size(), stride = n > 0 ? n : 1;
    std::vector<int32_t> idx((size_t)ns * stride, 0), cnt(ns, 0);
    for (int i = 0; i < n; i++) idx[(size_t)k.stream() * stride + i] = points[i];
    cnt[k.stream()] = n;
    check(c, vslam_set_lists(c, &idx[0], &cnt[0], stride));
    check(c, vslam_refind(c, 4, 8));                                   // "Very tight search radius!" (jni/MapMaker.cc:1007), 8 sub-pixel iterations (:1018)
    std::vector<int32_t> fl(3 * (size_t)stride); std::vector<double> pos(2 * (size_t)stride); int got = 0;
    check(c, vslam_get_refind_results(c, k.stream(), &fl[0], &pos[0], n, &got));
    out.assign(n, Measurement()); found.assign(n, 0);
    int nFoundNow = 0;
    for (int i = 0; i < n; i++) {
      if (!fl[3 * i]) continue;
      found[i] = 1; nFoundNow++;
      out[i].nLevel = fl[3 * i + 1]; out[i].bSubPix = fl[3 * i + 2] != 0; out[i].v2RootPos = Eigen::Vector2d(pos[2 * i], pos[2 * i + 1]); out[i].Source = Measurement::SRC_REFIND;
    }
    return nFoundNow;
  }

  // The search of MapMaker::AddPointEpipolar (jni/MapMaker.cc:525-640) for candidates (level pixels, e.g. Level::vCandidates of kSrc)
  // of level nLevel of source keyframe `src_kf_id` (uploaded with Tracker::SetSourceKeyFrame) in the target keyframe kTarget.
  // found[i] != 0: a match converged; out[i] = the SRC_EPIPOLAR measurement in the target (v2RootPos refined to sub-pixel).
  int AddPointsEpipolar(int src_kf_id, const SE3& srcCfromW, double dSceneDepthMean, double dSceneDepthSigma, KeyFrame& kTarget, const SE3& targetCfromW, int nLevel,
                        const std::vector<Eigen::Vector2d>& candidates, std::vector<Measurement>& out, std::vector<char>& found) {
    vslam_ctx* c = ctx_.get();
    double ps[12], pt[12];
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) { ps[4 * i + j] = srcCfromW.R(i, j); pt[4 * i + j] = targetCfromW.R(i, j); } ps[4 * i + 3] = srcCfromW.t(i); pt[4 * i + 3] = targetCfromW.t(i); }
    const int n = (int)candidates.size();
    std::vector<int32_t> xy(2 * (size_t)n + 2), f(n + 1); std::vector<double> pos(2 * (size_t)n + 2);
    for (int i = 0; i < n; i++) { xy[2 * i] = (int32_t)candidates[i](0); xy[2 * i + 1] = (int32_t)candidates[i](1); }
    check(c, vslam_epipolar_search(c, kTarget.stream(), src_kf_id, nLevel, n, &xy[0], ps, pt, dSceneDepthMean, dSceneDepthSigma, mdWiggleScale, &f[0], &pos[0], 0, 0));
    out.assign(n, Measurement()); found.assign(n, 0);
    int nAdded = 0;
    for (int i = 0; i < n; i++) {
      if (!f[i]) continue;
      found[i] = 1; nAdded++;
      out[i].nLevel = nLevel; out[i].bSubPix = true; out[i].v2RootPos = Eigen::Vector2d(pos[2 * i], pos[2 * i + 1]); out[i].Source = Measurement::SRC_EPIPOLAR;
    }
    return nAdded;
  }

  // The MapPoint that MapMaker::AddPointEpipolar creates for a converged candidate (jni/MapMaker.cc:646-690): fields the tracker reads.
  struct NewMapPoint {
    Eigen::Vector3d v3WorldPos, v3PixelRight_W, v3PixelDown_W;
    Eigen::Vector2d irCenter;   // level pixels in the source keyframe
    int nSourceLevel;
    Measurement root;           // SRC_ROOT measurement in the source keyframe
  };
  // Triangulate (MapMaker::ReprojectPoint, jni/MapMaker.cc:176-200) every candidate with found[i] != 0 of an AddPointsEpipolar call
  // and fill its patch-source fields + RefreshPixelVectors; indices[k] = which candidate points[k] came from.
  void MakeEpipolarPoints(const SE3& srcCfromW, const SE3& targetCfromW, int nLevel, const std::vector<Eigen::Vector2d>& candidates, const std::vector<Measurement>& meas,
                          const std::vector<char>& found, std::vector<NewMapPoint>& points, std::vector<int>& indices) {
    vslam_ctx* c = ctx_.get();
    double ps[12], pt[12];
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) { ps[4 * i + j] = srcCfromW.R(i, j); pt[4 * i + j] = targetCfromW.R(i, j); } ps[4 * i + 3] = srcCfromW.t(i); pt[4 * i + 3] = targetCfromW.t(i); }
    indices.clear();
    for (size_t i = 0; i < candidates.size(); i++) if (found[i]) indices.push_back((int)i);
    const size_t n = indices.size();
    std::vector<int32_t> xy(2 * n + 2), irc(2 * n + 2), lvl(n + 1); std::vector<double> pos(2 * n + 2), w(3 * n + 3), r(3 * n + 3), d(3 * n + 3);
    for (size_t k = 0; k < n; k++) {
      const int i = indices[k];
      xy[2 * k] = (int32_t)candidates[i](0); xy[2 * k + 1] = (int32_t)candidates[i](1); pos[2 * k] = meas[i].v2RootPos(0); pos[2 * k + 1] = meas[i].v2RootPos(1);
    }
    check(c, vslam_epipolar_make_points(c, nLevel, (int)n, &xy[0], &pos[0], ps, pt, &w[0], &r[0], &d[0], &irc[0], &lvl[0]));
    points.assign(n, NewMapPoint());
    for (size_t k = 0; k < n; k++) {
      NewMapPoint& p = points[k];
      p.v3WorldPos = Eigen::Vector3d(w[3 * k], w[3 * k + 1], w[3 * k + 2]); p.v3PixelRight_W = Eigen::Vector3d(r[3 * k], r[3 * k + 1], r[3 * k + 2]);
      p.v3PixelDown_W = Eigen::Vector3d(d[3 * k], d[3 * k + 1], d[3 * k + 2]);
      p.irCenter = Eigen::Vector2d(irc[2 * k], irc[2 * k + 1]); p.nSourceLevel = lvl[k];
      p.root.Source = Measurement::SRC_ROOT; p.root.nLevel = nLevel; p.root.bSubPix = true;
      p.root.v2RootPos = Eigen::Vector2d((irc[2 * k] + 0.5) * (1 << nLevel) - 0.5, (irc[2 * k + 1] + 0.5) * (1 << nLevel) - 0.5);
    }
  }
 private:
  Context& ctx_;
};

}  // namespace vslam_b200
#endif

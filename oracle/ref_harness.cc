// C-ABI harness around the UNMODIFIED reference tracking front-end (jni/*.cc compiled by
// oracle/build_ref.sh).  TEST INFRASTRUCTURE ONLY: used by tests/ to pin oracle/vslam_oracle.cc
// (the restatement) and by bench.py's cpu_baseline / --impl reference legs.  Never linked into
// the product library.
//
// What is NOT the reference here (and why):
//   * MapMaker's THREAD.  jni/MapMaker.cc, Bundle.cc and HomographyInit.cc are compiled as they are (the Eigen stand-in provides a
//     JacobiSVD / EigenSolver, oracle/shim/Eigen/Dense), but the reference never starts the map-maker thread
//     (jni/MapMaker.cc:55-56), so (a) the reset handshake of Tracker::Reset (jni/Tracker.cc:67-69) is answered by a
//     short-lived helper thread that does what MapMaker::run does on a reset request, and (b) keyframes the tracker queues with
//     MapMaker::AddKeyFrame are taken off the queue after every frame and appended to the map with their SmallBlurryImage -- the
//     map-building work of AddKeyFrameFromTopOfQueue (new points, bundle adjustment) is out of scope.
//   * ref_cam_fix_radius(): optional run-time overwrite of ATANCamera::mdLargestRadius/mdMaxR
//     with the value the code at jni/ATANCamera.cc:70-82 evidently intended (double, not int,
//     temporaries).  As shipped both are 0 and TrackerData::Project rejects every point
//     (SURVEY.md F5); the flag is off unless a test turns it on.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <queue>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#include <unistd.h>

#include <Eigen/Dense>
#include <opencv2/core/core.hpp>

// Reach the tracker's protected state from the harness (access specifiers do not change layout).
#define protected public
#define private public
// TrackerData.h defines two non-inline helpers and is meant to be included by Tracker.cc only;
// rename them in this translation unit so the link does not see duplicates.
#define myProject_TrackerData harness_myProject_TrackerData
#define myUnproject_TrackerData harness_myUnproject_TrackerData
#include "Tracker.h"
#include "TrackerData.h"
#include "MEstimator.h"
#include "SmallBlurryImage.h"
#undef protected
#undef private

#include "MapMaker.h"
#include <atomic>
#include <thread>

namespace { int g_keyframes_added = 0; }

namespace {

mySE3 pose_from12(const double* p) {  // row-major 3x4 [R|t]
  mySE3 s;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) s.get_rotation().get_matrix()(i, j) = p[4 * i + j];
    s.get_translation()(i) = p[4 * i + 3];
  }
  return s;
}
void pose_to12(const mySE3& s, double* p) {
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) p[4 * i + j] = s.get_rotation().get_matrix()(i, j);
    p[4 * i + 3] = s.get_translation()(i);
  }
}

struct RefTracker {
  Map* map; MapMaker* mm; Tracker* tr;
};
// Keyframes queued by the real MapMaker::AddKeyFrame (jni/MapMaker.cc: a copy with pSBI = NULL) join the map at once, as if the
// map-maker thread had taken them off its queue before the next frame, with the relocaliser's SmallBlurryImage as
// KeyFrame::MakeKeyFrame_Rest (jni/KeyFrame.cc:98) and the map maker's MakeJacs leave it.
void drain_keyframe_queue(RefTracker* t) {
  std::vector<KeyFrame*>& q = t->mm->mvpKeyFrameQueue;
  for (size_t i = 0; i < q.size(); i++) {
    KeyFrame* kf = q[i];
    kf->pSBI = new SmallBlurryImage(*kf);
    kf->pSBI->MakeJacs();
    t->map->vpKeyFrames.push_back(kf);
    g_keyframes_added++;
  }
  q.clear();
}

cv::Mat wrap_gray(const uint8_t* g, int w, int h, int stride) { return cv::Mat(h, w, CV_8UC1, (void*)g, (size_t)stride); }

uint8_t g_dummy_rgba[4] = {0, 0, 0, 0};

void fix_radius(ATANCamera& c, int on) {
  if (!on) { c.RefreshParams(); return; }
  const Eigen::VectorXd& p = c.mgvvCameraParams;
  Eigen::Vector2d v2;
  v2(0) = std::max(p[2], 1.0 - p[2]) / p[0];
  v2(1) = std::max(p[3], 1.0 - p[3]) / p[1];
  c.mdLargestRadius = c.invrtrans(sqrt(v2.dot(v2)));
  c.mdMaxR = 1.5 * c.mdLargestRadius;
}

}  // namespace

extern "C" {

void ref_srand(unsigned seed) { srand(seed); }
int ref_rand() { return rand(); }

// ------------------------------------------------------------------ KeyFrame (jni/KeyFrame.cc)
void* ref_kf_create() { return new KeyFrame(); }
void ref_kf_destroy(void* kf) { delete (KeyFrame*)kf; }

void ref_kf_make_lite(void* kf_, const uint8_t* gray, int w, int h, int stride, const uint8_t* rgba) {
  KeyFrame* kf = (KeyFrame*)kf_;
  cv::Mat im = wrap_gray(gray, w, h, stride);
  cv::Mat col = rgba ? cv::Mat(h, w, CV_8UC4, (void*)rgba) : cv::Mat(1, 1, CV_8UC4, g_dummy_rgba);
  kf->MakeKeyFrame_Lite(im, col);
}
void ref_kf_make_rest(void* kf_) { ((KeyFrame*)kf_)->MakeKeyFrame_Rest(); }

void ref_kf_set_pose(void* kf_, const double* pose12) { ((KeyFrame*)kf_)->se3CfromW = pose_from12(pose12); }
void ref_kf_level_dims(void* kf_, int l, int* w, int* h) { Level& L = ((KeyFrame*)kf_)->aLevels[l]; *w = L.im.cols; *h = L.im.rows; }
void ref_kf_level_pixels(void* kf_, int l, uint8_t* out) {
  Level& L = ((KeyFrame*)kf_)->aLevels[l];
  for (int y = 0; y < L.im.rows; y++) memcpy(out + (size_t)y * L.im.cols, L.im.ptr<uint8_t>(y), L.im.cols);
}
int ref_kf_num_corners(void* kf_, int l) { return (int)((KeyFrame*)kf_)->aLevels[l].vCorners.size(); }
void ref_kf_corners(void* kf_, int l, int32_t* xy) {
  std::vector<Eigen::Vector2d>& v = ((KeyFrame*)kf_)->aLevels[l].vCorners;
  for (size_t i = 0; i < v.size(); i++) { xy[2 * i] = (int)v[i](0); xy[2 * i + 1] = (int)v[i](1); }
}
int ref_kf_row_lut(void* kf_, int l, int32_t* out) {
  std::vector<int>& v = ((KeyFrame*)kf_)->aLevels[l].vCornerRowLUT;
  for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
  return (int)v.size();
}
int ref_kf_num_max_corners(void* kf_, int l) { return (int)((KeyFrame*)kf_)->aLevels[l].vMaxCorners.size(); }
void ref_kf_max_corners(void* kf_, int l, int32_t* xy) {
  std::vector<Eigen::Vector2d>& v = ((KeyFrame*)kf_)->aLevels[l].vMaxCorners;
  for (size_t i = 0; i < v.size(); i++) { xy[2 * i] = (int)v[i](0); xy[2 * i + 1] = (int)v[i](1); }
}
int ref_kf_num_candidates(void* kf_, int l) { return (int)((KeyFrame*)kf_)->aLevels[l].vCandidates.size(); }
void ref_kf_candidates(void* kf_, int l, int32_t* xy, double* score) {
  std::vector<Candidate>& v = ((KeyFrame*)kf_)->aLevels[l].vCandidates;
  for (size_t i = 0; i < v.size(); i++) { xy[2 * i] = (int)v[i].irLevelPos(0); xy[2 * i + 1] = (int)v[i].irLevelPos(1); score[i] = v[i].dSTScore; }
}

// FAST pieces on a bare image (jni/vision/cvfast.cpp)
int ref_fast10(const uint8_t* gray, int w, int h, int stride, int thr, int32_t* xy, int cap) {
  cv::Mat im = wrap_gray(gray, w, h, stride);
  std::vector<Eigen::Vector2d> c;
  cvCornerFast_10(im, c, thr);
  for (size_t i = 0; i < c.size() && (int)i < cap; i++) { xy[2 * i] = (int)c[i](0); xy[2 * i + 1] = (int)c[i](1); }
  return (int)c.size();
}
double ref_shi_tomasi(const uint8_t* gray, int w, int h, int stride, int nsize, int px, int py) {
  cv::Mat im = wrap_gray(gray, w, h, stride);
  return FindShiTomasiScoreAtPoint(im, nsize, px, py);
}

// ------------------------------------------------------------------ ATANCamera (jni/ATANCamera.cc)
void* ref_cam_create(double w, double h, int fix) {
  ATANCamera* c = new ATANCamera("Camera");
  Eigen::Vector2d sz(w, h);
  c->SetImageSize(sz);
  fix_radius(*c, fix);
  return c;
}
void ref_cam_destroy(void* c) { delete (ATANCamera*)c; }
void ref_cam_fix_radius(void* c, int on) { fix_radius(*(ATANCamera*)c, on); }
// out[13]: focal xy, center xy, W, Winv, 2Tan, OneOver2Tan, DistortionEnabled, LargestRadius, MaxR, OnePixelDist, (pad)
void ref_cam_scalars(void* c_, double* out) {
  ATANCamera* c = (ATANCamera*)c_;
  out[0] = c->mvFocal(0); out[1] = c->mvFocal(1); out[2] = c->mvCenter(0); out[3] = c->mvCenter(1);
  out[4] = c->mdW; out[5] = c->mdWinv; out[6] = c->md2Tan; out[7] = c->mdOneOver2Tan; out[8] = c->mdDistortionEnabled;
  out[9] = c->mdLargestRadius; out[10] = c->mdMaxR; out[11] = c->mdOnePixelDist; out[12] = 0;
}
void ref_cam_project(void* c_, const double* cam2, double* im2, int* invalid, double* derivs4) {
  ATANCamera* c = (ATANCamera*)c_;
  Eigen::Vector2d v(cam2[0], cam2[1]);
  Eigen::Vector2d r = c->Project(v);
  im2[0] = r(0); im2[1] = r(1);
  if (invalid) *invalid = c->Invalid();
  if (derivs4) { Eigen::Matrix2d d = c->GetProjectionDerivs_Eigen(); derivs4[0] = d(0, 0); derivs4[1] = d(0, 1); derivs4[2] = d(1, 0); derivs4[3] = d(1, 1); }
}
void ref_cam_unproject(void* c_, const double* im2, double* cam2) {
  Eigen::Vector2d v(im2[0], im2[1]);
  Eigen::Vector2d r = ((ATANCamera*)c_)->UnProject(v);
  cam2[0] = r(0); cam2[1] = r(1);
}

// ------------------------------------------------------------------ SE3 (jni/RT.h)
void ref_se3_exp(const double* mu6, double* pose12) {
  Eigen::VectorXd mu(6);
  for (int i = 0; i < 6; i++) mu(i) = mu6[i];
  pose_to12(mySE3::exp(mu), pose12);
}
void ref_se3_ln(const double* pose12, double* mu6) {
  Eigen::VectorXd mu = mySE3::ln(pose_from12(pose12));
  for (int i = 0; i < 6; i++) mu6[i] = mu(i);
}
void ref_se3_mul(const double* a12, const double* b12, double* out12) { pose_to12(pose_from12(a12) * pose_from12(b12), out12); }
void ref_se3_inverse(const double* a12, double* out12) { pose_to12(pose_from12(a12).inverse(), out12); }

// ------------------------------------------------------------------ Map / MapPoint (jni/Map.h, jni/MapPoint.cc)
void* ref_map_create() { Map* m = new Map(); return m; }
void ref_map_set_good(void* m, int good) { ((Map*)m)->bGood = good != 0; }
int ref_map_add_keyframe(void* m_, void* kf) { Map* m = (Map*)m_; m->vpKeyFrames.push_back((KeyFrame*)kf); return (int)m->vpKeyFrames.size() - 1; }
int ref_map_num_points(void* m_) { return (int)((Map*)m_)->vpPoints.size(); }
// MapPoint built with the AddPointEpipolar recipe's fields (jni/MapMaker.cc:650-686); vectors are given by the caller.
int ref_map_add_point(void* m_, void* srckf, int level, const double* ircenter2, const double* world3, const double* center_nc3,
                      const double* oneright_nc3, const double* onedown_nc3, const double* normal_nc3) {
  Map* m = (Map*)m_;
  MapPoint* p = new MapPoint();
  p->pPatchSourceKF = (KeyFrame*)srckf;
  p->nSourceLevel = level;
  p->irCenter = Eigen::Vector2d(ircenter2[0], ircenter2[1]);
  for (int i = 0; i < 3; i++) {
    p->v3WorldPos(i) = world3[i]; p->v3Center_NC(i) = center_nc3[i]; p->v3OneRightFromCenter_NC(i) = oneright_nc3[i];
    p->v3OneDownFromCenter_NC(i) = onedown_nc3[i]; p->v3Normal_NC(i) = normal_nc3[i];
  }
  p->RefreshPixelVectors();
  m->vpPoints.push_back(p);
  return (int)m->vpPoints.size() - 1;
}
void ref_map_point_pixel_vectors(void* m_, int i, double* right3, double* down3) {
  MapPoint* p = ((Map*)m_)->vpPoints[i];
  for (int k = 0; k < 3; k++) { right3[k] = p->v3PixelRight_W(k); down3[k] = p->v3PixelDown_W(k); }
}
void ref_map_point_counts(void* m_, int i, int* outlier, int* inlier) {
  MapPoint* p = ((Map*)m_)->vpPoints[i];
  *outlier = p->nMEstimatorOutlierCount; *inlier = p->nMEstimatorInlierCount;
}

// ------------------------------------------------------------------ PatchFinder, one object at a time (jni/PatchFinder.cc)
void* ref_pf_create(int P) { return new PatchFinder(P); }
void ref_pf_destroy(void* pf) { delete (PatchFinder*)pf; }
int ref_pf_max_ssd(void* pf) { return ((PatchFinder*)pf)->mnMaxSSD; }
int ref_pf_calc_level_warp(void* pf_, void* map, int pt, const double* pose12, const double* derivs4, double* warpinv4) {
  PatchFinder* pf = (PatchFinder*)pf_;
  Eigen::Matrix2d d; d(0, 0) = derivs4[0]; d(0, 1) = derivs4[1]; d(1, 0) = derivs4[2]; d(1, 1) = derivs4[3];
  int l = pf->CalcSearchLevelAndWarpMatrix(*((Map*)map)->vpPoints[pt], pose_from12(pose12), d);
  const Eigen::Matrix2d& w = pf->mm2WarpInverse;
  warpinv4[0] = w(0, 0); warpinv4[1] = w(0, 1); warpinv4[2] = w(1, 0); warpinv4[3] = w(1, 1);
  return l;
}
void ref_pf_set_level_warp(void* pf_, int level, const double* warpinv4) {
  PatchFinder* pf = (PatchFinder*)pf_;
  pf->mnSearchLevel = level;
  pf->mm2WarpInverse(0, 0) = warpinv4[0]; pf->mm2WarpInverse(0, 1) = warpinv4[1];
  pf->mm2WarpInverse(1, 0) = warpinv4[2]; pf->mm2WarpInverse(1, 1) = warpinv4[3];
}
// returns TemplateBad(); *regenerated tells whether the reuse cache (jni/PatchFinder.cc:91-102) let it through
int ref_pf_make_template(void* pf_, void* map, int pt, uint8_t* tmpl, int* sum, int* sumsq) {
  PatchFinder* pf = (PatchFinder*)pf_;
  pf->mbTemplateBad = false;
  pf->MakeTemplateCoarseCont(*((Map*)map)->vpPoints[pt]);
  const int P = pf->mnPatchSize;
  for (int r = 0; r < P; r++) memcpy(tmpl + r * P, pf->mimTemplate.ptr<uint8_t>(r), P);
  *sum = pf->mnTemplateSum; *sumsq = pf->mnTemplateSumSq;
  return pf->TemplateBad();
}
void ref_pf_set_template(void* pf_, const uint8_t* tmpl) {
  PatchFinder* pf = (PatchFinder*)pf_;
  const int P = pf->mnPatchSize;
  for (int r = 0; r < P; r++) memcpy(pf->mimTemplate.ptr<uint8_t>(r), tmpl + r * P, P);
  // PatchFinder::MakeTemplateSums is declared inline in PatchFinder.cc (no external symbol): same two sums here.
  int sum = 0, sumsq = 0;
  for (int k = 0; k < P * P; k++) { sum += tmpl[k]; sumsq += tmpl[k] * tmpl[k]; }
  pf->mnTemplateSum = sum; pf->mnTemplateSumSq = sumsq;
  pf->mbTemplateBad = false;
}
int ref_pf_make_template_nowarp(void* pf_, void* kf, int level, int x, int y, uint8_t* tmpl, int* sum, int* sumsq) {
  PatchFinder* pf = (PatchFinder*)pf_;
  pf->MakeTemplateCoarseNoWarp(*(KeyFrame*)kf, level, x, y);
  const int P = pf->mnPatchSize;
  if (!pf->TemplateBad()) {
    for (int r = 0; r < P; r++) memcpy(tmpl + r * P, pf->mimTemplate.ptr<uint8_t>(r), P);
    *sum = pf->mnTemplateSum; *sumsq = pf->mnTemplateSumSq;
  }
  return pf->TemplateBad();
}
int ref_pf_zmssd(void* pf_, void* kf, int level, int x, int y) {
  return ((PatchFinder*)pf_)->ZMSSDAtPoint(((KeyFrame*)kf)->aLevels[level].im, x, y);
}
int ref_pf_find_coarse(void* pf_, double x, double y, void* kf, unsigned range, double* pos2) {
  PatchFinder* pf = (PatchFinder*)pf_;
  bool f = pf->FindPatchCoarse(Eigen::Vector2d(x, y), *(KeyFrame*)kf, range);
  if (f) { Eigen::Vector2d p = pf->GetCoarsePosAsVector(); pos2[0] = p(0); pos2[1] = p(1); }
  return f;
}
// MakeSubPixTemplate + IterateSubPixToConvergence from a given coarse position (L0 coords)
int ref_pf_subpix(void* pf_, void* kf, const double* coarse2, int max_its, double* pos2, double* hinv9) {
  PatchFinder* pf = (PatchFinder*)pf_;
  pf->mv2CoarsePos = Eigen::Vector2d(coarse2[0], coarse2[1]);
  pf->MakeSubPixTemplate();
  if (hinv9) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) hinv9[3 * i + j] = pf->mm3HInv(i, j);
  bool ok = pf->IterateSubPixToConvergence(*(KeyFrame*)kf, max_its);
  Eigen::Vector2d p = pf->GetSubPixPos(); pos2[0] = p(0); pos2[1] = p(1);
  return ok;
}

// ------------------------------------------------------------------ MiniPatch (jni/MiniPatch.cc)
void* ref_mp_create() { return new MiniPatch(); }
void ref_mp_destroy(void* mp) { delete (MiniPatch*)mp; }
void ref_mp_set_max_ssd(int v) { MiniPatch::mnMaxSSD = v; }
void ref_mp_sample(void* mp, void* kf, int x, int y, uint8_t* out81) {
  MiniPatch* m = (MiniPatch*)mp;
  m->SampleFromImage(Eigen::Vector2d(x, y), ((KeyFrame*)kf)->aLevels[0].im);
  if (out81) for (int r = 0; r < m->mimOrigPatch.rows; r++) memcpy(out81 + r * m->mimOrigPatch.cols, m->mimOrigPatch.ptr<uint8_t>(r), m->mimOrigPatch.cols);
}
// (MiniPatch::SSDAtPoint is declared inline in MiniPatch.cc: no external symbol to call; FindPatch covers it.)
int ref_mp_find(void* mp, void* kf, double* pos2, int range, int use_lut) {
  Level& L = ((KeyFrame*)kf)->aLevels[0];
  Eigen::Vector2d p(pos2[0], pos2[1]);
  bool f = ((MiniPatch*)mp)->FindPatch(p, L.im, range, L.vCorners, use_lut ? &L.vCornerRowLUT : NULL);
  pos2[0] = p(0); pos2[1] = p(1);
  return f;
}

// ------------------------------------------------------------------ Tracker (jni/Tracker.cc)
void* ref_tracker_create(int w, int h, void* cam, void* map_, int fix) {
  RefTracker* t = new RefTracker();
  t->map = (Map*)map_;
  // MapMaker::MapMaker -> Reset() and Tracker::Tracker -> Reset() -> MapMaker::RequestReset both wipe the map (Map::Reset deletes the
  // points): the caller's map is set aside for the duration and put back afterwards.  The reset request is answered by a helper
  // thread the way MapMaker::run answers it (CHECK_RESET, jni/MapMaker.cc:80), because the reference's own thread is never started.
  std::vector<MapPoint*> points; points.swap(t->map->vpPoints);
  std::vector<KeyFrame*> keyframes; keyframes.swap(t->map->vpKeyFrames);
  const bool good = t->map->bGood;
  t->mm = new MapMaker(*t->map, *(ATANCamera*)cam);
  {
    std::atomic<bool> stop(false);
    MapMaker* mm = t->mm;
    std::thread helper([&stop, mm] { while (!stop.load()) { if (mm->mbResetRequested) mm->Reset(); usleep(20); } });
    t->tr = new Tracker(w, h, *(ATANCamera*)cam, *t->map, *t->mm);
    stop.store(true);
    helper.join();
  }
  t->map->vpPoints.swap(points); t->map->vpKeyFrames.swap(keyframes); t->map->bGood = good;
  t->mm->mdWiggleScale = 1e30; t->mm->mdWiggleScaleDepthNormalized = 1e30;   // keyframe heuristics off until ref_set_keyframe_policy (both members are uninitialised before InitFromStereo)
  fix_radius(t->tr->mCamera, fix);
  t->tr->mbDraw = false;
  return t;
}
void ref_tracker_set_pose(void* t, const double* pose12) { ((RefTracker*)t)->tr->mse3CamFromWorld = pose_from12(pose12); }
void ref_tracker_get_pose(void* t, double* pose12) { pose_to12(((RefTracker*)t)->tr->mse3CamFromWorld, pose12); }
void ref_tracker_set_velocity(void* t, const double* v6, double msd_scaled_mag) {
  Tracker* tr = ((RefTracker*)t)->tr;
  for (int i = 0; i < 6; i++) tr->mv6CameraVelocity_eigen(i) = v6[i];
  tr->mdMSDScaledVelocityMagnitude = msd_scaled_mag;
}
void ref_tracker_get_velocity(void* t, double* v6, double* msd_scaled_mag) {
  Tracker* tr = ((RefTracker*)t)->tr;
  for (int i = 0; i < 6; i++) v6[i] = tr->mv6CameraVelocity_eigen(i);
  *msd_scaled_mag = tr->mdMSDScaledVelocityMagnitude;
}
void ref_tracker_set_scene_depth(void* t, double mean, double sigma) {
  Tracker* tr = ((RefTracker*)t)->tr; tr->mCurrentKF.dSceneDepthMean = mean; tr->mCurrentKF.dSceneDepthSigma = sigma;
}
void ref_tracker_get_scene_depth(void* t, double* mean, double* sigma) {
  Tracker* tr = ((RefTracker*)t)->tr; *mean = tr->mCurrentKF.dSceneDepthMean; *sigma = tr->mCurrentKF.dSceneDepthSigma;
}
void* ref_tracker_current_kf(void* t) { return &((RefTracker*)t)->tr->mCurrentKF; }
void ref_tracker_make_current_kf(void* t, const uint8_t* gray, int w, int h, int stride) {
  Tracker* tr = ((RefTracker*)t)->tr;
  cv::Mat im = wrap_gray(gray, w, h, stride);
  cv::Mat col(1, 1, CV_8UC4, g_dummy_rgba);
  tr->mCurrentKF.mMeasurements.clear();
  tr->mCurrentKF.MakeKeyFrame_Lite(im, col);
}
void ref_tracker_track_map(void* t) {
  Tracker* tr = ((RefTracker*)t)->tr;
  cv::Mat col(1, 1, CV_8UC4, g_dummy_rgba);
  tr->TrackMap(col);
}
void ref_tracker_track_frame(void* t, const uint8_t* gray, int w, int h, int stride) {
  Tracker* tr = ((RefTracker*)t)->tr;
  cv::Mat im = wrap_gray(gray, w, h, stride);
  cv::Mat col(1, 1, CV_8UC4, g_dummy_rgba);
  tr->TrackFrame(im, col, false);
  drain_keyframe_queue((RefTracker*)t);
}
// f3 pins.  MapMaker.cc is not part of this build (it needs Eigen's JacobiSVD / EigenSolver and the Bundle / HomographyInit link
// surface), so the two MapMaker searches are DRIVEN here on the reference's own objects: every arithmetic step below is a call into
// reference code (mySE3 / mySO3 operators, ATANCamera::Project / UnProject / GetProjectionDerivs_Eigen / OnePixelDist,
// PatchFinder::MakeTemplateCoarse / MakeTemplateCoarseNoWarp / FindPatchCoarse / ZMSSDAtPoint / MakeSubPixTemplate /
// IterateSubPixToConvergence, LevelZeroPos); what this harness supplies is the ORDER of those calls, written from the description of
// MapMaker::ReFind_Common (jni/MapMaker.cc:967-1036) and of the search part of MapMaker::AddPointEpipolar (:525-640) in SURVEY.md /
// DESIGN.md.  Bookkeeping (measurement sets, new map points, triangulation) is left out.

namespace {
struct PixelOnPlane { bool ok; Eigen::Vector2d image; };
// z = 1 projection of a camera-frame point, then the camera model with its validity checks
PixelOnPlane to_pixel(ATANCamera& cam, const Eigen::Vector3d& in_camera, int cols, int rows) {
  PixelOnPlane r; r.ok = false;
  if (in_camera(2) < 0.001) return r;
  Eigen::Vector2d plane; plane(0) = in_camera(0) / in_camera(2); plane(1) = in_camera(1) / in_camera(2);
  const double limit = cam.LargestRadiusInImage();
  if (plane.dot(plane) > limit * limit) return r;
  r.image = cam.Project(plane);
  if (cam.Invalid()) return r;
  if (r.image[0] < 0 || r.image[1] < 0 || r.image[0] > cols || r.image[1] > rows) return r;
  r.ok = true;
  return r;
}
}  // namespace

// Re-find listed map points in the tracker's current keyframe (pose = the tracker's pose).  ONE PatchFinder for the whole list, as
// the reference keeps a function-static one.  out3 = {found, search level, refined}, pos2 = measurement position.
void ref_refind(void* t, const int32_t* idx, int n, int range, int subpix_its, int32_t* out3, double* pos2) {
  Tracker* tracker = ((RefTracker*)t)->tr;
  Map* map = ((RefTracker*)t)->map;
  KeyFrame& frame = tracker->mCurrentKF;
  frame.se3CfromW = tracker->mse3CamFromWorld;
  ATANCamera& cam = tracker->mCamera;
  static PatchFinder finder;
  for (int q = 0; q < n; q++) {
    int32_t* o = out3 + 3 * q; double* xy = pos2 + 2 * q;
    o[0] = 0; o[1] = -1; o[2] = 0; xy[0] = xy[1] = 0.0;
    MapPoint& point = *map->vpPoints[idx[q]];
    const PixelOnPlane px = to_pixel(cam, frame.se3CfromW * point.v3WorldPos, frame.aLevels[0].im.cols, frame.aLevels[0].im.rows);
    if (!px.ok) continue;
    Eigen::Matrix2d derivs = cam.GetProjectionDerivs_Eigen();
    finder.MakeTemplateCoarse(point, frame.se3CfromW, derivs);
    o[1] = finder.GetLevel();
    if (finder.TemplateBad() || !finder.FindPatchCoarse(px.image, frame, range)) continue;
    o[0] = 1;
    const bool refine = finder.GetLevel() > 0;
    if (refine) { finder.MakeSubPixTemplate(); finder.IterateSubPixToConvergence(frame, subpix_its); }
    const Eigen::Vector2d where = refine ? finder.GetSubPixPos() : finder.GetCoarsePosAsVector();
    o[2] = refine ? 1 : 0; xy[0] = where(0); xy[1] = where(1);
  }
}

// Epipolar search of one Shi-Tomasi candidate of `source` (after MakeKeyFrame_Rest) in `target`; poses row-major 3x4.
// out3 = {converged match, index of the best corner of target's level (-1: none), its ZMSSD}; pos2 = refined position.
void ref_epipolar_search(void* t, void* source_kf, void* target_kf, const double* src_pose12, const double* tgt_pose12, double depth_mean, double depth_sigma,
                         double wiggle, int level, int candidate_index, int32_t* out3, double* pos2) {
  ATANCamera& cam = ((RefTracker*)t)->tr->mCamera;
  KeyFrame& source = *(KeyFrame*)source_kf; KeyFrame& target = *(KeyFrame*)target_kf;
  source.se3CfromW = pose_from12(src_pose12); target.se3CfromW = pose_from12(tgt_pose12);
  out3[0] = 0; out3[1] = -1; out3[2] = 0; pos2[0] = pos2[1] = 0.0;
  // image-plane position of every integer pixel (the reference caches this table once per image size)
  static Eigen::MatrixXd plane_x, plane_y; static int tw = 0, th = 0;
  const int cols = source.aLevels[0].im.cols, rows = source.aLevels[0].im.rows;
  if (tw != cols || th != rows) {
    tw = cols; th = rows; plane_x.resize(rows, cols); plane_y.resize(rows, cols);
    for (int u = 0; u < cols; u++) for (int v = 0; v < rows; v++) { const Eigen::Vector2d p = cam.UnProject(Eigen::Vector2d(u, v)); plane_x(v, u) = p(0); plane_y(v, u) = p(1); }
  }
  const int scale = LevelScale(level);
  Eigen::Vector2d level_pos = source.aLevels[level].vCandidates[candidate_index].irLevelPos;   // LevelZeroPos takes a mutable reference
  const Eigen::Vector2d root = LevelZeroPos(level_pos, level);
  // viewing ray of the candidate in the source camera, then as a direction in the target camera
  const Eigen::Vector2d ray_plane = cam.UnProject(root);
  Eigen::Vector3d ray; ray(0) = ray_plane(0); ray(1) = ray_plane(1); ray(2) = 1.0; ray.normalize();
  const Eigen::Vector3d dir = target.se3CfromW.get_rotation() * (source.se3CfromW.get_rotation().inverse() * ray);
  const double near_depth = std::max(wiggle, depth_mean - depth_sigma), far_depth = std::min(40 * wiggle, depth_mean + depth_sigma);
  const Eigen::Vector3d origin = target.se3CfromW * source.se3CfromW.inverse().get_translation();
  Eigen::Vector3d near_pt = origin + near_depth * dir;
  const Eigen::Vector3d far_pt = origin + far_depth * dir;
  if (far_pt(2) <= near_pt(2) || far_pt(2) <= 0.0) return;
  if (near_pt(2) <= 0.0) near_pt += dir * (0.001 - near_pt(2) / dir(2));
  Eigen::Vector2d a; a(0) = near_pt(0) / near_pt(2); a(1) = near_pt(1) / near_pt(2);
  Eigen::Vector2d b; b(0) = far_pt(0) / far_pt(2); b(1) = far_pt(1) / far_pt(2);
  Eigen::Vector2d along = a - b;
  if (along.dot(along) < 0.00000001) return;
  along.normalize();
  Eigen::Vector2d across; across(0) = along(1); across(1) = -along(0);
  const double offset = a.dot(across);
  if (fabs(offset) > cam.LargestRadiusInImage()) return;
  double lo = std::min(along.dot(a), along.dot(b)) - 0.05, hi = std::max(along.dot(a), along.dot(b)) + 0.05;
  if (lo < -2.0) lo = -2.0;
  if (hi < -2.0) hi = -2.0;
  if (lo > 2.0) lo = 2.0;
  if (hi > 2.0) hi = 2.0;
  PatchFinder finder;
  finder.MakeTemplateCoarseNoWarp(source, level, (int)level_pos(0), (int)level_pos(1));
  if (finder.TemplateBad()) return;
  std::vector<Eigen::Vector2d>& corners = target.aLevels[level].vCorners;
  const double band = cam.OnePixelDist() * (4.0 + 1.0 * scale), band_sq = band * band;
  int best = -1, best_score = finder.mnMaxSSD + 1;
  for (unsigned int i = 0; i < corners.size(); i++) {
    const Eigen::Vector2d z = LevelZeroPos(corners[i], level);
    const Eigen::Vector2d on_plane(plane_x(z(1), z(0)), plane_y(z(1), z(0)));       // table indexed with the truncated position
    const double off = offset - on_plane.dot(across);
    if (off * off > band_sq) continue;
    if (on_plane.dot(along) < lo) continue;
    if (on_plane.dot(along) > hi) continue;
    const int score = finder.ZMSSDAtPoint(target.aLevels[level].im, (int)corners[i](0), (int)corners[i](1));
    if (score < best_score) { best = (int)i; best_score = score; }
  }
  out3[1] = best; out3[2] = best_score;
  if (best < 0) return;
  finder.MakeSubPixTemplate();
  finder.SetSubPixPos(LevelZeroPos(corners[best], level));
  out3[0] = finder.IterateSubPixToConvergence(target, 10) ? 1 : 0;
  const Eigen::Vector2d refined = finder.GetSubPixPos();
  pos2[0] = refined(0); pos2[1] = refined(1);
}
// Patch-source fields of a map point created from a candidate of `source` (as the tail of MapMaker::AddPointEpipolar fills them,
// jni/MapMaker.cc:655-684), then the reference's own MapPoint::RefreshPixelVectors.  out15 = centre ray, one-right ray, one-down
// ray (source camera frame, unit length), pixel-right and pixel-down vectors (world frame).
void ref_epipolar_point_fields(void* t, void* source_kf, const double* src_pose12, int level, int candidate_index, const double* world3, double* out15) {
  ATANCamera& cam = ((RefTracker*)t)->tr->mCamera;
  KeyFrame& source = *(KeyFrame*)source_kf;
  source.se3CfromW = pose_from12(src_pose12);
  Eigen::Vector2d level_pos = source.aLevels[level].vCandidates[candidate_index].irLevelPos;
  const Eigen::Vector2d root = LevelZeroPos(level_pos, level);
  const double step = LevelScale(level);
  MapPoint point;
  point.pPatchSourceKF = &source; point.nSourceLevel = level;
  point.v3WorldPos = Eigen::Vector3d(world3[0], world3[1], world3[2]);
  point.v3Normal_NC = Eigen::Vector3d(0, 0, -1);
  Eigen::Vector3d* rays[3] = {&point.v3Center_NC, &point.v3OneRightFromCenter_NC, &point.v3OneDownFromCenter_NC};
  const Eigen::Vector2d at[3] = {root, root + Eigen::Vector2d(step, 0), root + Eigen::Vector2d(0, step)};
  for (int k = 0; k < 3; k++) { const Eigen::Vector2d p = cam.UnProject(at[k]); (*rays[k])(0) = p(0); (*rays[k])(1) = p(1); (*rays[k])(2) = 1.0; rays[k]->normalize(); }
  point.RefreshPixelVectors();
  for (int k = 0; k < 3; k++) for (int q = 0; q < 3; q++) out15[3 * k + q] = (*rays[k])(q);
  for (int q = 0; q < 3; q++) { out15[9 + q] = point.v3PixelRight_W(q); out15[12 + q] = point.v3PixelDown_W(q); }
}
void ref_set_keyframe_policy(void* t, int enable, double wiggle, double wiggle_dn, double mult) {
  // the reference's own MapMaker::NeedNewKeyFrame / IsDistanceToNearestKeyFrameExcessive (jni/MapMaker.cc:763-773,1098-1101) compare against
  // these two members; MaxKFDistWiggleMult is the constant 0.2 of :768 (`mult` must be that value)
  MapMaker* mm = ((RefTracker*)t)->mm;
  assert(!enable || mult == 0.2); (void)mult;
  mm->mdWiggleScale = enable ? wiggle : 1e30; mm->mdWiggleScaleDepthNormalized = enable ? wiggle_dn : 1e30;
}

// ------------------------------------------------------------------ the reference's own MapMaker functions (jni/MapMaker.cc), called directly
static void ensure_mm_data(MapPoint& p) { if (!p.pMMData) p.pMMData = new MapMakerData(); }
// MapMaker::ReFind_Common (jni/MapMaker.cc:967-1036) of the listed map points in the tracker's current keyframe (pose = the tracker's pose).
// out4 = {found, measurement level, sub-pixel, never-retry}; pos2 = Measurement::v2RootPos.  The bookkeeping the call leaves behind is undone.
void ref_mm_refind(void* t, const int32_t* idx, int n, int32_t* out4, double* pos2) {
  RefTracker* r = (RefTracker*)t;
  KeyFrame& frame = r->tr->mCurrentKF;
  frame.se3CfromW = r->tr->mse3CamFromWorld;
  for (int q = 0; q < n; q++) {
    MapPoint& p = *r->map->vpPoints[idx[q]];
    ensure_mm_data(p);
    p.pMMData->sMeasurementKFs.erase(&frame); p.pMMData->sNeverRetryKFs.erase(&frame); frame.mMeasurements.erase(&p);
    const bool found = r->mm->ReFind_Common(frame, p);
    int32_t* o = out4 + 4 * q; o[0] = found; o[1] = -1; o[2] = 0; o[3] = (int)p.pMMData->sNeverRetryKFs.count(&frame);
    pos2[2 * q] = pos2[2 * q + 1] = 0.0;
    if (found) {
      const Measurement& m = frame.mMeasurements[&p];
      o[1] = m.nLevel; o[2] = m.bSubPix; pos2[2 * q] = m.v2RootPos(0); pos2[2 * q + 1] = m.v2RootPos(1);
    }
    p.pMMData->sMeasurementKFs.erase(&frame); p.pMMData->sNeverRetryKFs.erase(&frame); frame.mMeasurements.erase(&p);
  }
}
// MapMaker::AddPointEpipolar (jni/MapMaker.cc:525-703) for candidate `candidate_index` of `source`'s level (after MakeKeyFrame_Rest) against
// `target`.  Returns its result; on success out27 = the new MapPoint: world position, centre / one-right / one-down rays, pixel-right /
// pixel-down vectors (6 x 3), irCenter (2), source level, the two measurements' root positions (source, target: 2 + 2), 2 spare.  The
// point is taken out of the map again (the caller's map stays as it was).
int ref_mm_add_point_epipolar(void* t, void* source_kf, void* target_kf, const double* src_pose12, const double* tgt_pose12, double depth_mean, double depth_sigma,
                              double wiggle, int level, int candidate_index, double* out27) {
  RefTracker* r = (RefTracker*)t;
  KeyFrame& source = *(KeyFrame*)source_kf; KeyFrame& target = *(KeyFrame*)target_kf;
  source.se3CfromW = pose_from12(src_pose12); target.se3CfromW = pose_from12(tgt_pose12);
  source.dSceneDepthMean = depth_mean; source.dSceneDepthSigma = depth_sigma;
  const double saved = r->mm->mdWiggleScale; r->mm->mdWiggleScale = wiggle;
  const size_t before = r->map->vpPoints.size();
  const bool ok = r->mm->AddPointEpipolar(source, target, level, candidate_index);
  r->mm->mdWiggleScale = saved;
  for (int k = 0; k < 27; k++) out27[k] = 0.0;
  if (!ok) return 0;
  assert(r->map->vpPoints.size() == before + 1);
  MapPoint* p = r->map->vpPoints.back();
  const Eigen::Vector3d* v[6] = {&p->v3WorldPos, &p->v3Center_NC, &p->v3OneRightFromCenter_NC, &p->v3OneDownFromCenter_NC, &p->v3PixelRight_W, &p->v3PixelDown_W};
  for (int k = 0; k < 6; k++) for (int c = 0; c < 3; c++) out27[3 * k + c] = (*v[k])(c);
  out27[18] = p->irCenter(0); out27[19] = p->irCenter(1); out27[20] = p->nSourceLevel;
  out27[21] = source.mMeasurements[p].v2RootPos(0); out27[22] = source.mMeasurements[p].v2RootPos(1);
  out27[23] = target.mMeasurements[p].v2RootPos(0); out27[24] = target.mMeasurements[p].v2RootPos(1);
  source.mMeasurements.erase(p); target.mMeasurements.erase(p);
  r->map->vpPoints.pop_back();
  while (!r->mm->mqNewQueue.empty()) r->mm->mqNewQueue.pop();
  delete p->pMMData; delete p;
  return 1;
}
// MapMaker::ReprojectPoint (jni/MapMaker.cc:176-200): point in frame B from the two z = 1 projections
void ref_mm_reproject_point(void* t, const double* a_from_b12, const double* plane_a2, const double* plane_b2, double* out3) {
  RefTracker* r = (RefTracker*)t;
  const Eigen::Vector3d p = r->mm->ReprojectPoint(pose_from12(a_from_b12), Eigen::Vector2d(plane_a2[0], plane_a2[1]), Eigen::Vector2d(plane_b2[0], plane_b2[1]));
  for (int c = 0; c < 3; c++) out3[c] = p(c);
}
// MapMaker::NeedNewKeyFrame / IsDistanceToNearestKeyFrameExcessive / DistToNearestKeyFrame on the tracker's current keyframe and pose
void ref_mm_keyframe_heuristics(void* t, int* need_new, int* excessive, double* dist) {
  RefTracker* r = (RefTracker*)t;
  KeyFrame& frame = r->tr->mCurrentKF;
  frame.se3CfromW = r->tr->mse3CamFromWorld;
  *need_new = r->mm->NeedNewKeyFrame(frame); *excessive = r->mm->IsDistanceToNearestKeyFrameExcessive(frame); *dist = r->mm->DistToNearestKeyFrame(frame);
}
void ref_keyframe_info(void* t, int* n_keyframes, int* added_total, int* n_frame, int* last_dropped) {
  RefTracker* r = (RefTracker*)t;
  *n_keyframes = (int)r->map->vpKeyFrames.size(); *added_total = g_keyframes_added; *n_frame = r->tr->mnFrame; *last_dropped = r->tr->mnLastKeyFrameDropped;
}
int ref_kf_num_candidates_l(void* kf, int l) { return (int)((KeyFrame*)kf)->aLevels[l].vCandidates.size(); }

// Relocaliser support: a map keyframe gets its SmallBlurryImage the way KeyFrame::MakeKeyFrame_Rest (jni/KeyFrame.cc:98) and the
// map maker (MakeJacs before the keyframe is used as an alignment target) leave it
void ref_kf_make_sbi(void* kf_) {
  KeyFrame* kf = (KeyFrame*)kf_;
  if (kf->pSBI) delete kf->pSBI;
  kf->pSBI = new SmallBlurryImage(*kf);
  kf->pSBI->MakeJacs();
}
void ref_tracker_set_lost(void* t, int lost_frames, int quality) { Tracker* tr = ((RefTracker*)t)->tr; tr->mnLostFrames = lost_frames; tr->mTrackingQuality = (decltype(tr->mTrackingQuality))quality; }

// Trail tracking for the initial map (jni/Tracker.cc:264-346) on the tracker's current keyframe (ref_tracker_make_current_kf first)
int ref_tracker_trail_start(void* t) {
  Tracker* tr = ((RefTracker*)t)->tr;
  tr->mlTrails.clear();
  tr->TrailTracking_Start();
  return (int)tr->mlTrails.size();
}
int ref_tracker_trail_advance(void* t, int max_ssd) {
  Tracker* tr = ((RefTracker*)t)->tr;
  MiniPatch::mnMaxSSD = max_ssd;
  cv::Mat col(1, 1, CV_8UC4, g_dummy_rgba);
  return tr->TrailTracking_Advance(col);
}
int ref_tracker_trail_count(void* t) { return (int)((RefTracker*)t)->tr->mlTrails.size(); }
void ref_tracker_trails(void* t, double* init_cur4) {
  Tracker* tr = ((RefTracker*)t)->tr;
  int k = 0;
  for (std::list<Trail>::iterator i = tr->mlTrails.begin(); i != tr->mlTrails.end(); ++i, ++k) {
    init_cur4[4 * k] = i->irInitialPos(0); init_cur4[4 * k + 1] = i->irInitialPos(1); init_cur4[4 * k + 2] = i->irCurrentPos(0); init_cur4[4 * k + 3] = i->irCurrentPos(1);
  }
}
// Tracker::TrackFrame's good-map branch (jni/Tracker.cc:76-112) driven piece by piece WITHOUT the SmallBlurryImage steps
// (jni/Tracker.cc:86-97,105-106: the f1 "next" row of SURVEY.md §8): the SBI rotation is whatever ref_tracker_set_sbi_rot set.
void ref_tracker_track_frame_nosbi(void* t, const uint8_t* gray, int w, int h, int stride) {
  Tracker* tr = ((RefTracker*)t)->tr;
  cv::Mat im = wrap_gray(gray, w, h, stride);
  cv::Mat col(1, 1, CV_8UC4, g_dummy_rgba);
  tr->mCurrentKF.mMeasurements.clear();
  tr->mCurrentKF.MakeKeyFrame_Lite(im, col);
  tr->mnFrame++;
  if (tr->mnLostFrames < 3) {
    tr->ApplyMotionModel();
    tr->TrackMap(col);
    tr->UpdateMotionModel();
    tr->AssessTrackingQuality();
  }
}
void ref_tracker_motion_model(void* t, int apply_not_update) {
  Tracker* tr = ((RefTracker*)t)->tr;
  if (apply_not_update) tr->ApplyMotionModel(); else tr->UpdateMotionModel();
}
void ref_tracker_set_sbi_rot(void* t, const double* v6, int use_sbi) {
  Tracker* tr = ((RefTracker*)t)->tr;
  for (int i = 0; i < 6; i++) tr->mv6SBIRot_eigen(i) = v6[i];
  tr->mbUseSBIInit = use_sbi != 0;
}
void ref_tracker_get_sbi_rot(void* t, double* v6) { Tracker* tr = ((RefTracker*)t)->tr; for (int i = 0; i < 6; i++) v6[i] = tr->mv6SBIRot_eigen(i); }
void ref_tracker_counters(void* t, int* attempted4, int* found4, int* quality, int* lost, int* did_coarse) {
  Tracker* tr = ((RefTracker*)t)->tr;
  for (int i = 0; i < LEVELS; i++) { attempted4[i] = tr->manMeasAttempted[i]; found4[i] = tr->manMeasFound[i]; }
  *quality = (int)tr->mTrackingQuality; *lost = tr->mnLostFrames; *did_coarse = tr->mbDidCoarse;
}
int ref_tracker_message(void* t, char* buf, int cap) {
  std::string s = ((RefTracker*)t)->tr->GetMessageForUser();
  int n = (int)s.size() < cap - 1 ? (int)s.size() : cap - 1;
  memcpy(buf, s.data(), n); buf[n] = 0; return (int)s.size();
}

// The first loop of TrackMap (jni/Tracker.cc:369-392) on its own: Project / GetDerivsUnsafe /
// CalcSearchLevelAndWarpMatrix for every map point, with the tracker's current pose.
void ref_tracker_project_all(void* t) {
  Tracker* tr = ((RefTracker*)t)->tr;
  for (unsigned i = 0; i < tr->mMap.vpPoints.size(); i++) {
    MapPoint& p = *tr->mMap.vpPoints[i];
    if (!p.pTData) p.pTData = new TrackerData(&p);
    TrackerData& TD = *p.pTData;
    TD.nSearchLevel = -1; TD.bSearched = false; TD.bFound = false; TD.bDidSubPix = false;
    TD.Project(tr->mse3CamFromWorld, tr->mCamera);
    if (!TD.bInImage) continue;
    TD.GetDerivsUnsafe(tr->mCamera);
    TD.nSearchLevel = TD.Finder.CalcSearchLevelAndWarpMatrix(TD.Point, tr->mse3CamFromWorld, TD.m2CamDerivs);
  }
}
// Per-point TrackerData dump.  ints[8]: inImage, searchLevel, searched, found, didSubPix, templateBad, hasTData, pad
// dbl[32]: v2Image(2) v2Found(2) derivs(4 row-major) v3Cam(3) warpInv(4 row-major) sqrtInvNoise(1) jac(12 row-major 2x6) err(2) coarse(2)
void ref_tracker_point_state(void* t, int i, int32_t* ints, double* dbl) {
  Tracker* tr = ((RefTracker*)t)->tr;
  MapPoint& p = *tr->mMap.vpPoints[i];
  memset(ints, 0, 8 * sizeof(int32_t)); memset(dbl, 0, 32 * sizeof(double));
  if (!p.pTData) return;
  TrackerData& TD = *p.pTData;
  ints[0] = TD.bInImage; ints[1] = TD.nSearchLevel; ints[2] = TD.bSearched; ints[3] = TD.bFound; ints[4] = TD.bDidSubPix;
  ints[5] = TD.Finder.mbTemplateBad; ints[6] = 1;
  dbl[0] = TD.v2Image(0); dbl[1] = TD.v2Image(1); dbl[2] = TD.v2Found(0); dbl[3] = TD.v2Found(1);
  dbl[4] = TD.m2CamDerivs(0, 0); dbl[5] = TD.m2CamDerivs(0, 1); dbl[6] = TD.m2CamDerivs(1, 0); dbl[7] = TD.m2CamDerivs(1, 1);
  dbl[8] = TD.v3Cam(0); dbl[9] = TD.v3Cam(1); dbl[10] = TD.v3Cam(2);
  const Eigen::Matrix2d& w = TD.Finder.mm2WarpInverse;
  dbl[11] = w(0, 0); dbl[12] = w(0, 1); dbl[13] = w(1, 0); dbl[14] = w(1, 1);
  dbl[15] = TD.dSqrtInvNoise;
  for (int r = 0; r < 2; r++) for (int c = 0; c < 6; c++) dbl[16 + 6 * r + c] = TD.m26Jacobian(r, c);
  dbl[28] = TD.v2Error_CovScaled(0); dbl[29] = TD.v2Error_CovScaled(1);
  dbl[30] = TD.Finder.mv2CoarsePos(0); dbl[31] = TD.Finder.mv2CoarsePos(1);
}
void ref_tracker_point_template(void* t, int i, uint8_t* tmpl, int* sum, int* sumsq) {
  Tracker* tr = ((RefTracker*)t)->tr;
  PatchFinder& F = tr->mMap.vpPoints[i]->pTData->Finder;
  for (int r = 0; r < F.mnPatchSize; r++) memcpy(tmpl + r * F.mnPatchSize, F.mimTemplate.ptr<uint8_t>(r), F.mnPatchSize);
  *sum = F.mnTemplateSum; *sumsq = F.mnTemplateSumSq;
}
// SearchForPoints over an explicit index list (jni/Tracker.cc:629-674).  Counters are NOT cleared.
int ref_tracker_search_for_points(void* t, const int32_t* idx, int n, int range, int subpix_its) {
  Tracker* tr = ((RefTracker*)t)->tr;
  std::vector<TrackerData*> v;
  for (int i = 0; i < n; i++) v.push_back(tr->mMap.vpPoints[idx[i]]->pTData);
  return tr->SearchForPoints(v, range, subpix_its);
}
void ref_tracker_clear_counters(void* t) {
  Tracker* tr = ((RefTracker*)t)->tr;
  for (int i = 0; i < LEVELS; i++) tr->manMeasAttempted[i] = tr->manMeasFound[i] = 0;
}
void ref_tracker_calc_jacobians(void* t, const int32_t* idx, int n) {
  Tracker* tr = ((RefTracker*)t)->tr;
  for (int i = 0; i < n; i++) { TrackerData* TD = tr->mMap.vpPoints[idx[i]]->pTData; if (TD->bFound) TD->CalcJacobian(); }
}
void ref_tracker_project_and_derivs(void* t, const int32_t* idx, int n, int only_found) {
  Tracker* tr = ((RefTracker*)t)->tr;
  for (int i = 0; i < n; i++) { TrackerData* TD = tr->mMap.vpPoints[idx[i]]->pTData; if (!only_found || TD->bFound) TD->ProjectAndDerivs(tr->mse3CamFromWorld, tr->mCamera); }
}
void ref_tracker_linear_update(void* t, const int32_t* idx, int n, const double* v6) {
  Tracker* tr = ((RefTracker*)t)->tr;
  Eigen::VectorXd v(6); for (int i = 0; i < 6; i++) v(i) = v6[i];
  for (int i = 0; i < n; i++) { TrackerData* TD = tr->mMap.vpPoints[idx[i]]->pTData; if (TD->bFound) TD->LinearUpdate(v); }
}
// CalcPoseUpdate (jni/Tracker.cc:683-774); optionally applies pose = exp(update) * pose like TrackMap does.
void ref_tracker_calc_pose_update(void* t, const int32_t* idx, int n, double override_sigma, int mark_outliers, int apply, double* out6) {
  Tracker* tr = ((RefTracker*)t)->tr;
  std::vector<TrackerData*> v;
  for (int i = 0; i < n; i++) v.push_back(tr->mMap.vpPoints[idx[i]]->pTData);
  Eigen::VectorXd u = tr->CalcPoseUpdate(v, override_sigma, mark_outliers != 0);
  for (int i = 0; i < 6; i++) out6[i] = u(i);
  if (apply) tr->mse3CamFromWorld = mySE3::exp(u) * tr->mse3CamFromWorld;
}
double ref_tukey_sigma_squared(const double* err_sq, int n) {
  std::vector<double> v(err_sq, err_sq + n);
  return Tukey::FindSigmaSquared(v);
}
// Measurements written by TrackMap (jni/Tracker.cc:594-607), in map-point order: found flag per point.
int ref_tracker_num_measurements(void* t) { return (int)((RefTracker*)t)->tr->mCurrentKF.mMeasurements.size(); }

// SmallBlurryImage pieces (jni/SmallBlurryImage.cc) — the f1 "next" row.
void* ref_sbi_create(void* kf, double blur) { return new SmallBlurryImage(*(KeyFrame*)kf, blur); }
void ref_sbi_destroy(void* s) { delete (SmallBlurryImage*)s; }
void ref_sbi_dims(int* w, int* h) { *w = (int)SmallBlurryImage::mirSize(0); *h = (int)SmallBlurryImage::mirSize(1); }
void ref_sbi_reset_size() { SmallBlurryImage::mirSize = Eigen::Vector2d(-1, -1); }
void ref_sbi_template(void* s_, float* out) {
  SmallBlurryImage* s = (SmallBlurryImage*)s_;
  for (int y = 0; y < s->mimTemplate.rows; y++) memcpy(out + (size_t)y * s->mimTemplate.cols, s->mimTemplate.ptr<float>(y), sizeof(float) * s->mimTemplate.cols);
}
void ref_sbi_small(void* s_, uint8_t* out) {
  SmallBlurryImage* s = (SmallBlurryImage*)s_;
  for (int y = 0; y < s->mimSmall.rows; y++) memcpy(out + (size_t)y * s->mimSmall.cols, s->mimSmall.ptr<uint8_t>(y), s->mimSmall.cols);
}
// this.IteratePosRelToTarget(other) then SE3fromSE2(...).ln()  ==  Tracker::CalcSBIRotation (jni/Tracker.cc:885-893)
double ref_sbi_rotation(void* this_, void* other_, void* cam, int its, double* se2_3, double* v6) {
  SmallBlurryImage* a = (SmallBlurryImage*)this_; SmallBlurryImage* b = (SmallBlurryImage*)other_;
  b->MakeJacs();
  std::pair<mySE2, double> r = a->IteratePosRelToTarget(*b, its);
  if (se2_3) { se2_3[0] = r.first.get_translation()(0); se2_3[1] = r.first.get_translation()(1);
    se2_3[2] = atan2(r.first.get_rotation().get_matrix()(1, 0), r.first.get_rotation().get_matrix()(0, 0)); }
  mySE3 adj = SmallBlurryImage::SE3fromSE2(r.first, *(ATANCamera*)cam);
  Eigen::VectorXd l = adj.ln();
  for (int i = 0; i < 6; i++) v6[i] = l(i);
  return r.second;
}

}  // extern "C"

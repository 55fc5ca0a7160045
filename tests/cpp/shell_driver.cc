// Test driver for include/vslam_b200_shell.hpp: runs the reference-shaped C++ API (KeyFrame, Tracker, MiniPatch, PatchFinder) on a
// scene file written by tests/test_gpu_shell.py and prints the results as text; the Python side compares them with the oracle.
//   shell_driver <scene.bin> trails|track|reftypes|handoff|stages|mapsearch [map file to write, mapsearch only]
// scene.bin: int32 W,H,N,F; double params5[5]; u8 src[W*H]; double world[3N], right[3N], down[3N]; int32 irCenter[2N]; int32 level[N];
//            double pose0[12]; u8 frames[F][W*H]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include "vslam_b200_shell.hpp"

using namespace vslam_b200;

// Types shaped like the reference's (jni/ATANCamera.h, jni/MapPoint.h, jni/KeyFrame.h, jni/Map.h, jni/MapMaker.h), as far as the tracker reads
// them: what jni/jni_part.cpp:27-46 hands to `new Tracker(width, height, *mpCamera, *mpMap, *mpMapMaker)`.
namespace ref_like {
struct ATANCamera { explicit ATANCamera(const std::string&) : mvDefaultParams(5) {} Eigen::VectorXd mvDefaultParams; };
struct KeyFrame { Level aLevels[LEVELS]; };
struct MapPoint { Eigen::Vector3d v3WorldPos, v3PixelDown_W, v3PixelRight_W; Eigen::Vector2d irCenter; int nSourceLevel; KeyFrame* pPatchSourceKF; bool bBad; };
struct Map { Map() : bGood(false) {} bool IsGood() { return bGood; } std::vector<MapPoint*> vpPoints; std::vector<KeyFrame*> vpKeyFrames; bool bGood; };
struct MapMaker { MapMaker(Map& m, const ATANCamera&) : mMap(m) {} Map& mMap; };
}  // namespace ref_like

template <class T> static void rd(std::ifstream& f, T* p, size_t n) { f.read((char*)p, sizeof(T) * n); if (!f) { fprintf(stderr, "short scene file\n"); exit(2); } }

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  int32_t hdr[4]; rd(f, hdr, 4);
  const int W = hdr[0], H = hdr[1], N = hdr[2], F = hdr[3];
  double p5[5]; rd(f, p5, 5);
  std::vector<unsigned char> src((size_t)W * H); rd(f, &src[0], src.size());
  std::vector<double> world(3 * (size_t)N), right(3 * (size_t)N), down(3 * (size_t)N); rd(f, &world[0], world.size()); rd(f, &right[0], right.size()); rd(f, &down[0], down.size());
  std::vector<int32_t> irc(2 * (size_t)N), lvl(N), kf0(N, 0); rd(f, &irc[0], irc.size()); rd(f, &lvl[0], lvl.size());
  double pose0[12]; rd(f, pose0, 12);
  std::vector<std::vector<unsigned char> > frames(F, std::vector<unsigned char>((size_t)W * H));
  for (int k = 0; k < F; k++) rd(f, &frames[k][0], frames[k].size());
  const std::string mode = argv[2];
  double cam[13], cam_sbi[13];
  vslam_camera_from_params(p5, W, H, 0, cam);
  vslam_camera_from_params(p5, W / 16, H / 16, 0, cam_sbi);
  cv::Mat colour(1, 1, CV_8UC4);
  try {
    if (mode == "reftypes") {   // the reference's constructor shape on reference-shaped types, map read from Map::vpPoints (jni/jni_part.cpp:27-46)
      ref_like::ATANCamera* mpCamera = new ref_like::ATANCamera("Camera");
      for (int k = 0; k < 5; k++) mpCamera->mvDefaultParams(k) = p5[k];
      ref_like::Map* mpMap = new ref_like::Map;
      ref_like::MapMaker* mpMapMaker = new ref_like::MapMaker(*mpMap, *mpCamera);
      typedef TrackerOnReferenceTypes<ref_like::ATANCamera, ref_like::Map, ref_like::MapMaker> RefTracker;
      RefTracker* mpTracker = new RefTracker(W, H, *mpCamera, *mpMap, *mpMapMaker, N > 0 ? N : 1, 4);
      ref_like::KeyFrame* kf = new ref_like::KeyFrame;
      kf->aLevels[0].im = cv::Mat(H, W, CV_8UC1, &src[0]);
      mpMap->vpKeyFrames.push_back(kf);
      const int n_first = N - N / 4;                      // the last quarter of the map arrives later, like points a map maker adds while tracking
      for (int k = 0; k < N; k++) {
        ref_like::MapPoint* p = new ref_like::MapPoint;
        p->v3WorldPos = Eigen::Vector3d(world[3 * k], world[3 * k + 1], world[3 * k + 2]);
        p->v3PixelRight_W = Eigen::Vector3d(right[3 * k], right[3 * k + 1], right[3 * k + 2]);
        p->v3PixelDown_W = Eigen::Vector3d(down[3 * k], down[3 * k + 1], down[3 * k + 2]);
        p->irCenter = Eigen::Vector2d(irc[2 * k], irc[2 * k + 1]); p->nSourceLevel = lvl[k]; p->pPatchSourceKF = kf; p->bBad = false;
        if (k < n_first) mpMap->vpPoints.push_back(p); else delete p;
      }
      mpMap->bGood = true;
      SE3 start0; for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) start0.R(i, j) = pose0[4 * i + j]; start0.t(i) = pose0[4 * i + 3]; }
      mpTracker->SyncMap();
      mpTracker->SetCurrentPose(start0);
      for (int k = 0; k < F; k++) {
        if (k == F / 2) for (int q = n_first; q < N; q++) {
          ref_like::MapPoint* p = new ref_like::MapPoint;
          p->v3WorldPos = Eigen::Vector3d(world[3 * q], world[3 * q + 1], world[3 * q + 2]);
          p->v3PixelRight_W = Eigen::Vector3d(right[3 * q], right[3 * q + 1], right[3 * q + 2]);
          p->v3PixelDown_W = Eigen::Vector3d(down[3 * q], down[3 * q + 1], down[3 * q + 2]);
          p->irCenter = Eigen::Vector2d(irc[2 * q], irc[2 * q + 1]); p->nSourceLevel = lvl[q]; p->pPatchSourceKF = kf; p->bBad = false;
          mpMap->vpPoints.push_back(p);
        }
        cv::Mat g(H, W, CV_8UC1, &frames[k][0]);
        mpTracker->TrackFrame(g, colour, false);
        const SE3 p = mpTracker->GetCurrentPose();
        printf("pose");
        for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) printf(" %.17g", p.R(i, j)); printf(" %.17g", p.t(i)); }
        printf("\nmsg %s\n", mpTracker->GetMessageForUser().c_str());
      }
      // SystemPTAM::onTouchScreen (jni/jni_part.cpp:49-51) writes the member; a C-ABI binding posts the event instead
      mpTracker->mbUserPressedSpacebar = true;
      printf("spacebar %d\n", mpTracker->mbUserPressedSpacebar ? 1 : 0);
      delete mpTracker;
      return 0;
    }
    Context ctx(W, H, 1, N > 0 ? N : 1, 11, 0, 4);
    Tracker tracker(ctx, 0, cam);
    SE3 start; for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) start.R(i, j) = pose0[4 * i + j]; start.t(i) = pose0[4 * i + 3]; }

    if (mode == "trails") {   // Tracker::TrackForInitialMap over the frames: spacebar on frame 0
      for (int k = 0; k < F; k++) {
        cv::Mat g(H, W, CV_8UC1, &frames[k][0]);
        if (k == 0) { tracker.TrackFrame(g, colour, false); printf("msg %s", tracker.GetMessageForUser().c_str()); tracker.PressSpacebar(); }
        if (k == F - 1) tracker.PressSpacebar();
        tracker.TrackFrame(g, colour, false);
        printf("frame %d stage %d trails %d\n", k, tracker.mnInitialStage, (int)tracker.mlTrails.size());
        for (std::list<Trail>::iterator i = tracker.mlTrails.begin(); i != tracker.mlTrails.end(); ++i)
          printf("t %.17g %.17g %.17g %.17g\n", i->irInitialPos(0), i->irInitialPos(1), i->irCurrentPos(0), i->irCurrentPos(1));
      }
      printf("matches %d\n", (int)tracker.vInitMatches.size());
      // the per-object MiniPatch path on the first surviving trail: sample in the current frame, find it again in the same frame
      if (!tracker.mlTrails.empty()) {
        MiniPatch mp; Eigen::Vector2d pos = tracker.mlTrails.front().irCurrentPos;
        mp.SampleFromImage(pos, tracker.mCurrentKF);
        Eigen::Vector2d q = pos; const bool ok = mp.FindPatch(q, tracker.mCurrentKF, 10);
        printf("minipatch %d %.17g %.17g %.17g %.17g\n", ok ? 1 : 0, pos(0), pos(1), q(0), q(1));
      }
      return 0;
    }

    cv::Mat s(H, W, CV_8UC1, &src[0]);
    tracker.SetSourceKeyFrame(0, s);
    tracker.SetMap(N, &world[0], &right[0], &down[0], &irc[0], &lvl[0], &kf0[0]);
    tracker.SetCurrentPose(start);

    if (mode == "track") {    // Tracker::TrackFrame with the map good (SmallBlurryImage on, like the reference)
      tracker.EnableSBI(cam_sbi);
      for (int k = 0; k < F; k++) {
        cv::Mat g(H, W, CV_8UC1, &frames[k][0]);
        tracker.TrackFrame(g, colour, false);
        const SE3 p = tracker.GetCurrentPose();
        printf("pose");
        for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) printf(" %.17g", p.R(i, j)); printf(" %.17g", p.t(i)); }
        printf("\nmsg %s\n", tracker.GetMessageForUser().c_str());
      }
      return 0;
    }

    if (mode == "handoff") {  // TrackFrame with the keyframe policy: the shell adds keyframes when the device asks for one (jni/Tracker.cc:127-132)
      tracker.EnableSBI(cam_sbi);
      const int32_t id0 = 0; double eye12[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
      tracker.SetRelocKeyFrames(1, &id0, eye12);
      tracker.SetKeyFramePolicy(1, 0.1, 0.1);
      for (int k = 0; k < F; k++) {
        cv::Mat g(H, W, CV_8UC1, &frames[k][0]);
        tracker.TrackFrame(g, colour, false);
        printf("msg %s\n", tracker.GetMessageForUser().c_str());
        printf("kf %d %d\n", tracker.mnKeyFrames, tracker.mnLastKeyFrameDropped);
      }
      return 0;
    }

    if (mode == "mapsearch") {   // MapMaker's searches: frames[0] is the target keyframe at pose0; the source keyframe is keyframe 0 (identity pose)
      cv::Mat g(H, W, CV_8UC1, &frames[0][0]);
      MapSearch ms(ctx);
      // candidates of the source keyframe: make it the stream's keyframe once to run MakeKeyFrame_Rest on it
      tracker.mCurrentKF.MakeKeyFrame_Lite(s, colour);
      tracker.mCurrentKF.MakeKeyFrame_Rest();
      std::vector<std::vector<Eigen::Vector2d> > cands(LEVELS);
      for (int l = 0; l < LEVELS; l++) for (size_t i = 0; i < tracker.mCurrentKF.aLevels[l].vCandidates.size(); i++) cands[l].push_back(tracker.mCurrentKF.aLevels[l].vCandidates[i].irLevelPos);
      tracker.mCurrentKF.MakeKeyFrame_Lite(g, colour);
      SE3 eye; for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) eye.R(i, j) = i == j ? 1.0 : 0.0; eye.t(i) = 0.0; }
      std::vector<int> pts; for (int pt = 0; pt < N; pt++) pts.push_back(pt);
      std::vector<Measurement> meas; std::vector<char> found;
      const int nf = ms.ReFindInSingleKeyFrame(tracker.mCurrentKF, start, pts, meas, found);
      printf("refind %d\n", nf);
      for (int pt = 0; pt < N; pt++) if (found[pt]) printf("m %d %d %d %.17g %.17g\n", pt, meas[pt].nLevel, meas[pt].bSubPix ? 1 : 0, meas[pt].v2RootPos(0), meas[pt].v2RootPos(1));
      for (int l = 0; l < LEVELS; l++) {
        const int na = ms.AddPointsEpipolar(0, eye, 1.0, 0.3, tracker.mCurrentKF, start, l, cands[l], meas, found);
        printf("epipolar %d %d of %d\n", l, na, (int)cands[l].size());
        for (size_t i = 0; i < cands[l].size(); i++) if (found[i]) printf("e %d %d %d %.17g %.17g\n", l, (int)cands[l][i](0), (int)cands[l][i](1), meas[i].v2RootPos(0), meas[i].v2RootPos(1));
        std::vector<MapSearch::NewMapPoint> np; std::vector<int> from;
        ms.MakeEpipolarPoints(eye, start, l, cands[l], meas, found, np, from);
        for (size_t k = 0; k < np.size(); k++)
          printf("p %d %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", l, (int)np[k].irCenter(0), (int)np[k].irCenter(1), np[k].v3WorldPos(0), np[k].v3WorldPos(1), np[k].v3WorldPos(2),
                 np[k].v3PixelRight_W(0), np[k].v3PixelRight_W(1), np[k].v3PixelRight_W(2), np[k].v3PixelDown_W(0), np[k].v3PixelDown_W(1), np[k].v3PixelDown_W(2));
      }
      if (argc > 3) {   // map file round trip through the shell: save, load into a second context, compare what it reports
        ctx.SaveMap(argv[3]);
        vslam_map_file_info_t info;
        if (vslam_map_file_info(argv[3], &info) != VSLAM_OK) return 4;
        Context other(W, H, 1, N > 0 ? N : 1);
        other.LoadMap(argv[3]);
        printf("mapfile %d %d %d %d\n", info.n_points, info.n_keyframes, info.n_reloc_keyframes, other.MapSize());
      }
      return 0;
    }

    if (mode == "stages") {   // the protected stage functions and the per-object PatchFinder on frame 0
      cv::Mat g(H, W, CV_8UC1, &frames[0][0]);
      tracker.mCurrentKF.MakeKeyFrame_Lite(g, colour);
      PatchFinder finder(ctx, 0);
      for (int pt = 0; pt < N; pt += N / 12 > 0 ? N / 12 : 1) {
        const int level = finder.CalcSearchLevelAndWarpMatrix(pt, start);
        const Eigen::Vector2d v2 = finder.GetProjection();
        const Eigen::Matrix2d wi = finder.GetWarpInverse();
        printf("pf %d level %d v2 %.17g %.17g warp %.17g %.17g %.17g %.17g", pt, level, v2(0), v2(1), wi(0, 0), wi(0, 1), wi(1, 0), wi(1, 1));
        if (level >= 0) {
          finder.MakeTemplateCoarseCont(pt);
          const bool found = finder.FindPatchCoarseAndSubPix(v2, tracker.mCurrentKF, 10, 8);
          const Eigen::Vector2d c = finder.GetCoarsePosAsVector(), sp = finder.GetSubPixPos();
          printf(" bad %d found %d coarse %.17g %.17g subpix %.17g %.17g", finder.TemplateBad() ? 1 : 0, found ? 1 : 0, c(0), c(1), sp(0), sp(1));
          if (found) {   // the same in the reference's separate steps: ZMSSDAtPoint at the coarse hit, MakeSubPixTemplate, IterateSubPix x n == IterateSubPixToConvergence
            cv::Mat& lim = tracker.mCurrentKF.aLevels[level].im;
            const int cx = (int)((c(0) + 0.5) / (1 << level) - 0.5 + 0.5), cy = (int)((c(1) + 0.5) / (1 << level) - 0.5 + 0.5);
            const int z = finder.ZMSSDAtPoint(lim, cx, cy);
            finder.MakeSubPixTemplate();
            const bool conv = finder.IterateSubPixToConvergence(tracker.mCurrentKF, 8);
            const Eigen::Vector2d a = finder.GetSubPixPos();
            finder.SetSubPixPos(c); finder.MakeSubPixTemplate();
            int its = 0; double u = 1.0;
            for (; its < 8; its++) { u = finder.IterateSubPix(tracker.mCurrentKF); if (u < 0 || u < 0.03 * 0.03) break; }
            const Eigen::Vector2d b2 = finder.GetSubPixPos();
            printf(" zmssd %d maxssd %d conv %d steps %.17g %.17g same %d", z, finder.mnMaxSSD, conv ? 1 : 0, a(0), a(1), (a(0) == b2(0) && a(1) == b2(1)) ? 1 : 0);
          }
        }
        printf("\n");
      }
      // SearchForPoints + CalcPoseUpdate on every third point
      finder.CalcSearchLevelAndWarpMatrix(0, start);   // (re-project the whole map at the start pose)
      std::vector<int> list;
      for (int pt = 0; pt < N; pt += 3) if (finder.LevelOf(pt) >= 0) list.push_back(pt);
      const int nfound = tracker.SearchForPoints(list, 12, 4);
      printf("search %d of %d\n", nfound, (int)list.size());
      tracker.CalcJacobians(list);
      const Eigen::VectorXd mu = tracker.CalcPoseUpdate(list);
      printf("update %.17g %.17g %.17g %.17g %.17g %.17g\n", mu(0), mu(1), mu(2), mu(3), mu(4), mu(5));
      const Eigen::VectorXd mu2 = tracker.CalcPoseUpdate(list, 16.0, true);
      printf("update16 %.17g %.17g %.17g %.17g %.17g %.17g\n", mu2(0), mu2(1), mu2(2), mu2(3), mu2(4), mu2(5));
      return 0;
    }
  } catch (const std::exception& e) { fprintf(stderr, "%s\n", e.what()); return 3; }
  return 2;
}

// Shared by the tracking kernels (track.cu, pose_fast.cu, search_fast.cu): the by-value kernel argument block and the per-point
// projection helpers (TrackerData::Project, PatchFinder::CalcSearchLevelAndWarpMatrix).
#pragma once
#include "geometry.cuh"

namespace {

constexpr int kPT = 256;  // threads per CTA in per-stream kernels
constexpr int kSearchWarps = 4;

struct Dev {   // everything the kernels need, passed by value
  LevelDesc lev[VS_LEVELS];
  const uint8_t* const* l0_ptr; const int* l0_stride;
  CamDev cam; MapDev map; SourceKF src; PointState ps;
  StreamState* ss; int* lists; int list_cap; int* pvs; double* sort_scratch; int sort_cap;
  unsigned long long* evals;
  int S, N, P, truncate;
  int s0;                  // first stream of this launch (stream groups of vs_launch_frame; 0 otherwise)
  int* coarse_hint;        // [S] host-visible: did the stream try the coarse stage in this frame (vs_launch_track_map_rest reads it as a layout hint)
  int chain;               // vs_launch_track_map_rest's two launch chains: -1 every stream, 1 only the streams that try the coarse stage this frame, 0 only the others
  vslam_params prm;
  // keyframe policy: poses of the map's keyframes (the relocaliser registration), 0 keyframes = policy off
  const double* kf_pose; int kf_n, kf_min_frames; double kf_excess_dist, kf_need_dist; int* kf_req;
};

__device__ __forceinline__ int LevelScale(int l) { return 1 << l; }
// true: this stream belongs to the other launch chain of the frame (Dev::chain)
__device__ __forceinline__ bool other_chain(const Dev& D, const StreamState* st) { return D.chain >= 0 && (st->try_coarse != 0) != (D.chain != 0); }

// ------------------------------------------------------------------------------------------------
// TrackerData::Project (jni/TrackerData.h:69-86).  Returns true if Cam.Project was reached (cache valid).
__device__ inline bool td_project(const Dev& D, const double* pose, int i, size_t gi, size_t SN, CamCache& cc, int& flags) {
  flags &= ~F_INIMAGE;
  const double* w = D.map.world + 3 * (size_t)i;
  const double wp[3] = {w[0], w[1], w[2]};
  double c[3]; se3_apply(pose, wp, c);
  D.ps.v3cam[gi] = c[0]; D.ps.v3cam[SN + gi] = c[1]; D.ps.v3cam[2 * SN + gi] = c[2];
  if (c[2] < 0.001) return false;
  const double px = c[0] / c[2], py = c[1] / c[2];
  double d = 0; d += px * px; d += py * py;
  if (d > D.cam.largestRadius * D.cam.largestRadius) return false;
  double im[2]; cam_project(D.cam, px, py, im, cc);
  D.ps.v2image[gi] = im[0]; D.ps.v2image[SN + gi] = im[1];
  if (cc.invalid) return true;
  if (im[0] < 0 || im[1] < 0 || im[0] > D.cam.width || im[1] > D.cam.height) return true;
  flags |= F_INIMAGE;
  return true;
}

// PatchFinder::CalcSearchLevelAndWarpMatrix (jni/PatchFinder.cc:31-68)
__device__ inline int calc_level_warp(const Dev& D, const double* pose, int i, size_t gi, size_t SN, const double* dv, int& flags) {
  const double c[3] = {D.ps.v3cam[gi], D.ps.v3cam[SN + gi], D.ps.v3cam[2 * SN + gi]};
  const double invz = 1.0 / c[2];
  const double* rp = D.map.right + 3 * (size_t)i; const double* dp = D.map.down + 3 * (size_t)i;
  const double r3[3] = {rp[0], rp[1], rp[2]}, d3[3] = {dp[0], dp[1], dp[2]};
  double mr[3], md[3]; rot_apply(pose, r3, mr); rot_apply(pose, d3, md);
  double a[2], b[2];
  for (int k = 0; k < 2; k++) { a[k] = mr[k] - c[k] * mr[2] * invz; b[k] = md[k] - c[k] * md[2] * invz; }
  double aux1[2], aux2[2];
  for (int r = 0; r < 2; r++) {
    double s = dv[2 * r] * a[0]; s += dv[2 * r + 1] * a[1]; aux1[r] = s * invz;
    double t = dv[2 * r] * b[0]; t += dv[2 * r + 1] * b[1]; aux2[r] = t * invz;
  }
  const double w00 = aux1[0], w01 = aux2[0], w10 = aux1[1], w11 = aux2[1];
  D.ps.warpinv[gi] = w00; D.ps.warpinv[SN + gi] = w01; D.ps.warpinv[2 * SN + gi] = w10; D.ps.warpinv[3 * SN + gi] = w11;
  double det = w00 * w11 - w01 * w10;
  int level = 0;
  while (det > 3 && level < VS_LEVELS - 1) { level++; det *= 0.25; }
  // m2 = inverse(mm2WarpInverse) * LevelScale (jni/PatchFinder.cc:82-83, 2x2 adjugate inverse as frozen in the oracle), here
  // instead of on a single lane of k_search.  Also for rejected warps, with the level the loop reached (mnSearchLevel keeps that
  // value): MapMaker::ReFind_Common goes on to MakeTemplateCoarseCont after a rejection (jni/PatchFinder.cc:72-76).
  const double invdet = 1.0 / (w00 * w11 - w01 * w10);
  const int sc = LevelScale(level);
  D.ps.m2[gi] = (w11 * invdet) * sc; D.ps.m2[SN + gi] = (-w01 * invdet) * sc; D.ps.m2[2 * SN + gi] = (-w10 * invdet) * sc; D.ps.m2[3 * SN + gi] = (w00 * invdet) * sc;
  D.ps.rlevel[gi] = level;
  if (det > 3 || det < 0.25) { flags |= F_TBAD; return -1; }
  return level;
}

inline Dev make_dev(const vslam_ctx* ctx) {
  Dev D;
  for (int l = 0; l < VS_LEVELS; l++) D.lev[l] = ctx->lev[l];
  D.l0_ptr = ctx->l0_ptr; D.l0_stride = ctx->l0_stride;
  D.cam = ctx->cam; D.map = ctx->map; D.src = ctx->src; D.ps = ctx->ps; D.ss = ctx->ss; D.lists = ctx->lists; D.list_cap = ctx->list_cap;
  D.pvs = ctx->pvs; D.sort_scratch = ctx->sort_scratch; D.sort_cap = ctx->sort_cap; D.evals = ctx->evals;
  D.S = ctx->S; D.N = ctx->N; D.P = ctx->P; D.truncate = ctx->cfg.truncate_error; D.prm = ctx->params; D.s0 = ctx->cur_s0; D.chain = ctx->cur_chain; D.coarse_hint = ctx->coarse_hint_dev;
  const bool kf = ctx->kf_policy && ctx->reloc_n > 0;
  D.kf_pose = ctx->reloc_pose; D.kf_n = kf ? ctx->reloc_n : 0; D.kf_min_frames = ctx->kf_min_frames;
  D.kf_excess_dist = ctx->kf_wiggle * 10.0; D.kf_need_dist = ctx->kf_mult * ctx->kf_wiggle_dn; D.kf_req = ctx->kf_req;
  return D;
}

}  // namespace

// PatchFinder one object at a time: the per-object methods of jni/PatchFinder.h:45-121 that the batched kernels (search_fast.cu) fuse
// away -- MakeTemplateCoarseNoWarp, ZMSSDAtPoint at caller-chosen positions, MakeSubPixTemplate + IterateSubPix from a caller-set start.
// A "PatchFinder object" is the finder state the tracker keeps per (stream, map point): template, sums, search level.  This is the slow
// path (one small launch per call, results read back synchronously); it exists so that code written against the reference's per-object
// interface (include/vslam_b200_shell.hpp, class PatchFinder) runs unchanged.  Same arithmetic and operation order as the batched
// kernels, hence as the oracle.
#include "track_dev.cuh"

namespace {

// PatchFinder::MakeTemplateCoarseNoWarp(KeyFrame&, nLevel, x, y) (jni/PatchFinder.cc:130-143): the P x P pixels around (x, y) of level
// `level` of source keyframe `kf`, un-warped; also MakeTemplateSums (:153-164).  out[0] = mbTemplateBad.
__global__ void k_pf_template_nowarp(Dev D, int s, int i, int kf, int level, int x, int y, int* out) {
  const int P = D.P, lane = threadIdx.x;
  const size_t SN = (size_t)D.S * D.N, gi = (size_t)s * D.N + i;
  if (kf < 0) { kf = D.map.srckf[i]; level = D.map.srclevel[i]; x = D.map.ircenter[2 * i]; y = D.map.ircenter[2 * i + 1]; }   // MakeTemplateCoarseNoWarp(MapPoint&) (:146-149)
  const int w = D.src.w[level], h = D.src.h[level], pitch = D.src.pitch[level];
  const uint8_t* img = D.src.img[level] + (size_t)kf * h * pitch;
  const int bd = P / 2 + 1;
  int flags = D.ps.flags[gi];
  const bool bad = !(x >= bd && y >= bd && x < w - bd && y < h - bd);      // in_image_with_border(im, x, y, mnPatchSize / 2 + 1)
  if (lane == 0) { D.ps.level[gi] = level; D.ps.rlevel[gi] = level; }       // mnSearchLevel = nLevel, before the border test like the reference
  if (bad) {
    if (lane == 0) { D.ps.flags[gi] = (flags | F_TBAD) & ~F_HAVELAST; out[0] = 1; }
    return;
  }
  uint8_t* t = D.ps.tmpl + gi * VS_TMPL_BYTES;                              // rows of 12 bytes, zero padded (the dp4a operand layout)
  int sum = 0, sumsq = 0;
  for (int k = lane; k < 12 * P; k += 32) {
    const int r = k / 12, c = k - 12 * r;
    const int v = c < P ? img[(size_t)(y - P / 2 + r) * pitch + (x - P / 2 + c)] : 0;
    t[k] = (uint8_t)v; sum += v; sumsq += v * v;
  }
  for (int k = 12 * P + lane; k < VS_TMPL_BYTES; k += 32) t[k] = 0;
#pragma unroll
  for (int d = 16; d; d >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, d); sumsq += __shfl_xor_sync(0xffffffffu, sumsq, d); }
  if (lane == 0) {
    D.ps.tsum[gi] = sum; D.ps.tsum[SN + gi] = sumsq;
    // the template no longer is the warped one of the last MakeTemplateCoarseCont: the next one regenerates
    D.ps.flags[gi] = (flags & ~(F_TBAD | F_HAVELAST)) | F_NEWTMPL; out[0] = 0;
  }
}

// PatchFinder::ZMSSDAtPoint (jni/PatchFinder.cc:352-380) of point i's template at n positions of level `level` of stream s's current
// keyframe; one thread per position, integer arithmetic.
__global__ void k_pf_zmssd_at(Dev D, int s, int i, int level, int n, const int* __restrict__ xy, int* __restrict__ ssd) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int P = D.P, b = P / 2;
  const size_t SN = (size_t)D.S * D.N, gi = (size_t)s * D.N + i;
  const LevelDesc& L = D.lev[level];
  const uint8_t* img; int pitch;
  if (level == 0) { img = D.l0_ptr[s]; pitch = D.l0_stride[s]; } else { img = L.img + (size_t)s * L.h * L.pitch; pitch = L.pitch; }
  const int cx = xy[2 * k], cy = xy[2 * k + 1];
  const int maxSSD = P * P * 500;
  if (!(cx >= b && cy >= b && cx < L.w - b && cy < L.h - b)) { ssd[k] = maxSSD + 1; return; }
  const uint8_t* t = D.ps.tmpl + gi * VS_TMPL_BYTES;
  int sum = 0, sumsq = 0, cross = 0;
  for (int r = 0; r < P; r++) {
    const uint8_t* ip = img + (size_t)(cy - b + r) * pitch + (cx - b);
    for (int c = 0; c < P; c++) { const int v = ip[c]; sum += v; sumsq += v * v; cross += v * t[12 * r + c]; }
  }
  const int SA = D.ps.tsum[gi], SB = sum;
  ssd[k] = (2 * SA * SB - SA * SA - SB * SB) / (P * P) + sumsq + D.ps.tsum[SN + gi] - 2 * cross;
}

// MakeSubPixTemplate's JtJ^-1 (jni/PatchFinder.cc:242-267) and `max_its` x IterateSubPix (:290-350) of point i's template in level
// D.ps.level of stream s's current keyframe, from io[0..1] = mv2SubPixPos (level-zero pixels) and io[2] = mdMeanDiff.  One thread: the
// reference's loop order, term by term.  io[3] = the last iteration's squared pixel update (negative: off the image), io[4] = 1 when an
// iteration came in under the convergence limit (IterateSubPixToConvergence, :272-285), io[5] = iterations run.
__global__ void k_pf_subpix(Dev D, int s, int i, int max_its, double* io) {
  if (threadIdx.x != 0) return;
  const int P = D.P, Q = P - 2;
  const size_t gi = (size_t)s * D.N + i;
  const int level = D.ps.level[gi];
  if (level < 0 || level >= VS_LEVELS) { io[3] = -1.0; io[4] = 0.0; io[5] = 0.0; return; }
  const LevelDesc& L = D.lev[level];
  const uint8_t* img; int pitch;
  if (level == 0) { img = D.l0_ptr[s]; pitch = D.l0_stride[s]; } else { img = L.img + (size_t)s * L.h * L.pitch; pitch = L.pitch; }
  const uint8_t* tmpl = D.ps.tmpl + gi * VS_TMPL_BYTES;
  // JtJ of (gx, gy, 1): sums of multiples of 0.25 below 2^53, exact in any order
  double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int x = 1; x <= Q; x++)
    for (int y = 1; y <= Q; y++) {
      const double gx = 0.5 * (tmpl[y * 12 + x + 1] - tmpl[y * 12 + x - 1]), gy = 0.5 * (tmpl[(y + 1) * 12 + x] - tmpl[(y - 1) * 12 + x]);
      H[0] += gx * gx; H[1] += gx * gy; H[2] += gx; H[4] += gy * gy; H[5] += gy; H[8] += 1.0;
    }
  H[3] = H[1]; H[6] = H[2]; H[7] = H[5];
  // 3x3 inverse = adjugate * (1 / det), cofactor expansion along the first column (the stand-in Eigen of the reference build; csrc/search_fast.cu)
  double hinv[9];
  {
    const double c00 = H[4] * H[8] - H[5] * H[7], c10 = H[5] * H[6] - H[3] * H[8], c20 = H[3] * H[7] - H[4] * H[6];
    double det = H[0] * c00; det += H[1] * c10; det += H[2] * c20;
    const double invdet = 1.0 / det;
    hinv[0] = c00 * invdet; hinv[3] = c10 * invdet; hinv[6] = c20 * invdet;
    hinv[1] = (H[2] * H[7] - H[1] * H[8]) * invdet; hinv[4] = (H[0] * H[8] - H[2] * H[6]) * invdet; hinv[7] = (H[1] * H[6] - H[0] * H[7]) * invdet;
    hinv[2] = (H[1] * H[5] - H[2] * H[4]) * invdet; hinv[5] = (H[2] * H[3] - H[0] * H[5]) * invdet; hinv[8] = (H[0] * H[4] - H[1] * H[3]) * invdet;
  }
  const int nLevelScale = LevelScale(level);
  const double invScale = 1.0 / nLevelScale;
  double sp0 = io[0], sp1 = io[1], meanDiff = io[2], last = -1.0;
  int ok = 0, it = 0;
  for (; it < max_its && !ok; it++) {
    const double c0 = (sp0 + 0.5) * invScale - 0.5, c1 = (sp1 + 0.5) * invScale - 0.5;   // LevelNPos
    const int xb = (c0 > 0.0 ? c0 + 0.5 : c0 - 0.5), yb = (c1 > 0.0 ? c1 + 0.5 : c1 - 0.5);
    const int bd = P / 2 + 1;
    if (!(xb >= bd && yb >= bd && xb < L.w - bd && yb < L.h - bd)) { last = -1.0; it++; break; }   // went off the edge of the image
    const double b0 = c0 - (double)(P / 2), b1 = c1 - (double)(P / 2);
    const double dX = b0 - floor(b0), dY = b1 - floor(b1);
    const float fTL = (1.0 - dX) * (1.0 - dY), fTR = (dX) * (1.0 - dY), fBL = (1.0 - dX) * (dY), fBR = (dX) * (dY);
    double a0 = 0, a1 = 0, a2 = 0;
    for (int y = 1; y <= Q; y++)
      for (int x = 1; x <= Q; x++) {
        const uint8_t* tl = img + (size_t)((int)b1 + y) * pitch + ((int)b0 + x);
        const float fPixel = fTL * tl[0] + fTR * tl[1] + fBL * tl[pitch] + fBR * tl[pitch + 1];
        const double dDiff = fPixel - tmpl[y * 12 + x] + meanDiff;
        const double gx = 0.5 * (tmpl[y * 12 + x + 1] - tmpl[y * 12 + x - 1]), gy = 0.5 * (tmpl[(y + 1) * 12 + x] - tmpl[(y - 1) * 12 + x]);
        a0 += dDiff * gx; a1 += dDiff * gy; a2 += dDiff;
      }
    double upd[3];
    for (int r = 0; r < 3; r++) { double sacc = hinv[3 * r] * a0; sacc += hinv[3 * r + 1] * a1; sacc += hinv[3 * r + 2] * a2; upd[r] = sacc; }
    sp0 -= upd[0] * nLevelScale; sp1 -= upd[1] * nLevelScale;
    meanDiff -= upd[2];
    double d = 0; d += upd[0] * upd[0]; d += upd[1] * upd[1];
    last = d;
    const double lim = 0.03;
    if (d < lim * lim) ok = 1;
  }
  io[0] = sp0; io[1] = sp1; io[2] = meanDiff; io[3] = last; io[4] = (double)ok; io[5] = (double)it;
}

int check_point(vslam_ctx* ctx, int s, int i) {
  if (!ctx) return VSLAM_E_INVALID;
  if (s < 0 || s >= ctx->S) { ctx->err = "stream index out of range"; return VSLAM_E_INVALID; }
  if (i < 0 || i >= ctx->map.n) { ctx->err = "map point index out of range"; return VSLAM_E_INVALID; }
  return VSLAM_OK;
}

// small device scratch for the results of these calls (grown on demand, freed with the context's epipolar scratch)
int scratch(vslam_ctx* ctx, size_t bytes, void** out) {
  if (bytes > ctx->pf_cap) {
    if (ctx->pf_buf) cudaFree(ctx->pf_buf);
    ctx->pf_buf = nullptr; ctx->pf_cap = 0;
    VS_CUDA(cudaMalloc(&ctx->pf_buf, bytes));
    ctx->pf_cap = bytes;
  }
  *out = ctx->pf_buf;
  return VSLAM_OK;
}

}  // namespace

extern "C" {

// PatchFinder::MakeTemplateCoarseCont(p) (jni/PatchFinder.cc:79-125) for one point, with the warp of the last projection
// (vslam_project_all / CalcSearchLevelAndWarpMatrix): the template stage of k_search_fast on a one-entry list, re-use rule included.
int vslam_pf_make_template(vslam_ctx* ctx, int s, int i, int* template_bad) {
  int rc = check_point(ctx, s, i); if (rc) return rc;
  std::vector<int32_t> idx(ctx->S, 0), cnt(ctx->S, 0);
  idx[s] = i; cnt[s] = 1;
  if ((rc = vslam_set_lists(ctx, idx.data(), cnt.data(), 1))) return rc;
  if ((rc = vs_launch_search_fast(ctx, 0, 0, 0, 2 /* template only */))) return rc;
  int flags = 0;
  VS_CUDA(cudaMemcpyAsync(&flags, ctx->ps.flags + (size_t)s * ctx->N + i, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (template_bad) *template_bad = (flags & F_TBAD) ? 1 : 0;
  return VSLAM_OK;
}

int vslam_pf_make_template_nowarp(vslam_ctx* ctx, int s, int i, int src_kf, int level, int x, int y, int* template_bad) {
  int rc = check_point(ctx, s, i); if (rc) return rc;
  if (src_kf >= 0 && (level < 0 || level >= VS_LEVELS || src_kf >= ctx->n_src || !ctx->src_have[src_kf])) { ctx->err = "bad source keyframe / level"; return VSLAM_E_INVALID; }
  void* buf; if ((rc = scratch(ctx, 64, &buf))) return rc;
  k_pf_template_nowarp<<<1, 32, 0, ctx->stream>>>(make_dev(ctx), s, i, src_kf, level, x, y, (int*)buf);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  int bad = 0;
  VS_CUDA(cudaMemcpyAsync(&bad, buf, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (template_bad) *template_bad = bad;
  return VSLAM_OK;
}

int vslam_pf_zmssd_at(vslam_ctx* ctx, int s, int i, int level, int n, const int32_t* xy, int32_t* ssd) {
  int rc = check_point(ctx, s, i); if (rc) return rc;
  if (level < 0 || level >= VS_LEVELS || n < 0 || (n > 0 && (!xy || !ssd))) return VSLAM_E_INVALID;
  if (n == 0) return VSLAM_OK;
  void* buf; if ((rc = scratch(ctx, sizeof(int) * 3 * (size_t)n, &buf))) return rc;
  int* dxy = (int*)buf; int* dssd = dxy + 2 * (size_t)n;
  VS_CUDA(cudaMemcpyAsync(dxy, xy, sizeof(int) * 2 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  k_pf_zmssd_at<<<(n + 127) / 128, 128, 0, ctx->stream>>>(make_dev(ctx), s, i, level, n, dxy, dssd);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  VS_CUDA(cudaMemcpyAsync(ssd, dssd, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  return VSLAM_OK;
}

int vslam_pf_subpix(vslam_ctx* ctx, int s, int i, int max_its, double* pos2, double* mean_diff, int* converged, double* last_update_sq) {
  int rc = check_point(ctx, s, i); if (rc) return rc;
  if (!pos2 || !mean_diff || max_its < 0) return VSLAM_E_INVALID;
  void* buf; if ((rc = scratch(ctx, sizeof(double) * 8, &buf))) return rc;
  double io[6] = {pos2[0], pos2[1], *mean_diff, -1.0, 0.0, 0.0};
  VS_CUDA(cudaMemcpyAsync(buf, io, sizeof(io), cudaMemcpyHostToDevice, ctx->stream));
  k_pf_subpix<<<1, 32, 0, ctx->stream>>>(make_dev(ctx), s, i, max_its, (double*)buf);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  VS_CUDA(cudaMemcpyAsync(io, buf, sizeof(io), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  pos2[0] = io[0]; pos2[1] = io[1]; *mean_diff = io[2];
  if (last_update_sq) *last_update_sq = io[3];
  if (converged) *converged = io[4] != 0.0;
  return VSLAM_OK;
}

// The one user event of the reference's shell: SystemPTAM::onTouchScreen sets Tracker::mbUserPressedSpacebar (jni/jni_part.cpp:49-51,
// :125-129), consumed by TrackForInitialMap (jni/Tracker.cc:232-253).  The flag lives with the stream so that a binding without the C++
// shell (JNI straight onto this header) has somewhere to put it; vslam_take_user_event returns and clears it.
int vslam_user_event(vslam_ctx* ctx, int s, int event) {
  if (!ctx || s < 0 || s >= ctx->S) return VSLAM_E_INVALID;
  if (event != VSLAM_EVENT_SPACEBAR) { ctx->err = "unknown user event"; return VSLAM_E_INVALID; }
  ctx->user_events[s] |= 1;
  return VSLAM_OK;
}
int vslam_take_user_event(vslam_ctx* ctx, int s, int* pending) {
  if (!ctx || s < 0 || s >= ctx->S || !pending) return VSLAM_E_INVALID;
  *pending = ctx->user_events[s];
  ctx->user_events[s] = 0;
  return VSLAM_OK;
}

}  // extern "C"

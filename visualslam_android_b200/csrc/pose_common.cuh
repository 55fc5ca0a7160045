// Pieces of the pose-update kernels shared by k_pose (track.cu, any list length) and k_pose_fast (pose_fast.cu).
#pragma once
#include "track_dev.cuh"

namespace {

// mu = inverse(C) * b with inverse = partial-pivot LU, column by column (jni/myWLS.h:53-62; oracle inverse_lu): same operations per
// element as the serial routine, spread over six lanes of ONE warp (rows during elimination, columns during substitution).
// sums: the 21 upper-triangle terms of C (without the prior) + the 6 right-hand sides; lu/inv: 36 doubles, piv: 6 ints of shared memory.
__device__ __forceinline__ void wls_solve_warp(const double* sums, double* lu, double* inv, int* piv, double* mu) {
  const int lane = threadIdx.x & 31;
    __syncwarp();   // reconverge first: with diverged lanes every shuffle below takes the slow WARPSYNC.COLLECTIVE path (~300 cycles each)
    // lane r (< 6) keeps row r of C in registers; pivot search, row swap and the pivot-row broadcast go through shuffles
    const int r6 = lane < 6 ? lane : 5;
    double row[6];
#pragma unroll
    for (int c = 0; c < 6; c++) {
      const int lo = r6 < c ? r6 : c, hi = r6 < c ? c : r6;
      const int q = lo * 6 - (lo * (lo - 1)) / 2 + (hi - lo);          // index of (lo,hi) in the packed upper triangle
      row[c] = sums[q] + (lo == hi ? 100.0 : 0.0);                    // prior 100*I (:734)
    }
    int mypiv = r6;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      // partial pivoting: first row i >= k with the largest |a[i][k]| (strict >, as the serial routine)
      double best = (lane >= k && lane < 6) ? fabs(row[k]) : -1.0; int bi = lane;
#pragma unroll
      for (int d = 4; d; d >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, d); const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      const int p = __shfl_sync(0xffffffffu, bi, 0);
      // swap rows k and p (and their pivot labels)
      const int src = lane == k ? p : (lane == p ? k : lane);
#pragma unroll
      for (int c = 0; c < 6; c++) row[c] = __shfl_sync(0xffffffffu, row[c], src);
      mypiv = __shfl_sync(0xffffffffu, mypiv, src);
      double prow[6];
#pragma unroll
      for (int c = 0; c < 6; c++) prow[c] = __shfl_sync(0xffffffffu, row[c], k);
      if (lane > k && lane < 6) {
        const double f = row[k] / prow[k];
        row[k] = f;
#pragma unroll
        for (int j = k + 1; j < 6; j++) row[j] -= f * prow[j];
      }
    }
    if (lane < 6) {
#pragma unroll
      for (int c = 0; c < 6; c++) lu[lane * 6 + c] = row[c];
      piv[lane] = mypiv;
    }
    __syncwarp();
    if (lane < 6) {   // column `lane` of the inverse
      const int c = lane; double x[6];
#pragma unroll
      for (int i = 0; i < 6; i++) x[i] = (piv[i] == c) ? 1.0 : 0.0;
#pragma unroll
      for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < i; j++) x[i] -= lu[i * 6 + j] * x[j];
#pragma unroll
      for (int i = 5; i >= 0; i--) {
#pragma unroll
        for (int j = i + 1; j < 6; j++) x[i] -= lu[i * 6 + j] * x[j];
        x[i] /= lu[i * 6 + i];
      }
#pragma unroll
      for (int i = 0; i < 6; i++) inv[i * 6 + c] = x[i];
    }
    __syncwarp();
    if (lane < 6) { const int i = lane; double sacc = inv[6 * i] * sums[21]; for (int j = 1; j < 6; j++) sacc += inv[6 * i + j] * sums[21 + j]; mu[i] = sacc; }
}

// The same elimination with a shorter dependency chain (k_pose_fast): instead of a three-level shuffle tournament for the pivot and a
// broadcast of its index, every lane fetches column k of all six rows (six independent shuffles) and finds the pivot itself; the row swap
// and the pivot-row broadcast are issued together.  Operation for operation the arithmetic of wls_solve_warp (bit-identical results).
__device__ __forceinline__ void wls_solve_warp_fast(const double* sums, double* lu, double* inv, int* piv, double* mu, long long* tmark = nullptr) {
  const int lane = threadIdx.x & 31;
  long long t0_ = 0; if (tmark) t0_ = clock64();
  __syncwarp();
  const int r6 = lane < 6 ? lane : 5;
  double row[6];
#pragma unroll
  for (int c = 0; c < 6; c++) {
    const int lo = r6 < c ? r6 : c, hi = r6 < c ? c : r6;
    const int q = lo * 6 - (lo * (lo - 1)) / 2 + (hi - lo);
    row[c] = sums[q] + (lo == hi ? 100.0 : 0.0);
  }
  int mypiv = r6;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    double col[6];
#pragma unroll
    for (int j = 0; j < 6; j++) col[j] = fabs(__shfl_sync(0xffffffffu, row[k], j));
    int p = k; double best = col[k];
#pragma unroll
    for (int i = k + 1; i < 6; i++) if (col[i] > best) { best = col[i]; p = i; }      // first row with the largest |a[i][k]| (strict >)
    const int src = lane == k ? p : (lane == p ? k : lane);
    double prow[6];
#pragma unroll
    for (int c = 0; c < 6; c++) { prow[c] = __shfl_sync(0xffffffffu, row[c], p); row[c] = __shfl_sync(0xffffffffu, row[c], src); }   // old row p = new row k
    mypiv = __shfl_sync(0xffffffffu, mypiv, src);
    // branch-free update (rows <= k keep their values): no divergence between the shuffles of consecutive steps
    const bool act = lane > k;
    const double f = row[k] / prow[k];
    row[k] = act ? f : row[k];
#pragma unroll
    for (int j = k + 1; j < 6; j++) { const double t = row[j] - f * prow[j]; row[j] = act ? t : row[j]; }
  }
  if (lane < 6) {
#pragma unroll
    for (int c = 0; c < 6; c++) lu[lane * 6 + c] = row[c];
    piv[lane] = mypiv;
  }
  __syncwarp();
  if (tmark && lane == 0) { const long long t1_ = clock64(); tmark[0] += t1_ - t0_; t0_ = t1_; }
  if (lane < 6) {   // column `lane` of the inverse
    const int c = lane; double x[6];
#pragma unroll
    for (int i = 0; i < 6; i++) x[i] = (piv[i] == c) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j < i; j++) x[i] -= lu[i * 6 + j] * x[j];
#pragma unroll
    for (int i = 5; i >= 0; i--) {
#pragma unroll
      for (int j = i + 1; j < 6; j++) x[i] -= lu[i * 6 + j] * x[j];
      x[i] /= lu[i * 6 + i];
    }
#pragma unroll
    for (int i = 0; i < 6; i++) inv[i * 6 + c] = x[i];
  }
  __syncwarp();
  if (lane < 6) { const int i = lane; double sacc = inv[6 * i] * sums[21]; for (int j = 1; j < 6; j++) sacc += inv[6 * i + j] * sums[21 + j]; mu[i] = sacc; }
  if (tmark && lane == 0) { const long long t1_ = clock64(); tmark[1] += t1_ - t0_; }
}

// End of the fine stage, one thread: scene depth from the tracked features (jni/Tracker.cc:610-625), then -- when `tail` -- Tracker::UpdateMotionModel,
// AssessTrackingQuality and the keyframe request.  dSum / dSumSq / dNum: sums of z, z^2 and 1 over the found points; pose: the stage's final pose.
__device__ inline void pose_stage_tail(const Dev& D, StreamState* st, int s, const double* pose, double dSum, double dSumSq, double dNum, int tail) {
    const int nNum = (int)dNum;
    if (nNum > 20) { st->depth_mean = dSum / nNum; st->depth_sigma = sqrt((dSumSq / nNum) - (st->depth_mean) * (st->depth_mean)); }
    if (tail) {
      if (!st->recovered) {
      // Tracker::UpdateMotionModel (jni/Tracker.cc:802-820); not after a relocalisation (jni/Tracker.cc:136-139)
      double inv[12], nfo[12], m[6];
      se3_inverse(st->start_pose, inv); se3_mul(pose, inv, nfo); se3_ln(nfo, m);
      double sacc = 0;
      for (int k = 0; k < 6; k++) { st->velocity[k] = 0.9 * (0.5 * m[k] + 0.5 * st->velocity[k]); }
      for (int k = 0; k < 6; k++) sacc += st->velocity[k] * st->velocity[k];
      st->vel_mag = sqrt(sacc);
      double v[6]; for (int k = 0; k < 6; k++) v[k] = st->velocity[k];
      for (int k = 0; k < 3; k++) v[k] *= 1.0 / st->depth_mean;
      sacc = 0; for (int k = 0; k < 6; k++) sacc += v[k] * v[k];
      st->msd_scaled_vel = sqrt(sacc);
      }
      // Tracker::AssessTrackingQuality (jni/Tracker.cc:832-878)
      int nTA = 0, nTF = 0, nLA = 0, nLF = 0;
      for (int l = 0; l < VS_LEVELS; l++) { nTA += st->attempted[l]; nTF += st->found[l]; if (l >= 2) { nLA += st->attempted[l]; nLF += st->found[l]; } }
      int q;
      if (nTF == 0 || nTA == 0) q = 0;
      else {
        const double tot = (double)nTF / nTA, lg = (nLA > 10) ? (double)nLF / nLA : tot;
        q = (tot > 0.3) ? 2 : (lg < 0.13 ? 0 : 1);
      }
      // MapMaker::ClosestKeyFrame / KeyFrameLinearDist (jni/MapMaker.cc:705-712,736-754) over the registered keyframes: only the two
      // callers below need it (DODGY: "has the pose run miles away", GOOD: "is a new keyframe due")
      double kfd = 9999999999.9; int kfc = -1;
      if (D.kf_n > 0 && q != 0) {
        double inv[12]; se3_inverse(pose, inv);
        for (int k = 0; k < D.kf_n; k++) {
          double ki[12]; se3_inverse(D.kf_pose + 12 * k, ki);
          const double d0 = ki[3] - inv[3], d1 = ki[7] - inv[7], d2 = ki[11] - inv[11];
          double dd = d0 * d0; dd += d1 * d1; dd += d2 * d2;
          const double dist = sqrt(dd);
          if (dist < kfd) { kfd = dist; kfc = k; }
        }
        st->kf_dist = kfd; st->kf_closest = kfc;
        if (q == 1 && kfd > D.kf_excess_dist) q = 0;          // IsDistanceToNearestKeyFrameExcessive (jni/MapMaker.cc:1098-1101)
      }
      st->quality = q;
      if (q == 0) st->lost_frames++; else st->lost_frames = 0;
      // jni/Tracker.cc:127-132 (not in the recovery branch): GOOD && MapMaker::NeedNewKeyFrame (jni/MapMaker.cc:763-773) && enough
      // frames since the last one.  The queue-length term (QueueSize() < 3) is the caller's: it owns the queue.
      if (kfc >= 0 && q == 2 && !st->recovered) {
        double dDist = kfd; dDist *= (1.0 / st->depth_mean);
        if (dDist > D.kf_need_dist && st->frame_no - st->last_kf_dropped > D.kf_min_frames) { st->kf_request = 1; D.kf_req[s] = 1; }
      }
    }
}

}  // namespace

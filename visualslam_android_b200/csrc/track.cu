// Tracker::TrackMap on the GPU (reference: jni/Tracker.cc:358-626), batched over all streams of a context.
//
//   k_project_lists   one CTA per stream: TrackerData::Project + GetDerivsUnsafe + CalcSearchLevelAndWarpMatrix for
//                     every map point, ordered per-level PVS lists, std::random_shuffle with the stream's glibc rand()
//                     state, coarse / fine list selection (all of the reference's rand() use happens here).
//   k_search          one warp per (stream, list entry): MakeTemplateCoarseCont, FindPatchCoarse (ZMSSD with dp4a over
//                     the FAST corners of the row-LUT window), sub-pixel inverse-compositional refinement.
//   k_pose            one CTA per stream: the ten Gauss-Newton iterations of a stage — re-projection / linear update,
//                     Jacobians, Tukey sigma (radix-select median), weights, the 27-term normal-equation reduction,
//                     6x6 LU solve, pose = exp(mu) * pose — plus scene depth, motion model and quality assessment.
// FP64 throughout, no FMA contraction (-fmad=false): integer results are bit-exact, poses agree to ~1e-12.
#include "track_dev.cuh"
#include "pose_common.cuh"
#include "search_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// First loop of TrackMap + list building.  mode 0: stage API (flags reset for every point, no lists); mode 1: TrackMap.
// The swaps of one std::random_shuffle over v[0..n) (libstdc++ bits/stl_algo.h:4581-4597: element k swaps with element jv[k] = rand() % (k+1),
// k = 1..n-1), in order: a serial chain through shared memory, with the next partner index and v[k+1] (untouched until step k+1) fetched ahead.
__device__ __forceinline__ void shuffle_swaps(int* v, const int* jv, int n) {
  if (n <= 1) return;
  int jn = jv[1], vkn = v[1];
  for (int k = 1; k < n; k++) {
    const int j = jn, vk = vkn;
    if (k + 1 < n) { jn = jv[k + 1]; vkn = v[k + 1]; }
    const int vj = v[j];
    v[k] = vj; v[j] = vk;
  }
}
// (Measured and dropped: requesting the partner value of step k + 1 before the stores of step k and patching it from registers when step k wrote that
// position -- the index of the partner then becomes the chain, and the patch costs what the overlap saves: 43-47 cycles per step either way.)

// glibc_rand_fill (geometry.cuh) by many threads.  The generator is linear over Z / 2^32: 31 draws map the (rotated) ring r to T r, and kRandJumpBlocks
// such blocks to M r with M = T^kRandJumpBlocks -- a 31 x 31 matrix the host computes once (g_rand_jump, row-major, rows padded to 32).  Warp 0
// steps the ring from chunk to chunk of kRandChunk draws (31 multiply-adds per lane and chunk, operands by shuffle) and parks every chunk's start
// ring where that chunk's draws will go; then thread c produces chunk c from registers, exactly like the serial routine.  Same draws, same final
// ring / indices as n calls of rand().
constexpr int kRandJumpBlocks = 8, kRandChunk = 31 * kRandJumpBlocks;
__device__ uint32_t g_rand_jump[31 * 32];
__device__ void glibc_rand_fill_parallel(int* ring, int& f, int& b, int* out, int n) {
  __shared__ uint32_t s_tail[31];                                   // start ring of a last chunk shorter than 31 draws
  const int tid = threadIdx.x, lane = tid & 31;
  if (n <= 0) return;                                               // (uniform)
  const int b0 = b, C = (n + kRandChunk - 1) / kRandChunk;
  if (tid < 32) {
    uint32_t m[31];
#pragma unroll
    for (int j = 0; j < 31; j++) m[j] = lane < 31 ? g_rand_jump[lane * 32 + j] : 0u;
    int q = b0 + lane; if (q >= 31) q -= 31;
    uint32_t sv = lane < 31 ? (uint32_t)ring[q] : 0u;
    for (int c = 0; c < C; c++) {
      if (lane < 31) { if (c * kRandChunk + 31 <= n) out[c * kRandChunk + lane] = (int)sv; else s_tail[lane] = sv; }
      if (c + 1 < C) {
        uint32_t acc = 0u;
#pragma unroll
        for (int j = 0; j < 31; j++) acc += m[j] * __shfl_sync(0xffffffffu, sv, j);
        sv = acc;
      }
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += kPT) {
    const int base = c * kRandChunk, cnt = min(kRandChunk, n - base);
    uint32_t r[31];
#pragma unroll
    for (int j = 0; j < 31; j++) r[j] = base + 31 <= n ? (uint32_t)out[base + j] : s_tail[j];
    int k = 0;
    for (; cnt - k >= 31; k += 31) {
#pragma unroll
      for (int j = 0; j < 31; j++) { r[(j + 3) % 31] += r[j]; out[base + k + j] = (int)(r[(j + 3) % 31] >> 1); }
    }
    const int rem = cnt - k;
#pragma unroll
    for (int j = 0; j < 31; j++) if (j < rem) { r[(j + 3) % 31] += r[j]; out[base + k + j] = (int)(r[(j + 3) % 31] >> 1); }
    if (c == C - 1) {
#pragma unroll
      for (int j = 0; j < 31; j++) { int q = b0 + j; if (q >= 31) q -= 31; ring[q] = (int)r[j]; }
      b = (b0 + n) % 31; f = (b + 3) % 31;
    }
  }
  __syncthreads();
}

// The same swaps WITHOUT the chain.  Step k (k = 1..n-1, in order) swaps v[k] with v[t_k], t_k <= k; position k is untouched before step k.  Hence
//   * a step with t_k = q < k leaves v0[k] (the original element k) at q: position q ends up holding v0[k*], k* = the LAST step that targets q;
//   * what step k moves INTO position k -- W_k, the value at q = t_k at that time -- is v0[k'] with k' = the last step BEFORE k that targets q, or, if
//     there is none, what q's own step left there: W_q (W_q = v0[q] for t_q = q and for a position without a step);
//   * a position no later step targets keeps W_p.
// So: bucket the steps by target (counting sort, buckets unordered -- a bucket holds ~ln(n / q) steps, scanned linearly), give every step either its
// source element or a link to an earlier step's W (res), and every wanted output position p < m follows at most a few links (a link goes to a uniformly
// drawn earlier position that was not targeted again: ~1 on average).  All threads of the CTA; any number of independent segments at once (a segment's
// first element has t = itself).  tgt (n), ptr (n + 1), ent (n), res (n): scratch, shared or global; tgt is consumed.  Result: v[0..m).
__device__ void shuffle_parallel(int* v, int* tgt, int n, int m, int* ptr, int* ent, int* res) {
  __shared__ int s_part[kPT / 32], s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (n <= 1 || m <= 0) return;                                   // (uniform)
  for (int q = tid; q <= n; q += kPT) ptr[q] = 0;
  __syncthreads();
  for (int k = tid; k < n; k += kPT) { const int q = tgt[k]; if (q != k) atomicAdd(&ptr[q], 1); }
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int b = 0; b <= n; b += kPT) {                             // inclusive prefix sums: ptr[q] = end of bucket q; ptr[n] = number of real steps
    const int q = b + tid;
    const int c = q <= n ? ptr[q] : 0;
    int inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) s_part[warp] = inc;
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; w++) off += s_part[w];
    if (q <= n) ptr[q] = off + inc;
    __syncthreads();
    if (tid == kPT - 1) s_base = off + inc;
  }
  __syncthreads();
  for (int k = tid; k < n; k += kPT) { const int q = tgt[k]; if (q != k) ent[atomicSub(&ptr[q], 1) - 1] = k; }   // afterwards ptr[q] = start of bucket q, ptr[q + 1] its end
  __syncthreads();
  for (int k = tid; k < n; k += kPT) {
    const int q = tgt[k];
    int r = k;                                                     // t_k = k (or no step): W_k = v0[k]
    if (q != k) {
      int best = -1;
      for (int e = ptr[q], e1 = ptr[q + 1]; e < e1; e++) { const int x = ent[e]; if (x < k && x > best) best = x; }
      r = best >= 0 ? best : -(q + 1);                             // the last earlier step onto q, else a link to W_q
    }
    res[k] = r;
    int last = -1;                                                 // the last step that targets position k
    for (int e = ptr[k], e1 = ptr[k + 1]; e < e1; e++) { const int x = ent[e]; if (x > last) last = x; }
    tgt[k] = last;                                                 // (tgt[k] was read by this thread alone)
  }
  __syncthreads();
  for (int p = tid; p < m; p += kPT) {
    int r = tgt[p];
    if (r < 0) { r = res[p]; while (r < 0) r = res[-(r + 1)]; }
    ent[p] = v[r];                                                 // (ent is free: the buckets are not needed any more)
  }
  __syncthreads();
  for (int p = tid; p < m; p += kPT) v[p] = ent[p];
  __syncthreads();
}

// Tracker::ApplyMotionModel (jni/Tracker.cc:781-798): the pose a frame starts from.  sbi (6 doubles or null): Tracker::mv6SBIRot of THIS frame.
__device__ __forceinline__ void motion_model_pose(const Dev& D, const double* velocity, const double* pose, const double* sbi, double* np) {
  double v[6]; for (int k = 0; k < 6; k++) v[k] = velocity[k];
  if (D.prm.use_sbi) { v[0] = 0.0; v[1] = 0.0; v[3] = sbi[3]; v[4] = sbi[4]; v[5] = sbi[5]; }
  double e[12]; se3_exp(v, e); se3_mul(e, pose, np);
}

// Large maps (vs_launch_project_all: more than kSplitProjectN points): the per-point half of k_project_lists -- TrackerData::Project, GetDerivsUnsafe,
// CalcSearchLevelAndWarpMatrix -- on a grid of (point chunks, streams) instead of one CTA per stream walking the whole map (4K, 20000 points: 79 rounds of
// 256 points with three barriers each).  Every CTA derives the frame's start pose from the stream's state with the arithmetic of ApplyMotionModel and
// writes none of it: k_project_lists, launched behind it with pre_projected = 1, installs the same pose (same operations, same bits) and builds the lists
// from the flags / levels left here.
__global__ void __launch_bounds__(kPT) k_project_points(Dev D, int apply_motion, const double* __restrict__ sbi_rot_in) {
  cudaGridDependencySynchronize(); cudaTriggerProgrammaticLaunchCompletion();
  __shared__ double s_pose[12];
  const int s = blockIdx.y + D.s0, tid = threadIdx.x;
  const StreamState* st = D.ss + s;
  if (st->lost_frames >= 3 && !st->recovered) return;
  if (tid == 0) {
    if (apply_motion && !st->recovered) {
      double np[12]; motion_model_pose(D, st->velocity, st->pose, sbi_rot_in ? sbi_rot_in + 6 * (size_t)s : st->sbi_rot, np);
      for (int k = 0; k < 12; k++) s_pose[k] = np[k];
    } else for (int k = 0; k < 12; k++) s_pose[k] = st->pose[k];
  }
  __syncthreads();
  const int i = blockIdx.x * kPT + tid;
  if (i >= D.map.n) return;
  const size_t SN = (size_t)D.S * D.N, gi = (size_t)s * D.N + i;
  int flags = D.ps.flags[gi] | F_HASTD;
  CamCache cc;
  td_project(D, s_pose, i, gi, SN, cc, flags);
  if (flags & F_INIMAGE) {
    double dv[4]; cam_derivs(D.cam, cc, dv);
    for (int k = 0; k < 4; k++) D.ps.derivs[k * SN + gi] = dv[k];
    const int level = calc_level_warp(D, s_pose, i, gi, SN, dv, flags);
    D.ps.level[gi] = level;
    if (level >= 0) flags &= ~(F_SEARCHED | F_FOUND);
  }
  D.ps.flags[gi] = flags;
}

// par_shuffle: 0 = the shuffles as serial swap chains; 1 = shuffle_parallel with its scratch behind the two shared-memory arrays (6 N + 1 ints in all);
// 2 = ... with its scratch in the stream's global scratch (sort scratch + the PVS lists, which are free by then); needs use_smem
__global__ void __launch_bounds__(kPT) k_project_lists(Dev D, int mode, int apply_motion, int use_smem, const double* __restrict__ sbi_rot_in, int pre_projected, int par_shuffle) {
  cudaGridDependencySynchronize(); cudaTriggerProgrammaticLaunchCompletion();   // programmatic dependent launch (vs_launch_pdl); no-ops otherwise
  extern __shared__ int sh_i[];          // [N] packed level lists (L3|L2|L1|L0), [N] random draws
#ifdef VS_PROJ_TIMING   // instrumented build (scratch experiments): cycles per phase of one CTA, printed by stream 7
  long long tm_[12]; int tn_ = 0;
#define PJ_MARK() do { __syncthreads(); if (tn_ < 12) tm_[tn_++] = clock64(); } while (0)
#else
#define PJ_MARK() do { } while (0)
#endif
  __shared__ double s_pose[12];
  __shared__ int s_cnt[kPT / 32][VS_LEVELS], s_run[VS_LEVELS], s_off[VS_LEVELS + 1];
  __shared__ int s_seg[8];
  const int s = blockIdx.x + D.s0, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = D.map.n;
  const size_t SN = (size_t)D.S * D.N;
  StreamState* st = D.ss + s;
  if (tid == 0 && apply_motion) { st->frame_no++; st->kf_request = 0; D.kf_req[s] = 0; }   // mnFrame++ (jni/Tracker.cc:100): every TrackFrame, lost or not
  if (mode == 1 && st->lost_frames >= 3 && !st->recovered) return;   // lost and not relocalised this frame (k_relocalise): nothing to do, jni/Tracker.cc:104,134-140
  if (tid == 0 && apply_motion && !st->recovered) {   // Tracker::ApplyMotionModel (jni/Tracker.cc:781-798); a relocalised stream starts from the recovered pose
    for (int k = 0; k < 12; k++) st->start_pose[k] = st->pose[k];
    if (D.prm.use_sbi && sbi_rot_in) for (int k = 0; k < 6; k++) st->sbi_rot[k] = sbi_rot_in[6 * (size_t)s + k];   // Tracker::mv6SBIRot of THIS frame (k_sbi, front end of the frame)
    double np[12]; motion_model_pose(D, st->velocity, st->start_pose, st->sbi_rot, np);
    for (int k = 0; k < 12; k++) st->pose[k] = np[k];
  }
  __syncthreads();
  if (tid < 12) s_pose[tid] = st->pose[tid];
  if (tid < VS_LEVELS) s_run[tid] = 0;
  if (mode == 1 && tid < VS_LEVELS) { st->attempted[tid] = 0; st->found[tid] = 0; }
  __syncthreads();
  int* pvs = D.pvs + (size_t)s * VS_LEVELS * D.N;
  PJ_MARK();   // 0: motion model done

  if (pre_projected) {
    // k_project_points has been here: a point is potentially visible if it is in the image and its warp was accepted.  Ordered append to the
    // per-level PVS lists (jni/Tracker.cc:391) without a barrier per 256 points: every warp owns a contiguous range of the map, counts its points
    // per level, and -- once the counts of the warps before it are known -- writes them in index order.
    const int C = ((N + kPT - 1) / kPT) * 32, w0 = warp * C, w1 = min(N, w0 + C);
    // (a warp requests the flags / levels of kPvsBatch x 32 points before it looks at any of them: one round trip to L2 per batch, not two per 32 points)
    constexpr int kPvsBatch = 8;
    int cnt[VS_LEVELS] = {0, 0, 0, 0};
    for (int i0 = w0; i0 < w1; i0 += 32 * kPvsBatch) {
      int fl[kPvsBatch], lv[kPvsBatch];
#pragma unroll
      for (int u = 0; u < kPvsBatch; u++) { const int i = i0 + 32 * u + lane; const size_t gi = (size_t)s * D.N + (i < w1 ? i : w0); fl[u] = D.ps.flags[gi]; lv[u] = D.ps.level[gi]; }
#pragma unroll
      for (int u = 0; u < kPvsBatch; u++) {
        const int i = i0 + 32 * u + lane; const int l_ = (i < w1 && (fl[u] & F_INIMAGE)) ? lv[u] : -1;
#pragma unroll
        for (int l = 0; l < VS_LEVELS; l++) cnt[l] += __popc(__ballot_sync(0xffffffffu, l_ == l));
      }
    }
    if (lane == 0) for (int l = 0; l < VS_LEVELS; l++) s_cnt[warp][l] = cnt[l];
    __syncthreads();
    int off[VS_LEVELS];
#pragma unroll
    for (int l = 0; l < VS_LEVELS; l++) { off[l] = 0; for (int w = 0; w < warp; w++) off[l] += s_cnt[w][l]; }
    for (int i0 = w0; i0 < w1; i0 += 32 * kPvsBatch) {
      int fl[kPvsBatch], lv[kPvsBatch];
#pragma unroll
      for (int u = 0; u < kPvsBatch; u++) { const int i = i0 + 32 * u + lane; const size_t gi = (size_t)s * D.N + (i < w1 ? i : w0); fl[u] = D.ps.flags[gi]; lv[u] = D.ps.level[gi]; }
#pragma unroll
      for (int u = 0; u < kPvsBatch; u++) {
        const int i = i0 + 32 * u + lane; const int l_ = (i < w1 && (fl[u] & F_INIMAGE)) ? lv[u] : -1;
#pragma unroll
        for (int l = 0; l < VS_LEVELS; l++) {
          const unsigned bal = __ballot_sync(0xffffffffu, l_ == l);
          if (l_ == l) pvs[l * D.N + off[l] + __popc(bal & ((1u << lane) - 1u))] = i;
          off[l] += __popc(bal);
        }
      }
    }
    if (tid < VS_LEVELS) { int a = 0; for (int w = 0; w < kPT / 32; w++) a += s_cnt[w][tid]; s_run[tid] = a; }
    __syncthreads();
  }
  for (int base = 0; base < N && !pre_projected; base += kPT) {
    const int i = base + tid;
    int level = -1; bool pv = false;
    if (i < N) {
      const size_t gi = (size_t)s * D.N + i;
      int flags = D.ps.flags[gi] | F_HASTD;
      if (mode == 0) { flags &= ~(F_SEARCHED | F_FOUND | F_SUBPIX); }
      CamCache cc;
      td_project(D, s_pose, i, gi, SN, cc, flags);
      if (flags & F_INIMAGE) {
        double dv[4]; cam_derivs(D.cam, cc, dv);
        for (int k = 0; k < 4; k++) D.ps.derivs[k * SN + gi] = dv[k];
        level = calc_level_warp(D, s_pose, i, gi, SN, dv, flags);
        D.ps.level[gi] = level;
        if (level >= 0) { pv = true; flags &= ~(F_SEARCHED | F_FOUND); }
      } else if (mode == 0) D.ps.level[gi] = -1;
      D.ps.flags[gi] = flags;
    }
    if (mode == 1) {   // ordered append to the per-level PVS lists (jni/Tracker.cc:391)
      unsigned bal[VS_LEVELS];
#pragma unroll
      for (int l = 0; l < VS_LEVELS; l++) { bal[l] = __ballot_sync(0xffffffffu, pv && level == l); if (lane == 0) s_cnt[warp][l] = __popc(bal[l]); }
      __syncthreads();
      if (pv) {
        int off = s_run[level];
        for (int w = 0; w < warp; w++) off += s_cnt[w][level];
        off += __popc(bal[level] & ((1u << lane) - 1u));
        pvs[level * D.N + off] = i;
      }
      __syncthreads();
      if (tid < VS_LEVELS) { int a = s_run[tid]; for (int w = 0; w < kPT / 32; w++) a += s_cnt[w][tid]; s_run[tid] = a; }
      __syncthreads();
    }
  }
  if (mode == 0) return;
  PJ_MARK();   // 1: projection / PVS rounds

  // packed layout in shared memory: [L3][L2][L1][L0]
  if (tid == 0) { s_off[0] = 0; s_off[1] = s_run[3]; s_off[2] = s_off[1] + s_run[2]; s_off[3] = s_off[2] + s_run[1]; s_off[4] = s_off[3] + s_run[0]; }
  __syncthreads();
  // (maps too large for 2 x N ints of shared memory keep the two arrays in the stream's global sort scratch: sort_cap doubles >= 2 N ints)
  int* list = use_smem ? sh_i : (int*)(D.sort_scratch + (size_t)s * D.sort_cap); int* rnd = list + D.N;
  for (int q = 0; q < VS_LEVELS; q++) {   // q-th packed segment holds level 3-q
    const int l = 3 - q, n = s_run[l];
    for (int k = tid; k < n; k += kPT) list[s_off[q] + k] = pvs[l * D.N + k];
  }
  const int total = s_off[4];
  // random draws for the four level shuffles, in level order 0..3 (jni/Tracker.cc:396-397)
  int n_draws = 0;
#pragma unroll
  for (int l = 0; l < VS_LEVELS; l++) n_draws += max(s_run[l] - 1, 0);
  PJ_MARK();   // 2: packed copy
  if (par_shuffle) glibc_rand_fill_parallel(st->rng_ring, st->rng_f, st->rng_b, rnd, n_draws);
  else if (tid == 0) glibc_rand_fill(st->rng_ring, st->rng_f, st->rng_b, rnd, n_draws);
  __syncthreads();
  PJ_MARK();   // 3: rand fill
  // draw -> swap partner of std::random_shuffle (libstdc++ bits/stl_algo.h:4581-4597: i-th element swaps with rand() % (i+1)), all threads
  for (int d = tid; d < n_draws; d += kPT) {
    int k = d;
#pragma unroll
    for (int l = 0; l < VS_LEVELS - 1; l++) { const int m = max(s_run[l] - 1, 0); if (k >= 0 && k >= m) k -= m; else break; }
    rnd[d] = rnd[d] % (k + 2);
  }
  __syncthreads();
  PJ_MARK();   // 4: modulo
  int* const sc_tgt = par_shuffle == 1 ? sh_i + 2 * D.N : (int*)(D.sort_scratch + (size_t)s * D.sort_cap);
  int* const sc_ptr = par_shuffle == 1 ? sc_tgt + D.N : pvs;       // (the PVS lists have been copied into `list`)
  int* const sc_ent = sc_ptr + D.N + 1; int* const sc_res = sc_ent + D.N;
  if (par_shuffle == 1) {
    // the four level shuffles as ONE shuffle_parallel problem over the packed list: the partner of packed position pos = s_off[q] + k (level 3 - q,
    // k >= 1) is s_off[q] + rnd[draws of the lower levels + k - 1]; a level's first element has no step.  Only with the scratch in shared memory
    // (5000 points: 60 k cycles against 96 k for the longest level's chain; 1000 points: 14 k against 19.5 k); with the scratch in global memory the
    // dozen passes over all levels wait for L2 and lose against the chains (20000 points: 642 k against 392 k).
    for (int pos = tid; pos < total; pos += kPT) {
      const int q = pos >= s_off[3] ? 3 : (pos >= s_off[2] ? 2 : (pos >= s_off[1] ? 1 : 0)), l = 3 - q, k = pos - s_off[q];
      int roff = 0; for (int u = 0; u < l; u++) roff += max(s_run[u] - 1, 0);
      sc_tgt[pos] = k == 0 ? pos : s_off[q] + rnd[roff + k - 1];
    }
    __syncthreads();
    shuffle_parallel(list, sc_tgt, total, total, sc_ptr, sc_ent, sc_res);
  } else
  if (lane == 0 && warp < VS_LEVELS) {   // the swaps of level `warp`, in order
    const int l = warp, n = s_run[l];
    int roff = 0; for (int k = 0; k < l; k++) roff += max(s_run[k] - 1, 0);
    int* v = list + s_off[3 - l];
    const int* jv = rnd + roff - 1;
    shuffle_swaps(v, jv, n);
  }
  __syncthreads();
  PJ_MARK();   // 5: level swaps
  if (tid == 0) {   // coarse / fine selection (jni/Tracker.cc:399-527)
    const int n3 = s_run[3], n2 = s_run[2];
    unsigned nCoarseMax = D.prm.coarse_max, nCoarseRange = D.prm.coarse_range;
    bool tryCoarse = !(st->msd_scaled_vel < D.prm.coarse_min_vel || nCoarseMax == 0);
    if (st->just_recovered) { tryCoarse = true; nCoarseMax *= 2; nCoarseRange *= 2; st->just_recovered = 0; }
    int a0 = 0, a1 = 0, t0 = 0, t1 = n3, f0 = n3;   // A = [a0,a1), top = [t0,t1), fine = [f0,total)
    if (tryCoarse && (unsigned)(n3 + n2) > D.prm.coarse_min) {
      if ((unsigned)n3 > nCoarseMax) { a0 = 0; a1 = (int)nCoarseMax; t0 = a1; t1 = n3; f0 = n3; }
      else {
        a0 = 0; a1 = n3; t0 = t1 = 0; f0 = n3;
        if ((unsigned)n3 < nCoarseMax) {
          const int more = (int)nCoarseMax - n3;
          if (n2 <= more) { a0 = n3; a1 = n3 + n2; f0 = n3 + n2; }    // vNextToSearch overwritten by the L2 list (quirk, :454-456)
          else { a1 = n3 + more; f0 = n3 + more; }
        }
      }
    } else tryCoarse = false;
    st->try_coarse = tryCoarse ? 1 : 0; st->coarse_range = (int)nCoarseRange; st->did_coarse = 0;
    const int nFine = total - f0;
    int use = D.prm.max_patches_per_frame - ((a1 - a0) + (t1 - t0));
    if (use < 0) use = 0;
    s_seg[0] = a0; s_seg[1] = a1 - a0; s_seg[2] = t0; s_seg[3] = t1 - t0; s_seg[4] = f0; s_seg[5] = nFine;
    s_seg[6] = nFine > use ? 1 : 0; s_seg[7] = use;
  }
  __syncthreads();
  if (s_seg[6]) {
    // more fine candidates than MaxPatchesPerFrame allows: the fifth std::random_shuffle of the frame, over the whole fine list, then
    // truncation (jni/Tracker.cc:518-527).  Same three steps as the level shuffles: draws in bulk (one thread, ring in registers),
    // `% (k+1)` by all threads, then the swap chain -- the only serial part -- with the next partner fetched ahead.
    const int nF = s_seg[5];
    PJ_MARK();   // 6: selection
    if (par_shuffle) glibc_rand_fill_parallel(st->rng_ring, st->rng_f, st->rng_b, rnd, nF - 1);
    else if (tid == 0) glibc_rand_fill(st->rng_ring, st->rng_f, st->rng_b, rnd, nF - 1);
    __syncthreads();
    PJ_MARK();   // 7: fifth rand fill
    if (par_shuffle) {   // only the first `use` positions of the shuffled fine list are kept
      for (int k = tid; k < nF; k += kPT) sc_tgt[k] = k == 0 ? 0 : rnd[k - 1] % (k + 1);
      __syncthreads();
      shuffle_parallel(list + s_seg[4], sc_tgt, nF, min(nF, s_seg[7]), sc_ptr, sc_ent, sc_res);
    } else {
      for (int d = tid; d < nF - 1; d += kPT) rnd[d] = rnd[d] % (d + 2);
      __syncthreads();
      if (tid == 0) shuffle_swaps(list + s_seg[4], rnd - 1, nF);
    }
    __syncthreads();
    PJ_MARK();   // 8: fifth modulo + swaps
  }
  if (tid == 0) {
    const int nFine = s_seg[6] ? s_seg[7] : s_seg[5];
    s_seg[5] = nFine;
    st->nA = s_seg[1]; st->nB_top = s_seg[3]; st->nB = s_seg[3] + nFine; st->n_updates = 0;
  }
  __syncthreads();
  int* out = D.lists + (size_t)s * D.list_cap;
  int dst = 0;
  for (int g = 0; g < 3; g++) {
    const int src0 = s_seg[2 * g], n = s_seg[2 * g + 1];
    for (int k = tid; k < n; k += kPT) out[dst + k] = list[src0 + k];
    dst += n;
  }
#ifdef VS_PROJ_TIMING
  PJ_MARK();
  if (tid == 0 && s == 7) { printf("k_project_lists cycles (N=%d total=%d):", N, total); for (int q = 1; q < tn_; q++) printf(" %lld", tm_[q] - tm_[q - 1]); printf("\n"); }
#endif
}

// mode 0: explicit list [0,nA) with (range, subpix) arguments; 1: coarse set A; 2: fine set B
// sflags (mode 0): kSearchRefind = MapMaker::ReFind_Common's variant (jni/MapMaker.cc:967-1036): a cold PatchFinder per point (the
// template is always regenerated and its bad flag comes from the warp alone), the level the warp loop reached even for rejected
// warps, sub-pixel refinement only on levels > 0 and its position kept whether or not it converged.
// PT: compile-time template side (8 or 11), 0 = use the run-time D.P
constexpr int kSearchRefind = 1;
template <int PT>
__global__ void __launch_bounds__(kSearchWarps * 32, 10) k_search(Dev D, int mode, int range_arg, int subpix_arg, int sflags) {
  __shared__ SearchSmem sm_all[kSearchWarps];
  const int s = blockIdx.y + D.s0, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * kSearchWarps + warp;
  StreamState* st = D.ss + s;
  if (mode != 0 && st->lost_frames >= 3 && !st->recovered) return;
  int first, count, range, subpix;
  if (mode == 0) { first = 0; count = st->nA; range = range_arg; subpix = subpix_arg; }
  else if (mode == 1) { if (!st->try_coarse) return; first = 0; count = st->nA; range = st->coarse_range; subpix = D.prm.coarse_subpix_its; }
  else { first = st->nA; count = st->nB; range = st->did_coarse ? D.prm.fine_range_after_coarse : D.prm.fine_range; subpix = (e < st->nB_top) ? D.prm.fine_subpix_its_top_level : 0; }
  if (e >= count) return;
  SearchSmem& sm = sm_all[warp];
  const int i = D.lists[(size_t)s * D.list_cap + first + e];
  const size_t SN = (size_t)D.S * D.N, gi = (size_t)s * D.N + i;
  const int P = PT ? PT : D.P, PP = P * P;
  uint8_t* const tmpl = (uint8_t*)sm.tmpl_w;
  int flags = D.ps.flags[gi];
  const bool refind = (sflags & kSearchRefind) != 0;
  const int level = refind ? D.ps.rlevel[gi] : D.ps.level[gi];
  if (refind) {
    flags &= ~(F_SEARCHED | F_FOUND | F_SUBPIX);
    if (!(flags & F_INIMAGE)) { if (lane == 0) D.ps.flags[gi] = flags; return; }   // not in this keyframe's image: "never retry"
    if (level == 0) subpix = 0;
  }
  // (the fine set was re-projected by k_reproject_fine if the coarse stage moved the pose)

  // ---- MakeTemplateCoarseCont (jni/PatchFinder.cc:79-125)
  double m2[4]; int refresh = 0, inside = 0, tsum = 0, tsumsq = 0;
  // every per-point scalar this warp will need, loaded up front (uniform addresses: one transaction each) so the misses overlap
  m2[0] = D.ps.m2[gi]; m2[1] = D.ps.m2[SN + gi]; m2[2] = D.ps.m2[2 * SN + gi]; m2[3] = D.ps.m2[3 * SN + gi];
  const double lw0 = D.ps.lastwarp[gi], lw1 = D.ps.lastwarp[SN + gi], lw2 = D.ps.lastwarp[2 * SN + gi], lw3 = D.ps.lastwarp[3 * SN + gi];
  const double v2i0 = D.ps.v2image[gi], v2i1 = D.ps.v2image[SN + gi];
  const int tsum_old = D.ps.tsum[gi], tsumsq_old = D.ps.tsum[SN + gi];
  const uint32_t* tmpl_g = (const uint32_t*)(D.ps.tmpl + gi * VS_TMPL_BYTES);
  const uint32_t tmpl_old = tmpl_g[lane], tmpl_old2 = lane < VS_TMPL_BYTES / 4 - 32 ? tmpl_g[32 + lane] : 0u;   // coalesced
  if (lane == 0) {
    refresh = refind || !(flags & F_HAVELAST);
    // (the two columns written out: indexing m2 with a loop variable would put it in local memory -- 28 sectors of stores per warp)
    if (!refresh) { const double d0 = m2[0] - lw0, d1 = m2[2] - lw2; double dd = 0; dd += d0 * d0; dd += d1 * d1; if (dd > 0.07 * 0.07) refresh = 1; }
    if (!refresh) { const double d0 = m2[1] - lw1, d1 = m2[3] - lw3; double dd = 0; dd += d0 * d0; dd += d1 * d1; if (dd > 0.07 * 0.07) refresh = 1; }
  }
  refresh = __shfl_sync(0xffffffffu, refresh, 0);
  const int kf = D.map.srckf[i], sl = D.map.srclevel[i];
  if (refresh) {
    const uint8_t* simg = D.src.img[sl] + (size_t)kf * D.src.h[sl] * D.src.pitch[sl];
    const int iw = D.src.w[sl], ih = D.src.h[sl], sp = D.src.pitch[sl];
    if (lane == 0) {   // transform_image (jni/vision/ImageHandler.cpp:21-113): sample positions by sequential accumulation
      const double across0 = m2[0], across1 = m2[2], down0 = m2[1], down1 = m2[3];
      const double o = (double)(P / 2);
      double a = m2[0] * o; a += m2[1] * o; double b = m2[2] * o; b += m2[3] * o;
      const double p00 = (double)D.map.ircenter[2 * i] - a, p01 = (double)D.map.ircenter[2 * i + 1] - b;
      double min_x = p00, min_y = p01, max_x = min_x, max_y = min_y;
      if (across0 < 0) min_x += P * across0; else max_x += P * across0;
      if (down0 < 0) min_x += P * down0; else max_x += P * down0;
      if (across1 < 0) min_y += P * across1; else max_y += P * across1;
      if (down1 < 0) min_y += P * down1; else max_y += P * down1;
      const double cr0 = down0 - P * across0, cr1 = down1 - P * across1;
      inside = (min_x >= 0 && min_y >= 0 && max_x < iw - 1 && max_y < ih - 1);
      double px = p00, py = p01;
      for (int r = 0; r < P; ++r, px += cr0, py += cr1)
        for (int c = 0; c < P; ++c, px += across0, py += across1) { sm.pos[2 * (r * P + c)] = px; sm.pos[2 * (r * P + c) + 1] = py; }
    }
    inside = __shfl_sync(0xffffffffu, inside, 0);
    __syncwarp();
    const float x_bound = iw - 1, y_bound = ih - 1;
    int outside = 0;
    sm.tmpl_w[lane] = 0u; if (lane < VS_TMPL_BYTES / 4 - 32) sm.tmpl_w[32 + lane] = 0u;   // row padding must be zero
    __syncwarp();
    for (int k = lane; k < PP; k += 32) {
      double x = sm.pos[2 * k], y = sm.pos[2 * k + 1];
      uint8_t v = 0;
      if (inside || (0 <= x && 0 <= y && x < x_bound && y < y_bound)) {   // sample(u8) (jni/vision/ImageHandler.cpp:12-19)
        const int lx = (int)x, ly = (int)y;
        x -= lx; y -= ly;
        const uint8_t* r0 = simg + (size_t)ly * sp + lx; const uint8_t* r1 = r0 + sp;
        v = (uint8_t)((1 - y) * ((1 - x) * r0[0] + x * r0[1]) + y * ((1 - x) * r1[0] + x * r1[1]));
      } else outside++;
      const int r = k / P;
      tmpl[k + r * (12 - P)] = v;
    }
    outside = warp_sum(outside);
    __syncwarp();
    int ts = 0, tq = 0;
    for (int k = lane; k < 3 * P; k += 32) { const uint32_t w = sm.tmpl_w[k]; ts += (int)__dp4a(w, 0x01010101u, 0u); tq += (int)__dp4a(w, w, 0u); }   // padding is zero
    { uint32_t* g = (uint32_t*)(D.ps.tmpl + gi * VS_TMPL_BYTES); g[lane] = sm.tmpl_w[lane]; if (lane < VS_TMPL_BYTES / 4 - 32) g[32 + lane] = sm.tmpl_w[32 + lane]; }
    ts = warp_sum(ts); tq = warp_sum(tq);
    tsum = ts; tsumsq = tq;
    flags = outside ? (flags | F_TBAD) : (flags & ~F_TBAD);
    flags |= F_HAVELAST | F_NEWTMPL;
    if (lane == 0) {
      atomicAdd(D.evals + 2, 1ull);
      D.ps.tsum[gi] = ts; D.ps.tsum[SN + gi] = tq;
      D.ps.lastwarp[gi] = m2[0]; D.ps.lastwarp[SN + gi] = m2[1]; D.ps.lastwarp[2 * SN + gi] = m2[2]; D.ps.lastwarp[3 * SN + gi] = m2[3];
    }
  } else {
    flags &= ~F_NEWTMPL;
    sm.tmpl_w[lane] = tmpl_old; if (lane < VS_TMPL_BYTES / 4 - 32) sm.tmpl_w[32 + lane] = tmpl_old2;
    tsum = tsum_old; tsumsq = tsumsq_old;
  }
  __syncwarp();
  if (flags & F_TBAD) {   // jni/Tracker.cc:637-640
    if (lane == 0) D.ps.flags[gi] = flags & ~(F_INIMAGE | F_FOUND);
    return;
  }
  if (lane == 0) atomicAdd(&st->attempted[level], 1);

  // ---- FindPatchCoarse (jni/PatchFinder.cc:170-235)
  const LevelDesc& L = D.lev[level];
  const uint8_t* img; int pitch;
  if (level == 0) { img = D.l0_ptr[s]; pitch = D.l0_stride[s]; } else { img = L.img + (size_t)s * L.h * L.pitch; pitch = L.pitch; }
  const int maxSSD = PP * 500;
  const int nLevelScale = LevelScale(level);
  const double invScale = 1.0 / nLevelScale;                    // 2^-level: x / 2^l == x * 2^-l exactly
  const double ix = v2i0 * invScale, iy = v2i1 * invScale;
  const unsigned nRange = ((unsigned)range + nLevelScale - 1) / nLevelScale;
  int nTop = iy - nRange;
  const int nBottomPlusOne = iy + nRange + 1;
  const int nLeft = ix - nRange, nRight = ix + nRange;
  unsigned long long best = ((unsigned long long)(unsigned)(maxSSD + 1) << 32) | 0xffffffffull;
  flags |= F_SEARCHED;
  bool searched_any = true;
  if (nTop < 0) nTop = 0;
  if (nTop >= L.h || nBottomPlusOne <= 0) searched_any = false;
  int nevals = 0;
  if (searched_any) {
    const int* lut = L.lut + (size_t)s * (L.h + 1);
    const int begin = lut[nTop], end = (nBottomPlusOne >= L.h) ? lut[L.h] : lut[nBottomPlusOne];
    const uint32_t* corners = L.corners + (size_t)s * L.cap;
    int c0 = begin;
    // the corner words of the next two 32-corner steps are requested before the current step is examined (the kernel is bound by
    // the latency of these dependent-looking, in fact independent, loads)
    uint32_t cw1 = (c0 + lane < end) ? __ldg(corners + c0 + lane) : 0u, cw2 = (c0 + 32 + lane < end) ? __ldg(corners + c0 + 32 + lane) : 0u;
    while (c0 < end) {   // rounds: gather up to kCandCap candidates, then score them with (candidate,row) work items
      int ncand = 0;
      for (; c0 < end && ncand <= kCandCap - 32; c0 += 32) {
        const int ci = c0 + lane;
        bool pass = false; const uint32_t cw = cw1;
        cw1 = cw2; cw2 = (c0 + 64 + lane < end) ? __ldg(corners + c0 + 64 + lane) : 0u;
        if (ci < end) {
          const int cx = cw & 0xffff, cy = cw >> 16;
          pass = !((double)cx < nLeft || (double)cx > nRight);
          if (pass) { const double dx = ix - (double)cx, dy = iy - (double)cy; double d2 = 0; d2 += dx * dx; d2 += dy * dy; pass = !(d2 > nRange * nRange); }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        if (pass) { const int slot = ncand + __popc(bal & ((1u << lane) - 1u)); sm.cand_cw[slot] = cw; sm.cand_idx[slot] = ci; }
        ncand += __popc(bal);
      }
      nevals += ncand;   // (every lane holds the same count; reduced once below)
      __syncwarp();
      best = score_candidates<PT>(sm, ncand, img, pitch, L.w, L.h, P, tsum, tsumsq, maxSSD, best);
      __syncwarp();
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d); best = o < best ? o : best; }
  if (lane == 0 && nevals) atomicAdd(D.evals, (unsigned long long)nevals);
  const int bestSSD = (int)(best >> 32);
  if (!(bestSSD < maxSSD)) {
    if (lane == 0) D.ps.flags[gi] = flags & ~F_FOUND;
    return;
  }
  const uint32_t bc = (L.corners + (size_t)s * L.cap)[(unsigned)best];
  const double coarse0 = ((double)(bc & 0xffff) + 0.5) * nLevelScale - 0.5, coarse1 = ((double)(bc >> 16) + 0.5) * nLevelScale - 0.5;  // LevelZeroPos
  flags |= F_FOUND;
  double found0 = coarse0, found1 = coarse1;
  int ok = 1;
  if (subpix > 0) {
    flags |= F_SUBPIX;
    if (lane == 0) atomicAdd(D.evals + 3, 1ull);
    ok = subpix_refine(sm, tmpl, img, pitch, L.w, L.h, level, P, subpix, coarse0, coarse1, found0, found1);
  }
  if (lane == 0) {
    D.ps.coarse[gi] = coarse0; D.ps.coarse[SN + gi] = coarse1;
    D.ps.sqrtinv[gi] = invScale;
    if (ok || refind) { D.ps.v2found[gi] = found0; D.ps.v2found[SN + gi] = found1; atomicAdd(&st->found[level], 1); if (subpix <= 0) flags &= ~F_SUBPIX; }
    else flags &= ~F_FOUND;   // sub-pixel iteration did not converge (jni/Tracker.cc:660-666)
    D.ps.flags[gi] = flags;
  }
}

// ------------------------------------------------------------------------------------------------
// The search of MapMaker::AddPointEpipolar (jni/MapMaker.cc:525-640), one warp per candidate of the source keyframe: un-warped
// template (PatchFinder::MakeTemplateCoarseNoWarp, jni/PatchFinder.cc:130-143), every FAST corner of the target level whose
// image-plane position (the reference's per-pixel UnProject table, index = truncated level-zero position) lies in the epipolar
// band, ZMSSD, best candidate (first wins ties; accepted up to and including mnMaxSSD), then MakeSubPixTemplate +
// IterateSubPixToConvergence(10).  The per-candidate line geometry is computed on the host (vslam_epipolar_search, api.cu).
__global__ void __launch_bounds__(kSearchWarps * 32) k_epipolar(Dev D, int s, int src_kf, int level, int n, const EpiCand* __restrict__ cand,
                                                                const double* __restrict__ unproj, int subpix_its, int* __restrict__ out_int, double* __restrict__ out_pos) {
  __shared__ SearchSmem sm_all[kSearchWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * kSearchWarps + warp;
  if (e >= n) return;
  SearchSmem& sm = sm_all[warp];
  const EpiCand C = cand[e];
  const int P = D.P, PP = P * P, maxSSD = PP * 500;
  int found = 0, best_idx = -1, best_ssd = maxSSD + 1; double pos0 = 0, pos1 = 0;
  const int sw = D.src.w[level], sh = D.src.h[level], sp = D.src.pitch[level];
  const int bd = P / 2 + 1;
  const bool tmpl_ok = C.valid && C.x >= bd && C.y >= bd && C.x < sw - bd && C.y < sh - bd;
  if (tmpl_ok) {
    const uint8_t* simg = D.src.img[level] + (size_t)src_kf * sh * sp;
    uint8_t* const tmpl = (uint8_t*)sm.tmpl_w;
    sm.tmpl_w[lane] = 0u; if (lane < VS_TMPL_BYTES / 4 - 32) sm.tmpl_w[32 + lane] = 0u;
    __syncwarp();
    for (int k = lane; k < PP; k += 32) { const int r = k / P, c = k - r * P; tmpl[r * 12 + c] = simg[(size_t)(C.y - P / 2 + r) * sp + (C.x - P / 2 + c)]; }
    __syncwarp();
    int ts = 0, tq = 0;
    for (int k = lane; k < 3 * P; k += 32) { const uint32_t w = sm.tmpl_w[k]; ts += (int)__dp4a(w, 0x01010101u, 0u); tq += (int)__dp4a(w, w, 0u); }
    ts = warp_sum(ts); tq = warp_sum(tq);
    const LevelDesc& L = D.lev[level];
    const uint8_t* img; int pitch;
    if (level == 0) { img = D.l0_ptr[s]; pitch = D.l0_stride[s]; } else { img = L.img + (size_t)s * L.h * L.pitch; pitch = L.pitch; }
    const int* lut = L.lut + (size_t)s * (L.h + 1);
    const uint32_t* corners = L.corners + (size_t)s * L.cap;
    const int end = lut[L.h], W0 = D.lev[0].w, scale = LevelScale(level);
    unsigned long long best = ((unsigned long long)(unsigned)(maxSSD + 1) << 32) | 0xffffffffull;
    int c0 = 0;
    while (c0 < end) {
      int ncand = 0;
      for (; c0 < end && ncand <= kCandCap - 32; c0 += 32) {
        const int ci = c0 + lane;
        bool pass = false; uint32_t cw = 0;
        if (ci < end) {
          cw = corners[ci];
          const int cx = cw & 0xffff, cy = cw >> 16;
          const int zx = (int)(((double)cx + 0.5) * scale - 0.5), zy = (int)(((double)cy + 0.5) * scale - 0.5);   // LevelZeroPos, truncated by the table index
          const double ux = unproj[2 * ((size_t)zy * W0 + zx)], uy = unproj[2 * ((size_t)zy * W0 + zx) + 1];
          double dn = ux * C.nx; dn += uy * C.ny;
          const double dDistDiff = C.normDist - dn;
          double da = ux * C.ax; da += uy * C.ay;
          pass = !(dDistDiff * dDistDiff > C.maxDistSq) && !(da < C.minLen) && !(da > C.maxLen);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        if (pass) { const int slot = ncand + __popc(bal & ((1u << lane) - 1u)); sm.cand_cw[slot] = cw; sm.cand_idx[slot] = ci; }
        ncand += __popc(bal);
      }
      __syncwarp();
      best = score_candidates<0>(sm, ncand, img, pitch, L.w, L.h, P, ts, tq, maxSSD, best);
      __syncwarp();
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d); best = o < best ? o : best; }
    best_ssd = (int)(best >> 32);
    if (best_ssd < maxSSD + 1) {   // nBest != -1
      best_idx = (int)(unsigned)best;
      const uint32_t bc = corners[best_idx];
      const double c0d = ((double)(bc & 0xffff) + 0.5) * scale - 0.5, c1d = ((double)(bc >> 16) + 0.5) * scale - 0.5;
      found = subpix_refine(sm, tmpl, img, pitch, L.w, L.h, level, P, subpix_its, c0d, c1d, pos0, pos1);
    }
  }
  if (lane == 0) { out_int[3 * e] = found; out_int[3 * e + 1] = best_idx; out_int[3 * e + 2] = best_ssd; out_pos[2 * e] = pos0; out_pos[2 * e + 1] = pos1; }
}

// ------------------------------------------------------------------------------------------------
// Pose-update kernel, one CTA per stream.
// k-th smallest (0-based) of n non-negative doubles: MSB-first 11-bit radix select on the IEEE bit patterns (monotonic for
// values >= 0), stopping as soon as the selected bin holds a single element.  All threads of the CTA call it;
// `hist` is 2048 ints of shared memory, `sel` three 64-bit words.
__device__ inline double block_radix_select(const double* a, int n, int k, int* hist, unsigned long long* sel) {
  const int tid = threadIdx.x;
  unsigned long long prefix = 0;   // bits decided so far
  int rank = k, sh = 64;
  while (sh > 0) {
    const int bits = sh >= 11 ? 11 : sh, nsh = sh - bits, nb = 1 << bits;
    for (int b = tid; b < nb; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    for (int t = tid; t < n; t += blockDim.x) {
      const unsigned long long key = (unsigned long long)__double_as_longlong(a[t]);
      if (sh == 64 || (key >> sh) == (prefix >> sh)) atomicAdd(&hist[(int)((key >> nsh) & (unsigned long long)(nb - 1))], 1);
    }
    __syncthreads();
    if (tid < 32) {   // warp 0: find the bin that holds rank `rank`
      const int per = nb >> 5;
      int mine = 0;
      for (int q = 0; q < per; q++) mine += hist[tid * per + ((q + tid) & (per - 1))];   // rotated start per lane: no bank conflicts
      int incl = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (tid >= d) incl += v; }
      const int excl = incl - mine;
      if (rank >= excl && rank < incl) {
        int r = rank - excl, q = 0;
        while (r >= hist[tid * per + q]) { r -= hist[tid * per + q]; q++; }
        sel[0] = prefix | ((unsigned long long)(tid * per + q) << nsh); sel[1] = (unsigned long long)r; sel[2] = (unsigned long long)hist[tid * per + q];
      }
    }
    __syncthreads();
    prefix = sel[0]; rank = (int)sel[1];
    const bool single = sel[2] == 1ull;
    sh = nsh;
    __syncthreads();
    if (single && sh > 0) {   // exactly one element carries this prefix: fetch it
      for (int t = tid; t < n; t += blockDim.x) {
        const unsigned long long key = (unsigned long long)__double_as_longlong(a[t]);
        if ((key >> sh) == (prefix >> sh)) sel[0] = key;
      }
      __syncthreads();
      prefix = sel[0];
      __syncthreads();
      break;
    }
  }
  return __longlong_as_double((long long)prefix);
}

#ifdef VS_POSE_TIMING
#define PT_MARK(k) do { __syncthreads(); const long long t_ = clock64(); const int z_ = threadIdx.x == 0 ? 0 : 16; const long long d_ = t_ - sm.tlast[z_ ? 1 : 0]; sm.tacc[(k) + z_] += d_; __syncthreads(); sm.tlast[z_ ? 1 : 0] = t_; } while (0)
#else
#define PT_MARK(k) do { } while (0)
#endif
struct PoseSmem {
#ifdef VS_POSE_TIMING
  long long tacc[32], tlast[2];   // [0..15]: thread 0's phase counters, [16..31] and tlast[1]: dummy slots for the other threads (branch-free marks)
#endif
  double pose[12];
  double red[kPT / 32][28];
  double sums[28];
  double mu[6], last[6];
  double sigma;
  int nerr, cnt;
  unsigned long long sel[3];
  int warpcnt[kPT / 32];
  double lu[36], inv[36]; int piv[6];
};

// One CalcPoseUpdate over list entries [0,n) (jni/Tracker.cc:683-774).  Leaves mu in sm.mu (zeros if nothing found).
__device__ void calc_pose_update(const Dev& D, PoseSmem& sm, double* sortbuf, int sortcap, int* hist, double* part, const int* list, int n, int s, double overrideSigma, bool mark) {
  const size_t SN = (size_t)D.S * D.N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // the errors (:703-709) and the squared errors for the median were written by pose_points_pass
  const int nerr = n;
  if (nerr == 0) { if (tid < 6) sm.mu[tid] = 0.0; if (tid == 0) sm.sigma = 0.0; __syncthreads(); return; }
  if (overrideSigma > 0) { if (tid == 0) sm.sigma = overrideSigma; }
  else {   // Tukey::FindSigmaSquared (jni/MEstimator.h:67-77): the sort there only serves to pick v[n/2]
    const double med = block_radix_select(sortbuf, nerr, nerr / 2, hist, sm.sel);
    if (tid == 0) {
      const unsigned long long den = (unsigned long long)nerr * 2ull - 6ull;   // size_t arithmetic of the reference
      double sigma = 1.4826 * (1 + 5.0 / (double)den) * sqrt(med);
      sigma = 4.6851 * sigma;
      sm.sigma = sigma * sigma;
    }
  }
  __syncthreads();
  PT_MARK(5);
  const double sig2 = sm.sigma;
  // weighted normal equations: 21 upper-triangle terms + 6 right-hand sides (jni/myWLS.h:39-50)
  double acc[27];
#pragma unroll
  for (int k = 0; k < 27; k++) acc[k] = 0.0;
  for (int k = tid; k < n; k += kPT) {
    const int i = list[k];
    const size_t gi = (size_t)s * D.N + i;
    // everything the point needs is requested before the weight is known (one round trip to L2 instead of two or three; a point
    // whose weight turns out to be zero wastes its 12 Jacobian loads, which is rare after the first iterations).  Requesting the
    // NEXT point's values as well was measured slower (register spills at the 128-register limit).
    const double e0 = D.ps.err[gi], e1 = D.ps.err[SN + gi], si = D.ps.sqrtinv[gi];
    double Jraw[12];
#pragma unroll
    for (int c = 0; c < 12; c++) Jraw[c] = D.ps.jac[(size_t)c * SN + gi];
    double e2 = 0; e2 += e0 * e0; e2 += e1 * e1;
    const double sq = (e2 > sig2) ? 0.0 : 1.0 - (e2 / sig2);
    const double w = sq * sq;
    if (w == 0.0) { if (mark) D.ps.counts[gi]++; continue; }
    else if (mark) D.ps.counts[SN + gi]++;
#pragma unroll
    for (int row = 0; row < 2; row++) {
      double J[6];
#pragma unroll
      for (int c = 0; c < 6; c++) J[c] = si * Jraw[6 * row + c];
      const double e = row ? e1 : e0;
      const double m = D.truncate ? (double)(int)e : e;   // (int) cast of jni/Tracker.cc:766-767
      int q = 0;
#pragma unroll
      for (int r = 0; r < 6; r++) {
        const double Jw = w * J[r];
        // fused multiply-add: the sums over points are a parallel reduction (order differs from the reference's serial loop
        // anyway), so contracting here costs no parity and halves the FP64 instructions of the phase
        acc[21 + r] = __fma_rn(m, Jw, acc[21 + r]);
#pragma unroll
        for (int c = r; c < 6; c++) { acc[q] = __fma_rn(Jw, J[c], acc[q]); q++; }
      }
    }
  }
  // fixed-shape reduction: partials transposed through shared memory ([27][256]), then one warp-row per sum
#pragma unroll
  for (int k = 0; k < 27; k++) part[k * kPT + tid] = acc[k];
  __syncthreads();
  for (int k = warp; k < 27; k += kPT / 32) {
    double v = 0;
#pragma unroll
    for (int q = 0; q < kPT / 32; q++) v += part[k * kPT + q * 32 + lane];
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) sm.sums[k] = v;
  }
  __syncthreads();
  PT_MARK(6);
  // mu = inverse(C) * b with inverse = partial-pivot LU, column by column (jni/myWLS.h:53-62; oracle inverse_lu): same operations per
  // element as the serial routine, spread over six lanes of warp 0 (rows during elimination, columns during substitution).
  if (warp == 0) wls_solve_warp(sm.sums, sm.lu, sm.inv, sm.piv, sm.mu);
  __syncthreads();
  PT_MARK(7);
}

__device__ void reproject_found(const Dev& D, const double* pose, const int* list, int n, int s, int only_found, int* quirk) {
  const size_t SN = (size_t)D.S * D.N;
  for (int k = threadIdx.x; k < n; k += kPT) {
    const int i = list[k];
    const size_t gi = (size_t)s * D.N + i;
    int flags = D.ps.flags[gi];
    if (only_found && !(flags & F_FOUND)) continue;
    CamCache cc;
    const bool projected = td_project(D, pose, i, gi, SN, cc, flags);
    if (flags & F_FOUND) {   // ProjectAndDerivs refreshes the derivatives `if(bFound)` (jni/TrackerData.h:98-102)
      if (projected) { double dv[4]; cam_derivs(D.cam, cc, dv); for (int c = 0; c < 4; c++) D.ps.derivs[c * SN + gi] = dv[c]; }
      else atomicAdd(quirk, 1);   // reference would read the camera's cache of another point here; we keep the old derivatives
    }
    D.ps.flags[gi] = flags;
  }
}
// TrackerData::CalcJacobian (jni/TrackerData.h:107-123)
__device__ void calc_jacobians(const Dev& D, const int* list, int n, int s, bool all_found) {
  const size_t SN = (size_t)D.S * D.N;
  for (int k = threadIdx.x; k < n; k += kPT) {
    const size_t gi = (size_t)s * D.N + list[k];
    if (!all_found && !(D.ps.flags[gi] & F_FOUND)) continue;
    const double c[3] = {D.ps.v3cam[gi], D.ps.v3cam[SN + gi], D.ps.v3cam[2 * SN + gi]};
    const double dv[4] = {D.ps.derivs[gi], D.ps.derivs[SN + gi], D.ps.derivs[2 * SN + gi], D.ps.derivs[3 * SN + gi]};
    const double invz = 1.0 / c[2];
    const double pos[4] = {c[0], c[1], c[2], 1.0};
#pragma unroll
    for (int m = 0; m < 6; m++) {
      double v4[3] = {0, 0, 0};
      if (m < 3) v4[m] = pos[3];
      else { v4[(m + 1) % 3] = -pos[(m + 2) % 3]; v4[(m + 2) % 3] = pos[(m + 1) % 3]; }
      const double c0 = (v4[0] - c[0] * v4[2] * invz) * invz, c1 = (v4[1] - c[1] * v4[2] * invz) * invz;
      double a0 = dv[0] * c0; a0 += dv[1] * c1;
      double a1 = dv[2] * c0; a1 += dv[3] * c1;
      D.ps.jac[(size_t)m * SN + gi] = a0; D.ps.jac[(size_t)(6 + m) * SN + gi] = a1;
    }
  }
}
// One pass over the found points of a Gauss-Newton iteration, everything a point needs before the median in one go and in
// registers (no barrier, no reload between the steps):
//   action 1: TrackerData::ProjectAndDerivs (jni/TrackerData.h:91-103) with the current pose      (non-linear iterations > 0)
//   action 2: TrackerData::LinearUpdate (jni/TrackerData.h:126-132) with the last update `v6`     (linear iterations)
//   action 0: nothing (iteration 0: the projection of k_project_lists / k_reproject_fine stands)
//   jacobian: TrackerData::CalcJacobian (jni/TrackerData.h:107-123)                               (non-linear iterations)
//   then the error of jni/Tracker.cc:703-709 and its square (for the Tukey median, order irrelevant).
__device__ void pose_points_pass(const Dev& D, const double* pose, const int* flist, int n, int s, int action, bool jacobian, const double* v6,
                                 double* sortbuf, int sortcap, bool want_sort, int* quirk) {
  const size_t SN = (size_t)D.S * D.N;
  if (action == 2 && !jacobian) {
    // linear iteration (7 of the 10 of a fine stage): two points per thread and step, both points' 17 values requested before either
    // is used, so that a thread exposes one round trip to L2 per two points
    for (int k0 = threadIdx.x; k0 < n; k0 += 2 * kPT) {
      double si[2], f0[2], f1[2], im0[2], im1[2], J[2][12]; size_t gi[2]; bool ok[2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const int k = k0 + u * kPT;
        ok[u] = k < n;
        gi[u] = (size_t)s * D.N + flist[ok[u] ? k : k0];
        si[u] = D.ps.sqrtinv[gi[u]]; f0[u] = D.ps.v2found[gi[u]]; f1[u] = D.ps.v2found[SN + gi[u]];
        im0[u] = D.ps.v2image[gi[u]]; im1[u] = D.ps.v2image[SN + gi[u]];
#pragma unroll
        for (int q = 0; q < 12; q++) J[u][q] = D.ps.jac[(size_t)q * SN + gi[u]];
      }
#pragma unroll
      for (int u = 0; u < 2; u++) {
        if (!ok[u]) continue;
        double a0 = J[u][0] * v6[0], a1 = J[u][6] * v6[0];
#pragma unroll
        for (int q = 1; q < 6; q++) { a0 += J[u][q] * v6[q]; a1 += J[u][6 + q] * v6[q]; }
        const double n0 = im0[u] + a0, n1 = im1[u] + a1;
        D.ps.v2image[gi[u]] = n0; D.ps.v2image[SN + gi[u]] = n1;
        const double e0 = (f0[u] - n0) * si[u], e1 = (f1[u] - n1) * si[u];
        D.ps.err[gi[u]] = e0; D.ps.err[SN + gi[u]] = e1;
        double e2 = 0; e2 += e0 * e0; e2 += e1 * e1;
        const int k = k0 + u * kPT;
        if (want_sort && k < sortcap) sortbuf[k] = e2;
      }
    }
    return;
  }
  for (int k = threadIdx.x; k < n; k += kPT) {
    const int i = flist[k];
    const size_t gi = (size_t)s * D.N + i;
    const double si = D.ps.sqrtinv[gi], f0 = D.ps.v2found[gi], f1 = D.ps.v2found[SN + gi];
    double im0, im1, c[3], dv[4];
    bool have_cd = false;
    if (action == 1) {
      int flags = D.ps.flags[gi];
      CamCache cc;
      const bool projected = td_project(D, pose, i, gi, SN, cc, flags);
      // (flist holds found points only; ProjectAndDerivs refreshes the derivatives `if(bFound)`)
      if (projected) { cam_derivs(D.cam, cc, dv); for (int q = 0; q < 4; q++) D.ps.derivs[q * SN + gi] = dv[q]; }
      else { atomicAdd(quirk, 1); for (int q = 0; q < 4; q++) dv[q] = D.ps.derivs[q * SN + gi]; }   // reference reads another point's cache here; we keep the old derivatives
      D.ps.flags[gi] = flags;
      c[0] = D.ps.v3cam[gi]; c[1] = D.ps.v3cam[SN + gi]; c[2] = D.ps.v3cam[2 * SN + gi];
      im0 = D.ps.v2image[gi]; im1 = D.ps.v2image[SN + gi];
      have_cd = true;
    } else if (action == 2) {
      im0 = D.ps.v2image[gi]; im1 = D.ps.v2image[SN + gi];
      double a0 = D.ps.jac[gi] * v6[0], a1 = D.ps.jac[(size_t)6 * SN + gi] * v6[0];
#pragma unroll
      for (int q = 1; q < 6; q++) { a0 += D.ps.jac[(size_t)q * SN + gi] * v6[q]; a1 += D.ps.jac[(size_t)(6 + q) * SN + gi] * v6[q]; }
      im0 += a0; im1 += a1;
      D.ps.v2image[gi] = im0; D.ps.v2image[SN + gi] = im1;
    } else {
      im0 = D.ps.v2image[gi]; im1 = D.ps.v2image[SN + gi];
    }
    if (jacobian) {
      if (!have_cd) {
        c[0] = D.ps.v3cam[gi]; c[1] = D.ps.v3cam[SN + gi]; c[2] = D.ps.v3cam[2 * SN + gi];
        for (int q = 0; q < 4; q++) dv[q] = D.ps.derivs[q * SN + gi];
      }
      const double invz = 1.0 / c[2];
      const double pos[4] = {c[0], c[1], c[2], 1.0};
#pragma unroll
      for (int m = 0; m < 6; m++) {
        double v4[3] = {0, 0, 0};
        if (m < 3) v4[m] = pos[3];
        else { v4[(m + 1) % 3] = -pos[(m + 2) % 3]; v4[(m + 2) % 3] = pos[(m + 1) % 3]; }
        const double c0 = (v4[0] - c[0] * v4[2] * invz) * invz, c1 = (v4[1] - c[1] * v4[2] * invz) * invz;
        double a0 = dv[0] * c0; a0 += dv[1] * c1;
        double a1 = dv[2] * c0; a1 += dv[3] * c1;
        D.ps.jac[(size_t)m * SN + gi] = a0; D.ps.jac[(size_t)(6 + m) * SN + gi] = a1;
      }
    }
    const double e0 = (f0 - im0) * si, e1 = (f1 - im1) * si;
    D.ps.err[gi] = e0; D.ps.err[SN + gi] = e1;
    double e2 = 0; e2 += e0 * e0; e2 += e1 * e1;
    if (want_sort && k < sortcap) sortbuf[k] = e2;
  }
}

// Found entries of list[0,n), written to flist in ASCENDING POINT INDEX (a bitmap of the set, then an ordered expansion): the
// per-point SoA arrays are then read with consecutive addresses by consecutive threads.  The order of the set does not matter
// to the algorithm (sums, median).  `bitmap` needs (N+31)/32 words.  Returns the number of found entries (all threads).
__device__ int build_found_list(const Dev& D, PoseSmem& sm, const int* list, int n, int s, int* flist, unsigned* bitmap) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int words = (D.map.n + 31) >> 5;
  for (int w = tid; w < words; w += kPT) bitmap[w] = 0u;
  __syncthreads();
  for (int k = tid; k < n; k += kPT) {
    const int idx = list[k];
    if (D.ps.flags[(size_t)s * D.N + idx] & F_FOUND) atomicOr(&bitmap[idx >> 5], 1u << (idx & 31));
  }
  __syncthreads();
  int base = 0;
  for (int w0 = 0; w0 < words; w0 += kPT) {
    const int w = w0 + tid;
    unsigned m = w < words ? bitmap[w] : 0u;
    const int mine = __popc(m);
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    if (lane == 31) sm.warpcnt[warp] = incl;
    __syncthreads();
    int off = base, tot = 0;
    for (int q = 0; q < kPT / 32; q++) { if (q < warp) off += sm.warpcnt[q]; tot += sm.warpcnt[q]; }
    int pos = off + incl - mine;
    while (m) { const int b = __ffs(m) - 1; m &= m - 1; flist[pos++] = (w << 5) + b; }
    base += tot;
    __syncthreads();
  }
  return base;
}

// mode 0: one CalcPoseUpdate over [0,nA) (stage API);  1: coarse stage;  2: fine stage (+ scene depth; + motion model / quality if tail)
__global__ void __launch_bounds__(kPT, 2) k_pose(Dev D, int mode, double sigma_arg, int mark_arg, int apply_arg, int tail) {
  extern __shared__ double sh_sort[];
  __shared__ PoseSmem sm;
  const int s = blockIdx.x + D.s0, tid = threadIdx.x;
  StreamState* st = D.ss + s;
  if (mode != 0 && st->lost_frames >= 3 && !st->recovered) return;
  const int* list = D.lists + (size_t)s * D.list_cap;
  // dynamic shared memory: [2048 doubles sort keys][2048 ints found list][2048 ints radix histogram][27*256 doubles partial sums]
  double* sortbuf = sh_sort; int sortcap = 2048;
  int* flist = (int*)(sh_sort + 2048);
  int* hist = flist + 2048;
  double* part = sh_sort + 2048 + 2048;          // [27][kPT] partial sums of the normal equations
  const int nA = st->nA, nAll = st->nA + st->nB;
  const int nlist = (mode == 2) ? nAll : nA;
  if (nlist > 2048) { sortbuf = D.sort_scratch + (size_t)s * D.sort_cap; sortcap = D.sort_cap; flist = D.pvs + (size_t)s * VS_LEVELS * D.N; }
  if (tid < 12) sm.pose[tid] = st->pose[tid];
  if (tid < 6) sm.last[tid] = 0.0;
  __syncthreads();
  if (mode == 1 && !st->try_coarse) return;
  // every point that enters the iterations was searched before this kernel: the found set is fixed from here on
#ifdef VS_POSE_TIMING
  if (tid < 32) sm.tacc[tid] = 0;
  if (tid < 2) sm.tlast[tid] = clock64();
  __syncthreads();
#endif
  const int n = build_found_list(D, sm, list, nlist, s, flist, (unsigned*)hist);   // (the histogram buffer doubles as the bitmap: N <= 65536)
  PT_MARK(0);

  if (mode == 0) {
    pose_points_pass(D, sm.pose, flist, n, s, 0, false, sm.last, sortbuf, sortcap, sigma_arg <= 0, &st->quirk_stale_cache);
    __syncthreads();
    calc_pose_update(D, sm, sortbuf, sortcap, hist, part, flist, n, s, sigma_arg, mark_arg != 0);
    if (tid == 0) {
      const int u = st->n_updates < VS_MAX_UPDATES ? st->n_updates : VS_MAX_UPDATES - 1;
      for (int k = 0; k < 6; k++) st->updates[6 * u + k] = sm.mu[k];
      st->sigmas[u] = sm.sigma; st->n_updates = u + 1;
      if (apply_arg) { double e[12], np[12]; se3_exp(sm.mu, e); se3_mul(e, sm.pose, np); for (int k = 0; k < 12; k++) st->pose[k] = np[k]; }
    }
    return;
  }

  const int iters = 10;
  if (mode == 1) {   // coarse stage (jni/Tracker.cc:464-489): needs nFound >= CoarseMin
    if ((unsigned)n < D.prm.coarse_min) return;
    if (tid == 0) st->did_coarse = 1;
  }

  for (int iter = 0; iter < iters; iter++) {
    bool nonlinear = true;
    if (mode == 2) nonlinear = (iter == 0 || iter == 4 || iter == 9);
    const double ov = (iter > 5) ? (mode == 1 ? 1.0 : 16.0) : 0.0;
    pose_points_pass(D, sm.pose, flist, n, s, iter == 0 ? 0 : (nonlinear ? 1 : 2), nonlinear, sm.last, sortbuf, sortcap, ov <= 0, &st->quirk_stale_cache);
    __syncthreads();
    PT_MARK(1);
    calc_pose_update(D, sm, sortbuf, sortcap, hist, part, flist, n, s, ov, mode == 2 && iter == 9);
    if (tid == 0) {
      double e[12], np[12]; se3_exp(sm.mu, e); se3_mul(e, sm.pose, np);
      for (int k = 0; k < 12; k++) sm.pose[k] = np[k];
      for (int k = 0; k < 6; k++) sm.last[k] = sm.mu[k];
      const int u = st->n_updates;
      if (u < VS_MAX_UPDATES) { for (int k = 0; k < 6; k++) st->updates[6 * u + k] = sm.mu[k]; st->sigmas[u] = sm.sigma; st->n_updates = u + 1; }
    }
    __syncthreads();
    PT_MARK(8);
  }
#ifdef VS_POSE_TIMING
  if (tid == 0 && s == 7 && mode == 2) printf("k_pose cycles: found_list %lld reproject %lld linear %lld jac %lld err %lld median %lld accum %lld solve %lld exp %lld (n=%d)\n", sm.tacc[0], sm.tacc[1], sm.tacc[2], sm.tacc[3], sm.tacc[4], sm.tacc[5], sm.tacc[6], sm.tacc[7], sm.tacc[8], n);
#endif
  if (tid < 12) st->pose[tid] = sm.pose[tid];
  if (mode != 2) return;

  // scene depth from the tracked features (jni/Tracker.cc:610-625); fixed-shape parallel sums
  {
    const size_t SN2 = 2 * (size_t)D.S * D.N;
    double a0 = 0, a1 = 0, a2 = 0;
    for (int k = tid; k < n; k += kPT) {
      const size_t gi = (size_t)s * D.N + flist[k];
      const double z = D.ps.v3cam[SN2 + gi]; a0 += z; a1 += z * z; a2 += 1.0;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, d); a1 += __shfl_xor_sync(0xffffffffu, a1, d); a2 += __shfl_xor_sync(0xffffffffu, a2, d); }
    if ((tid & 31) == 0) { sm.red[tid >> 5][0] = a0; sm.red[tid >> 5][1] = a1; sm.red[tid >> 5][2] = a2; }
  }
  __syncthreads();
  if (tid == 0) {
    double dSum = 0, dSumSq = 0, dNum = 0;
    for (int w = 0; w < kPT / 32; w++) { dSum += sm.red[w][0]; dSumSq += sm.red[w][1]; dNum += sm.red[w][2]; }
    pose_stage_tail(D, st, s, sm.pose, dSum, dSumSq, dNum, tail);
  }
}

// ProjectAndDerivs of the fine set [nA, nA+nB) before its search (jni/Tracker.cc:501-503,530-532).  Nothing in that set is
// `found` yet, so only the projection is refreshed; when the coarse stage did not run the pose is unchanged and the
// projection of k_project_lists is still exact, so the CTA returns at once.
__global__ void __launch_bounds__(kPT) k_reproject_fine(Dev D) {
  cudaGridDependencySynchronize(); cudaTriggerProgrammaticLaunchCompletion();   // programmatic dependent launch (vs_launch_pdl); no-ops otherwise
  const int s = blockIdx.x + D.s0;
  StreamState* st = D.ss + s;
  // The layout hint of vs_launch_track_map_rest: did this stream try the coarse stage in this frame?  Written to mapped HOST memory, from here because
  // this kernel is off the critical path whenever the answer matters (in the two-chain layout it belongs to the coarse chain, which is empty while the
  // hint reads 0): a kernel that has written to system memory completes ~3 us later (measured in k_project_lists: 0.0375 -> 0.0412 ms at 32 streams).
  const bool lost = st->lost_frames >= 3 && !st->recovered;
  if (threadIdx.x == 0) D.coarse_hint[s] = lost ? 0 : st->try_coarse;
  if (lost || !st->did_coarse || other_chain(D, st)) return;
  const size_t SN = (size_t)D.S * D.N;
  const int* list = D.lists + (size_t)s * D.list_cap + st->nA;
  for (int k = threadIdx.x; k < st->nB; k += kPT) {
    const int i = list[k];
    const size_t gi = (size_t)s * D.N + i;
    int flags = D.ps.flags[gi];
    CamCache cc; td_project(D, st->pose, i, gi, SN, cc, flags);
    D.ps.flags[gi] = flags;
  }
}

__global__ void __launch_bounds__(kPT) k_project_and_derivs(Dev D, int only_found) {
  const int s = blockIdx.x + D.s0;
  StreamState* st = D.ss + s;
  reproject_found(D, st->pose, D.lists + (size_t)s * D.list_cap, st->nA, s, only_found, &st->quirk_stale_cache);
}
__global__ void __launch_bounds__(kPT) k_calc_jacobians(Dev D) {
  const int s = blockIdx.x + D.s0;
  calc_jacobians(D, D.lists + (size_t)s * D.list_cap, D.ss[s].nA, s, false);
}

__global__ void k_atan(const double* x, double* y, int n, int dd_only) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) y[i] = dd_only ? atan_cr_dd(x[i]) : atan_cr(x[i]); }

}  // namespace

// M = T^kRandJumpBlocks of glibc_rand_fill_parallel: column j = the ring after kRandJumpBlocks blocks of 31 draws started from the unit vector e_j
static int upload_rand_jump(vslam_ctx* ctx) {
  static const std::vector<uint32_t> M = [] {
    std::vector<uint32_t> m(31 * 32, 0u);
    for (int j = 0; j < 31; j++) {
      uint32_t r[31] = {}; r[j] = 1u;
      for (int blk = 0; blk < kRandJumpBlocks; blk++) for (int q = 0; q < 31; q++) r[(q + 3) % 31] += r[q];
      for (int i = 0; i < 31; i++) m[i * 32 + j] = r[i];
    }
    return m;
  }();
  VS_CUDA(cudaMemcpyToSymbolAsync(g_rand_jump, M.data(), M.size() * sizeof(uint32_t), 0, cudaMemcpyHostToDevice, ctx->stream));
  ctx->rand_jump_ready = true;
  return VSLAM_OK;
}
constexpr int kSplitProjectN = 2048;   // maps above this many points project on a (point chunks, streams) grid (k_project_points)
int vs_launch_project_all(vslam_ctx* ctx, int mode) {
  const Dev D = make_dev(ctx);
  // packed level lists + random draws: 2 x N ints of shared memory when that fits beside the static part, the stream's global scratch otherwise
  size_t smem = (size_t)2 * ctx->N * sizeof(int);
  const int use_smem = smem <= 200 * 1024;
  if (!use_smem) smem = 0;
  // std::random_shuffle without the swap chain (shuffle_parallel): scratch behind the two arrays when 6 N + 1 ints fit, else in the stream's global
  // scratch; maps too large for the two arrays in shared memory keep the serial chains
  static const int par_env = getenv("VSLAM_PAR_SHUFFLE") ? atoi(getenv("VSLAM_PAR_SHUFFLE")) : -1;
  int par_shuffle = 0;
  if (!ctx->rand_jump_ready) { const int rc_ = upload_rand_jump(ctx); if (rc_) return rc_; }
  if (use_smem && par_env != 0) {
    const size_t smem6 = ((size_t)6 * ctx->N + 1) * sizeof(int);
    if (smem6 <= 200 * 1024 && par_env != 2) { par_shuffle = 1; smem = smem6; } else par_shuffle = 2;
  }
  if (smem > ctx->smem_attr[0]) { VS_CUDA(cudaFuncSetAttribute(k_project_lists, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); ctx->smem_attr[0] = smem; }
  // large maps: the per-point projection on its own grid (k_project_points), the list building behind it
  static const int split_env = getenv("VSLAM_SPLIT_PROJECT") ? atoi(getenv("VSLAM_SPLIT_PROJECT")) : -1;
  const int pre = (mode & 1) && (split_env >= 0 ? split_env != 0 : ctx->map.n > kSplitProjectN);
  const bool pdl = ctx->pdl && !ctx->timing;
  vs_time_begin(ctx, VS_ST_PROJECT);
  if (pre) { VS_CUDA(vs_launch_pdl(k_project_points, dim3((ctx->map.n + kPT - 1) / kPT, ctx->cur_cnt), dim3(kPT), 0, ctx->stream, pdl, D, (mode >> 1) & 1, ctx->cur_sbi_rot)); ctx->launches++; }
  VS_CUDA(vs_launch_pdl(k_project_lists, dim3(ctx->cur_cnt), dim3(kPT), smem, ctx->stream, pdl, D, mode & 1, (mode >> 1) & 1, use_smem, ctx->cur_sbi_rot, pre, par_shuffle));
  vs_time_end(ctx);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  return VSLAM_OK;
}

int vs_launch_search(vslam_ctx* ctx, int which, int range, int subpix, int sflags) {
  if (ctx->params.search_kernel == 0) return vs_launch_search_fast(ctx, which, range, subpix, sflags);
  { const int rc_ = vs_ensure_lists(ctx); if (rc_) return rc_; }
  const Dev D = make_dev(ctx);
  const int max_entries = which == 1 ? (int)(2 * ctx->params.coarse_max) : ctx->list_cap;
  if (max_entries <= 0) return VSLAM_OK;   // nCoarseMax == 0: the reference skips the coarse stage (jni/Tracker.cc:425)
  dim3 grid((max_entries + kSearchWarps - 1) / kSearchWarps, ctx->cur_cnt);
  vs_time_begin(ctx, which == 2 ? VS_ST_SEARCH_FINE : VS_ST_SEARCH_COARSE);
  if (ctx->P == 11) k_search<11><<<grid, kSearchWarps * 32, 0, ctx->stream>>>(D, which, range, subpix, sflags);
  else if (ctx->P == 8) k_search<8><<<grid, kSearchWarps * 32, 0, ctx->stream>>>(D, which, range, subpix, sflags);
  else k_search<0><<<grid, kSearchWarps * 32, 0, ctx->stream>>>(D, which, range, subpix, sflags);
  vs_time_end(ctx);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  return VSLAM_OK;
}

int vs_launch_pose(vslam_ctx* ctx, int mode, double sigma, int mark, int apply) {
  if ((mode & 3) != 0 && ctx->params.pose_kernel == 0) return vs_launch_pose_fast(ctx, mode);   // the TrackMap stages; mode 0 (one CalcPoseUpdate, stage API) stays here
  const Dev D = make_dev(ctx);
  const size_t smem = 2048 * sizeof(double) + 2 * 2048 * sizeof(int) + 27 * kPT * sizeof(double);   // sort keys + found list + radix histogram + partial sums
  if (smem > ctx->smem_attr[1]) { VS_CUDA(cudaFuncSetAttribute(k_pose, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); ctx->smem_attr[1] = smem; }
  vs_time_begin(ctx, (mode & 3) == 2 ? VS_ST_POSE_FINE : VS_ST_POSE_COARSE);
  k_pose<<<ctx->cur_cnt, kPT, smem, ctx->stream>>>(D, mode & 3, sigma, mark, apply, (mode >> 2) & 1);
  vs_time_end(ctx);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  return VSLAM_OK;
}

int vs_launch_project_and_derivs(vslam_ctx* ctx, int only_found) {
  k_project_and_derivs<<<ctx->cur_cnt, kPT, 0, ctx->stream>>>(make_dev(ctx), only_found);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  return VSLAM_OK;
}
int vs_launch_calc_jacobians(vslam_ctx* ctx) {
  k_calc_jacobians<<<ctx->cur_cnt, kPT, 0, ctx->stream>>>(make_dev(ctx));
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  return VSLAM_OK;
}

// Tracker::TrackMap for all streams: 6 launches, no host synchronisation in between.
// Two launch chains.  Whether a stream runs the coarse stage is decided on the device (k_project_lists: StreamState::try_coarse), so the host
// launches it for every frame: three searches / pose kernels and the re-projection that do nothing at ordinary camera speed, each waiting for
// the one before -- 0.018 ms of a frame's latency (32 streams: 7 % of the step).  Instead the frame forks: the streams that try the coarse stage
// take the whole chain (coarse search, coarse pose, re-projection, fine search, fine pose) on a stream of its own, the others the fine stage alone
// on the calling stream (Dev::chain tells a kernel which streams are its own); the chains join at the end.  A chain without streams is a row of
// empty kernels BESIDE the other chain's work instead of in front of it.  Streams are independent: no result depends on the split.
static int track_map_rest_chain(vslam_ctx* ctx, int with_motion_model);
constexpr int kDualChainMaxList = 2048, kDualChainMinStreams = 48;
int vs_launch_track_map_rest(vslam_ctx* ctx, int with_motion_model) {
  static const int env = getenv("VSLAM_COARSE_CHAIN") ? atoi(getenv("VSLAM_COARSE_CHAIN")) : -1;   // A/B runs: overrides vslam_params.coarse_chain
  const int want = env >= 0 ? env : ctx->params.coarse_chain;
  bool dual = want != 0 && !ctx->timing && ctx->params.search_kernel == 0 && ctx->params.pose_kernel == 0;
  if (dual && want < 0) {
    // Worth it only while the coarse chain is empty and its empty kernels are small: with a fast camera the work would merely move to the
    // other stream and pay the fork / join (measured: + 2 %), and the empty fine search of a 20000-point map is 185 k CTAs (4K: + 1.2 %).  Which
    // streams tried the coarse stage in their latest frame is a HINT read without synchronisation (k_project_lists writes it to mapped host
    // memory): it chooses the layout, never the result -- both layouts run every stream through the stages it asks for.
    // ... and while the host keeps up: the second chain costs three more launches and two event pairs per frame, and a context of a few streams finishes a
    // frame in the time the host needs to enqueue one (32 streams per GPU: 0.250 -> 0.245 ms with one process driving one GPU, but 0.248 -> 0.257 ms
    // with eight processes driving eight GPUs of one host)
    if (ctx->list_cap > kDualChainMaxList || ctx->S < kDualChainMinStreams) dual = false;
    else { const volatile int* h = ctx->coarse_hint_host; for (int s = ctx->cur_s0; s < ctx->cur_s0 + ctx->cur_cnt && dual; s++) if (h[s]) dual = false; }
  }
  if (!dual) return track_map_rest_chain(ctx, with_motion_model);
  const int g = ctx->cur_group;
  const bool hi = ctx->back_stream && ctx->stream == ctx->back_stream;     // look-ahead back end: keep its priority
  cudaStream_t* slot = hi ? &ctx->chain_stream_hi : &ctx->chain_stream[g];
  if (!*slot) {
    int lo_p = 0, hi_p = 0; VS_CUDA(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
    VS_CUDA(cudaStreamCreateWithPriority(slot, cudaStreamNonBlocking, hi ? hi_p : 0));
  }
  if (!ctx->ev_chain_fork[g]) { VS_CUDA(cudaEventCreateWithFlags(&ctx->ev_chain_fork[g], cudaEventDisableTiming)); VS_CUDA(cudaEventCreateWithFlags(&ctx->ev_chain_join[g], cudaEventDisableTiming)); }
  cudaStream_t main_stream = ctx->stream, side = *slot;
  struct Guard { vslam_ctx* c; cudaStream_t m; ~Guard() { c->stream = m; c->cur_chain = -1; } } guard{ctx, main_stream};
  VS_CUDA(cudaEventRecord(ctx->ev_chain_fork[g], main_stream));
  VS_CUDA(cudaStreamWaitEvent(side, ctx->ev_chain_fork[g], 0));
  int rc;
  ctx->stream = side; ctx->cur_chain = 1;
  if ((rc = track_map_rest_chain(ctx, with_motion_model))) return rc;
  VS_CUDA(cudaEventRecord(ctx->ev_chain_join[g], side));
  ctx->stream = main_stream; ctx->cur_chain = 0;
  if ((rc = vs_launch_search(ctx, 2, 0, 0, 0))) return rc;
  if ((rc = vs_launch_pose(ctx, 2 | (with_motion_model ? 4 : 0), 0.0, 0, 0))) return rc;
  VS_CUDA(cudaStreamWaitEvent(main_stream, ctx->ev_chain_join[g], 0));
  return VSLAM_OK;
}
static int track_map_rest_chain(vslam_ctx* ctx, int with_motion_model) {
  int rc;
  if ((rc = vs_launch_search(ctx, 1, 0, 0, 0))) return rc;
  if ((rc = vs_launch_pose(ctx, 1, 0.0, 0, 0))) return rc;
  vs_time_begin(ctx, VS_ST_OTHER);
  VS_CUDA(vs_launch_pdl(k_reproject_fine, dim3(ctx->cur_cnt), dim3(kPT), 0, ctx->stream, ctx->pdl && !ctx->timing, make_dev(ctx)));
  vs_time_end(ctx);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  if ((rc = vs_launch_search(ctx, 2, 0, 0, 0))) return rc;
  if ((rc = vs_launch_pose(ctx, 2 | (with_motion_model ? 4 : 0), 0.0, 0, 0))) return rc;
  return VSLAM_OK;
}
int vs_launch_track_map(vslam_ctx* ctx, int with_motion_model) {
  const int rc = vs_launch_project_all(ctx, 1 | (with_motion_model ? 2 : 0));
  return rc ? rc : vs_launch_track_map_rest(ctx, with_motion_model);
}

// Tracker::TrackFrame for all streams (jni/Tracker.cc:68-160, map-good branch).
// Launch graph of one stream group: the level images exist once the level-0 launch is done, so SmallBlurryImage + relocaliser +
// motion model + projection (side stream; one CTA per stream, latency-bound) run beside the FAST pass of levels 1..3 (main stream)
// and join before the first patch search, which needs the corner bitmasks of all levels.
static int launch_frame_group(vslam_ctx* ctx, int g, bool fork) {
  int rc;
  cudaStream_t main_stream = ctx->stream;
  // k_project_lists takes this frame's SmallBlurryImage rotation from the frame set's slot of sbi_rot_buf (null: StreamState::sbi_rot as the host set it)
  struct RotGuard { vslam_ctx* c; ~RotGuard() { c->cur_sbi_rot = nullptr; } } rot_guard{ctx};
  ctx->cur_sbi_rot = (ctx->sbi_on && ctx->params.use_sbi) ? ctx->sbi_rot_buf + (size_t)ctx->cur_set * ctx->S * 6 : nullptr;
  if ((rc = vs_launch_pyramid_l0(ctx, ctx->cur_s0, ctx->cur_cnt))) return rc;
  if (fork) {
    VS_CUDA(cudaEventRecord(ctx->ev_fork[g], main_stream));
    VS_CUDA(cudaStreamWaitEvent(ctx->side_stream[g], ctx->ev_fork[g], 0));
    ctx->stream = ctx->side_stream[g];
  }
  rc = vs_launch_sbi(ctx);
  if (!rc) rc = vs_launch_relocalise(ctx);
  if (!rc) rc = vs_launch_project_all(ctx, 3);
  ctx->stream = main_stream;
  if (fork) VS_CUDA(cudaEventRecord(ctx->ev_join[g], ctx->side_stream[g]));
  if (rc) return rc;
  if ((rc = vs_launch_fast_levels(ctx, ctx->cur_s0, ctx->cur_cnt))) return rc;
  // The corner lists / row LUTs (k_corner_lists) are for the API and for the kernels that walk lists (k_search of round 1, MapMaker's
  // epipolar search, MiniPatch, MakeKeyFrame_Rest): k_search_fast reads the corner bitmasks.  With it a tracked frame does not build the
  // lists at all; whoever needs them next builds them from the bitmasks first (vs_ensure_lists).
  if (ctx->params.search_kernel == 0) ctx->lists_stale = true;
  else if ((rc = vs_launch_corner_lists(ctx, ctx->cur_s0, ctx->cur_cnt))) return rc;
  if (fork) VS_CUDA(cudaStreamWaitEvent(main_stream, ctx->ev_join[g], 0));
  return vs_launch_track_map_rest(ctx, 1);
}
// Optionally (vslam_params.stream_groups > 1) the streams of a context are split into groups whose launch graphs run on separate
// CUDA streams between one fork and one join on ctx->stream.  Streams are independent, so the split changes no result.  Measured on
// B200 (256 VGA streams): 2-4 groups are 5 % slower than one (smaller grids, more tails; the hoped-for overlap of one group's
// latency-bound k_pose with another group's k_search does not materialise because k_pose CTAs hold half an SM's registers each),
// hence the default of 1.  With per-stage timing on (vslam_set_timing) everything is serialised on ctx->stream.
// Frame look-ahead (vs_begin_frame chose it: ctx->la_frame, ctx->front == ctx->front_stream, the frame's set is current): the front end
//   front_stream: k_pyramid_fast -> fork { front_side: k_sbi } || k_fast_levels -> join
// depends on the frame alone and runs beside the previous frame's back end, which is still in flight on ctx->stream; the back end
//   ctx->stream:  (front end done) -> k_relocalise -> k_project_lists -> searches and pose iterations
// follows in stream order.  ev_back_done[set] tells the front end of the frame after next that this set may be overwritten.
static int launch_frame_lookahead(vslam_ctx* ctx) {
  int rc;
  cudaStream_t main_stream = ctx->stream, F = ctx->front_stream;
  struct Guard { vslam_ctx* c; cudaStream_t m; ~Guard() { c->stream = m; c->cur_sbi_rot = nullptr; } } guard{ctx, main_stream};
  ctx->cur_sbi_rot = (ctx->sbi_on && ctx->params.use_sbi) ? ctx->sbi_rot_buf + (size_t)ctx->cur_set * ctx->S * 6 : nullptr;
  ctx->stream = F;
  if ((rc = vs_launch_pyramid_l0(ctx, 0, ctx->S))) return rc;
  if (ctx->sbi_on) {
    VS_CUDA(cudaEventRecord(ctx->ev_la_fork, F));
    VS_CUDA(cudaStreamWaitEvent(ctx->front_side, ctx->ev_la_fork, 0));
    ctx->stream = ctx->front_side;
    rc = vs_launch_sbi(ctx);
    ctx->stream = F;
    VS_CUDA(cudaEventRecord(ctx->ev_la_join, ctx->front_side));
    if (rc) return rc;
  }
  if ((rc = vs_launch_fast_levels(ctx, 0, ctx->S))) return rc;
  if (ctx->sbi_on) VS_CUDA(cudaStreamWaitEvent(F, ctx->ev_la_join, 0));
  VS_CUDA(cudaEventRecord(ctx->ev_front_done, F));
  // back end on the library's high-priority stream: after whatever the caller has enqueued on ctx->stream so far (that includes the previous
  // frame's join below) and after this frame's front end; ctx->stream then waits for it, so the caller sees ordinary stream order
  cudaStream_t B = ctx->back_stream;
  VS_CUDA(cudaEventRecord(ctx->ev_user, main_stream));
  VS_CUDA(cudaStreamWaitEvent(B, ctx->ev_user, 0));
  VS_CUDA(cudaStreamWaitEvent(B, ctx->ev_front_done, 0));
  ctx->stream = B;
  if (ctx->params.search_kernel == 0) ctx->lists_stale = true;
  else if ((rc = vs_launch_corner_lists(ctx, 0, ctx->S))) return rc;
  if ((rc = vs_launch_relocalise(ctx))) return rc;
  if ((rc = vs_launch_project_all(ctx, 3))) return rc;
  if ((rc = vs_launch_track_map_rest(ctx, 1))) return rc;
  VS_CUDA(cudaEventRecord(ctx->ev_back_done[ctx->cur_set], B));
  ctx->stream = main_stream;
  VS_CUDA(cudaStreamWaitEvent(main_stream, ctx->ev_back_done[ctx->cur_set], 0));
  ctx->launches_after_frame = ctx->launches;
  return VSLAM_OK;
}

int vs_launch_frame(vslam_ctx* ctx) {
  if (ctx->la_frame) { ctx->cur_s0 = 0; ctx->cur_cnt = ctx->S; ctx->cur_group = 0; return launch_frame_lookahead(ctx); }
  int G = ctx->params.stream_groups;
  if (G < 1) G = 1;
  if (G > VS_MAX_GROUPS) G = VS_MAX_GROUPS;
  if (G > ctx->S) G = ctx->S;
  if (ctx->timing) { ctx->cur_s0 = 0; ctx->cur_cnt = ctx->S; ctx->cur_group = 0; return launch_frame_group(ctx, 0, false); }
  cudaStream_t user_stream = ctx->stream;
  int rc = VSLAM_OK;
  if (G > 1) VS_CUDA(cudaEventRecord(ctx->ev_begin, user_stream));
  for (int g = 0; g < G && !rc; g++) {
    const int s0 = (int)((long long)ctx->S * g / G), s1 = (int)((long long)ctx->S * (g + 1) / G);
    ctx->cur_s0 = s0; ctx->cur_cnt = s1 - s0; ctx->cur_group = g;
    if (g > 0) { if (cudaStreamWaitEvent(ctx->group_stream[g], ctx->ev_begin, 0) != cudaSuccess) rc = VSLAM_E_CUDA; ctx->stream = ctx->group_stream[g]; }
    if (!rc) rc = launch_frame_group(ctx, g, true);
    ctx->stream = user_stream;
    if (g > 0 && !rc) {
      if (cudaEventRecord(ctx->ev_end[g], ctx->group_stream[g]) != cudaSuccess || cudaStreamWaitEvent(user_stream, ctx->ev_end[g], 0) != cudaSuccess) rc = VSLAM_E_CUDA;
    }
  }
  ctx->cur_s0 = 0; ctx->cur_cnt = ctx->S; ctx->cur_group = 0;
  if (rc == VSLAM_E_CUDA && ctx->err.empty()) ctx->err = "vs_launch_frame: CUDA stream/event error";
  return rc;
}

// Test hook: atan_cr over a device-side copy of x (used to compare against the host libm).
static int debug_atan(const double* x_host, double* y_host, int n, int dd_only);
extern "C" int vslam_debug_atan(const double* x_host, double* y_host, int n) { return debug_atan(x_host, y_host, n, 0); }
extern "C" int vslam_debug_atan_dd(const double* x_host, double* y_host, int n) { return debug_atan(x_host, y_host, n, 1); }   // the double-double routine alone
static int debug_atan(const double* x_host, double* y_host, int n, int dd_only) {
  double *dx = nullptr, *dy = nullptr;
  if (cudaMalloc(&dx, sizeof(double) * n) != cudaSuccess || cudaMalloc(&dy, sizeof(double) * n) != cudaSuccess) return VSLAM_E_CUDA;
  cudaMemcpy(dx, x_host, sizeof(double) * n, cudaMemcpyHostToDevice);
  k_atan<<<(n + 255) / 256, 256>>>(dx, dy, n, dd_only);
  const cudaError_t e = cudaMemcpy(y_host, dy, sizeof(double) * n, cudaMemcpyDeviceToHost);
  cudaFree(dx); cudaFree(dy);
  return e == cudaSuccess ? VSLAM_OK : VSLAM_E_CUDA;
}

// The line geometry of MapMaker::AddPointEpipolar (jni/MapMaker.cc:543-591), one thread per candidate: the candidate's viewing ray (z = 1
// plane coordinates `rays`, from ATANCamera::UnProject on the host) rotated into the target camera, the depth range cut to the part in front
// of the camera, the projected segment A-B in normal form.  The reference's operations in the reference's order (no contraction: -fmad=false).
namespace {
__global__ void k_epipolar_geometry(int n, EpiGeom G, const double* __restrict__ rays, const int* __restrict__ xy, EpiCand* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  EpiCand C;
  C.nx = C.ny = C.ax = C.ay = C.normDist = C.minLen = C.maxLen = C.maxDistSq = 0.0; C.valid = 0; C.pad = 0;
  C.x = xy[2 * k]; C.y = xy[2 * k + 1];
  // v3CamCenter_TC = kTarget.se3CfromW * kSrc.se3CfromW.inverse().get_translation()
  const double* S = G.src_pose; const double* T = G.tgt_pose;
  const double ts[3] = {S[3], S[7], S[11]};
  double rts[3];
  for (int i = 0; i < 3; i++) { double a = S[i] * ts[0]; a += S[4 + i] * ts[1]; a += S[8 + i] * ts[2]; rts[i] = a; }      // R_s^T t_s
  const double sinv_t[3] = {-rts[0], -rts[1], -rts[2]};
  double center[3]; rot_apply(T, sinv_t, center); center[0] = center[0] + T[3]; center[1] = center[1] + T[7]; center[2] = center[2] + T[11];
  double ray[3] = {rays[2 * k], rays[2 * k + 1], 1.0};
  { double nn = ray[0] * ray[0]; nn += ray[1] * ray[1]; nn += ray[2] * ray[2]; const double nrm = sqrt(nn); ray[0] /= nrm; ray[1] /= nrm; ray[2] /= nrm; }
  double tmp[3], dirn[3];
  for (int i = 0; i < 3; i++) { double a = S[i] * ray[0]; a += S[4 + i] * ray[1]; a += S[8 + i] * ray[2]; tmp[i] = a; }     // R_s^T ray
  rot_apply(T, tmp, dirn);
  double start[3], end[3];
  for (int q = 0; q < 3; q++) { start[q] = center[q] + G.start_depth * dirn[q]; end[q] = center[q] + G.end_depth * dirn[q]; }
  bool ok = !(end[2] <= start[2]) && !(end[2] <= 0.0);
  if (ok) {
    if (start[2] <= 0.0) { const double f = 0.001 - start[2] / dirn[2]; for (int q = 0; q < 3; q++) start[q] += dirn[q] * f; }
    const double A[2] = {start[0] / start[2], start[1] / start[2]}, B[2] = {end[0] / end[2], end[1] / end[2]};
    double al[2] = {A[0] - B[0], A[1] - B[1]};
    double aa = al[0] * al[0]; aa += al[1] * al[1];
    if (!(aa < 0.00000001)) {
      { const double nrm = sqrt(aa); al[0] /= nrm; al[1] /= nrm; }
      const double nrml[2] = {al[1], -al[0]};
      double dNormDist = A[0] * nrml[0]; dNormDist += A[1] * nrml[1];
      if (!(fabs(dNormDist) > G.largest_radius)) {
        double aA = al[0] * A[0]; aA += al[1] * A[1];
        double aB = al[0] * B[0]; aB += al[1] * B[1];
        double dMinLen = fmin(aA, aB) - 0.05, dMaxLen = fmax(aA, aB) + 0.05;
        if (dMinLen < -2.0) dMinLen = -2.0;
        if (dMaxLen < -2.0) dMaxLen = -2.0;
        if (dMinLen > 2.0) dMinLen = 2.0;
        if (dMaxLen > 2.0) dMaxLen = 2.0;
        C.nx = nrml[0]; C.ny = nrml[1]; C.ax = al[0]; C.ay = al[1]; C.normDist = dNormDist; C.minLen = dMinLen; C.maxLen = dMaxLen; C.maxDistSq = G.max_dist_sq; C.valid = 1;
      }
    }
  }
  out[k] = C;
}
}  // namespace
int vs_launch_epipolar_geometry(vslam_ctx* ctx, int n, const EpiGeom& G, const double* rays_dev, const int* xy_dev, EpiCand* cand_dev) {
  if (n <= 0) return VSLAM_OK;
  k_epipolar_geometry<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, G, rays_dev, xy_dev, cand_dev);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  return VSLAM_OK;
}

int vs_launch_epipolar(vslam_ctx* ctx, int stream, int src_kf, int level, int n, const EpiCand* cand_dev, const double* unproj_dev, int subpix_its, int* out_int_dev, double* out_pos_dev) {
  if (n <= 0) return VSLAM_OK;
  { const int rc_ = vs_ensure_lists(ctx); if (rc_) return rc_; }
  k_epipolar<<<(n + kSearchWarps - 1) / kSearchWarps, kSearchWarps * 32, 0, ctx->stream>>>(make_dev(ctx), stream, src_kf, level, n, cand_dev, unproj_dev, subpix_its, out_int_dev, out_pos_dev);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  return VSLAM_OK;
}

// k_pose_fast: the ten Gauss-Newton iterations of a TrackMap stage (jni/Tracker.cc:464-489 coarse, :543-577 fine; CalcPoseUpdate :683-774)
// with the found points RESIDENT on the SM for the whole stage.
//
// k_pose (track.cu) re-reads every found point's 12 Jacobian entries, error and weight inputs from L2 in each of the ten iterations
// (about 136 KB per stream and iteration) and spends its time waiting for those round trips.  Here a CTA (one per stream, 256 threads)
// keeps, for the first 1024 found points of its stream:
//   * the Jacobian rows, already scaled by dSqrtInvNoise (= 2^-level, an exact scaling, so J.v and the normal equations see bit-identical
//     values), in shared memory: 12 x 1024 doubles = 96 KB, two CTAs per SM;
//   * projected position, found position and the noise scale in registers, four points per thread.
// Global memory is only WRITTEN inside the iterations (the per-point state other entry points and the tests read back); found points
// beyond the 1024th (maps much larger than MaxPatchesPerFrame) take a spill path that works on global memory like k_pose.
// Median: 11-bit radix select straight from the registers.  Normal equations: per-thread partial sums in point order, a 28-shuffle
// transposing warp reduction, then the eight warp partials in warp order -- a fixed shape, hence run-to-run deterministic.
// `serial` (vslam_params.serial_normal_equations): the 27 sums are instead accumulated by 27 lanes in LIST order without multiply-add
// contraction, i.e. in the reference's own order of operations (jni/myWLS.h:39-50) -- bit-identical to the oracle given identical inputs,
// and slower (a chain of 2 n dependent additions per iteration).
#include "pose_common.cuh"
#include <cstdio>
#include <cstdlib>

// Two CTA shapes of the same code (pose_fast_impl.cuh):
//   pf256: 256 threads x 4 resident points (1024 per stream), 96 KB of shared memory, two CTAs per SM -- contexts with more streams than SMs;
//   pf512: 512 threads x 4 resident points (2048 per stream), 192 KB, one CTA per SM -- for the streams of a context with at most one stream per
//          SM whose iteration list is longer than 1024 entries (the reference's MaxPatchesPerFrame quirk lets through ~2000 points at 4K with a
//          20000-point map): they stay off the spill path (4K: fine pose 1.10 -> 0.51 ms).  Measured and NOT done: pf512 for every stream of such a
//          context -- at the tracker's ~1000 found points it is no faster (32 streams: 0.123 against 0.122 ms; the stage is a chain of barriers and
//          serial sections, not of per-thread work) and, filling the SM, it leaves no room for the next frame's front end (128 streams with
//          look-ahead: 0.412 -> 0.481 ms per step).
// Both shapes are launched where lists can be that long (vs_launch_pose_fast); a CTA whose stream belongs to the other shape returns at once.  The shapes differ in the fixed shape of the parallel sums (16 instead of 8 warp partials), i.e. in the last bits of the normal equations.
#define PF_NS pf256
#define PF_FT 256
#define PF_PP 4
#define PF_MINB 2
#include "pose_fast_impl.cuh"
#undef PF_NS
#undef PF_FT
#undef PF_PP
#undef PF_MINB
#define PF_NS pf512
#define PF_FT 512
#define PF_PP 4
#define PF_MINB 1
#include "pose_fast_impl.cuh"
#undef PF_NS
#undef PF_FT
#undef PF_PP
#undef PF_MINB

// Stages 1 (coarse) and 2 (fine) of TrackMap; mode bit 2 = with motion model / quality (the tail of TrackFrame)
template <class Smem, class Kernel>
static int launch_pose_fast(vslam_ctx* ctx, int mode, Kernel kernel, int threads, int slot, int sel) {
  const Dev D = make_dev(ctx);
  const size_t smem = sizeof(Smem);
  if (smem > ctx->smem_attr[slot]) {
    VS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));   // needs the whole 228 KB
    ctx->smem_attr[slot] = smem;
    if (getenv("VSLAM_DEBUG_OCCUPANCY")) {
      int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem);
      fprintf(stderr, "k_pose_fast (%d threads): %zu bytes of shared memory, %d CTAs per SM\n", threads, smem, nb);
    }
  }
  vs_time_begin(ctx, (mode & 3) == 2 ? VS_ST_POSE_FINE : VS_ST_POSE_COARSE);
  VS_CUDA(vs_launch_pdl(kernel, dim3(ctx->cur_cnt), dim3(threads), smem, ctx->stream, ctx->pdl && !ctx->timing, D, mode & 3, (mode >> 2) & 1, ctx->params.serial_normal_equations, sel));
  vs_time_end(ctx);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  return VSLAM_OK;
}
int vs_launch_pose_fast(vslam_ctx* ctx, int mode) {
  static int n_sm = 0;
  if (!n_sm) { int dev = 0; cudaGetDevice(&dev); if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm < 1) n_sm = 1; }
  // (by the context's stream count and map size, not the launch's: every frame of a context is launched the same way)
  // pf512 beside pf256 only where lists longer than 1024 entries are likely -- the fine stage of maps of more than 8192 points (every point
  // searched at level 3 is in the list, the rest fills it up to MaxPatchesPerFrame, jni/Tracker.cc:499-527) -- because even its CTAs that return
  // at once need an SM all to themselves to start: beside the next frame's front end that wait costs 0.05 ms per launch (1080p x 148 streams,
  // 5000 points, lists of 1000: 1.76 -> 1.86 ms per step with the second launch)
  const bool both = ctx->S <= n_sm && ctx->map.n > 8192 && (mode & 3) == 2;
  if (const char* e = getenv("VSLAM_POSE_SHAPE")) {   // A/B runs: one shape for every stream
    if (atoi(e) == 512) return launch_pose_fast<pf512::FastSmem>(ctx, mode, pf512::k_pose_fast, 512, 3, 0);
    return launch_pose_fast<pf256::FastSmem>(ctx, mode, pf256::k_pose_fast, 256, 2, 0);
  }
  const int rc = launch_pose_fast<pf256::FastSmem>(ctx, mode, pf256::k_pose_fast, 256, 2, both ? 1 : 0);
  if (rc || !both) return rc;
  return launch_pose_fast<pf512::FastSmem>(ctx, mode, pf512::k_pose_fast, 512, 3, 2);
}

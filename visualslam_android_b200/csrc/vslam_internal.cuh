// Internal layout of a vslam_ctx and device helpers shared by the kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <cstring>
#include <vector>
#include "vslam_b200.h"

#define VS_LEVELS VSLAM_LEVELS
#define VS_MAXP VSLAM_MAX_PATCH
#define VS_TMPL_BYTES 144           // template rows padded to 12 bytes (3 words, the dp4a operand layout): 11 * 12 = 132, rounded to 16
#define VS_MAX_STRIP_ROWS 64        // upper bound of LevelDesc::strip_rows
#define VS_MAX_GROUPS 4             // stream groups of vs_launch_frame
#define VS_MAX_UPDATES 20           // 10 coarse + 10 fine CalcPoseUpdate calls per TrackMap

// ------------------------------------------------------------------------------------------------
// Device-visible description of one pyramid level of the per-stream "current keyframe".
struct LevelDesc {
  int w, h, pitch;            // pitch: bytes between rows of the ctx-owned level image (multiple of 128)
  int cap;                    // corner capacity per stream
  int strip_rows;             // rows of this level handled by one CTA of the pyramid / FAST kernels (vs_strip_rows)
  int n_strips;               // ceil(h / strip_rows)
  // 32-bit reciprocals (0xffffffff / d + 1; 0 when d == 1) for the small divisions of the FAST kernels: d = n_strips, ceil(w / 16), ceil(w / 32), w / 2, w
  uint32_t mg_strips, mg_cpr, mg_wpr, mg_hw, mg_w;
  uint8_t* img;               // [S][h][pitch]      (level 0: only used by the host-input path, see l0_ptr)
  uint32_t* corners;          // [S][cap]           packed (y << 16 | x), raster order
  int* lut;                   // [S][h + 1]         lut[y] = #corners with row < y ; lut[h] = total
  uint32_t* cbits;            // [S][h][ceil(w / 32)] corner bitmask of the frame being processed (bit x & 31 of word x >> 5; cleared per frame, inside ctx->sync_words)
};

struct CamDev { double fx, fy, cx, cy, W, Winv, twoTan, oneOver2Tan, distEnabled, largestRadius, maxR, width, height; };

// Per (stream, point) tracker state, SoA with index s * N + i   (TrackerData + PatchFinder, jni/TrackerData.h, jni/PatchFinder.h)
struct PointState {
  double* v3cam;      // [3][S*N]
  double* v2image;    // [2][S*N]
  double* derivs;     // [4][S*N]  row-major 2x2
  double* warpinv;    // [4][S*N]
  double* m2;         // [4][S*N]  inverse(warpinv) * 2^level (valid when level >= 0)
  double* lastwarp;   // [4][S*N]
  double* v2found;    // [2][S*N]
  double* coarse;     // [2][S*N]
  double* jac;        // [12][S*N]
  double* err;        // [2][S*N]
  double* sqrtinv;    // [S*N]
  int* flags;         // [S*N] bit0 inImage, bit1 searched, bit2 found, bit3 didSubPix, bit4 templateBad, bit5 haveLast, bit6 hasTData
  int* level;         // [S*N] nSearchLevel (-1 = rejected)
  int* rlevel;        // [S*N] the level the warp loop reached (PatchFinder::mnSearchLevel, also for rejected warps)
  uint8_t* tmpl;      // [S*N][VS_TMPL_BYTES]
  int* tsum;          // [2][S*N] sum, sumsq
  int* counts;        // [2][S*N] outlier, inlier
};
enum { F_INIMAGE = 1, F_SEARCHED = 2, F_FOUND = 4, F_SUBPIX = 8, F_TBAD = 16, F_HAVELAST = 32, F_HASTD = 64,
       F_NEWTMPL = 128 /* the last search regenerated the template (statistics only) */ };

struct MapDev {
  int n;
  double* world;   // [n][3]
  double* right;   // [n][3]
  double* down;    // [n][3]
  int* ircenter;   // [n][2]
  int* srclevel;   // [n]
  int* srckf;      // [n]
};

struct SourceKF {   // device pyramids of the map's source keyframes: level l of keyframe k at img[l] + k*h*pitch
  int w[VS_LEVELS], h[VS_LEVELS], pitch[VS_LEVELS];
  uint8_t* img[VS_LEVELS];
};

// One candidate of MapMaker::AddPointEpipolar, line geometry resolved on the host (api.cu vslam_epipolar_search)
struct EpiCand {
  double nx, ny, ax, ay;        // v2Normal, v2AlongProjectedLine
  double normDist, minLen, maxLen, maxDistSq;
  int x, y, valid, pad;         // candidate position in the source level; valid = 0: rejected before the search
};

// Inputs of the epipolar line geometry shared by all candidates of one vslam_epipolar_search call
struct EpiGeom { double src_pose[12], tgt_pose[12]; double start_depth, end_depth, max_dist_sq, largest_radius; };

// Per-stream tracker scalars (Tracker members)
struct StreamState {
  double pose[12], start_pose[12];
  double velocity[6], sbi_rot[6];
  double msd_scaled_vel, vel_mag, depth_mean, depth_sigma;
  int attempted[VS_LEVELS], found[VS_LEVELS];
  int quality, lost_frames, did_coarse, just_recovered;
  int rng_ring[31]; int rng_f, rng_b;
  int nA, nB, nB_top;           // iteration list = [0,nA) coarse set, [nA, nA+nB) fine set (first nB_top entries: top level, sub-pixel)
  int try_coarse, coarse_range, n_updates, quirk_stale_cache;
  double updates[VS_MAX_UPDATES * 6], sigmas[VS_MAX_UPDATES];
  // relocalisation (k_relocalise): set for the frame in which the stream was recovered; best keyframe / final ESM score of the last attempt
  int recovered, reloc_best, n_recoveries, pad_; double reloc_score;
  // keyframe hand-off (vslam_set_keyframe_policy): Tracker::mnFrame / mnLastKeyFrameDropped, and what the last frame decided
  int frame_no, last_kf_dropped, kf_request, kf_closest; double kf_dist;
};

// The buffers one frame's pose-independent front end (pyramid, FAST) writes and its pose-dependent back end (projection, patch search) reads.
// A context has one set, or two when frame look-ahead is on (vs_begin_frame): frame k lives in set k & 1, so that the front end of frame
// k + 1 can run beside the back end of frame k.  ctx->lev[l].img / .cbits, ctx->l0_ptr / l0_stride (+ host mirrors) always alias the CURRENT set.
struct FrameSet {
  uint8_t* img[VS_LEVELS];          // level images ([0]: the ctx-owned level-0 buffer of the host-input paths; may be null in set 1 until needed)
  uint32_t* cbits[VS_LEVELS];       // corner bitmasks
  unsigned long long* cbits_block;  // their allocation (set 1 only; set 0's live in ctx->sync_words)
  const uint8_t** l0_ptr; int* l0_stride; const uint8_t** l0_ptr_host; int* l0_stride_host;
};

struct vslam_ctx {
  vslam_config cfg;
  vslam_params params;
  cudaStream_t stream;
  bool own_stream;
  int S, N, P;
  LevelDesc lev[VS_LEVELS];
  const uint8_t** l0_ptr;        // [S] device: level-0 image of each stream (ctx-owned or adopted user buffer)
  int* l0_stride;                // [S] device
  const uint8_t** l0_ptr_host; int* l0_stride_host;
  // double-buffered host-input pipeline (vslam_track_frame_async): level-0 buffer 0 is lev[0].img, buffer 1 is l0_alt
  uint8_t* l0_alt; cudaStream_t copy_stream; cudaEvent_t ev_copied[2], ev_computed[2], ev_done[2]; long long step; bool pipe_ready; int* status_pin;   // [2][4] pinned copy of `status` per slot
  // vs_launch_frame: stream groups (group 0 runs on ctx->stream) and, per group, a side stream on which SmallBlurryImage +
  // projection run beside the FAST pass of levels 1..3
  cudaStream_t group_stream[VS_MAX_GROUPS], side_stream[VS_MAX_GROUPS]; cudaEvent_t ev_fork[VS_MAX_GROUPS], ev_join[VS_MAX_GROUPS], ev_end[VS_MAX_GROUPS], ev_begin;
  int cur_s0, cur_cnt, cur_group;   // stream range / group the launchers act on (0, S, 0 outside vs_launch_frame)
  unsigned long long* sync_words; size_t sync_words_n;   // corner bitmasks (LevelDesc::cbits) of levels 0..3, then the tickets: one allocation, one memset per frame
  unsigned* tickets;             // [2 * VS_MAX_GROUPS] device (inside sync_words)
  int* status;                   // [4] device: [0] capacity overflow flag
  CamDev cam;
  MapDev map;
  SourceKF src; int n_src;
  std::vector<char> src_have;    // [n_src] which source keyframes have been uploaded (map files, mapfile.cu)
  std::vector<int> reloc_ids; std::vector<double> reloc_poses_host;   // host copy of the vslam_set_reloc_keyframes registration
  PointState ps;
  StreamState* ss;               // [S] device
  int* lists;                    // [S][list_cap] iteration / search lists
  int list_cap;
  int* pvs;                      // [S][N] scratch for the per-level potentially-visible sets
  double* sort_scratch;          // [S][sort_cap]
  int sort_cap;
  unsigned long long* evals;     // device counter
  unsigned long long launches;
  void* scratch_host; size_t scratch_host_bytes;   // pinned staging
  // MakeKeyFrame_Rest results of one stream (lazy scratch) and the per-stream keyframe snapshot used by MiniPatch trail tracking
  int* rest_scores; uint32_t* rest_max; uint32_t* rest_cand; double* rest_cand_score; int* rest_counts; size_t rest_off[VS_LEVELS]; int rest_stream;
  uint8_t* snap_img; uint32_t* snap_corners; int* snap_lut;
  // relocaliser keyframes (vslam_set_reloc_keyframes): SmallBlurryImages with blur 2.5, their gradient images and poses
  int reloc_n; float reloc_taps[17]; float* reloc_tmpl; float* reloc_jac; float* reloc_tmp; uint8_t* reloc_small; double* reloc_pose; double* reloc_scores;
  // keyframe policy (vslam_set_keyframe_policy): the MapMaker heuristics the tracker consults, over the registered keyframes' poses
  bool kf_policy; double kf_wiggle, kf_wiggle_dn, kf_mult; int kf_min_frames;
  int* kf_req;                   // [S] device: compact mirror of StreamState::kf_request (one small D2H copy per poll)
  void* epi_buf; size_t epi_cap;    // scratch of vslam_epipolar_search (candidates, rays, results), grown on demand
  double* unproj_lut; bool unproj_ok;   // [H][W][2] ATANCamera::UnProject of every integer level-0 pixel (MapMaker::AddPointEpipolar's imUnProj), built on the host
  // on-device SmallBlurryImage (vslam_enable_sbi)
  int sbi_exact_half = 1; int* sbi_resize_tab = nullptr;   // cv::resize of level 3 to the SmallBlurryImage size: see sbi.cu ResizeTab
  bool sbi_on; float sbi_taps[9]; CamDev sbi_cam; double sbi_orig[2][3]; float* sbi_tmpl; float* sbi_scratch; float* sbi_jac; uint8_t* sbi_small; int* sbi_have;
  // optional per-kernel timing with CUDA events on ctx->stream (vslam_set_timing)
  bool timing; std::vector<cudaEvent_t> ev_pool; std::vector<int> ev_stage; size_t ev_used;
  size_t smem_attr[4] = {0, 0, 0, 0};   // dynamic shared memory already opted into, per kernel (cudaFuncSetAttribute once, not per frame)
  void* pf_buf = nullptr; size_t pf_cap = 0;   // scratch of the per-object PatchFinder calls (patchfinder_ops.cu)
  std::vector<int> user_events;                // [S] pending user events (vslam_user_event)
  int* list_counts = nullptr; size_t list_counts_cap = 0;   // per-chunk corner counts of k_corner_count / k_corner_lists
  // frame look-ahead (vslam_params.frame_lookahead, vs_begin_frame / vs_launch_frame): two frame sets, the front end of a frame on front_stream
  FrameSet sets[2] = {}; int cur_set = 0; bool have_set1 = false;
  cudaStream_t back_stream = nullptr;   // the back end of a look-ahead frame: a high-priority stream of the library's own (the front-end streams have the lowest priority), joined to ctx->stream
  cudaEvent_t ev_user = nullptr;
  cudaStream_t front_stream = nullptr, front_side = nullptr; cudaEvent_t ev_front_done = nullptr, ev_barrier = nullptr, ev_back_done[2] = {nullptr, nullptr}, ev_la_fork = nullptr, ev_la_join = nullptr;
  cudaStream_t front = nullptr;  // inside a vslam_track_frame* call: the stream the frame's input and front end are enqueued on (front_stream or ctx->stream)
  bool la_frame = false;         // the frame being enqueued runs with look-ahead
  bool la_unavailable = false;   // the second frame set could not be allocated
  unsigned long long launches_after_frame = ~0ull;   // ctx->launches when the last look-ahead frame had been enqueued: any kernel launched since forces a full barrier
  double* sbi_rot_buf = nullptr; const double* cur_sbi_rot = nullptr;   // [2][S][6] k_sbi's result per frame set; what k_project_lists of this frame reads (null: StreamState::sbi_rot as set by the host)
  float* reloc_frame_scratch = nullptr; uint8_t* reloc_frame_small = nullptr;      // k_relocalise's own scratch ([S][3n] / [S][n]): it may run beside the next frame's k_sbi
  // vs_launch_track_map_rest: the coarse-stage chain of a frame (streams that try the coarse stage) runs on a stream of its own beside the fine-only chain
  int* coarse_hint_host = nullptr; int* coarse_hint_dev = nullptr;   // [S] mapped pinned memory: k_project_lists leaves each stream's try_coarse of its latest frame here; read by the host WITHOUT synchronisation, as a hint for the launch layout only
  bool rand_jump_ready = false;  // g_rand_jump (track.cu) has been uploaded to this context's device
  int cur_chain = -1; cudaStream_t chain_stream[VS_MAX_GROUPS] = {}, chain_stream_hi = nullptr; cudaEvent_t ev_chain_fork[VS_MAX_GROUPS] = {}, ev_chain_join[VS_MAX_GROUPS] = {};
  bool pdl = true;               // programmatic dependent launch of a frame's kernels (vs_launch_pdl); VSLAM_PDL=0 turns it off
  bool lists_stale = false;      // the last tracked frame left corner bitmasks only: corner lists / row LUTs are built on demand (vs_ensure_lists)
  std::string err;
};

#define VS_CUDA(call)                                                                                 \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                  \
      return VSLAM_E_CUDA;                                                                            \
    }                                                                                                 \
  } while (0)

enum { VS_ST_PYR0 = 0, VS_ST_PYR1, VS_ST_PYR2, VS_ST_PYR3, VS_ST_PROJECT, VS_ST_SEARCH_COARSE, VS_ST_POSE_COARSE, VS_ST_SEARCH_FINE, VS_ST_POSE_FINE, VS_ST_H2D, VS_ST_OTHER };
// Bracket one launch with two events of the pool (no-ops unless timing is on).
inline void vs_time_begin(vslam_ctx* ctx, int stage) {
  if (!ctx->timing) return;
  if (ctx->ev_used + 2 > ctx->ev_pool.size()) { for (int k = 0; k < 64; k++) { cudaEvent_t e; cudaEventCreate(&e); ctx->ev_pool.push_back(e); } }
  ctx->ev_stage.push_back(stage);
  cudaEventRecord(ctx->ev_pool[ctx->ev_used++], ctx->stream);
}
inline void vs_time_end(vslam_ctx* ctx) { if (ctx->timing) cudaEventRecord(ctx->ev_pool[ctx->ev_used++], ctx->stream); }

// Launch with programmatic dependent launch (sm_90+): the kernel's CTAs may be scheduled while the previous kernel of the stream drains -- its
// launch latency and ramp-up hide behind that kernel's tail -- and the kernel itself waits for the previous kernel's completion and memory
// (cudaGridDependencySynchronize() as its FIRST statement; a no-op in an ordinary launch).  pdl = false: an ordinary launch.
template <class... KArgs, class... Args>
inline cudaError_t vs_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// kernels/launchers implemented in the .cu files
int vs_strip_rows(int level, int w, int pitch);
int vs_launch_pyramid_fast(vslam_ctx* ctx, int first_stream, int count);   // = vs_launch_pyramid_l0 + vs_launch_fast_levels + vs_launch_corner_lists
int vs_launch_pyramid_l0(vslam_ctx* ctx, int first_stream, int count);
int vs_launch_fast_levels(vslam_ctx* ctx, int first_stream, int count);
int vs_launch_corner_lists(vslam_ctx* ctx, int first_stream, int count);   // corner lists + row LUT of all levels from the corner bitmasks
int vs_ensure_lists(vslam_ctx* ctx);                                       // ... of every stream, if the last tracked frame skipped them
int vs_launch_source_pyramid(vslam_ctx* ctx, int kf_id);
int vs_launch_project_all(vslam_ctx* ctx, int build_lists);
int vs_launch_search(vslam_ctx* ctx, int which /*0 explicit list,1 coarse A,2 fine B*/, int range, int subpix, int sflags /*1: ReFind_Common variant*/);
int vs_launch_pose(vslam_ctx* ctx, int mode, double sigma, int mark, int apply);
int vs_launch_search_fast(vslam_ctx* ctx, int which, int range, int subpix, int sflags);   // search_fast.cu: eight lanes per map point + k_subpix
int vs_launch_pose_fast(vslam_ctx* ctx, int mode);   // pose_fast.cu: stages 1 / 2 with the found points resident on the SM
int vs_launch_track_map(vslam_ctx* ctx, int with_motion_model);
int vs_launch_track_map_rest(vslam_ctx* ctx, int with_motion_model);   // everything after vs_launch_project_all
int vs_launch_frame(vslam_ctx* ctx);                                   // pyramid + FAST, SmallBlurryImage, TrackMap of all streams
int vs_launch_project_and_derivs(vslam_ctx* ctx, int only_found);
int vs_launch_calc_jacobians(vslam_ctx* ctx);
int vs_launch_sbi(vslam_ctx* ctx);          // SmallBlurryImage pair + CalcSBIRotation of the frame (pose independent: part of the front end)
int vs_launch_relocalise(vslam_ctx* ctx);   // k_relocalise when relocaliser keyframes are registered (back end: needs the stream's lost state)
inline cudaStream_t vs_in_stream(vslam_ctx* ctx) { return ctx->front ? ctx->front : ctx->stream; }   // where a frame's input (copies, pointer tables) is enqueued
int vs_launch_reloc_make(vslam_ctx* ctx, const int* src_ids_dev);
int vs_launch_epipolar_geometry(vslam_ctx* ctx, int n, const EpiGeom& G, const double* rays_dev, const int* xy_dev, EpiCand* cand_dev);
int vs_launch_epipolar(vslam_ctx* ctx, int stream, int src_kf, int level, int n, const EpiCand* cand_dev, const double* unproj_dev, int subpix_its, int* out_int_dev, double* out_pos_dev);
int vs_keyframe_rest(vslam_ctx* ctx, int stream);
int vs_minipatch_sample(vslam_ctx* ctx, int stream, int which, const int* xy_dev, int n, uint8_t* patches_dev);
int vs_minipatch_find(vslam_ctx* ctx, int stream, int which, const uint8_t* patches_dev, int n, double* pos_dev, int* found_dev, int* best_dev, int range, int max_ssd);

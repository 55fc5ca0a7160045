/* vslam_b200.h — C-ABI of the B200-native PTAM tracking front-end.
 *
 * Drop-in boundary for the per-frame tracking path of ahcorde/visualSLAM_Android (jni/).  Plain
 * pointers and sizes only; no C++/torch types.  One context per GPU; a context tracks `n_streams`
 * independent cameras ("streams") of one image size against one shared, read-only map, and every
 * batched entry point processes all streams in one set of kernel launches.
 *
 * Reference interface each entry point stands in for (file:line under /root/reference):
 *   vslam_make_keyframe_lite[_dev]   KeyFrame::MakeKeyFrame_Lite           jni/KeyFrame.cc:5-51, jni/KeyFrame.h:89
 *                                    (cv::resize 2:1 pyramid :20-23, cvCornerFast_10 jni/vision/cvfast.cpp:6088-9241,
 *                                     row LUT :41-49)
 *   vslam_upload_source_keyframe     the map keyframe a MapPoint's patch comes from (MapPoint::pPatchSourceKF,
 *                                    jni/MapPoint.h:38); pyramid built like MakeKeyFrame_Lite
 *   vslam_set_map                    Map::vpPoints / MapPoint fields read by the tracker (jni/Map.h:29, jni/MapPoint.h:33-54)
 *   vslam_append_map_points          Map::vpPoints.push_back of new points during tracking (jni/MapMaker.cc:685, jni/Tracker.cc:372)
 *   vslam_set_camera                 ATANCamera scalars after RefreshParams (jni/ATANCamera.cc:37-129)
 *   vslam_project_all                first loop of Tracker::TrackMap       jni/Tracker.cc:369-392
 *                                    (TrackerData::Project jni/TrackerData.h:69-86, GetDerivsUnsafe :92-95,
 *                                     PatchFinder::CalcSearchLevelAndWarpMatrix jni/PatchFinder.cc:31-68)
 *   vslam_search_for_points          Tracker::SearchForPoints              jni/Tracker.cc:629-674
 *                                    (MakeTemplateCoarseCont jni/PatchFinder.cc:79-125, FindPatchCoarse :170-235,
 *                                     ZMSSDAtPoint :352-380, MakeSubPixTemplate/IterateSubPixToConvergence :242-350)
 *   vslam_calc_pose_update           Tracker::CalcPoseUpdate               jni/Tracker.cc:683-774 (+ Tukey, myWLS<6>)
 *   vslam_track_map                  Tracker::TrackMap                     jni/Tracker.cc:358-626
 *   vslam_track_frame[_dev|_async]   Tracker::TrackFrame (good-map branch) jni/Tracker.cc:76-140
 *                                    (ApplyMotionModel :781-798, UpdateMotionModel :802-820, AssessTrackingQuality :832-878;
 *                                     lost streams: AttemptRecovery :167-180 with vslam_set_reloc_keyframes)
 *   vslam_set_reloc_keyframes        Relocaliser::AttemptRecovery / ScoreKFs       jni/Relocaliser.cc:17-58
 *   vslam_reset_stream               Tracker::Reset                               jni/Tracker.cc:45-60
 *   vslam_set_keyframe_policy        MapMaker::NeedNewKeyFrame / IsDistanceToNearestKeyFrameExcessive as called from TrackFrame
 *                                    jni/Tracker.cc:127-132,869-871; jni/MapMaker.cc:705-773,1098-1101
 *   vslam_add_keyframe_from_stream   Tracker::AddNewKeyFrame + MapMaker::AddKeyFrame jni/Tracker.cc:823-827, jni/MapMaker.cc:470-478
 *   vslam_save_map_file / _load_     (no reference counterpart: the reference keeps its map in memory only)
 *   vslam_export_map_text            MapMaker::GUICommandHandler("SaveMap") dump  jni/MapMaker.cc:1254-1297
 *   vslam_refind                     MapMaker::ReFind_Common                      jni/MapMaker.cc:967-1036
 *   vslam_epipolar_search            MapMaker::AddPointEpipolar (the search)      jni/MapMaker.cc:525-640
 *   vslam_epipolar_make_points       MapMaker::AddPointEpipolar (the new point)   jni/MapMaker.cc:646-690 (ReprojectPoint :176-200, MapPoint::RefreshPixelVectors)
 *   vslam_make_keyframe_rest         KeyFrame::MakeKeyFrame_Rest           jni/KeyFrame.cc:53-95 (fast_nonmax jni/vision/cvfast.cpp:9243-9405,
 *                                    FindShiTomasiScoreAtPoint jni/vision/ImageHandler.cpp:124-155)
 *   vslam_minipatch_sample / _find   MiniPatch::SampleFromImage / FindPatch jni/MiniPatch.cc:32-83
 *   vslam_enable_sbi                 SmallBlurryImage + Tracker::CalcSBIRotation jni/SmallBlurryImage.cc, jni/Tracker.cc:86-97,885-893
 *   vslam_create / vslam_destroy     JNI native_createTest / native_disposeTest   jni/jni_part.cpp:114-123
 *   vslam_track_frame                JNI native_update                            jni/jni_part.cpp:132-145
 *
 * All functions return 0 on success and a negative VSLAM_E_* code on failure; vslam_last_error() gives the text.
 * Nothing throws across this boundary.  Work is enqueued on the context's CUDA stream; the vslam_get_* readers and
 * vslam_sync() wait for it.  Host frame buffers passed to the *_lite / track_frame calls are read asynchronously: keep them
 * valid until the next vslam_sync / vslam_get_* / vslam_wait_step; all other host inputs are consumed during the call.
 * A context is not thread-safe; distinct contexts are independent (no globals).
 */
#ifndef VSLAM_B200_H
#define VSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSLAM_LEVELS 4            /* jni/KeyFrame.h:31 */
#define VSLAM_MAX_PATCH 11        /* int32 ZMSSD arithmetic of jni/PatchFinder.cc:379 overflows above 11 */

#define VSLAM_OK 0
#define VSLAM_E_INVALID (-1)      /* bad argument */
#define VSLAM_E_CUDA (-2)         /* CUDA runtime error (text in vslam_last_error) */
#define VSLAM_E_CAPACITY (-3)     /* more corners than max_corner_frac allows: reported, never silently truncated */
#define VSLAM_E_NO_DEVICE (-4)    /* no usable CUDA device: there is no CPU fallback */
#define VSLAM_E_IO (-5)           /* map file cannot be opened / is truncated / fails its checksum */

typedef struct vslam_ctx vslam_ctx;

typedef struct vslam_config {
  int device;               /* CUDA device ordinal */
  int width, height;        /* level-0 size; width % 32 == 0, height % 8 == 0 (every level keeps even dimensions) */
  int n_streams;            /* independent cameras tracked by this context */
  int max_points;           /* map capacity N */
  int patch_size;           /* PatchFinder template side P: 8 or 11 (reference default 11, jni/PatchFinder.h:48) */
  int max_source_keyframes; /* device-resident source pyramids */
  float max_corner_frac;    /* corner-list capacity per level as a fraction of the level's pixels (0 => 0.5) */
  void* cuda_stream;        /* cudaStream_t to enqueue on (NULL => the context creates its own) */
  int truncate_error;       /* 1 = reproduce the (int) cast of jni/Tracker.cc:766-767 (reference behaviour) */
  unsigned rand_seed;       /* per-stream glibc rand() state starts as srand(rand_seed) would leave it (reference: 1) */
} vslam_config;

/* Tunables the reference hard-codes (jni/Tracker.cc:405-410,495-497,518); vslam_default_params() gives its values. */
typedef struct vslam_params {
  unsigned coarse_min, coarse_max, coarse_range;
  int coarse_subpix_its;
  double coarse_min_vel;
  int fine_range, fine_range_after_coarse, fine_subpix_its_top_level;
  int max_patches_per_frame;
  int use_sbi;              /* Tracker::mbUseSBIInit (jni/Tracker.cc:87-89): motion model takes rotation from sbi_rot */
  int stream_groups;        /* execution only, no effect on results: vslam_track_frame* can split the streams into 1..4 groups whose
                               kernels run on separate CUDA streams.  Default 1: on B200 with 256 VGA streams, 2-4 groups measured 5 %
                               SLOWER (1.16 vs 1.10 ms per step) — the smaller grids cost more in tails than the overlap of one
                               group's latency-bound kernels with another's throughput-bound ones gives back */
  int serial_normal_equations; /* 0 (default): the 27 sums of CalcPoseUpdate's normal equations (jni/myWLS.h:39-50) are a fixed-shape parallel
                               reduction with fused multiply-add.  1: accumulated serially in the reference's own order (list order, row 0
                               then row 1, no contraction): bit-identical sums given identical inputs, and a chain of 2 n dependent
                               additions per iteration (measured cost: DESIGN.md section 4.4) */
  int pose_kernel;          /* execution only: 0 (default) = k_pose_fast (found points resident on the SM), 1 = k_pose (the round-1 kernel) */
  int search_kernel;        /* execution only: 0 (default) = k_search_fast + k_subpix (eight lanes per map point), 1 = k_search (one warp per point, round 1) */
  int frame_lookahead;      /* execution only, no effect on results: 1 = vslam_track_frame* keep two sets of level images / corner bitmasks and run the
                               pose-independent front end of a frame (pyramid, FAST, SmallBlurryImage rotation) on a second CUDA stream, beside the
                               previous frame's projection / patch search / pose iterations, whenever the caller has issued that frame before the
                               previous one finished (the calls are asynchronous); 0 = off; -1 (default) = the library decides from the workload
                                */
  int coarse_chain;         /* execution only, no effect on results: whether a stream runs TrackMap's coarse stage is decided on the device, so its
                               kernels are launched for every frame and, at ordinary camera speed, find nothing to do -- each still waiting for the one
                               before.  1 = a frame's back end forks into two launch chains, coarse stage + fine stage for the streams that try the coarse
                               stage and the fine stage alone for the others, so that an empty chain runs BESIDE the other one instead of in front of it;
                               0 = one chain; -1 (default) = two chains while no stream of the context tried the coarse stage in its latest frame, for contexts
                               of at least 48 streams (fewer: the host's launch rate decides) and maps of at most 2040 points */
} vslam_params;

void vslam_default_config(vslam_config* cfg);
void vslam_default_params(vslam_params* p);

int vslam_create(const vslam_config* cfg, vslam_ctx** out);
void vslam_destroy(vslam_ctx* ctx);
const char* vslam_last_error(const vslam_ctx* ctx);   /* ctx may be NULL: error of the last failed vslam_create */
int vslam_sync(vslam_ctx* ctx);
int vslam_set_params(vslam_ctx* ctx, const vslam_params* p);

/* cam13: fx fy cx cy W Winv 2tan(W/2) 1/(2tan) distortionEnabled largestRadius maxR width height */
int vslam_set_camera(vslam_ctx* ctx, const double* cam13);
/* Host-side mirror of ATANCamera::RefreshParams: fills cam13 from the 5 normalised parameters and an image size.
 * as_shipped_radius = 1 reproduces the int-temporary bug of jni/ATANCamera.cc:70-82 (largestRadius = maxR = 0). */
void vslam_camera_from_params(const double* params5, int width, int height, int as_shipped_radius, double* cam13);

int vslam_upload_source_keyframe(vslam_ctx* ctx, int kf_id, const uint8_t* gray_host, int stride);
int vslam_set_map(vslam_ctx* ctx, int n, const double* world3, const double* pixel_right3, const double* pixel_down3,
                  const int32_t* ir_center2, const int32_t* src_level, const int32_t* src_kf /* NULL => all 0 */);
/* Map::vpPoints.push_back while streams run (what MapMaker::AddPointEpipolar does from its thread, jni/MapMaker.cc:685): new points are
 * appended behind the existing ones.  Unlike vslam_set_map the existing points keep their per-stream tracker state (template cache,
 * found flags, M-estimator counters); the new points start without, like a MapPoint whose TrackerData is created on first use. */
int vslam_append_map_points(vslam_ctx* ctx, int n_new, const double* world3, const double* pixel_right3, const double* pixel_down3,
                            const int32_t* ir_center2, const int32_t* src_level, const int32_t* src_kf /* NULL => all 0 */);

/* ---- KeyFrame::MakeKeyFrame_Lite for streams [first, first+count) -------------------------------------------- */
/* gray_host: frame k at gray_host + k*frame_stride, rows `stride` bytes apart (pinned memory recommended). */
int vslam_make_keyframe_lite(vslam_ctx* ctx, int first_stream, int count, const uint8_t* gray_host, int stride, size_t frame_stride);
/* gray_dev: same layout in device memory (16-byte aligned, stride % 16 == 0).  Zero-copy: the buffer becomes level 0 of
 * those streams' current keyframe and must stay unmodified until their next make_keyframe_lite / track_frame. */
int vslam_make_keyframe_lite_dev(vslam_ctx* ctx, int first_stream, int count, const uint8_t* gray_dev, int stride, size_t frame_stride);
/* Source keyframe kf_id as stream `stream`'s current keyframe (zero copy of its level-0 image, then pyramid + FAST): what a host
 * MapMaker does to make an OLD keyframe the target of vslam_epipolar_search / vslam_refind, which need the target's corners. */
int vslam_make_keyframe_from_source(vslam_ctx* ctx, int stream, int kf_id);
int vslam_level_dims(const vslam_ctx* ctx, int level, int* width, int* height);
int vslam_get_level(vslam_ctx* ctx, int stream, int level, uint8_t* out, int out_stride);
int vslam_get_num_corners(vslam_ctx* ctx, int stream, int level, int* n);
int vslam_get_corners(vslam_ctx* ctx, int stream, int level, int32_t* xy, int cap);     /* (x,y) pairs, raster order */
int vslam_get_row_lut(vslam_ctx* ctx, int stream, int level, int32_t* lut);             /* H_level entries */

/* ---- KeyFrame::MakeKeyFrame_Rest (jni/KeyFrame.cc:53-95: fast_nonmax + Shi-Tomasi candidates; the SmallBlurryImage tail is not
 * part of this path) for one stream's current keyframe.  Results stay readable until the next call. */
int vslam_make_keyframe_rest(vslam_ctx* ctx, int stream);
int vslam_get_max_corners(vslam_ctx* ctx, int stream, int level, int32_t* xy /* may be NULL: count only */, int cap, int* n);   /* Level::vMaxCorners */
int vslam_get_candidates(vslam_ctx* ctx, int stream, int level, int32_t* xy, double* st_score, int cap, int* n);                /* Level::vCandidates */

/* ---- MiniPatch (jni/MiniPatch.cc): trail tracking while the initial map is built (Tracker::TrailTracking_*, jni/Tracker.cc:264-346).
 * `which` selects the searched image: 0 = the stream's current keyframe, 1 = its snapshot (the "previous frame" of the married-match
 * check), taken with vslam_snapshot_keyframe.  Patches are 9x9 bytes each; positions are level-0 pixels (integer valued). */
int vslam_snapshot_keyframe(vslam_ctx* ctx, int stream);
int vslam_minipatch_sample(vslam_ctx* ctx, int stream, int which, int n, const int32_t* xy, uint8_t* patches81);            /* SampleFromImage */
int vslam_minipatch_find(vslam_ctx* ctx, int stream, int which, int n, const uint8_t* patches81, double* pos2 /* in/out */,
                         int32_t* found, int32_t* best_ssd /* may be NULL */, int range, int max_ssd);                      /* FindPatch */

/* ---- per-stream tracker state (Tracker members, jni/Tracker.h:105-133) ----------------------------------------- */
int vslam_set_pose(vslam_ctx* ctx, int stream, const double* pose12);   /* row-major 3x4 [R|t], camera-from-world */
int vslam_get_pose(vslam_ctx* ctx, int stream, double* pose12);
int vslam_get_poses(vslam_ctx* ctx, double* pose12_per_stream);         /* all streams, one D2H copy */
int vslam_set_motion(vslam_ctx* ctx, int stream, const double* velocity6, double msd_scaled_velocity, double scene_depth_mean,
                     double scene_depth_sigma);
int vslam_get_motion(vslam_ctx* ctx, int stream, double* velocity6, double* msd_scaled_velocity, double* scene_depth_mean,
                     double* scene_depth_sigma);
/* Tracker::Reset (jni/Tracker.cc:45-60), the tracker's own state of one stream: quality GOOD, lost-frame and coarse flags cleared,
 * velocity and scaled speed zero, scene depth 1 +- 1, counters zero.  The pose is left as it is (the reference's Reset does not
 * touch mse3CamFromWorld); the map is replaced with vslam_set_map (MapMaker's part of the reset is out of scope). */
int vslam_reset_stream(vslam_ctx* ctx, int stream);
int vslam_set_sbi_rotation(vslam_ctx* ctx, int stream, const double* rot6);  /* Tracker::mv6SBIRot supplied by the caller (when the on-device SBI is off) */
/* SmallBlurryImage on the device (jni/SmallBlurryImage.cc; Tracker::CalcSBIRotation jni/Tracker.cc:885-893): after this call every
 * vslam_track_frame* builds the 1/16-size blurred thumbnail of each stream, aligns it to the previous frame's (6 ESM iterations) and
 * writes the resulting rotation into mv6SBIRot before the motion model runs.  cam13_sbi = vslam_camera_from_params at (width/16, height/16). */
int vslam_enable_sbi(vslam_ctx* ctx, const double* cam13_sbi);
int vslam_get_sbi_rotation(vslam_ctx* ctx, int stream, double* rot6);
/* Relocaliser (jni/Relocaliser.cc:17-58 + Tracker::AttemptRecovery, jni/Tracker.cc:167-180): register the map's keyframes (ids of
 * uploaded source keyframes + their poses; needs vslam_enable_sbi).  From then on vslam_track_frame* runs the lost branch of
 * Tracker::TrackFrame (jni/Tracker.cc:134-140) on the device for every stream with lost_frames >= 3: SmallBlurryImage of the frame
 * (blur 2.5), SSD against every keyframe's, ESM alignment to the best, SE3fromSE2 * keyframe pose; if the score < 9e6 the stream
 * continues from that pose (velocity zero, doubled coarse stage) with TrackMap + AssessTrackingQuality in the same frame.  Without
 * registered keyframes a lost stream waits (vslam_set_pose / vslam_reset_stream / vslam_set_lost). */
int vslam_set_reloc_keyframes(vslam_ctx* ctx, int n, const int32_t* src_kf_ids, const double* poses12);
int vslam_get_reloc_info(vslam_ctx* ctx, int stream, int* best_keyframe, double* score, int* n_recoveries, int* recovered_last_frame);
int vslam_set_lost(vslam_ctx* ctx, int stream, int lost_frames, int quality);

/* ---- Keyframe hand-off: the two MapMaker heuristics Tracker::TrackFrame consults, evaluated on the device over the registered
 * keyframes' poses (vslam_set_reloc_keyframes / vslam_add_keyframe_from_stream), so that a host MapMaker only hears about streams
 * that need a keyframe.  With the policy enabled every vslam_track_frame* also (1) turns quality DODGY into BAD when the camera is
 * further than 10 x wiggle_scale from the nearest keyframe (MapMaker::IsDistanceToNearestKeyFrameExcessive, jni/MapMaker.cc:1098-1101,
 * called at jni/Tracker.cc:869-871) and (2) raises a keyframe request when quality is GOOD, the distance to the nearest keyframe
 * divided by the scene depth exceeds max_kf_dist_wiggle_mult x wiggle_scale_depth_normalized (MapMaker::NeedNewKeyFrame,
 * jni/MapMaker.cc:763-773; the reference's values: 0.2, mdWiggleScaleDepthNormalized) and more than min_frames_between (20) frames
 * passed since the stream's last keyframe (jni/Tracker.cc:127-129).  The remaining term, QueueSize() < 3, is the caller's.
 * vslam_add_keyframe_from_stream = Tracker::AddNewKeyFrame + MapMaker::AddKeyFrame (jni/Tracker.cc:823-827, jni/MapMaker.cc:470-478):
 * the stream's current keyframe is copied device-to-device into source keyframe slot kf_id at the stream's pose and registered. */
int vslam_set_keyframe_policy(vslam_ctx* ctx, int enable, double wiggle_scale, double wiggle_scale_depth_normalized, double max_kf_dist_wiggle_mult, int min_frames_between);
int vslam_get_keyframe_requests(vslam_ctx* ctx, int32_t* request /* [n_streams] */, int32_t* closest_keyframe /* may be NULL */, double* distance /* may be NULL */);
int vslam_add_keyframe_from_stream(vslam_ctx* ctx, int stream, int kf_id);

/* ---- Map files: camera + source keyframes (level-0 images) + map points + relocaliser registration in one checksummed binary file
 * (layout in csrc/mapfile.cu), so a service can restart or several processes / GPUs can serve streams of one map.
 * vslam_save_map_file writes what this context currently holds.  vslam_load_map_file verifies the whole file first (size, version,
 * checksum, capacity: the context is untouched on any error), then uploads the keyframes (pyramids are rebuilt on the device) and the
 * points like vslam_upload_source_keyframe / vslam_set_map would; VSLAM_MAP_LOAD_CAMERA also installs the file's camera,
 * VSLAM_MAP_LOAD_RELOC also re-registers the relocaliser keyframes (needs vslam_enable_sbi).  vslam_map_file_info reads only the
 * header and needs no GPU.  vslam_export_map_text writes the reference's debug dump layout (jni/MapMaker.cc:1254-1297): <dir>/map.dump
 * with world position + source level per point and <dir>/keyframes/<i>.info with the pose of the i-th relocaliser keyframe. */
#define VSLAM_MAP_LOAD_CAMERA 1
#define VSLAM_MAP_LOAD_RELOC 2
typedef struct vslam_map_file_info_t { int width, height, n_points, n_keyframes, n_reloc_keyframes; double cam13[13]; } vslam_map_file_info_t;
int vslam_map_file_info(const char* path, vslam_map_file_info_t* out);
int vslam_save_map_file(vslam_ctx* ctx, const char* path);
int vslam_load_map_file(vslam_ctx* ctx, const char* path, int flags);
int vslam_export_map_text(vslam_ctx* ctx, const char* dir);
/* attempted[4], found[4], quality (0 BAD,1 DODGY,2 GOOD), lost_frames, did_coarse */
int vslam_get_counters(vslam_ctx* ctx, int stream, int32_t* attempted4, int32_t* found4, int* quality, int* lost_frames, int* did_coarse);
/* Per-point TrackerData dump, layout of oracle/ref_harness.cc ref_tracker_point_state: ints[n][8], dbl[n][32]. */
int vslam_get_point_states(vslam_ctx* ctx, int stream, int32_t* ints, double* dbl);
int vslam_get_point_template(vslam_ctx* ctx, int stream, int point, uint8_t* tmpl /* P*P */, int* sum, int* sumsq);
int vslam_get_point_counts(vslam_ctx* ctx, int stream, int32_t* outlier_inlier /* n*2 */);
/* 6-vectors and sigma^2 of every CalcPoseUpdate of the last track_map (<= 20), in order. */
int vslam_get_updates(vslam_ctx* ctx, int stream, double* upd6, double* sigma_sq, int cap, int* n);
int vslam_get_zmssd_evals(vslam_ctx* ctx, unsigned long long* total);   /* candidates scored since create (all streams) */
/* search counters since create (all streams): [0] ZMSSD candidates scored, [1] reserved (0), [2] templates generated
 * (MakeTemplateCoarseCont calls that did not reuse the cached template), [3] sub-pixel refinements run */
int vslam_get_search_stats(vslam_ctx* ctx, unsigned long long* stats4);

/* ---- stages, all streams at once ------------------------------------------------------------------------------- */
int vslam_project_all(vslam_ctx* ctx);
/* Stage isolation for parity tests: overwrite what vslam_project_all computed for `stream` with caller-supplied
 * projected pixels (n*2), warp matrices mm2WarpInverse (n*4, row-major) and search levels (n), n = map size. */
int vslam_set_point_projection(vslam_ctx* ctx, int stream, const double* v2image, const double* warp_inverse, const int32_t* level);
/* idx: per-stream lists, stream s uses idx[s*idx_stride .. + n[s]); the same list is used by vslam_calc_pose_update. */
int vslam_set_lists(vslam_ctx* ctx, const int32_t* idx, const int32_t* n, int idx_stride);
int vslam_clear_counters(vslam_ctx* ctx);
int vslam_search_for_points(vslam_ctx* ctx, int range, int subpix_its);
/* ---- PatchFinder, one object at a time: the per-object methods of jni/PatchFinder.h:45-121 that the batched calls fuse away.  A "PatchFinder
 * object" is the finder state kept per (stream, map point): template, sums, search level (vslam_get_point_states / _point_template).  Slow
 * path: one small launch and a synchronous read-back per call; include/vslam_b200_shell.hpp's class PatchFinder is written on these. */
/* MakeTemplateCoarseCont(p) (jni/PatchFinder.cc:79-125) with the warp of the last projection, re-use rule included */
int vslam_pf_make_template(vslam_ctx* ctx, int stream, int point, int* template_bad);
/* MakeTemplateCoarseNoWarp(KeyFrame&, nLevel, x, y) (:130-143): P x P pixels of level `level` of source keyframe `src_kf` around (x, y); sets
 * the point's search level to `level`.  src_kf < 0: MakeTemplateCoarseNoWarp(MapPoint&) (:146-149), the point's own source keyframe, level and
 * irCenter (level, x, y are ignored). */
int vslam_pf_make_template_nowarp(vslam_ctx* ctx, int stream, int point, int src_kf, int level, int x, int y, int* template_bad);
/* ZMSSDAtPoint(img, icol, irow) (:352-380) at n positions (x, y) of level `level` of the stream's current keyframe; mnMaxSSD + 1 outside the border */
int vslam_pf_zmssd_at(vslam_ctx* ctx, int stream, int point, int level, int n, const int32_t* xy, int32_t* ssd);
/* MakeSubPixTemplate's JtJ^-1 (:242-267) and up to max_its x IterateSubPix (:290-350) on the point's search level, from pos2 = mv2SubPixPos
 * (level-zero pixels, in/out) and *mean_diff = mdMeanDiff (in/out).  max_its = 1: one IterateSubPix, *last_update_sq = its return value
 * (negative: off the image).  max_its = n: IterateSubPixToConvergence(kf, n) (:272-285), *converged = its return value. */
int vslam_pf_subpix(vslam_ctx* ctx, int stream, int point, int max_its, double* pos2, double* mean_diff, int* converged, double* last_update_sq);

/* The user event of the reference's shell: SystemPTAM::onTouchScreen -> Tracker::mbUserPressedSpacebar (jni/jni_part.cpp:49-51, 125-129),
 * consumed by Tracker::TrackForInitialMap (jni/Tracker.cc:232-253).  vslam_take_user_event returns the pending events of a stream and clears them. */
#define VSLAM_EVENT_SPACEBAR 1
int vslam_user_event(vslam_ctx* ctx, int stream, int event);
int vslam_take_user_event(vslam_ctx* ctx, int stream, int* pending);

/* MapMaker::ReFind_Common (jni/MapMaker.cc:967-1036), batched: every stream stands for one keyframe (its current keyframe image and
 * pose), its list (vslam_set_lists) for the map points to re-find in it.  Per pair: projection and in-image tests, a cold-finder
 * MakeTemplateCoarse (always regenerated; the reference's static PatchFinder does the same whenever consecutive calls name different
 * points), FindPatchCoarse with `range` (reference: 4), then on levels > 0 MakeSubPixTemplate + IterateSubPixToConvergence
 * (`subpix_its`, reference: 8) whose position is kept whether or not it converged.  The Measurement / sNeverRetryKFs bookkeeping
 * is the caller's.  vslam_get_refind_results: list order, flags3 = {found, level, bSubPix}, pos2 = Measurement::v2RootPos. */
int vslam_refind(vslam_ctx* ctx, int range, int subpix_its);
int vslam_get_refind_results(vslam_ctx* ctx, int stream, int32_t* flags3, double* pos2, int cap, int* n);
/* The search of MapMaker::AddPointEpipolar (jni/MapMaker.cc:525-640) for n candidates of level `level` of source keyframe `src_kf`
 * (vslam_upload_source_keyframe; positions cand_xy in level pixels, e.g. Level::vCandidates) in the CURRENT keyframe of `stream`
 * (the target): epipolar line from the two keyframe poses (row-major 3x4, camera-from-world) and the source keyframe's scene depth
 * (dSceneDepthMean +- dSceneDepthSigma, clamped by mdWiggleScale as the reference does), un-warped template, ZMSSD over the target
 * level's FAST corners inside the epipolar band, sub-pixel refinement (10 iterations).  found[k] = 1 iff a match was found and the
 * refinement converged; pos2[k] = the refined level-0 position (Measurement::v2RootPos in the target); best_corner / best_zmssd (may be
 * NULL): index into the target level's corner list and its score (-1 / mnMaxSSD+1 if no corner qualified).  Triangulation and the
 * new MapPoint are the caller's (MapMaker).  Synchronous. */
int vslam_epipolar_search(vslam_ctx* ctx, int stream, int src_kf, int level, int n, const int32_t* cand_xy, const double* src_pose12,
                          const double* target_pose12, double depth_mean, double depth_sigma, double wiggle_scale, int32_t* found, double* pos2,
                          int32_t* best_corner, int32_t* best_zmssd);
/* The rest of MapMaker::AddPointEpipolar for candidates whose search converged (jni/MapMaker.cc:646-690): triangulate the new point
 * from the candidate's level-zero position in the source keyframe and the refined position in the target (MapMaker::ReprojectPoint,
 * jni/MapMaker.cc:176-200, smallest right singular vector of the 4x4 two-view system), then derive the patch-source fields and
 * MapPoint::RefreshPixelVectors (jni/MapPoint.cc:4-29).  Host arithmetic (one 4x4 SVD per new point).  The outputs are the
 * vslam_set_map arrays of the new points; their source keyframe is the one `src_pose12` belongs to. */
int vslam_epipolar_make_points(vslam_ctx* ctx, int level, int n, const int32_t* cand_xy, const double* found_pos2, const double* src_pose12, const double* tgt_pose12,
                               double* world3, double* pixel_right3, double* pixel_down3, int32_t* ir_center2, int32_t* src_level);
int vslam_project_and_derivs(vslam_ctx* ctx, int only_found);
int vslam_calc_jacobians(vslam_ctx* ctx);
int vslam_calc_pose_update(vslam_ctx* ctx, double override_sigma, int mark_outliers, int apply, double* upd6_per_stream /* may be NULL */);
int vslam_track_map(vslam_ctx* ctx);
/* Tracker::TrackFrame, map-good branch (jni/Tracker.cc:76-140): MakeKeyFrame_Lite, SmallBlurryImage (vslam_enable_sbi), then per stream
 * either (lost_frames < 3) CalcSBIRotation, ApplyMotionModel, TrackMap, UpdateMotionModel, AssessTrackingQuality, or (lost, with
 * vslam_set_reloc_keyframes) AttemptRecovery and on success TrackMap + AssessTrackingQuality. */
int vslam_track_frame(vslam_ctx* ctx, const uint8_t* gray_host, int stride, size_t frame_stride);
int vslam_track_frame_dev(vslam_ctx* ctx, const uint8_t* gray_dev, int stride, size_t frame_stride);
/* Pipelined form of vslam_track_frame for a continuous feed: the host->device copy of this step runs on an internal copy
 * stream into one of two level-0 buffers and overlaps the kernels of the previous step; if poses_out is not NULL every
 * stream's pose (12 doubles each) is copied there when the step completes.  gray_host and poses_out (pinned memory) must
 * stay valid until vslam_wait_step(id) returns.  Returns a step id >= 0 (or a negative error); at most two steps are in
 * flight, a third call blocks until the oldest has finished. */
int vslam_track_frame_async(vslam_ctx* ctx, const uint8_t* gray_host, int stride, size_t frame_stride, double* poses_out);
int vslam_wait_step(vslam_ctx* ctx, int step_id);

/* Per-kernel device times measured with CUDA events on the context's stream.  Stages: 0-3 pyramid+FAST level 0-3,
 * 4 project+lists, 5 coarse search, 6 coarse pose iterations, 7 fine search, 8 fine pose iterations, 9 host->device copy,
 * 10 other.  vslam_get_stage_times synchronises, returns the milliseconds and launch counts accumulated since the last call
 * and resets them. */
#define VSLAM_N_STAGES 11
int vslam_set_timing(vslam_ctx* ctx, int on);
int vslam_get_stage_times(vslam_ctx* ctx, double* ms /* [VSLAM_N_STAGES] */, int* launches /* [VSLAM_N_STAGES] */);

/* Test hook: y[i] = the correctly-rounded device atan used by the camera model (csrc/atan_dd.cuh). */
int vslam_debug_atan(const double* x_host, double* y_host, int n);
int vslam_debug_atan_dd(const double* x_host, double* y_host, int n);   /* the slow double-double evaluation alone (test hook for the fast path's rounding test) */

/* Integer-pipe micro-benchmark: measured dp4a throughput of the current device in tera-MACs/s (the ZMSSD roofline denominator). */
int vslam_debug_dp4a_peak(double* tmacs_per_s);

/* Number of kernels this library has launched on the context since creation (bench.py's gpu_launches). */
unsigned long long vslam_kernel_launches(const vslam_ctx* ctx);
/* 1 if the next vslam_track_frame* call would run with frame look-ahead (vslam_params.frame_lookahead resolved against the workload, per-stage
   timing and stream groups), else 0.  Execution only: results do not depend on it. */
int vslam_frame_lookahead_active(const vslam_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif

"""ctypes binding of libvslam_b200.so — the C-ABI declared in include/vslam_b200.h.

This is the only way Python reaches the CUDA path.  There is no CPU fallback: a missing library or a
missing GPU raises.  (The reference's host language is C++; this module exists for tests and bench.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VSLAM_LIB") or os.path.join(HERE, "libvslam_b200.so")   # VSLAM_LIB: A/B runs of an experimental build of the same library
LEVELS = 4

OK, E_INVALID, E_CUDA, E_CAPACITY, E_NO_DEVICE, E_IO = 0, -1, -2, -3, -4, -5
MAP_LOAD_CAMERA, MAP_LOAD_RELOC = 1, 2


class VslamError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"vslam error {code}: {text}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("width", C.c_int), ("height", C.c_int), ("n_streams", C.c_int), ("max_points", C.c_int),
                ("patch_size", C.c_int), ("max_source_keyframes", C.c_int), ("max_corner_frac", C.c_float), ("cuda_stream", C.c_void_p),
                ("truncate_error", C.c_int), ("rand_seed", C.c_uint)]


class MapFileInfo(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("n_points", C.c_int), ("n_keyframes", C.c_int), ("n_reloc_keyframes", C.c_int),
                ("cam13", C.c_double * 13)]


class Params(C.Structure):
    _fields_ = [("coarse_min", C.c_uint), ("coarse_max", C.c_uint), ("coarse_range", C.c_uint), ("coarse_subpix_its", C.c_int),
                ("coarse_min_vel", C.c_double), ("fine_range", C.c_int), ("fine_range_after_coarse", C.c_int),
                ("fine_subpix_its_top_level", C.c_int), ("max_patches_per_frame", C.c_int), ("use_sbi", C.c_int), ("stream_groups", C.c_int),
                ("serial_normal_equations", C.c_int), ("pose_kernel", C.c_int), ("search_kernel", C.c_int), ("frame_lookahead", C.c_int), ("coarse_chain", C.c_int)]


# every symbol include/vslam_b200.h declares (tests/test_abi_cpu.py checks the library exports all of them)
ABI_SYMBOLS = [
    "vslam_default_config", "vslam_default_params", "vslam_create", "vslam_destroy", "vslam_last_error", "vslam_sync", "vslam_set_params",
    "vslam_set_camera", "vslam_camera_from_params", "vslam_upload_source_keyframe", "vslam_set_map", "vslam_make_keyframe_lite",
    "vslam_make_keyframe_lite_dev", "vslam_level_dims", "vslam_get_level", "vslam_get_num_corners", "vslam_get_corners", "vslam_get_row_lut",
    "vslam_make_keyframe_rest", "vslam_get_max_corners", "vslam_get_candidates", "vslam_snapshot_keyframe", "vslam_minipatch_sample", "vslam_minipatch_find",
    "vslam_set_pose", "vslam_get_pose", "vslam_get_poses", "vslam_set_motion", "vslam_get_motion", "vslam_reset_stream", "vslam_set_sbi_rotation", "vslam_enable_sbi", "vslam_get_sbi_rotation", "vslam_set_reloc_keyframes", "vslam_get_reloc_info", "vslam_set_lost", "vslam_get_counters",
    "vslam_get_point_states", "vslam_get_point_template", "vslam_get_point_counts", "vslam_get_updates", "vslam_get_zmssd_evals",
    "vslam_project_all", "vslam_set_point_projection", "vslam_set_lists", "vslam_clear_counters", "vslam_search_for_points", "vslam_refind", "vslam_get_refind_results", "vslam_epipolar_search", "vslam_project_and_derivs", "vslam_calc_jacobians",
    "vslam_calc_pose_update", "vslam_track_map", "vslam_track_frame", "vslam_track_frame_dev", "vslam_track_frame_async", "vslam_wait_step", "vslam_debug_atan", "vslam_debug_atan_dd", "vslam_debug_dp4a_peak", "vslam_kernel_launches", "vslam_frame_lookahead_active", "vslam_set_timing", "vslam_get_stage_times",
    "vslam_get_search_stats", "vslam_make_keyframe_from_source", "vslam_append_map_points", "vslam_set_keyframe_policy", "vslam_get_keyframe_requests", "vslam_add_keyframe_from_stream", "vslam_epipolar_make_points", "vslam_map_file_info", "vslam_save_map_file", "vslam_load_map_file", "vslam_export_map_text",
    "vslam_pf_make_template", "vslam_pf_make_template_nowarp", "vslam_pf_zmssd_at", "vslam_pf_subpix", "vslam_user_event", "vslam_take_user_event",
]

_lib = None


def map_file_info(path):
    """Header of a map file (needs no GPU): dict with width, height, n_points, n_keyframes, n_reloc_keyframes, cam13."""
    info = MapFileInfo()
    rc = load().vslam_map_file_info(os.fsencode(path), C.byref(info))
    if rc:
        raise VslamError(rc, "not a readable vslam map file: " + str(path))
    return dict(width=info.width, height=info.height, n_points=info.n_points, n_keyframes=info.n_keyframes,
                n_reloc_keyframes=info.n_reloc_keyframes, cam13=np.array(info.cam13[:], dtype=np.float64))


def load():
    """Load the CUDA library; raise if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(visualslam_android_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    pi, pd = C.POINTER(C.c_int), C.POINTER(C.c_double)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("vslam_default_config", None, C.POINTER(Config))
    sig("vslam_default_params", None, C.POINTER(Params))
    sig("vslam_create", i, C.POINTER(Config), C.POINTER(vp))
    sig("vslam_destroy", None, vp)
    sig("vslam_last_error", C.c_char_p, vp)
    sig("vslam_sync", i, vp)
    sig("vslam_set_params", i, vp, C.POINTER(Params))
    sig("vslam_set_camera", i, vp, vp)
    sig("vslam_camera_from_params", None, vp, i, i, i, vp)
    sig("vslam_upload_source_keyframe", i, vp, i, vp, i)
    sig("vslam_set_map", i, vp, i, vp, vp, vp, vp, vp, vp)
    sig("vslam_make_keyframe_lite", i, vp, i, i, vp, i, C.c_size_t)
    sig("vslam_make_keyframe_lite_dev", i, vp, i, i, vp, i, C.c_size_t)
    sig("vslam_level_dims", i, vp, i, pi, pi)
    sig("vslam_get_level", i, vp, i, i, vp, i)
    sig("vslam_get_num_corners", i, vp, i, i, pi)
    sig("vslam_get_corners", i, vp, i, i, vp, i)
    sig("vslam_get_row_lut", i, vp, i, i, vp)
    sig("vslam_make_keyframe_rest", i, vp, i)
    sig("vslam_get_max_corners", i, vp, i, i, vp, i, pi)
    sig("vslam_get_candidates", i, vp, i, i, vp, vp, i, pi)
    sig("vslam_snapshot_keyframe", i, vp, i)
    sig("vslam_minipatch_sample", i, vp, i, i, i, vp, vp)
    sig("vslam_minipatch_find", i, vp, i, i, i, vp, vp, vp, vp, i, i)
    sig("vslam_set_pose", i, vp, i, vp)
    sig("vslam_get_pose", i, vp, i, vp)
    sig("vslam_get_poses", i, vp, vp)
    sig("vslam_set_motion", i, vp, i, vp, d, d, d)
    sig("vslam_reset_stream", i, vp, i)
    sig("vslam_set_reloc_keyframes", i, vp, i, vp, vp)
    sig("vslam_get_search_stats", i, vp, vp)
    sig("vslam_make_keyframe_from_source", i, vp, i, i)
    sig("vslam_append_map_points", i, vp, i, vp, vp, vp, vp, vp, vp)
    sig("vslam_set_keyframe_policy", i, vp, i, d, d, d, i)
    sig("vslam_get_keyframe_requests", i, vp, vp, vp, vp)
    sig("vslam_add_keyframe_from_stream", i, vp, i, i)
    sig("vslam_epipolar_make_points", i, vp, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp)
    sig("vslam_map_file_info", i, C.c_char_p, C.POINTER(MapFileInfo))
    sig("vslam_save_map_file", i, vp, C.c_char_p)
    sig("vslam_load_map_file", i, vp, C.c_char_p, i)
    sig("vslam_export_map_text", i, vp, C.c_char_p)
    sig("vslam_get_reloc_info", i, vp, i, pi, pd, pi, pi)
    sig("vslam_set_lost", i, vp, i, i, i)
    sig("vslam_get_motion", i, vp, i, vp, pd, pd, pd)
    sig("vslam_set_sbi_rotation", i, vp, i, vp)
    sig("vslam_enable_sbi", i, vp, vp)
    sig("vslam_get_sbi_rotation", i, vp, i, vp)
    sig("vslam_get_counters", i, vp, i, vp, vp, pi, pi, pi)
    sig("vslam_get_point_states", i, vp, i, vp, vp)
    sig("vslam_get_point_template", i, vp, i, i, vp, pi, pi)
    sig("vslam_get_point_counts", i, vp, i, vp)
    sig("vslam_get_updates", i, vp, i, vp, vp, i, pi)
    sig("vslam_get_zmssd_evals", i, vp, C.POINTER(C.c_ulonglong))
    sig("vslam_project_all", i, vp)
    sig("vslam_set_point_projection", i, vp, i, vp, vp, vp)
    sig("vslam_set_lists", i, vp, vp, vp, i)
    sig("vslam_clear_counters", i, vp)
    sig("vslam_search_for_points", i, vp, i, i)
    sig("vslam_refind", i, vp, i, i)
    sig("vslam_pf_make_template", i, vp, i, i, pi)
    sig("vslam_pf_make_template_nowarp", i, vp, i, i, i, i, i, i, pi)
    sig("vslam_pf_zmssd_at", i, vp, i, i, i, i, vp, vp)
    sig("vslam_pf_subpix", i, vp, i, i, i, vp, pd, pi, pd)
    sig("vslam_user_event", i, vp, i, i)
    sig("vslam_take_user_event", i, vp, i, pi)
    sig("vslam_epipolar_search", i, vp, i, i, i, i, vp, vp, vp, d, d, d, vp, vp, vp, vp)
    sig("vslam_get_refind_results", i, vp, i, vp, vp, i, pi)
    sig("vslam_project_and_derivs", i, vp, i)
    sig("vslam_calc_jacobians", i, vp)
    sig("vslam_calc_pose_update", i, vp, d, i, i, vp)
    sig("vslam_track_map", i, vp)
    sig("vslam_track_frame", i, vp, vp, i, C.c_size_t)
    sig("vslam_track_frame_dev", i, vp, vp, i, C.c_size_t)
    sig("vslam_kernel_launches", C.c_ulonglong, vp)
    sig("vslam_frame_lookahead_active", i, vp)
    sig("vslam_track_frame_async", i, vp, vp, i, C.c_size_t, vp)
    sig("vslam_wait_step", i, vp, i)
    sig("vslam_debug_atan", i, vp, vp, i)
    sig("vslam_debug_atan_dd", i, vp, vp, i)
    sig("vslam_debug_dp4a_peak", i, pd)
    sig("vslam_set_timing", i, vp, i)
    sig("vslam_get_stage_times", i, vp, vp, vp)
    _lib = L
    return L


def camera_from_params(params5, width, height, as_shipped_radius=False) -> np.ndarray:
    L = load()
    p = np.ascontiguousarray(params5, dtype=np.float64)
    out = np.zeros(13)
    L.vslam_camera_from_params(p.ctypes.data, int(width), int(height), int(as_shipped_radius), out.ctypes.data)
    return out


def debug_atan(x, dd_only=False) -> np.ndarray:
    """The device atan of the camera model; dd_only: the double-double evaluation alone (without the fast path and its rounding test)."""
    L = load()
    a = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(a)
    rc = (L.vslam_debug_atan_dd if dd_only else L.vslam_debug_atan)(a.ctypes.data, out.ctypes.data, a.size)
    if rc != OK:
        raise VslamError(rc, 'vslam_debug_atan')
    return out


def dp4a_peak_tmacs() -> float:
    """Measured dp4a throughput (tera-MACs/s) of the current CUDA device."""
    L = load()
    v = C.c_double()
    rc = L.vslam_debug_dp4a_peak(C.byref(v))
    if rc != OK:
        raise VslamError(rc, 'vslam_debug_dp4a_peak')
    return v.value


def _ptr(a):
    return a.ctypes.data if isinstance(a, np.ndarray) else int(a)


class Context:
    """One GPU context: `n_streams` cameras of one size tracked against one map (see include/vslam_b200.h)."""

    def __init__(self, width, height, n_streams=1, max_points=1000, patch_size=11, device=0, cuda_stream=None, truncate_error=True,
                 max_corner_frac=0.5, rand_seed=1, max_source_keyframes=1):
        self.L = load()
        cfg = Config()
        self.L.vslam_default_config(C.byref(cfg))
        cfg.device, cfg.width, cfg.height, cfg.n_streams, cfg.max_points = device, width, height, n_streams, max_points
        cfg.patch_size, cfg.max_source_keyframes, cfg.max_corner_frac = patch_size, max_source_keyframes, max_corner_frac
        cfg.cuda_stream = cuda_stream
        cfg.truncate_error, cfg.rand_seed = int(truncate_error), rand_seed
        h = C.c_void_p()
        rc = self.L.vslam_create(C.byref(cfg), C.byref(h))
        if rc != OK:
            raise VslamError(rc, self.L.vslam_last_error(None).decode())
        self.h = h
        self.width, self.height, self.S, self.N, self.P = width, height, n_streams, max_points, patch_size
        self.n_points = 0
        self._keep = []
        # debugging aid: VSLAM_PARAMS="pose_kernel=1,serial_normal_equations=1" overrides vslam_params of every context of the process (A/B runs of the tests)
        env = os.environ.get("VSLAM_PARAMS")
        if env:
            self.set_params()

    def close(self):
        if self.h:
            self.L.vslam_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != OK:
            raise VslamError(rc, self.L.vslam_last_error(self.h).decode())

    # -- setup
    def set_params(self, **kw):
        p = Params()
        self.L.vslam_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        env = os.environ.get("VSLAM_PARAMS")       # debugging aid, see __init__
        if env:
            for k, v in (kv.split("=") for kv in env.split(",") if kv):
                setattr(p, k, int(v))
        self._ck(self.L.vslam_set_params(self.h, C.byref(p)))

    def set_camera(self, cam13):
        a = np.ascontiguousarray(cam13, dtype=np.float64)
        assert a.size == 13
        self._ck(self.L.vslam_set_camera(self.h, a.ctypes.data))

    def upload_source_keyframe(self, gray, kf_id=0):
        g = np.ascontiguousarray(gray, dtype=np.uint8)
        self._ck(self.L.vslam_upload_source_keyframe(self.h, kf_id, g.ctypes.data, g.shape[1]))

    def set_map(self, world, right, down, ir_center, src_level, src_kf=None):
        w = np.ascontiguousarray(world, dtype=np.float64)
        r = np.ascontiguousarray(right, dtype=np.float64)
        d = np.ascontiguousarray(down, dtype=np.float64)
        c = np.ascontiguousarray(ir_center, dtype=np.int32)
        lv = np.ascontiguousarray(src_level, dtype=np.int32)
        kf = None if src_kf is None else np.ascontiguousarray(src_kf, dtype=np.int32)
        n = w.shape[0]
        self._ck(self.L.vslam_set_map(self.h, n, w.ctypes.data, r.ctypes.data, d.ctypes.data, c.ctypes.data, lv.ctypes.data,
                                      None if kf is None else kf.ctypes.data))
        self.n_points = n

    # -- MakeKeyFrame_Lite
    def append_map_points(self, world, right, down, ir_center, src_level, src_kf=None):
        """New map points behind the existing ones; existing points keep their per-stream tracker state."""
        w = np.ascontiguousarray(world, dtype=np.float64); r = np.ascontiguousarray(right, dtype=np.float64); d = np.ascontiguousarray(down, dtype=np.float64)
        c = np.ascontiguousarray(ir_center, dtype=np.int32); lv = np.ascontiguousarray(src_level, dtype=np.int32)
        kf = None if src_kf is None else np.ascontiguousarray(src_kf, dtype=np.int32)
        self._ck(self.L.vslam_append_map_points(self.h, w.shape[0], w.ctypes.data, r.ctypes.data, d.ctypes.data, c.ctypes.data, lv.ctypes.data,
                                                None if kf is None else kf.ctypes.data))
        self.n_points += w.shape[0]

    def make_keyframe_lite(self, frames, first_stream=0):
        """frames: host uint8 array (count, H, W) or (H, W)."""
        f = np.ascontiguousarray(frames, dtype=np.uint8)
        if f.ndim == 2:
            f = f[None]
        self._ck(self.L.vslam_make_keyframe_lite(self.h, first_stream, f.shape[0], f.ctypes.data, f.shape[2], f.shape[1] * f.shape[2]))

    def make_keyframe_from_source(self, s, kf_id):
        """Source keyframe kf_id becomes stream s's current keyframe (zero copy + pyramid + FAST)."""
        self._ck(self.L.vslam_make_keyframe_from_source(self.h, s, kf_id))

    def make_keyframe_lite_ptr(self, ptr, count, stride, frame_stride, first_stream=0, device=False):
        fn = self.L.vslam_make_keyframe_lite_dev if device else self.L.vslam_make_keyframe_lite
        self._ck(fn(self.h, first_stream, count, int(ptr), stride, frame_stride))

    def level_dims(self, l):
        w, h = C.c_int(), C.c_int()
        self._ck(self.L.vslam_level_dims(self.h, l, C.byref(w), C.byref(h)))
        return w.value, h.value

    def level(self, s, l):
        w, h = self.level_dims(l)
        out = np.empty((h, w), dtype=np.uint8)
        self._ck(self.L.vslam_get_level(self.h, s, l, out.ctypes.data, w))
        return out

    def corners(self, s, l):
        n = C.c_int()
        self._ck(self.L.vslam_get_num_corners(self.h, s, l, C.byref(n)))
        out = np.empty((n.value, 2), dtype=np.int32)
        self._ck(self.L.vslam_get_corners(self.h, s, l, out.ctypes.data, n.value))
        return out

    def row_lut(self, s, l):
        _, h = self.level_dims(l)
        out = np.empty(h, dtype=np.int32)
        self._ck(self.L.vslam_get_row_lut(self.h, s, l, out.ctypes.data))
        return out

    # -- MakeKeyFrame_Rest / MiniPatch
    def make_keyframe_rest(self, s):
        self._ck(self.L.vslam_make_keyframe_rest(self.h, s))

    def max_corners(self, s, l):
        n = C.c_int()
        self._ck(self.L.vslam_get_max_corners(self.h, s, l, None, 0, C.byref(n)))
        out = np.empty((n.value, 2), dtype=np.int32)
        self._ck(self.L.vslam_get_max_corners(self.h, s, l, out.ctypes.data, n.value, C.byref(n)))
        return out

    def candidates(self, s, l):
        n = C.c_int()
        self._ck(self.L.vslam_get_candidates(self.h, s, l, None, None, 0, C.byref(n)))
        xy = np.empty((n.value, 2), dtype=np.int32)
        sc = np.empty(n.value, dtype=np.float64)
        self._ck(self.L.vslam_get_candidates(self.h, s, l, xy.ctypes.data, sc.ctypes.data, n.value, C.byref(n)))
        return xy, sc

    def snapshot_keyframe(self, s):
        self._ck(self.L.vslam_snapshot_keyframe(self.h, s))

    def minipatch_sample(self, s, xy, which=0):
        xy = np.ascontiguousarray(xy, dtype=np.int32)
        out = np.empty((xy.shape[0], 9, 9), dtype=np.uint8)
        self._ck(self.L.vslam_minipatch_sample(self.h, s, which, xy.shape[0], xy.ctypes.data, out.ctypes.data))
        return out

    def minipatch_find(self, s, patches, pos, rng=10, max_ssd=100000, which=0):
        p = np.ascontiguousarray(patches, dtype=np.uint8)
        pos = np.ascontiguousarray(pos, dtype=np.float64).copy()
        n = pos.shape[0]
        found = np.zeros(n, dtype=np.int32)
        best = np.zeros(n, dtype=np.int32)
        self._ck(self.L.vslam_minipatch_find(self.h, s, which, n, p.ctypes.data, pos.ctypes.data, found.ctypes.data, best.ctypes.data, rng, max_ssd))
        return pos, found, best

    # -- tracker state
    def set_pose(self, s, pose):
        p = np.ascontiguousarray(pose, dtype=np.float64).reshape(12)
        self._ck(self.L.vslam_set_pose(self.h, s, p.ctypes.data))

    def get_pose(self, s):
        out = np.empty(12)
        self._ck(self.L.vslam_get_pose(self.h, s, out.ctypes.data))
        return out.reshape(3, 4)

    def get_poses(self):
        out = np.empty((self.S, 12))
        self._ck(self.L.vslam_get_poses(self.h, out.ctypes.data))
        return out.reshape(self.S, 3, 4)

    def set_motion(self, s, velocity6, msd, depth_mean=1.0, depth_sigma=1.0):
        v = np.ascontiguousarray(velocity6, dtype=np.float64)
        self._ck(self.L.vslam_set_motion(self.h, s, v.ctypes.data, msd, depth_mean, depth_sigma))

    def set_reloc_keyframes(self, src_kf_ids, poses):
        """Relocaliser keyframes: ids of uploaded source keyframes and their poses (n x 3 x 4)."""
        ids = np.ascontiguousarray(src_kf_ids, dtype=np.int32); p = np.ascontiguousarray(poses, dtype=np.float64).reshape(len(ids), 12)
        self._ck(self.L.vslam_set_reloc_keyframes(self.h, len(ids), ids.ctypes.data, p.ctypes.data))

    def set_keyframe_policy(self, enable=True, wiggle_scale=0.1, wiggle_scale_depth_normalized=0.1, max_kf_dist_wiggle_mult=0.2, min_frames_between=20):
        """MapMaker::NeedNewKeyFrame / IsDistanceToNearestKeyFrameExcessive as Tracker::TrackFrame consults them, on the device."""
        self._ck(self.L.vslam_set_keyframe_policy(self.h, int(enable), wiggle_scale, wiggle_scale_depth_normalized, max_kf_dist_wiggle_mult, min_frames_between))

    def keyframe_requests(self, flags_only=False):
        """(request flag, index of the closest registered keyframe, distance to it) per stream after the last track_frame;
        flags_only: just the flags (the cheap per-frame poll: one copy of n_streams ints)."""
        r = np.zeros(self.S, dtype=np.int32); c = np.zeros(self.S, dtype=np.int32); dist = np.zeros(self.S)
        if flags_only:
            self._ck(self.L.vslam_get_keyframe_requests(self.h, r.ctypes.data, None, None))
            return r
        self._ck(self.L.vslam_get_keyframe_requests(self.h, r.ctypes.data, c.ctypes.data, dist.ctypes.data))
        return r, c, dist

    def add_keyframe_from_stream(self, s, kf_id):
        """Tracker::AddNewKeyFrame: the stream's current keyframe becomes source keyframe kf_id at the stream's pose (device copy)."""
        self._ck(self.L.vslam_add_keyframe_from_stream(self.h, s, kf_id))

    def save_map_file(self, path):
        """Camera + source keyframes + map points + relocaliser registration of this context -> one checksummed file."""
        self._ck(self.L.vslam_save_map_file(self.h, os.fsencode(path)))

    def load_map_file(self, path, flags=0):
        """Load a map file (verified before anything is touched); flags: MAP_LOAD_CAMERA | MAP_LOAD_RELOC."""
        self._ck(self.L.vslam_load_map_file(self.h, os.fsencode(path), flags))
        self.n_points = int(map_file_info(path)["n_points"])     # point_states() / point_counts() size their buffers from it

    def export_map_text(self, directory):
        """The reference's SaveMap debug dump layout (jni/MapMaker.cc:1254-1297): map.dump + keyframes/<i>.info."""
        self._ck(self.L.vslam_export_map_text(self.h, os.fsencode(directory)))

    def reloc_info(self, s):
        b, sc, n, r = C.c_int(), C.c_double(), C.c_int(), C.c_int()
        self._ck(self.L.vslam_get_reloc_info(self.h, s, C.byref(b), C.byref(sc), C.byref(n), C.byref(r)))
        return b.value, sc.value, n.value, r.value

    def set_lost(self, s, lost_frames, quality=0):
        self._ck(self.L.vslam_set_lost(self.h, s, lost_frames, quality))

    def reset_stream(self, s):
        """Tracker::Reset (jni/Tracker.cc:45-60) for the tracker state of stream s."""
        self._ck(self.L.vslam_reset_stream(self.h, s))

    def get_motion(self, s):
        v = np.empty(6)
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.L.vslam_get_motion(self.h, s, v.ctypes.data, C.byref(a), C.byref(b), C.byref(c)))
        return v, a.value, b.value, c.value

    def set_sbi_rotation(self, s, rot6):
        v = np.ascontiguousarray(rot6, dtype=np.float64)
        self._ck(self.L.vslam_set_sbi_rotation(self.h, s, v.ctypes.data))

    def enable_sbi(self, cam13_sbi):
        a = np.ascontiguousarray(cam13_sbi, dtype=np.float64)
        self._ck(self.L.vslam_enable_sbi(self.h, a.ctypes.data))

    def get_sbi_rotation(self, s):
        """mv6SBIRot of stream s (read from the stream state)."""
        out = np.zeros(6)
        self._ck(self.L.vslam_get_sbi_rotation(self.h, s, out.ctypes.data))
        return out

    def counters(self, s):
        a = np.zeros(4, dtype=np.int32)
        f = np.zeros(4, dtype=np.int32)
        q, lost, dc = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.vslam_get_counters(self.h, s, a.ctypes.data, f.ctypes.data, C.byref(q), C.byref(lost), C.byref(dc)))
        return a, f, q.value, lost.value, dc.value

    def point_states(self, s):
        ints = np.zeros((self.n_points, 8), dtype=np.int32)
        dbl = np.zeros((self.n_points, 32), dtype=np.float64)
        self._ck(self.L.vslam_get_point_states(self.h, s, ints.ctypes.data, dbl.ctypes.data))
        return ints, dbl

    def point_template(self, s, i):
        t = np.zeros(self.P * self.P, dtype=np.uint8)
        a, b = C.c_int(), C.c_int()
        self._ck(self.L.vslam_get_point_template(self.h, s, i, t.ctypes.data, C.byref(a), C.byref(b)))
        return t.reshape(self.P, self.P), a.value, b.value

    def point_counts(self, s):
        out = np.zeros((self.n_points, 2), dtype=np.int32)
        self._ck(self.L.vslam_get_point_counts(self.h, s, out.ctypes.data))
        return out

    def updates(self, s):
        u = np.zeros((20, 6))
        sg = np.zeros(20)
        n = C.c_int()
        self._ck(self.L.vslam_get_updates(self.h, s, u.ctypes.data, sg.ctypes.data, 20, C.byref(n)))
        return u[:n.value], sg[:n.value]

    def search_stats(self):
        """Counters since create, all streams: ZMSSD candidates scored, templates generated, sub-pixel refinements."""
        v = (C.c_ulonglong * 4)()
        self._ck(self.L.vslam_get_search_stats(self.h, v))
        return dict(zmssd_candidates=int(v[0]), templates_generated=int(v[2]), subpix_refinements=int(v[3]))

    def zmssd_evals(self):
        v = C.c_ulonglong()
        self._ck(self.L.vslam_get_zmssd_evals(self.h, C.byref(v)))
        return v.value

    # -- stages
    def project_all(self):
        self._ck(self.L.vslam_project_all(self.h))

    def set_point_projection(self, s, v2image, warp_inverse, level):
        a = np.ascontiguousarray(v2image, dtype=np.float64)
        b = np.ascontiguousarray(warp_inverse, dtype=np.float64)
        c = np.ascontiguousarray(level, dtype=np.int32)
        assert a.shape == (self.n_points, 2) and b.shape == (self.n_points, 4) and c.shape == (self.n_points,)
        self._ck(self.L.vslam_set_point_projection(self.h, s, a.ctypes.data, b.ctypes.data, c.ctypes.data))

    def set_lists(self, lists):
        """lists: one index sequence per stream."""
        n = np.array([len(x) for x in lists], dtype=np.int32)
        stride = max(1, int(n.max()))
        idx = np.zeros((self.S, stride), dtype=np.int32)
        for s, x in enumerate(lists):
            idx[s, :len(x)] = x
        self._ck(self.L.vslam_set_lists(self.h, idx.ctypes.data, n.ctypes.data, stride))

    def clear_counters(self):
        self._ck(self.L.vslam_clear_counters(self.h))

    def search_for_points(self, rng, subpix_its):
        self._ck(self.L.vslam_search_for_points(self.h, rng, subpix_its))

    # ---- PatchFinder, one object at a time (jni/PatchFinder.h:45-121; the slow path)
    def pf_make_template(self, s, point):
        """MakeTemplateCoarseCont of one point with the warp of the last projection; returns mbTemplateBad."""
        bad = C.c_int()
        self._ck(self.L.vslam_pf_make_template(self.h, s, point, C.byref(bad)))
        return bool(bad.value)

    def pf_make_template_nowarp(self, s, point, src_kf, level, x, y):
        """MakeTemplateCoarseNoWarp(KeyFrame&, nLevel, x, y); returns mbTemplateBad."""
        bad = C.c_int()
        self._ck(self.L.vslam_pf_make_template_nowarp(self.h, s, point, src_kf, level, int(x), int(y), C.byref(bad)))
        return bool(bad.value)

    def pf_zmssd_at(self, s, point, level, xy):
        """ZMSSDAtPoint of the point's template at the (x, y) rows of `xy` on `level` of the stream's current keyframe."""
        xy = np.ascontiguousarray(xy, dtype=np.int32).reshape(-1, 2)
        out = np.empty(len(xy), dtype=np.int32)
        self._ck(self.L.vslam_pf_zmssd_at(self.h, s, point, level, len(xy), xy.ctypes.data, out.ctypes.data))
        return out

    def pf_subpix(self, s, point, max_its, pos, mean_diff=0.0):
        """MakeSubPixTemplate's inverse + up to max_its IterateSubPix from `pos` (level-zero pixels); returns (pos, mean_diff, converged, last_update_sq)."""
        p = np.array(pos, dtype=np.float64)
        md, conv, last = C.c_double(mean_diff), C.c_int(), C.c_double()
        self._ck(self.L.vslam_pf_subpix(self.h, s, point, max_its, p.ctypes.data, C.byref(md), C.byref(conv), C.byref(last)))
        return p, md.value, bool(conv.value), last.value

    def user_event(self, s, event=1):
        self._ck(self.L.vslam_user_event(self.h, s, event))

    def take_user_event(self, s):
        v = C.c_int()
        self._ck(self.L.vslam_take_user_event(self.h, s, C.byref(v)))
        return v.value

    def refind(self, rng=4, subpix_its=8):
        """MapMaker::ReFind_Common for every (stream, listed point) pair (set_lists first)."""
        self._ck(self.L.vslam_refind(self.h, rng, subpix_its))

    def refind_results(self, s, cap):
        fl = np.zeros((max(cap, 1), 3), dtype=np.int32); pos = np.zeros((max(cap, 1), 2)); n = C.c_int()
        self._ck(self.L.vslam_get_refind_results(self.h, s, fl.ctypes.data, pos.ctypes.data, cap, C.byref(n)))
        return fl[:min(cap, n.value)], pos[:min(cap, n.value)]

    def epipolar_search(self, s, src_kf, level, cand_xy, src_pose, tgt_pose, depth_mean, depth_sigma, wiggle_scale):
        """The search of MapMaker::AddPointEpipolar for candidates of a source keyframe level in stream s's current keyframe."""
        xy = np.ascontiguousarray(cand_xy, dtype=np.int32).reshape(-1, 2); n = len(xy)
        sp = np.ascontiguousarray(src_pose, dtype=np.float64).reshape(12); tp = np.ascontiguousarray(tgt_pose, dtype=np.float64).reshape(12)
        found = np.zeros(max(n, 1), dtype=np.int32); pos = np.zeros((max(n, 1), 2)); bi = np.zeros(max(n, 1), dtype=np.int32); bs = np.zeros(max(n, 1), dtype=np.int32)
        self._ck(self.L.vslam_epipolar_search(self.h, s, src_kf, level, n, xy.ctypes.data, sp.ctypes.data, tp.ctypes.data, depth_mean, depth_sigma, wiggle_scale,
                                              found.ctypes.data, pos.ctypes.data, bi.ctypes.data, bs.ctypes.data))
        return found[:n], pos[:n], bi[:n], bs[:n]

    def epipolar_make_points(self, level, cand_xy, found_pos, src_pose, tgt_pose):
        """Tail of MapMaker::AddPointEpipolar for converged candidates: triangulated world points and the vslam_set_map fields."""
        xy = np.ascontiguousarray(cand_xy, dtype=np.int32).reshape(-1, 2); n = len(xy)
        fp = np.ascontiguousarray(found_pos, dtype=np.float64).reshape(-1, 2)
        assert len(fp) == n
        sp = np.ascontiguousarray(src_pose, dtype=np.float64).reshape(12); tp = np.ascontiguousarray(tgt_pose, dtype=np.float64).reshape(12)
        m = max(n, 1)
        world = np.zeros((m, 3)); right = np.zeros((m, 3)); down = np.zeros((m, 3)); irc = np.zeros((m, 2), dtype=np.int32); lvl = np.zeros(m, dtype=np.int32)
        self._ck(self.L.vslam_epipolar_make_points(self.h, level, n, xy.ctypes.data, fp.ctypes.data, sp.ctypes.data, tp.ctypes.data,
                                                   world.ctypes.data, right.ctypes.data, down.ctypes.data, irc.ctypes.data, lvl.ctypes.data))
        return world[:n], right[:n], down[:n], irc[:n], lvl[:n]

    def project_and_derivs(self, only_found=True):
        self._ck(self.L.vslam_project_and_derivs(self.h, int(only_found)))

    def calc_jacobians(self):
        self._ck(self.L.vslam_calc_jacobians(self.h))

    def calc_pose_update(self, override_sigma=0.0, mark_outliers=False, apply=False):
        out = np.zeros((self.S, 6))
        self._ck(self.L.vslam_calc_pose_update(self.h, float(override_sigma), int(mark_outliers), int(apply), out.ctypes.data))
        return out

    def track_map(self):
        self._ck(self.L.vslam_track_map(self.h))

    def track_frame(self, frames):
        f = np.ascontiguousarray(frames, dtype=np.uint8)
        assert f.shape[0] == self.S
        self._ck(self.L.vslam_track_frame(self.h, f.ctypes.data, f.shape[2], f.shape[1] * f.shape[2]))

    def track_frame_ptr(self, ptr, stride, frame_stride, device=False):
        fn = self.L.vslam_track_frame_dev if device else self.L.vslam_track_frame
        self._ck(fn(self.h, int(ptr), stride, frame_stride))

    def track_frame_async(self, ptr, stride, frame_stride, poses_out_ptr=None):
        """Pipelined host-input step; returns the step id to pass to wait_step."""
        rc = self.L.vslam_track_frame_async(self.h, int(ptr), stride, frame_stride, None if poses_out_ptr is None else int(poses_out_ptr))
        if rc < 0:
            self._ck(rc)
        return rc

    def wait_step(self, step):
        self._ck(self.L.vslam_wait_step(self.h, step))

    def sync(self):
        self._ck(self.L.vslam_sync(self.h))

    STAGES = ("pyrfast_l0", "pyrfast_l1", "pyrfast_l2", "pyrfast_l3", "project_lists", "search_coarse", "pose_coarse", "search_fine",
              "pose_fine", "h2d", "other")

    def set_timing(self, on=True):
        self._ck(self.L.vslam_set_timing(self.h, int(on)))

    def stage_times(self):
        """{stage: (ms, launches)} accumulated since the last call (synchronises)."""
        ms = np.zeros(len(self.STAGES))
        n = np.zeros(len(self.STAGES), dtype=np.int32)
        self._ck(self.L.vslam_get_stage_times(self.h, ms.ctypes.data, n.ctypes.data))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.STAGES)}

    def kernel_launches(self):
        return int(self.L.vslam_kernel_launches(self.h))

    def frame_lookahead_active(self):
        """Whether the next track_frame* call runs with frame look-ahead (execution only)."""
        return bool(self.L.vslam_frame_lookahead_active(self.h))

"""ctypes binding of oracle/_ref/libvslam_ref.so (the compiled reference; see build_ref.sh).

TEST INFRASTRUCTURE ONLY — imported by tests/, by tests/golden/make_golden.py and by bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libvslam_ref.so")

_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(LIB_PATH)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    vp, i, d, u = C.c_void_p, C.c_int, C.c_double, C.c_uint
    sig("ref_srand", None, u)
    sig("ref_rand", i)
    sig("ref_kf_create", vp)
    sig("ref_kf_destroy", None, vp)
    sig("ref_kf_make_lite", None, vp, _u8p, i, i, i, C.c_void_p)
    sig("ref_kf_make_rest", None, vp)
    sig("ref_kf_set_pose", None, vp, _f64p)
    sig("ref_kf_level_dims", None, vp, i, C.POINTER(i), C.POINTER(i))
    sig("ref_kf_level_pixels", None, vp, i, _u8p)
    sig("ref_kf_num_corners", i, vp, i)
    sig("ref_kf_corners", None, vp, i, _i32p)
    sig("ref_kf_row_lut", i, vp, i, _i32p)
    sig("ref_kf_num_max_corners", i, vp, i)
    sig("ref_kf_max_corners", None, vp, i, _i32p)
    sig("ref_kf_num_candidates", i, vp, i)
    sig("ref_kf_candidates", None, vp, i, _i32p, _f64p)
    sig("ref_fast10", i, _u8p, i, i, i, i, _i32p, i)
    sig("ref_shi_tomasi", d, _u8p, i, i, i, i, i, i)
    sig("ref_cam_create", vp, d, d, i)
    sig("ref_cam_destroy", None, vp)
    sig("ref_cam_fix_radius", None, vp, i)
    sig("ref_cam_scalars", None, vp, _f64p)
    sig("ref_cam_project", None, vp, _f64p, _f64p, C.POINTER(i), _f64p)
    sig("ref_cam_unproject", None, vp, _f64p, _f64p)
    sig("ref_se3_exp", None, _f64p, _f64p)
    sig("ref_se3_ln", None, _f64p, _f64p)
    sig("ref_se3_mul", None, _f64p, _f64p, _f64p)
    sig("ref_se3_inverse", None, _f64p, _f64p)
    sig("ref_map_create", vp)
    sig("ref_map_set_good", None, vp, i)
    sig("ref_map_add_keyframe", i, vp, vp)
    sig("ref_map_num_points", i, vp)
    sig("ref_map_add_point", i, vp, vp, i, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p)
    sig("ref_map_point_pixel_vectors", None, vp, i, _f64p, _f64p)
    sig("ref_map_point_counts", None, vp, i, C.POINTER(i), C.POINTER(i))
    sig("ref_pf_create", vp, i)
    sig("ref_pf_destroy", None, vp)
    sig("ref_pf_max_ssd", i, vp)
    sig("ref_pf_calc_level_warp", i, vp, vp, i, _f64p, _f64p, _f64p)
    sig("ref_pf_set_level_warp", None, vp, i, _f64p)
    sig("ref_pf_make_template", i, vp, vp, i, _u8p, C.POINTER(i), C.POINTER(i))
    sig("ref_pf_set_template", None, vp, _u8p)
    sig("ref_pf_make_template_nowarp", i, vp, vp, i, i, i, _u8p, C.POINTER(i), C.POINTER(i))
    sig("ref_pf_zmssd", i, vp, vp, i, i, i)
    sig("ref_pf_find_coarse", i, vp, d, d, vp, u, _f64p)
    sig("ref_pf_subpix", i, vp, vp, _f64p, i, _f64p, C.c_void_p)
    sig("ref_mp_create", vp)
    sig("ref_mp_destroy", None, vp)
    sig("ref_mp_set_max_ssd", None, i)
    sig("ref_mp_sample", None, vp, vp, i, i, C.c_void_p)
    sig("ref_mp_find", i, vp, vp, _f64p, i, i)
    sig("ref_tracker_create", vp, i, i, vp, vp, i)
    sig("ref_tracker_set_pose", None, vp, _f64p)
    sig("ref_tracker_get_pose", None, vp, _f64p)
    sig("ref_tracker_set_velocity", None, vp, _f64p, d)
    sig("ref_tracker_get_velocity", None, vp, _f64p, C.POINTER(d))
    sig("ref_tracker_set_scene_depth", None, vp, d, d)
    sig("ref_tracker_get_scene_depth", None, vp, C.POINTER(d), C.POINTER(d))
    sig("ref_tracker_current_kf", vp, vp)
    sig("ref_tracker_make_current_kf", None, vp, _u8p, i, i, i)
    sig("ref_tracker_track_map", None, vp)
    sig("ref_tracker_track_frame", None, vp, _u8p, i, i, i)
    sig("ref_refind", None, vp, _i32p, i, i, i, _i32p, _f64p)
    sig("ref_epipolar_search", None, vp, vp, vp, _f64p, _f64p, d, d, d, i, i, _i32p, _f64p)
    sig("ref_set_keyframe_policy", None, vp, i, d, d, d)
    sig("ref_keyframe_info", None, vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int))
    sig("ref_epipolar_point_fields", None, vp, vp, _f64p, i, i, _f64p, _f64p)
    # the reference's own MapMaker functions (jni/MapMaker.cc is part of the build)
    sig("ref_mm_refind", None, vp, _i32p, i, _i32p, _f64p)
    sig("ref_mm_add_point_epipolar", i, vp, vp, vp, _f64p, _f64p, d, d, d, i, i, _f64p)
    sig("ref_mm_reproject_point", None, vp, _f64p, _f64p, _f64p, _f64p)
    sig("ref_mm_keyframe_heuristics", None, vp, C.POINTER(i), C.POINTER(i), C.POINTER(d))
    sig("ref_kf_num_candidates_l", i, vp, i)
    sig("ref_kf_make_sbi", None, vp)
    sig("ref_tracker_set_lost", None, vp, i, i)
    sig("ref_tracker_trail_start", i, vp)
    sig("ref_tracker_trail_advance", i, vp, i)
    sig("ref_tracker_trail_count", i, vp)
    sig("ref_tracker_trails", None, vp, _f64p)
    sig("ref_tracker_motion_model", None, vp, i)
    sig("ref_tracker_track_frame_nosbi", None, vp, _u8p, i, i, i)
    sig("ref_tracker_set_sbi_rot", None, vp, _f64p, i)
    sig("ref_tracker_get_sbi_rot", None, vp, _f64p)
    sig("ref_tracker_counters", None, vp, _i32p, _i32p, C.POINTER(i), C.POINTER(i), C.POINTER(i))
    sig("ref_tracker_message", i, vp, C.c_char_p, i)
    sig("ref_tracker_project_all", None, vp)
    sig("ref_tracker_point_state", None, vp, i, _i32p, _f64p)
    sig("ref_tracker_point_template", None, vp, i, _u8p, C.POINTER(i), C.POINTER(i))
    sig("ref_tracker_search_for_points", i, vp, _i32p, i, i, i)
    sig("ref_tracker_clear_counters", None, vp)
    sig("ref_tracker_calc_jacobians", None, vp, _i32p, i)
    sig("ref_tracker_project_and_derivs", None, vp, _i32p, i, i)
    sig("ref_tracker_linear_update", None, vp, _i32p, i, _f64p)
    sig("ref_tracker_calc_pose_update", None, vp, _i32p, i, d, i, i, _f64p)
    sig("ref_tukey_sigma_squared", d, _f64p, i)
    sig("ref_tracker_num_measurements", i, vp)
    sig("ref_sbi_create", vp, vp, d)
    sig("ref_sbi_destroy", None, vp)
    sig("ref_sbi_dims", None, C.POINTER(i), C.POINTER(i))
    sig("ref_sbi_reset_size", None)
    sig("ref_sbi_template", None, vp, _f32p)
    sig("ref_sbi_small", None, vp, _u8p)
    sig("ref_sbi_rotation", d, vp, vp, vp, i, C.c_void_p, _f64p)
    _lib = L
    return L


class RefKeyFrame:
    """KeyFrame of the compiled reference (jni/KeyFrame.h)."""

    def __init__(self, handle=None):
        self.L = lib()
        self.owned = handle is None
        self.h = self.L.ref_kf_create() if handle is None else handle

    def make_lite(self, gray: np.ndarray, rgba: np.ndarray | None = None):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        h, w = gray.shape
        self._keep = (gray, rgba)
        self.L.ref_kf_make_lite(self.h, gray, w, h, w, None if rgba is None else rgba.ctypes.data)
        return self

    def make_rest(self):
        self.L.ref_kf_make_rest(self.h)

    def dims(self, l):
        w, h = C.c_int(), C.c_int()
        self.L.ref_kf_level_dims(self.h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def pixels(self, l) -> np.ndarray:
        w, h = self.dims(l)
        out = np.empty((h, w), dtype=np.uint8)
        self.L.ref_kf_level_pixels(self.h, l, out)
        return out

    def corners(self, l) -> np.ndarray:
        n = self.L.ref_kf_num_corners(self.h, l)
        out = np.empty((n, 2), dtype=np.int32)
        if n:
            self.L.ref_kf_corners(self.h, l, out)
        return out

    def row_lut(self, l) -> np.ndarray:
        _, h = self.dims(l)
        out = np.empty(h, dtype=np.int32)
        n = self.L.ref_kf_row_lut(self.h, l, out)
        assert n == h
        return out

    def max_corners(self, l) -> np.ndarray:
        n = self.L.ref_kf_num_max_corners(self.h, l)
        out = np.empty((n, 2), dtype=np.int32)
        if n:
            self.L.ref_kf_max_corners(self.h, l, out)
        return out

    def candidates(self, l):
        n = self.L.ref_kf_num_candidates(self.h, l)
        xy = np.empty((n, 2), dtype=np.int32)
        sc = np.empty(n, dtype=np.float64)
        if n:
            self.L.ref_kf_candidates(self.h, l, xy, sc)
        return xy, sc

    def set_pose(self, pose):
        self.L.ref_kf_set_pose(self.h, np.ascontiguousarray(pose, dtype=np.float64).reshape(12))


class RefWorld:
    """Camera + map + tracker of the compiled reference, filled from a SyntheticMap."""

    def __init__(self, width, height, src_gray, smap, fix_radius=True):
        L = self.L = lib()
        self.cam = L.ref_cam_create(float(width), float(height), int(fix_radius))
        self.map = L.ref_map_create()
        self.src_kf = RefKeyFrame().make_lite(src_gray)
        self.src_kf.set_pose(np.concatenate([np.eye(3), np.zeros((3, 1))], axis=1))
        L.ref_map_add_keyframe(self.map, self.src_kf.h)
        normal = np.array([0.0, 0.0, -1.0])
        for k in range(smap.n):
            L.ref_map_add_point(self.map, self.src_kf.h, int(smap.src_level[k]), smap.ir_center[k].astype(np.float64),
                                np.ascontiguousarray(smap.world[k]), np.ascontiguousarray(smap.center_nc[k]),
                                np.ascontiguousarray(smap.one_right_nc[k]), np.ascontiguousarray(smap.one_down_nc[k]), normal)
        L.ref_map_set_good(self.map, 1)
        self.n = smap.n
        self.tracker = L.ref_tracker_create(int(width), int(height), self.cam, self.map, int(fix_radius))
        self.width, self.height = width, height

    def append_points(self, smap):
        """Map::vpPoints.push_back of new points while the tracker runs."""
        normal = np.array([0.0, 0.0, -1.0])
        for k in range(smap.n):
            self.L.ref_map_add_point(self.map, self.src_kf.h, int(smap.src_level[k]), smap.ir_center[k].astype(np.float64),
                                     np.ascontiguousarray(smap.world[k]), np.ascontiguousarray(smap.center_nc[k]),
                                     np.ascontiguousarray(smap.one_right_nc[k]), np.ascontiguousarray(smap.one_down_nc[k]), normal)
        self.n += smap.n

    # -- tracker state
    def set_pose(self, pose):
        self.L.ref_tracker_set_pose(self.tracker, np.ascontiguousarray(pose, dtype=np.float64).reshape(12))

    def get_pose(self):
        out = np.empty(12)
        self.L.ref_tracker_get_pose(self.tracker, out)
        return out.reshape(3, 4)

    def make_current_kf(self, gray):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        self._cur = gray
        h, w = gray.shape
        self.L.ref_tracker_make_current_kf(self.tracker, gray, w, h, w)
        return RefKeyFrame(self.L.ref_tracker_current_kf(self.tracker))

    def point_states(self):
        ints = np.zeros((self.n, 8), dtype=np.int32)
        dbl = np.zeros((self.n, 32), dtype=np.float64)
        for k in range(self.n):
            self.L.ref_tracker_point_state(self.tracker, k, ints[k], dbl[k])
        return ints, dbl

    def counters(self):
        a = np.zeros(4, dtype=np.int32)
        f = np.zeros(4, dtype=np.int32)
        q, lost, dc = C.c_int(), C.c_int(), C.c_int()
        self.L.ref_tracker_counters(self.tracker, a, f, C.byref(q), C.byref(lost), C.byref(dc))
        return a, f, q.value, lost.value, dc.value

    def message(self):
        """Tracker::GetMessageForUser of the last frame."""
        buf = C.create_string_buffer(1024)
        self.L.ref_tracker_message(self.tracker, buf, 1024)
        return buf.value.decode(errors="replace")

    def pixel_vectors(self):
        r = np.empty((self.n, 3))
        d = np.empty((self.n, 3))
        for k in range(self.n):
            self.L.ref_map_point_pixel_vectors(self.map, k, r[k], d[k])
        return r, d

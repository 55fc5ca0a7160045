"""ctypes binding of oracle/libvslam_oracle.so (the CPU restatement, oracle/vslam_oracle.cc).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, as the checker.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvslam_oracle.so")

_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

_lib = None


def build():
    subprocess.check_call(["bash", os.path.join(HERE, "build_oracle.sh")], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    vp, i, d, u, pi, pd = C.c_void_p, C.c_int, C.c_double, C.c_uint, C.POINTER(C.c_int), C.POINTER(C.c_double)
    sig("orc_rand_create", vp, u)
    sig("orc_rand_next", i, vp)
    sig("orc_rand_destroy", None, vp)
    sig("orc_kf_create", vp)
    sig("orc_kf_destroy", None, vp)
    sig("orc_kf_make_lite", None, vp, _u8p, i, i, i)
    sig("orc_kf_make_rest", None, vp)
    sig("orc_kf_level_dims", None, vp, i, pi, pi)
    sig("orc_kf_level_pixels", None, vp, i, _u8p)
    sig("orc_kf_num_corners", i, vp, i)
    sig("orc_kf_corners", None, vp, i, _i32p)
    sig("orc_kf_row_lut", i, vp, i, _i32p)
    sig("orc_kf_num_max_corners", i, vp, i)
    sig("orc_kf_max_corners", None, vp, i, _i32p)
    sig("orc_kf_num_candidates", i, vp, i)
    sig("orc_kf_candidates", None, vp, i, _i32p, _f64p)
    sig("orc_kf_fast_scores", None, vp, i, i, _i32p)
    sig("orc_shi_tomasi", d, vp, i, i, i, i)
    sig("orc_cam_project", None, _f64p, _f64p, _f64p, pi, _f64p)
    sig("orc_cam_unproject", None, _f64p, _f64p, _f64p)
    sig("orc_se3_exp", None, _f64p, _f64p)
    sig("orc_se3_ln", None, _f64p, _f64p)
    sig("orc_se3_mul", None, _f64p, _f64p, _f64p)
    sig("orc_se3_inverse", None, _f64p, _f64p)
    sig("orc_resize_linear_u8", None, _u8p, i, i, _u8p, i, i)
    sig("orc_make_template", i, vp, i, _i32p, i, _f64p, _u8p, pi, pi)
    sig("orc_zmssd", i, vp, i, _u8p, i, i, i)
    sig("orc_find_patch_coarse", i, vp, i, _u8p, i, d, d, u, _f64p, pi, C.POINTER(C.c_long))
    sig("orc_subpix", i, vp, i, _u8p, i, _f64p, i, _f64p, C.c_void_p)
    sig("orc_minipatch_find", i, vp, i, i, vp, _f64p, i, i, i, pi)
    sig("orc_trails_create", vp)
    sig("orc_trails_destroy", None, vp)
    sig("orc_trails_start", i, vp, vp)
    sig("orc_trails_advance", i, vp, vp, i)
    sig("orc_trails_count", i, vp)
    sig("orc_trails_get", None, vp, _f64p)
    sig("orc_tracker_create", vp, _f64p, i)
    sig("orc_tracker_destroy", None, vp)
    sig("orc_tracker_seed", None, vp, u)
    sig("orc_tracker_set_truncate", None, vp, i)
    sig("orc_tracker_set_map", None, vp, vp, i, _f64p, _f64p, _f64p, _i32p, _i32p)
    sig("orc_tracker_set_point_source_kf", None, vp, i, vp)
    sig("orc_tracker_set_pose", None, vp, _f64p)
    sig("orc_tracker_get_pose", None, vp, _f64p)
    sig("orc_tracker_set_velocity", None, vp, _f64p, d)
    sig("orc_tracker_get_velocity", None, vp, _f64p, pd)
    sig("orc_tracker_set_scene_depth", None, vp, d, d)
    sig("orc_tracker_get_scene_depth", None, vp, pd, pd)
    sig("orc_tracker_set_sbi_rot", None, vp, _f64p, i)
    sig("orc_tracker_current_kf", vp, vp)
    sig("orc_tracker_make_current_kf", None, vp, _u8p, i, i, i)
    sig("orc_tracker_track_map", None, vp)
    sig("orc_tracker_motion_model", None, vp, i)
    sig("orc_tracker_assess_quality", None, vp)
    sig("orc_tracker_track_frame", None, vp, _u8p, i, i, i)
    sig("orc_tracker_counters", None, vp, _i32p, _i32p, pi, pi, pi)
    sig("orc_tracker_zmssd_evals", C.c_long, vp)
    sig("orc_tracker_num_updates", i, vp)
    sig("orc_tracker_updates", None, vp, _f64p, _f64p)
    sig("orc_tracker_project_all", None, vp)
    sig("orc_tracker_point_state", None, vp, i, _i32p, _f64p)
    sig("orc_tracker_point_template", None, vp, i, _u8p, pi, pi)
    sig("orc_tracker_point_counts", None, vp, i, pi, pi)
    sig("orc_tracker_search_for_points", i, vp, _i32p, i, i, i)
    sig("orc_tracker_clear_counters", None, vp)
    sig("orc_epipolar_search", None, vp, vp, vp, _f64p, _f64p, d, d, d, i, i, i, _i32p, _f64p, C.c_void_p)
    sig("orc_tracker_append_points", None, vp, vp, i, _f64p, _f64p, _f64p, _i32p, _i32p)
    sig("orc_tracker_set_keyframe_policy", None, vp, i, d, d, d, i)
    sig("orc_tracker_keyframe_info", None, vp, pi, pi, pi, pi)
    sig("orc_epipolar_point_fields", None, _f64p, _f64p, i, i, i, _f64p, _f64p)
    sig("orc_tracker_refind", None, vp, _i32p, i, i, i, i, _i32p, _f64p)
    sig("orc_tracker_calc_jacobians", None, vp, _i32p, i)
    sig("orc_tracker_project_and_derivs", None, vp, _i32p, i, i)
    sig("orc_tracker_linear_update", None, vp, _i32p, i, _f64p)
    sig("orc_tracker_calc_pose_update", None, vp, _i32p, i, d, i, i, _f64p)
    sig("orc_tukey_sigma_squared", d, _f64p, i)
    sig("orc_tracker_enable_sbi", None, vp, _f64p)
    sig("orc_tracker_add_reloc_keyframe", None, vp, vp, _f64p)
    sig("orc_tracker_reloc_info", None, vp, pi, pd, pi)
    sig("orc_tracker_set_lost", None, vp, i, i)
    sig("orc_tracker_get_sbi_rot", None, vp, _f64p)
    sig("orc_sbi_create", vp, vp, d)
    sig("orc_sbi_destroy", None, vp)
    sig("orc_sbi_dims", None, vp, pi, pi)
    sig("orc_sbi_template", None, vp, np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS"))
    sig("orc_sbi_small", None, vp, _u8p)
    sig("orc_sbi_rotation", d, vp, vp, _f64p, i, C.c_void_p, _f64p)
    _lib = L
    return L


def triangulate(cam13, src_pose, tgt_pose, root_pos, found_pos):
    """MapMaker::ReprojectPoint (jni/MapMaker.cc:176-200) as AddPointEpipolar calls it (:648): the homogeneous two-view system of the
    source ray A = UnProject(root) and the target ray B = UnProject(found), solved by the right singular vector of the smallest
    singular value (the reference uses Eigen::JacobiSVD -- Eigen is not under /root/reference, so LAPACK's SVD stands in: the vector
    is unique up to sign and the sign cancels), then moved from the target camera's frame to the world.  Parity unpinned (no Eigen)."""
    L = lib()
    cam13 = np.ascontiguousarray(cam13, dtype=np.float64)
    S = np.vstack([np.asarray(src_pose, dtype=np.float64).reshape(3, 4), [0, 0, 0, 1]])
    T = np.vstack([np.asarray(tgt_pose, dtype=np.float64).reshape(3, 4), [0, 0, 0, 1]])
    a, b = np.zeros(2), np.zeros(2)
    L.orc_cam_unproject(cam13, np.ascontiguousarray(root_pos, dtype=np.float64), a)
    L.orc_cam_unproject(cam13, np.ascontiguousarray(found_pos, dtype=np.float64), b)
    P = (S @ np.linalg.inv(T))[:3]                    # se3AfromB, A = source, B = target
    A = np.array([[-1.0, 0.0, b[0], 0.0], [0.0, -1.0, b[1], 0.0], a[0] * P[2] - P[0], a[1] * P[2] - P[1]])
    v = np.linalg.svd(A)[2][3].copy()
    if v[3] == 0.0:
        v[3] = 0.00001
    return (np.linalg.inv(T) @ np.append(v[:3] / v[3], 1.0))[:3]


def epipolar_point_fields(cam13, src_pose, level, cx, cy, world):
    out = np.zeros(15)
    lib().orc_epipolar_point_fields(np.ascontiguousarray(cam13, dtype=np.float64), np.ascontiguousarray(src_pose, dtype=np.float64).reshape(12), level, int(cx), int(cy),
                                    np.ascontiguousarray(world, dtype=np.float64), out)
    return out.reshape(5, 3)


class OrcKeyFrame:
    """KeyFrame of the restatement (MakeKeyFrame_Lite / _Rest)."""

    def __init__(self, handle=None):
        self.L = lib()
        self.h = self.L.orc_kf_create() if handle is None else handle

    def make_lite(self, gray: np.ndarray):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        h, w = gray.shape
        self.L.orc_kf_make_lite(self.h, gray, w, h, w)
        return self

    def make_rest(self):
        self.L.orc_kf_make_rest(self.h)

    def dims(self, l):
        w, h = C.c_int(), C.c_int()
        self.L.orc_kf_level_dims(self.h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def pixels(self, l):
        w, h = self.dims(l)
        out = np.empty((h, w), dtype=np.uint8)
        self.L.orc_kf_level_pixels(self.h, l, out)
        return out

    def corners(self, l):
        n = self.L.orc_kf_num_corners(self.h, l)
        out = np.empty((n, 2), dtype=np.int32)
        if n:
            self.L.orc_kf_corners(self.h, l, out)
        return out

    def row_lut(self, l):
        _, h = self.dims(l)
        out = np.empty(h, dtype=np.int32)
        assert self.L.orc_kf_row_lut(self.h, l, out) == h
        return out

    def max_corners(self, l):
        n = self.L.orc_kf_num_max_corners(self.h, l)
        out = np.empty((n, 2), dtype=np.int32)
        if n:
            self.L.orc_kf_max_corners(self.h, l, out)
        return out

    def candidates(self, l):
        n = self.L.orc_kf_num_candidates(self.h, l)
        xy = np.empty((n, 2), dtype=np.int32)
        sc = np.empty(n, dtype=np.float64)
        if n:
            self.L.orc_kf_candidates(self.h, l, xy, sc)
        return xy, sc

    def fast_scores(self, l, barrier=10):
        n = self.L.orc_kf_num_corners(self.h, l)
        out = np.empty(n, dtype=np.int32)
        if n:
            self.L.orc_kf_fast_scores(self.h, l, barrier, out)
        return out


class OrcTrails:
    """Trail list of the restatement (Tracker::TrailTracking_Start / _Advance)."""

    def __init__(self):
        self.L = lib()
        self.h = self.L.orc_trails_create()

    def start(self, kf: "OrcKeyFrame"):
        return self.L.orc_trails_start(self.h, kf.h)

    def advance(self, kf: "OrcKeyFrame", max_ssd=100000):
        return self.L.orc_trails_advance(self.h, kf.h, max_ssd)

    def trails(self):
        n = self.L.orc_trails_count(self.h)
        out = np.zeros((n, 4))
        if n:
            self.L.orc_trails_get(self.h, out)
        return out


class OrcWorld:
    """Camera scalars + map + tracker of the restatement, filled from a SyntheticMap."""

    def __init__(self, cam, src_gray, smap, P=11, pix_right=None, pix_down=None):
        L = self.L = lib()
        self.cam13 = cam.scalars()
        self.src_kf = OrcKeyFrame().make_lite(src_gray)
        self.tracker = L.orc_tracker_create(self.cam13, P)
        self.n = smap.n
        self.P = P
        right = smap.pix_right_w if pix_right is None else pix_right
        down = smap.pix_down_w if pix_down is None else pix_down
        self._keep = [np.ascontiguousarray(a) for a in (smap.world, right, down, smap.ir_center.astype(np.int32), smap.src_level.astype(np.int32))]
        L.orc_tracker_set_map(self.tracker, self.src_kf.h, self.n, *self._keep)

    def append_points(self, smap, pix_right=None, pix_down=None, src_kf=None):
        """New map points while tracking runs (existing points keep their TrackerData)."""
        right = smap.pix_right_w if pix_right is None else pix_right
        down = smap.pix_down_w if pix_down is None else pix_down
        arrs = [np.ascontiguousarray(a) for a in (smap.world, right, down, smap.ir_center.astype(np.int32), smap.src_level.astype(np.int32))]
        self._keep += arrs
        self.L.orc_tracker_append_points(self.tracker, (src_kf or self.src_kf).h, smap.n, *arrs)
        self.n += smap.n

    def set_pose(self, pose):
        self.L.orc_tracker_set_pose(self.tracker, np.ascontiguousarray(pose, dtype=np.float64).reshape(12))

    def get_pose(self):
        out = np.empty(12)
        self.L.orc_tracker_get_pose(self.tracker, out)
        return out.reshape(3, 4)

    def make_current_kf(self, gray):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        h, w = gray.shape
        self.L.orc_tracker_make_current_kf(self.tracker, gray, w, h, w)
        return OrcKeyFrame(self.L.orc_tracker_current_kf(self.tracker))

    def point_states(self):
        ints = np.zeros((self.n, 8), dtype=np.int32)
        dbl = np.zeros((self.n, 32), dtype=np.float64)
        for k in range(self.n):
            self.L.orc_tracker_point_state(self.tracker, k, ints[k], dbl[k])
        return ints, dbl

    def point_template(self, k):
        t = np.zeros(self.P * self.P, dtype=np.uint8)
        s, q = C.c_int(), C.c_int()
        self.L.orc_tracker_point_template(self.tracker, k, t, C.byref(s), C.byref(q))
        return t.reshape(self.P, self.P), s.value, q.value

    def counters(self):
        a = np.zeros(4, dtype=np.int32)
        f = np.zeros(4, dtype=np.int32)
        q, lost, dc = C.c_int(), C.c_int(), C.c_int()
        self.L.orc_tracker_counters(self.tracker, a, f, C.byref(q), C.byref(lost), C.byref(dc))
        return a, f, q.value, lost.value, dc.value

    def updates(self):
        n = self.L.orc_tracker_num_updates(self.tracker)
        u = np.zeros((n, 6))
        s = np.zeros(n)
        if n:
            self.L.orc_tracker_updates(self.tracker, u, s)
        return u, s

    def point_counts(self):
        out = np.zeros((self.n, 2), dtype=np.int32)
        for k in range(self.n):
            a, b = C.c_int(), C.c_int()
            self.L.orc_tracker_point_counts(self.tracker, k, C.byref(a), C.byref(b))
            out[k] = (a.value, b.value)
        return out

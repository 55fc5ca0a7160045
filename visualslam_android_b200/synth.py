"""Deterministic synthetic inputs for the tracking front-end.

Everything here produces INPUT BYTES (texture, rendered frames, a map built from a source
keyframe, camera poses) that are handed identically to the CPU oracle and to the CUDA path.
None of it is on the product path and none of it is checked for parity: it only has to be
deterministic.  The camera model used for rendering mirrors the reference's FOV model
(jni/ATANCamera.cc:37-164) so that the synthetic frames look like what its tracker expects.
"""
from __future__ import annotations

import dataclasses

import numpy as np

# jni/ATANCamera.cc:20-24 (hard-coded calibration)
CAMERA_PARAMS = (0.841906, 1.10893, 0.505171, 0.470265, -0.0133843)

TEXTURE_SEED = 20261018
MAP_SEED = 7
# SURVEY.md §8d "Config 1" frame pose twist (translation xyz, rotation xyz)
CONFIG1_TWIST = (0.020, -0.010, 0.010, 0.010, -0.015, 0.020)


@dataclasses.dataclass
class Camera:
    """Host scalars of the FOV camera for a given image size (mirrors ATANCamera::RefreshParams)."""

    width: int
    height: int
    fix_radius: bool = True
    params: tuple = CAMERA_PARAMS

    def __post_init__(self):
        p = self.params
        self.fx = self.width * p[0]
        self.fy = self.height * p[1]
        self.cx = self.width * p[2] - 0.5
        self.cy = self.height * p[3] - 0.5
        self.w = p[4]
        if self.w != 0.0:
            self.two_tan = 2.0 * np.tan(self.w / 2.0)
            self.one_over_two_tan = 1.0 / self.two_tan
            self.w_inv = 1.0 / self.w
            self.distortion = 1.0
        else:
            self.two_tan = self.one_over_two_tan = self.w_inv = self.distortion = 0.0
        if self.fix_radius:
            v0 = max(p[2], 1.0 - p[2]) / p[0]
            v1 = max(p[3], 1.0 - p[3]) / p[1]
        else:  # jni/ATANCamera.cc:70-82: int temporaries -> 0
            v0 = max(int(p[2]), int(1.0 - p[2])) / p[0]
            v1 = max(int(p[3]), int(1.0 - p[3])) / p[1]
        r = float(np.sqrt(v0 * v0 + v1 * v1))
        self.largest_radius = float(np.tan(r * self.w) * self.one_over_two_tan) if self.w != 0.0 else r
        self.max_r = 1.5 * self.largest_radius

    def scalars(self) -> np.ndarray:
        """fx fy cx cy w w_inv two_tan one_over_two_tan distortion largest_radius max_r width height (13 doubles)."""
        return np.array([self.fx, self.fy, self.cx, self.cy, self.w, self.w_inv, self.two_tan, self.one_over_two_tan,
                         self.distortion, self.largest_radius, self.max_r, float(self.width), float(self.height)], dtype=np.float64)

    def unproject(self, u, v):
        """pixel -> z=1 plane (vectorised)."""
        dx = (np.asarray(u, dtype=np.float64) - self.cx) / self.fx
        dy = (np.asarray(v, dtype=np.float64) - self.cy) / self.fy
        rd = np.sqrt(dx * dx + dy * dy)
        if self.w == 0.0:
            return dx, dy
        r = np.tan(rd * self.w) * self.one_over_two_tan
        f = np.where(rd > 0.01, r / np.maximum(rd, 1e-300), 1.0)
        return dx * f, dy * f

    def project(self, x, y):
        """z=1 plane -> pixel (vectorised)."""
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        r = np.sqrt(x * x + y * y)
        if self.w == 0.0:
            f = np.ones_like(r)
        else:
            f = np.where(r < 0.001, 1.0, self.w_inv * np.arctan(r * self.two_tan) / np.maximum(r, 1e-300))
        return self.cx + self.fx * f * x, self.cy + self.fy * f * y


def se3_exp(mu) -> np.ndarray:
    """SE3 exponential, 3x4 [R|t], twist = (translation, rotation)."""
    mu = np.asarray(mu, dtype=np.float64)
    t, w = mu[:3], mu[3:]
    th2 = float(w @ w)
    th = np.sqrt(th2)
    cr = np.cross(w, t)
    if th2 < 1e-8:
        A, B = 1.0 - th2 / 6.0, 0.5
        trans = t + 0.5 * cr
    else:
        if th2 < 1e-6:
            C = (1.0 - th2 / 20.0) / 6.0
            A = 1.0 - th2 * C
            B = 0.5 - 0.25 * th2 / 6.0
        else:
            A = np.sin(th) / th
            B = (1 - np.cos(th)) / th2
            C = (1 - A) / th2
        trans = t + B * cr + C * np.cross(w, cr)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    R = np.eye(3) + A * K + B * (K @ K)
    return np.concatenate([R, trans[:, None]], axis=1)


def se3_mul(a, b) -> np.ndarray:
    return np.concatenate([a[:, :3] @ b[:, :3], (a[:, :3] @ b[:, 3] + a[:, 3])[:, None]], axis=1)


IDENTITY_POSE = np.concatenate([np.eye(3), np.zeros((3, 1))], axis=1)


def make_texture(size: int = 4096, seed: int = TEXTURE_SEED) -> np.ndarray:
    """Blocky multi-octave random texture: FAST corners at every pyramid level."""
    rs = np.random.RandomState(seed)
    acc = np.full((size, size), 128.0, dtype=np.float32)
    for block, amp in ((128, 30), (64, 40), (32, 40), (16, 16), (8, 10)):
        n = size // block
        vals = rs.randint(-amp, amp + 1, size=(n, n)).astype(np.float32)
        acc += np.repeat(np.repeat(vals, block, axis=0), block, axis=1)
    acc += rs.randint(-2, 3, size=(size, size)).astype(np.float32)
    return np.clip(acc, 0, 255).astype(np.uint8)


def render_frame(tex: np.ndarray, cam: Camera, pose_cfw: np.ndarray, tex_scale: float | None = None) -> np.ndarray:
    """Render the textured plane z=1 (world frame) seen from camera pose `pose_cfw` (3x4 cam-from-world)."""
    H, W = cam.height, cam.width
    size = tex.shape[0]
    if tex_scale is None:
        tex_scale = cam.fx
    v, u = np.mgrid[0:H, 0:W]
    x, y = cam.unproject(u, v)
    R, t = pose_cfw[:, :3], pose_cfw[:, 3]
    o = -R.T @ t
    d = np.stack([x, y, np.ones_like(x)], axis=-1) @ R  # (R^T d) for every pixel
    s = (1.0 - o[2]) / d[..., 2]
    X = o[0] + s * d[..., 0]
    Y = o[1] + s * d[..., 1]
    tu = np.mod(X * tex_scale + size / 2.0, size - 1.0)
    tv = np.mod(Y * tex_scale + size / 2.0, size - 1.0)
    iu = np.floor(tu).astype(np.int64)
    iv = np.floor(tv).astype(np.int64)
    fu = (tu - iu).astype(np.float32)
    fv = (tv - iv).astype(np.float32)
    tf = tex.astype(np.float32)
    val = (1 - fv) * ((1 - fu) * tf[iv, iu] + fu * tf[iv, iu + 1]) + fv * ((1 - fu) * tf[iv + 1, iu] + fu * tf[iv + 1, iu + 1])
    return np.clip(np.floor(val + 0.5), 0, 255).astype(np.uint8)


@dataclasses.dataclass
class SyntheticMap:
    """SoA map, source keyframe = KF0 (pose identity).  Mirrors the MapPoint fields the tracker reads."""

    world: np.ndarray        # (N,3) f64  v3WorldPos
    pix_right_w: np.ndarray  # (N,3) f64  v3PixelRight_W
    pix_down_w: np.ndarray   # (N,3) f64  v3PixelDown_W
    ir_center: np.ndarray    # (N,2) i32  irCenter (source-level coordinates)
    src_level: np.ndarray    # (N,)  i32  nSourceLevel
    center_nc: np.ndarray    # (N,3) f64
    one_right_nc: np.ndarray  # (N,3) f64
    one_down_nc: np.ndarray  # (N,3) f64

    @property
    def n(self) -> int:
        return int(self.world.shape[0])


def refresh_pixel_vectors(center_nc, one_right_nc, one_down_nc, world, normal=(0.0, 0.0, -1.0)):
    """MapPoint::RefreshPixelVectors for a source keyframe at the identity pose (jni/MapPoint.cc:4-29)."""
    nrm = np.asarray(normal, dtype=np.float64)
    cam_height = np.abs(world @ nrm)
    pixel_rate = np.abs(center_nc @ nrm)
    right_rate = np.abs(one_right_nc @ nrm)
    down_rate = np.abs(one_down_nc @ nrm)
    c_on = center_nc * cam_height[:, None] / pixel_rate[:, None]
    r_on = one_right_nc * cam_height[:, None] / right_rate[:, None]
    d_on = one_down_nc * cam_height[:, None] / down_rate[:, None]
    return r_on - c_on, d_on - c_on


def build_map(cam: Camera, corners_per_level, level_dims, n_points: int, seed: int = MAP_SEED,
              split=(0.4, 0.3, 0.2, 0.1), border: int = 12) -> SyntheticMap:
    """Choose map points among KF0's FAST corners (>= `border` level-px from the image border)."""
    rs = np.random.RandomState(seed)
    want = [int(round(n_points * f)) for f in split]
    want[0] += n_points - sum(want)
    chosen = []
    spill = 0
    for l in (3, 2, 1, 0):  # coarse levels have few corners: spill the shortfall to finer levels
        c = np.asarray(corners_per_level[l], dtype=np.int64).reshape(-1, 2)
        w, h = level_dims[l]
        ok = (c[:, 0] >= border) & (c[:, 1] >= border) & (c[:, 0] < w - border) & (c[:, 1] < h - border)
        c = c[ok]
        perm = rs.permutation(len(c))
        k = min(len(c), want[l] + spill)
        spill = want[l] + spill - k
        chosen.append((l, c[perm[:k]]))
    chosen.reverse()
    lv = np.concatenate([np.full(len(c), l, dtype=np.int32) for l, c in chosen])
    ir = np.concatenate([c for _, c in chosen]).astype(np.int32)
    order = rs.permutation(len(lv))
    lv, ir = lv[order], ir[order]
    scale = (1 << lv).astype(np.float64)
    root = (ir.astype(np.float64) + 0.5) * scale[:, None] - 0.5  # LevelZeroPos

    def unit_ray(px, py):
        x, y = cam.unproject(px, py)
        v = np.stack([x, y, np.ones_like(x)], axis=-1)
        return v / np.linalg.norm(v, axis=-1, keepdims=True)

    center = unit_ray(root[:, 0], root[:, 1])
    right = unit_ray(root[:, 0] + scale, root[:, 1])
    down = unit_ray(root[:, 0], root[:, 1] + scale)
    world = center / center[:, 2:3]  # ray ∩ plane z = 1
    pr, pd = refresh_pixel_vectors(center, right, down, world)
    return SyntheticMap(world=world, pix_right_w=pr, pix_down_w=pd, ir_center=ir, src_level=lv,
                        center_nc=center, one_right_nc=right, one_down_nc=down)


def build_map_at_pose(cam: Camera, corners_per_level, level_dims, n_points: int, pose, seed: int = MAP_SEED + 1) -> SyntheticMap:
    """Like build_map for a source keyframe at `pose` (3x4 camera-from-world): the *_nc rays are in that keyframe's camera frame,
    world points are the rays' intersections with the scene plane z = 1 (world frame = frame of KF0).  pix_right_w / pix_down_w are
    left zero: MapPoint::RefreshPixelVectors needs the keyframe pose, so the tests take them from the compiled reference or from
    refresh_pixel_vectors_at_pose."""
    m = build_map(cam, corners_per_level, level_dims, n_points, seed=seed)
    R, t = np.asarray(pose)[:, :3], np.asarray(pose)[:, 3]
    o = -R.T @ t                                   # camera centre in the world
    d = m.center_nc @ R                            # R^T ray, row-wise
    lam = (1.0 - o[2]) / d[:, 2]
    world = o[None, :] + lam[:, None] * d
    pr, pd = refresh_pixel_vectors_at_pose(m.center_nc, m.one_right_nc, m.one_down_nc, world, pose)
    return SyntheticMap(world=world, pix_right_w=pr, pix_down_w=pd, ir_center=m.ir_center, src_level=m.src_level,
                        center_nc=m.center_nc, one_right_nc=m.one_right_nc, one_down_nc=m.one_down_nc)


def refresh_pixel_vectors_at_pose(center_nc, one_right_nc, one_down_nc, world, pose, normal=(0.0, 0.0, -1.0)):
    """MapPoint::RefreshPixelVectors (jni/MapPoint.cc:4-29) for a source keyframe at `pose` (synthetic-input helper; parity tests use
    the vectors computed by the compiled reference)."""
    R, t = np.asarray(pose)[:, :3], np.asarray(pose)[:, 3]
    nrm = np.asarray(normal, dtype=np.float64)
    cam_pts = world @ R.T + t[None, :]
    cam_height = np.abs(cam_pts @ nrm)
    c_on = center_nc * (cam_height / np.abs(center_nc @ nrm))[:, None]
    r_on = one_right_nc * (cam_height / np.abs(one_right_nc @ nrm))[:, None]
    d_on = one_down_nc * (cam_height / np.abs(one_down_nc @ nrm))[:, None]
    return (r_on - c_on) @ R, (d_on - c_on) @ R    # rotated into the world frame (R^T v, row-wise)


def stream_pose(k: int, stream: int = 0, twist=CONFIG1_TWIST) -> np.ndarray:
    """Camera pose of frame k of a synthetic sequence."""
    xi = np.asarray(twist, dtype=np.float64)
    if stream:
        rs = np.random.RandomState(1000 + stream)
        xi = xi * rs.uniform(0.5, 1.5, size=6) * rs.choice([-1.0, 1.0], size=6)
    wob = 0.01 * np.array([np.sin(0.07 * k), np.cos(0.05 * k) - 1.0, 0, 0, 0, np.sin(0.03 * k)])
    return se3_exp(k * xi / 20.0 + wob)

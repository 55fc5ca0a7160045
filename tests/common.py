"""Shared scene builders for the tests (synthetic inputs only; see visualslam_android_b200/synth.py)."""
from __future__ import annotations

import functools

import numpy as np

from visualslam_android_b200 import synth


@functools.lru_cache(maxsize=4)
def texture(size=2048):
    return synth.make_texture(size)


@functools.lru_cache(maxsize=8)
def scene(width=640, height=480, n_points=1000, tex_size=2048):
    """(cam, frame0, SyntheticMap) — KF0 rendered at the identity pose, map chosen from its corners."""
    from oracle import oraclebind

    cam = synth.Camera(width, height)
    f0 = synth.render_frame(texture(tex_size), cam, synth.IDENTITY_POSE)
    kf = oraclebind.OrcKeyFrame().make_lite(f0)
    corners = [kf.corners(l) for l in range(4)]
    dims = [kf.dims(l) for l in range(4)]
    smap = synth.build_map(cam, corners, dims, n_points)
    return cam, f0, smap


def frame_at(cam, twist, tex_size=2048):
    pose = synth.se3_exp(twist)
    return synth.render_frame(texture(tex_size), cam, pose), pose


@functools.lru_cache(maxsize=2)
def two_keyframe_scene(n0=500, n1=400):
    """A map whose points come from TWO source keyframes: KF0 at the identity and KF1 after a sideways motion.
    Returns (cam, [f0, f1], [pose0, pose1], merged SyntheticMap, src_kf ids)."""
    from oracle import oraclebind
    cam, f0, smap0 = scene(n_points=n0)
    pose1 = synth.se3_exp(np.array([0.10, 0.02, 0.01, 0.01, -0.04, 0.05]))
    f1 = synth.render_frame(texture(), cam, pose1)
    kf1 = oraclebind.OrcKeyFrame().make_lite(f1)
    smap1 = synth.build_map_at_pose(cam, [kf1.corners(l) for l in range(4)], [kf1.dims(l) for l in range(4)], n1, pose1)
    names = ("world", "pix_right_w", "pix_down_w", "ir_center", "src_level", "center_nc", "one_right_nc", "one_down_nc")
    merged = synth.SyntheticMap(**{k: np.concatenate([getattr(smap0, k), getattr(smap1, k)]) for k in names})
    src_kf = np.concatenate([np.zeros(smap0.n, dtype=np.int32), np.ones(smap1.n, dtype=np.int32)])
    return cam, [f0, f1], [synth.IDENTITY_POSE, pose1], merged, src_kf

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

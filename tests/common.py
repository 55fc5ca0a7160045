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


# ---- map files: an independent restatement of the layout documented in visualslam_android_b200/csrc/mapfile.cu --------------------
MAPFILE_MAGIC = b"VSLMAP\x00\x01"


def _fnv1a64(data: bytes) -> int:
    # FNV-1a is byte-serial; vectorise with the closed form over 8-bit lanes is not possible, so keep the files in the tests small.
    h = 1469598103934665603
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _pad8(buf: bytearray):
    while len(buf) % 8:
        buf.append(0)


def mapfile_pack(width, height, cam13, keyframes, points, reloc):
    """keyframes: list of (id, registered_index, HxW uint8); points: dict world/right/down (n,3) f64, ircenter (n,2), srclevel, srckf i32;
    reloc: (ids int32[n], poses f64[n,12])."""
    import struct
    n = len(points["srclevel"])
    out = bytearray()
    out += MAPFILE_MAGIC + struct.pack("<IIiiiiii", 1, 168, width, height, n, len(keyframes), len(reloc[0]), 0)
    out += np.asarray(cam13, dtype="<f8").tobytes() + bytes(24)
    assert len(out) == 168
    for kid, reg, img in keyframes:
        out += struct.pack("<ii", kid, reg) + np.ascontiguousarray(img, dtype=np.uint8).tobytes()
        _pad8(out)
    if n:
        for k in ("world", "right", "down"):
            out += np.ascontiguousarray(points[k], dtype="<f8").tobytes()
        for k in ("ircenter", "srclevel", "srckf"):
            out += np.ascontiguousarray(points[k], dtype="<i4").tobytes()
        _pad8(out)
    out += np.ascontiguousarray(reloc[0], dtype="<i4").tobytes()
    _pad8(out)
    out += np.ascontiguousarray(reloc[1], dtype="<f8").tobytes()
    out += struct.pack("<Q", _fnv1a64(bytes(out)))
    return bytes(out)


def mapfile_unpack(data: bytes):
    import struct
    assert data[:8] == MAPFILE_MAGIC
    ver, hb, w, h, n, nkf, nrel, _ = struct.unpack_from("<IIiiiiii", data, 8)
    assert (ver, hb) == (1, 168)
    cam13 = np.frombuffer(data, dtype="<f8", count=13, offset=40)
    pos = 168
    kfs = []
    for _k in range(nkf):
        kid, reg = struct.unpack_from("<ii", data, pos); pos += 8
        kfs.append((kid, reg, np.frombuffer(data, dtype=np.uint8, count=w * h, offset=pos).reshape(h, w))); pos += w * h
        pos += -pos % 8
    pts = {}
    if n:
        for k in ("world", "right", "down"):
            pts[k] = np.frombuffer(data, dtype="<f8", count=3 * n, offset=pos).reshape(n, 3); pos += 24 * n
        pts["ircenter"] = np.frombuffer(data, dtype="<i4", count=2 * n, offset=pos).reshape(n, 2); pos += 8 * n
        for k in ("srclevel", "srckf"):
            pts[k] = np.frombuffer(data, dtype="<i4", count=n, offset=pos); pos += 4 * n
        pos += -pos % 8
    rid = np.frombuffer(data, dtype="<i4", count=nrel, offset=pos); pos += 4 * nrel
    pos += -pos % 8
    rpose = np.frombuffer(data, dtype="<f8", count=12 * nrel, offset=pos).reshape(nrel, 12); pos += 96 * nrel
    (chk,) = struct.unpack_from("<Q", data, pos)
    assert pos + 8 == len(data), "trailing bytes"
    assert chk == _fnv1a64(data[:pos]), "checksum"
    return dict(width=w, height=h, cam13=cam13, keyframes=kfs, points=pts, reloc=(rid, rpose))


def render_sequence(cam, poses, tex_size=2048):
    """uint8 frames (len(poses), H, W) of the textured plane.  Input generation only: on a GPU box the torch renderer of bench.py is used
    (float32 ray arithmetic, so the bytes differ slightly from synth.render_frame -- every arm of a test gets the same bytes either way)."""
    poses = np.asarray(poses)
    try:
        import torch
        if torch.cuda.is_available():
            import bench
            dev = torch.device("cuda", 0)
            tex_t = torch.from_numpy(texture(tex_size).astype(np.float32)).to(dev)
            return np.concatenate([bench.render_frames_torch(tex_t, cam, poses[k0:k0 + 32], dev).cpu().numpy() for k0 in range(0, len(poses), 32)])
    except ImportError:
        pass
    return np.stack([synth.render_frame(texture(tex_size), cam, p) for p in poses])


def map_slice(smap, lo, hi):
    """Points [lo, hi) of a SyntheticMap as a map of their own."""
    names = ("world", "pix_right_w", "pix_down_w", "ir_center", "src_level", "center_nc", "one_right_nc", "one_down_nc")
    return synth.SyntheticMap(**{k: getattr(smap, k)[lo:hi] for k in names})

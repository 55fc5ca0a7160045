#!/usr/bin/env python
"""Generate tests/golden/ref_small.npz from the COMPILED REFERENCE (oracle/_ref/libvslam_ref.so, built from
/root/reference/jni by oracle/build_ref.sh).  Run in the build container:  python tests/golden/make_golden.py

The fixture pins the CPU restatement (oracle/vslam_oracle.cc) wherever oracle/_ref is not available (e.g. on the GPU box
before build, or for readers without /root/reference): tests/test_oracle_golden.py replays the same inputs through the
restatement and compares with what the reference produced here.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402
from visualslam_android_b200 import synth  # noqa: E402

W, H, N = 320, 240, 300


def _initialised_only(ints, dbl):
    """The reference leaves TrackerData fields of points outside the potentially-visible set (and v2Found of points not found)
    uninitialised: zero them so that the fixture is reproducible byte for byte.  The test compares the initialised entries only."""
    ints, dbl = ints.copy(), dbl.copy()
    pvs = ints[:, 1] >= 0
    ints[~pvs, 2:] = 0
    dbl[~(pvs & (ints[:, 3] == 1))] = 0.0
    return ints, dbl


def main():
    cam = synth.Camera(W, H)
    tex = synth.make_texture(1024)
    f0 = synth.render_frame(tex, cam, synth.IDENTITY_POSE)
    kf0 = refbind.RefKeyFrame().make_lite(f0)
    smap = synth.build_map(cam, [kf0.corners(l) for l in range(4)], [kf0.dims(l) for l in range(4)], N)
    out = {"f0": f0, "world": smap.world, "ir_center": smap.ir_center, "src_level": smap.src_level, "center_nc": smap.center_nc,
           "one_right_nc": smap.one_right_nc, "one_down_nc": smap.one_down_nc, "cam13": cam.scalars()}
    rw = refbind.RefWorld(W, H, f0, smap)
    pr, pd = rw.pixel_vectors()
    out["pix_right_w"], out["pix_down_w"] = pr, pd
    sc = np.zeros(13); rw.L.ref_cam_scalars(rw.cam, sc); out["ref_cam_scalars"] = sc

    # --- case A: MakeKeyFrame_Lite + MakeKeyFrame_Rest on frame 1
    tw = np.array(synth.CONFIG1_TWIST) * 0.4
    f1 = synth.render_frame(tex, cam, synth.se3_exp(tw))
    out["f1"] = f1
    rgba = np.repeat(f1[:, :, None], 4, axis=2).copy()
    kf = refbind.RefKeyFrame().make_lite(f1, rgba)
    kf.make_rest()
    for l in range(4):
        out[f"lvl{l}"] = kf.pixels(l); out[f"corners{l}"] = kf.corners(l); out[f"lut{l}"] = kf.row_lut(l)
        out[f"max{l}"] = kf.max_corners(l)
        xy, s = kf.candidates(l); out[f"cand{l}"] = xy; out[f"cand_score{l}"] = s

    # --- case B: TrackMap, fine stage only (start pose = 0.2 twist), then a second TrackMap with a coarse stage on frame 2
    rw.L.ref_srand(1)
    rw.make_current_kf(f1)
    start = synth.se3_exp(np.array(synth.CONFIG1_TWIST) * 0.2)
    rw.set_pose(start); out["start_pose"] = start
    rw.L.ref_tracker_track_map(rw.tracker)
    ints, dbl = rw.point_states()
    out["B_ints"], out["B_dbl"] = _initialised_only(ints[:, [0, 1, 2, 3, 5]], dbl[:, [0, 1, 2, 3, 15]])
    a, f, q, lost, dc = rw.counters()
    out["B_counters"] = np.concatenate([a, f, [dc]]); out["B_pose"] = rw.get_pose()
    f2 = synth.render_frame(tex, cam, synth.se3_exp(np.array(synth.CONFIG1_TWIST) * 0.9))
    out["f2"] = f2
    rw.make_current_kf(f2)
    rw.L.ref_tracker_set_velocity(rw.tracker, np.zeros(6), 0.05)
    rw.L.ref_tracker_track_map(rw.tracker)
    ints, dbl = rw.point_states()
    out["C_ints"], out["C_dbl"] = _initialised_only(ints[:, [0, 1, 2, 3, 5]], dbl[:, [0, 1, 2, 3, 15]])
    a, f, q, lost, dc = rw.counters()
    out["C_counters"] = np.concatenate([a, f, [dc]]); out["C_pose"] = rw.get_pose()
    cnt = np.zeros((smap.n, 2), dtype=np.int32)
    for k in range(smap.n):
        x, y = C.c_int(), C.c_int(); rw.L.ref_map_point_counts(rw.map, k, C.byref(x), C.byref(y)); cnt[k] = (x.value, y.value)
    out["C_counts"] = cnt

    # --- case D: SE3 exp / ln and Tukey known answers
    rs = np.random.RandomState(11)
    mus = np.concatenate([rs.uniform(-0.3, 0.3, (20, 6)), rs.uniform(-1e-4, 1e-4, (5, 6)), rs.uniform(-2.5, 2.5, (10, 6))])
    exps = np.zeros((len(mus), 12)); lns = np.zeros((len(mus), 6))
    for k, mu in enumerate(mus):
        rw.L.ref_se3_exp(np.ascontiguousarray(mu), exps[k]); rw.L.ref_se3_ln(exps[k], lns[k])
    out["D_mu"], out["D_exp"], out["D_ln"] = mus, exps, lns
    errs = [rs.uniform(0, 9, n) for n in (1, 2, 4, 5, 100, 1001)]
    out["D_tukey_in"] = np.concatenate(errs); out["D_tukey_n"] = np.array([len(e) for e in errs])
    out["D_tukey_out"] = np.array([rw.L.ref_tukey_sigma_squared(np.ascontiguousarray(e), len(e)) for e in errs])
    # libc rand() after srand(1): first draws
    rw.L.ref_srand(1)
    out["D_rand"] = np.array([rw.L.ref_rand() for _ in range(400)], dtype=np.int64)

    # --- case E: initial-map trail tracking (Tracker::TrailTracking_Start on f0, _Advance on three frames of a sideways motion)
    rw.make_current_kf(f0)
    out["E_start_n"] = np.array([rw.L.ref_tracker_trail_start(rw.tracker)])

    def trails():
        t = np.zeros((rw.L.ref_tracker_trail_count(rw.tracker), 4))
        if len(t):
            rw.L.ref_tracker_trails(rw.tracker, t)
        return t
    out["E_trails0"] = trails()
    for k in range(1, 4):
        fk = synth.render_frame(tex, cam, synth.se3_exp(np.array([0.02, 0.004, 0.0, 0.0, 0.0, 0.003]) * k))
        out[f"E_f{k}"] = fk
        rw.make_current_kf(fk)
        out[f"E_good{k}"] = np.array([rw.L.ref_tracker_trail_advance(rw.tracker, 100000)])
        out[f"E_trails{k}"] = trails()

    # --- case F: MapMaker::ReFind_Common's call sequence (oracle/ref_harness.cc ref_refind) on frame 2 with a slightly wrong pose
    rw.make_current_kf(f2)
    off = synth.se3_exp(np.array([0.0008, -0.0006, 0.0005, 0.0006, -0.0004, 0.0007]))
    true = synth.se3_exp(np.array(synth.CONFIG1_TWIST) * 0.9)
    kf_pose = (np.vstack([off, [0, 0, 0, 1]]) @ np.vstack([true, [0, 0, 0, 1]]))[:3]
    rw.set_pose(kf_pose); out["F_pose"] = kf_pose
    idx = np.arange(smap.n, dtype=np.int32)
    ro, rp = np.zeros((smap.n, 3), dtype=np.int32), np.zeros((smap.n, 2))
    rw.L.ref_refind(rw.tracker, idx, smap.n, 4, 8, ro, rp)
    out["F_flags"], out["F_pos"] = ro, rp

    # --- case G: the search of MapMaker::AddPointEpipolar (oracle/ref_harness.cc ref_epipolar_search): candidates of f0 searched in f3
    tw3 = np.array([0.12, 0.03, 0.02, 0.01, -0.03, 0.02])
    pose3 = synth.se3_exp(tw3)
    f3 = synth.render_frame(tex, cam, pose3)
    out["G_f3"], out["G_pose"] = f3, pose3
    rk0 = refbind.RefKeyFrame().make_lite(f0); rk0.make_rest()
    rk3 = refbind.RefKeyFrame().make_lite(f3)
    eye = np.ascontiguousarray(synth.IDENTITY_POSE, dtype=np.float64).reshape(12); p3 = np.ascontiguousarray(pose3, dtype=np.float64).reshape(12)
    rows = []
    for level in range(4):
        xy, _ = rk0.candidates(level)
        for k in range(0, len(xy), max(1, len(xy) // 40)):
            ro, rp = np.zeros(3, dtype=np.int32), np.zeros(2)
            rw.L.ref_epipolar_search(rw.tracker, rk0.h, rk3.h, eye, p3, 1.0, 0.3, 0.1, level, k, ro, rp)
            rows.append([level, xy[k, 0], xy[k, 1], ro[0], ro[1], ro[2], rp[0], rp[1]])
    out["G_rows"] = np.array(rows, dtype=np.float64)
    # new-point fields of the converged candidates' source pixels (oracle/ref_harness.cc ref_epipolar_point_fields: the reference's MapPoint / ATANCamera
    # objects), for the triangulated position and, to exercise the rotations, with the source keyframe at a non-identity pose
    src_pose = synth.se3_exp(np.array([0.02, -0.01, 0.03, 0.01, 0.02, -0.01]))
    sp = np.ascontiguousarray(src_pose, dtype=np.float64).reshape(12)
    cam13 = np.ascontiguousarray(cam.scalars(), dtype=np.float64)
    frows = []
    for level in range(4):
        xy, _ = rk0.candidates(level)
        for k in range(0, len(xy), max(1, len(xy) // 40)):
            r = [q for q in rows if q[0] == level and q[1] == xy[k, 0] and q[2] == xy[k, 1]][0]
            if not r[3]:
                continue
            root = (xy[k] + 0.5) * (1 << level) - 0.5
            ux, uy = cam.unproject(root[0], root[1])
            world = np.array([float(ux), float(uy), 1.0]) * (1.0 + 0.01 * level)       # about where the synthetic plane is
            for pose12, w in ((eye, world), (sp, world + np.array([0.05, -0.02, 0.4]))):
                ref15 = np.zeros(15)
                rw.L.ref_epipolar_point_fields(rw.tracker, rk0.h, pose12, level, k, np.ascontiguousarray(w), ref15)
                frows.append(np.concatenate([[level, xy[k, 0], xy[k, 1]], pose12, w, ref15]))
    out["G_fields"] = np.array(frows, dtype=np.float64)

    # --- case H: the unmodified Tracker::TrackFrame through a loss of tracking and two relocalisations (three map keyframes).
    # Noise frames are regenerated by the test from RandomState(5); rendered frames are stored.
    rw2 = refbind.RefWorld(W, H, f0, smap)
    kf_tw = [np.zeros(6), np.array([0.10, 0.02, 0.01, 0.01, -0.04, 0.05]), np.array([-0.08, -0.05, 0.02, -0.02, 0.03, -0.06])]
    keep = []
    for k, tw in enumerate(kf_tw):
        pose = synth.se3_exp(tw); fr = synth.render_frame(tex, cam, pose)
        out[f"H_kf{k}"], out[f"H_kfpose{k}"] = fr, pose
        if k == 0:
            rk = rw2.src_kf
        else:
            rk = refbind.RefKeyFrame().make_lite(fr); rk.set_pose(pose); rw2.L.ref_map_add_keyframe(rw2.map, rk.h)
        rw2.L.ref_kf_make_sbi(rk.h); keep.append(rk)
    rw2.set_pose(synth.IDENTITY_POSE)
    rw2.L.ref_srand(1)
    rs5 = np.random.RandomState(5)
    rend = lambda tw: synth.render_frame(tex, cam, synth.se3_exp(np.asarray(tw)))
    plan = [("r", np.array(synth.CONFIG1_TWIST) * 0.2)] + [("n", None)] * 4 + [("r", kf_tw[1] + np.array([0.004, -0.003, 0.002, 0.01, 0.008, -0.012])),
            ("r", kf_tw[1] + np.array([0.006, -0.002, 0.002, 0.012, 0.006, -0.01]))] + [("n", None)] * 4 + [("r", kf_tw[2] + np.array([-0.003, 0.004, 0.001, -0.008, 0.01, 0.009]))]
    poses, cnts, kinds, nr = [], [], [], 0
    for kind, tw in plan:
        if kind == "n":
            fr = rs5.randint(0, 255, (H, W)).astype(np.uint8)
        else:
            fr = rend(tw); out[f"H_r{nr}"] = fr; nr += 1
        kinds.append(0 if kind == "n" else 1)
        rw2.L.ref_tracker_track_frame(rw2.tracker, np.ascontiguousarray(fr), W, H, W)
        poses.append(rw2.get_pose()); a, f, q, lost, dc = rw2.counters(); cnts.append(np.concatenate([a, f, [q, lost, dc]]))
    out["H_kinds"], out["H_poses"], out["H_counters"] = np.array(kinds), np.stack(poses), np.stack(cnts)

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

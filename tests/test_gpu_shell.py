"""GPU tests of the reference-shaped C++ shell (include/vslam_b200_shell.hpp): a driver program (tests/cpp/shell_driver.cc) is built
against the stand-in cv::Mat / Eigen headers, run on the GPU through the shell's KeyFrame / Tracker / MiniPatch / PatchFinder
classes, and its output is compared with the oracle (oracle/: the CPU restatement, used here as the checker only)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import common
from visualslam_android_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = tmp_path_factory.mktemp("shell") / "shell_driver"
    libdir = os.path.join(ROOT, "visualslam_android_b200")
    cmd = ["g++", "-std=gnu++11", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "shim"),
           os.path.join(ROOT, "tests", "cpp", "shell_driver.cc"), "-o", str(exe), "-L", libdir, "-lvslam_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return str(exe)


def _write_scene(path, cam, f0, smap, pose0, frames):
    H, W = f0.shape
    n = 0 if smap is None else smap.n
    with open(path, "wb") as f:
        f.write(np.array([W, H, n, len(frames)], dtype=np.int32).tobytes())
        f.write(np.asarray(synth.CAMERA_PARAMS, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(f0, dtype=np.uint8).tobytes())
        if n:
            for a in (smap.world, smap.pix_right_w, smap.pix_down_w):
                f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
            f.write(np.ascontiguousarray(smap.ir_center, dtype=np.int32).tobytes())
            f.write(np.ascontiguousarray(smap.src_level, dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(pose0, dtype=np.float64).reshape(12).tobytes())
        for fr in frames:
            f.write(np.ascontiguousarray(fr, dtype=np.uint8).tobytes())


def _run(driver, scene, mode, *extra):
    r = subprocess.run([driver, scene, mode, *extra], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    return r.stdout.splitlines()


def test_shell_trail_tracking_matches_the_oracle(driver, tmp_path):
    """f2 host logic in the shell: Tracker::TrackForInitialMap / TrailTracking_Start / _Advance (jni/Tracker.cc:203-346) with the
    MakeKeyFrame_Rest and MiniPatch kernels underneath; trail lists after every frame equal the oracle's, which
    tests/test_oracle_vs_ref.py pins to the compiled reference."""
    from oracle import oraclebind
    cam, f0, _ = common.scene()
    tw = np.array([0.02, 0.004, 0.0, 0.0, 0.0, 0.003])
    frames = [f0] + [common.frame_at(cam, tw * k)[0] for k in range(1, 5)]
    scene = str(tmp_path / "trails.bin")
    _write_scene(scene, cam, f0, None, synth.IDENTITY_POSE, frames)
    out = _run(driver, scene, "trails")
    assert out[0].startswith("msg Point camera at planar scene and press spacebar")
    got, cur = {}, None
    for line in out[1:]:
        w = line.split()
        if w[0] == "frame":
            cur = int(w[1]); got[cur] = {"stage": int(w[3]), "n": int(w[5]), "t": []}
        elif w[0] == "t":
            got[cur]["t"].append([float(x) for x in w[1:]])
    ot = oraclebind.OrcTrails()
    ot.start(oraclebind.OrcKeyFrame().make_lite(frames[0]))
    for k in range(len(frames)):
        if k:
            good = ot.advance(oraclebind.OrcKeyFrame().make_lite(frames[k]), 100000)
            assert good >= 10
        exp = ot.trails()
        assert got[k]["n"] == len(exp) and np.array_equal(np.array(got[k]["t"]).reshape(-1, 4), exp), k
        assert got[k]["stage"] == (2 if k == len(frames) - 1 else 1)
    assert len(exp) > 30
    assert [l for l in out if l.startswith("matches")][0] == f"matches {len(exp)}"
    mp = [l for l in out if l.startswith("minipatch")][0].split()
    assert mp[1] == "1" and mp[2:4] == mp[4:6]            # a patch sampled at a corner is found again at that corner (SSD 0)


def test_shell_track_frame_matches_the_oracle(driver, tmp_path):
    """Tracker::TrackFrame / GetCurrentPose / GetMessageForUser of the shell over a short sequence, SmallBlurryImage on."""
    from oracle import oraclebind
    cam, f0, smap = common.scene(n_points=600)
    frames = [synth.render_frame(common.texture(), cam, synth.stream_pose(5 * k, 3)) for k in range(1, 5)]
    scene = str(tmp_path / "track.bin")
    _write_scene(scene, cam, f0, smap, synth.IDENTITY_POSE, frames)
    out = _run(driver, scene, "track")
    poses = [np.array([float(x) for x in l.split()[1:]]).reshape(3, 4) for l in out if l.startswith("pose")]
    msgs = [l[4:] for l in out if l.startswith("msg ")]
    ow = oraclebind.OrcWorld(cam, f0, smap)
    ow.set_pose(synth.IDENTITY_POSE)
    ow.L.orc_tracker_enable_sbi(ow.tracker, synth.Camera(cam.width // 16, cam.height // 16).scalars())
    for k, fr in enumerate(frames):
        ow.L.orc_tracker_track_frame(ow.tracker, np.ascontiguousarray(fr), cam.width, cam.height, cam.width)
        assert np.abs(poses[k] - ow.get_pose()).max() <= 1e-8, k
        a, f, q, lost, dc = ow.counters()
        exp = "Tracking Map, quality " + {2: "good.", 1: "poor.", 0: "bad."}[q] + " Found:" + "".join(f" {f[l]}/{a[l]}" for l in range(4)) + f" Map: {smap.n}P"
        assert msgs[k] == exp, (msgs[k], exp)


def test_shell_adds_keyframes_when_the_device_asks(driver, tmp_path):
    """Tracker::TrackFrame's keyframe heuristics through the shell (jni/Tracker.cc:127-132): messages, ' Adding key-frame.' on the frame
    where the restatement adds one, mnLastKeyFrameDropped."""
    from oracle import oraclebind
    cam, f0, smap = common.scene(n_points=600)
    step = np.array([0.004, 0.001, 0.0005, 0.0004, -0.0012, 0.0008])
    frames = [common.frame_at(cam, step * k)[0] for k in range(1, 9)]
    scene = str(tmp_path / "handoff.bin")
    _write_scene(scene, cam, f0, smap, synth.IDENTITY_POSE, frames)
    out = _run(driver, scene, "handoff")
    msgs = [l[4:] for l in out if l.startswith("msg ")]
    kfs = [[int(v) for v in l.split()[1:]] for l in out if l.startswith("kf ")]
    ow = oraclebind.OrcWorld(cam, f0, smap)
    ow.set_pose(synth.IDENTITY_POSE)
    ow.L.orc_tracker_enable_sbi(ow.tracker, synth.Camera(cam.width // 16, cam.height // 16).scalars())
    okf0 = oraclebind.OrcKeyFrame().make_lite(f0)
    ow.L.orc_tracker_add_reloc_keyframe(ow.tracker, okf0.h, np.ascontiguousarray(synth.IDENTITY_POSE, dtype=np.float64).reshape(12))
    ow.L.orc_tracker_set_keyframe_policy(ow.tracker, 1, 0.1, 0.1, 0.2, 20)
    n_kf, added = 1, []
    for k, fr in enumerate(frames):
        ow.L.orc_tracker_track_frame(ow.tracker, np.ascontiguousarray(fr), cam.width, cam.height, cam.width)
        a, f, q, lost, dc = ow.counters()
        v = [C.c_int() for _ in range(4)]
        ow.L.orc_tracker_keyframe_info(ow.tracker, *[C.byref(x) for x in v])
        exp = "Tracking Map, quality " + {2: "good.", 1: "poor.", 0: "bad."}[q] + " Found:" + "".join(f" {f[l]}/{a[l]}" for l in range(4)) + f" Map: {smap.n}P, {n_kf}KF"
        if v[1].value:
            exp += " Adding key-frame."; n_kf += 1; added.append(k + 1)
        assert msgs[k] == exp, (k, msgs[k], exp)
        assert kfs[k] == [v[0].value, v[3].value], (k, kfs[k])
    assert len(added) == 1 and added[0] in (5, 6)      # 0.004 per frame against 0.2 * 0.1 * scene depth; one keyframe, then the 20-frame gap


def test_shell_stage_functions_and_patchfinder_match_the_oracle(driver, tmp_path):
    """PatchFinder per object (CalcSearchLevelAndWarpMatrix, FindPatchCoarse + sub-pixel) and Tracker::SearchForPoints / CalcPoseUpdate
    on explicit point lists, through the shell."""
    from oracle import oraclebind
    cam, f0, smap = common.scene(n_points=600)
    start = synth.se3_exp(np.array(synth.CONFIG1_TWIST) * 0.9)
    f1, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST))
    scene = str(tmp_path / "stages.bin")
    _write_scene(scene, cam, f0, smap, start, [f1])
    out = _run(driver, scene, "stages")
    ow = oraclebind.OrcWorld(cam, f0, smap)
    ow.set_pose(start)
    ow.make_current_kf(f1)
    ow.L.orc_tracker_project_all(ow.tracker)
    ints, dbl = ow.point_states()
    n_pf = n_steps = 0
    for l in out:
        w = l.split()
        if w[0] != "pf":
            continue
        pt, level = int(w[1]), int(w[3])
        assert level == ints[pt, 1]
        v2 = np.array([float(w[5]), float(w[6])]); warp = np.array([float(x) for x in w[8:12]])
        if ints[pt, 0]:
            assert np.abs(v2 - dbl[pt, 0:2]).max() <= 1e-9 and np.abs(warp - dbl[pt, 11:15]).max() <= 1e-9 * max(1.0, np.abs(warp).max())
        if level >= 0:
            idx = np.array([pt], dtype=np.int32)
            ow.L.orc_tracker_clear_counters(ow.tracker)
            ow.L.orc_tracker_search_for_points(ow.tracker, idx, 1, 10, 8)
            i2, d2 = ow.point_states()
            k = w.index("bad")
            bad, found = int(w[k + 1]), int(w[k + 3])
            assert bad == i2[pt, 5] and found == i2[pt, 3], pt
            if found:
                coarse = np.array([float(w[k + 5]), float(w[k + 6])]); sub = np.array([float(w[k + 8]), float(w[k + 9])])
                assert np.array_equal(coarse, d2[pt, 30:32]) and np.abs(sub - d2[pt, 2:4]).max() <= 1e-6
                # the reference's separate steps (ZMSSDAtPoint, MakeSubPixTemplate, IterateSubPix / IterateSubPixToConvergence, SetSubPixPos)
                z = w.index("zmssd")
                assert 0 <= int(w[z + 1]) < int(w[z + 3]) and int(w[z + 5]) == 1 and int(w[z + 10]) == 1, l
                steps = np.array([float(w[z + 7]), float(w[z + 8])])
                assert np.abs(steps - sub).max() <= 1e-9, (steps, sub)
                n_steps += 1
            n_pf += 1
    assert n_pf >= 8 and n_steps >= 4
    # SearchForPoints / CalcPoseUpdate on every third point with a valid level
    ow2 = oraclebind.OrcWorld(cam, f0, smap)
    ow2.set_pose(start)
    ow2.make_current_kf(f1)
    ow2.L.orc_tracker_project_all(ow2.tracker)
    ints, _ = ow2.point_states()
    lst = np.array([p for p in range(0, smap.n, 3) if ints[p, 1] >= 0], dtype=np.int32)
    nf = ow2.L.orc_tracker_search_for_points(ow2.tracker, lst, len(lst), 12, 4)
    s = [l for l in out if l.startswith("search")][0].split()
    assert int(s[1]) == nf and int(s[3]) == len(lst) and nf > 50
    ow2.L.orc_tracker_calc_jacobians(ow2.tracker, lst, len(lst))
    mu = np.zeros(6); ow2.L.orc_tracker_calc_pose_update(ow2.tracker, lst, len(lst), 0.0, 0, 0, mu)
    got = np.array([float(x) for x in [l for l in out if l.startswith("update ")][0].split()[1:]])
    assert np.abs(got - mu).max() <= 1e-9 * max(1.0, np.abs(mu).max()), (got, mu)
    mu2 = np.zeros(6); ow2.L.orc_tracker_calc_pose_update(ow2.tracker, lst, len(lst), 16.0, 1, 0, mu2)
    got2 = np.array([float(x) for x in [l for l in out if l.startswith("update16")][0].split()[1:]])
    assert np.abs(got2 - mu2).max() <= 1e-9 * max(1.0, np.abs(mu2).max())


def test_shell_mapmaker_searches_match_the_oracle(driver, tmp_path):
    """MapSearch::ReFindInSingleKeyFrame (MapMaker::ReFind_Common for every map point) and MapSearch::AddPointsEpipolar (the search of
    MapMaker::AddPointEpipolar for every Shi-Tomasi candidate of every level) through the shell, against the oracle."""
    from oracle import oraclebind
    cam, f0, smap = common.scene(n_points=600)
    tw = np.array([0.12, 0.03, 0.02, 0.01, -0.03, 0.02])
    f1, pose1 = common.frame_at(cam, tw)
    scene = str(tmp_path / "mapsearch.bin")
    _write_scene(scene, cam, f0, smap, pose1, [f1])
    out = _run(driver, scene, "mapsearch", str(tmp_path / "shell.vsmap"))
    ow = oraclebind.OrcWorld(cam, f0, smap)
    ow.make_current_kf(f1); ow.set_pose(pose1)
    idx = np.arange(smap.n, dtype=np.int32)
    oo, op = np.zeros((smap.n, 3), dtype=np.int32), np.zeros((smap.n, 2))
    ow.L.orc_tracker_refind(ow.tracker, idx, smap.n, 4, 8, 1, oo, op)
    got = {int(l.split()[1]): l.split()[2:] for l in out if l.startswith("m ")}
    assert int([l for l in out if l.startswith("refind")][0].split()[1]) == int(oo[:, 0].sum()) == len(got) and len(got) > 100
    for pt, w in got.items():
        assert oo[pt, 0] == 1 and int(w[0]) == oo[pt, 1] and int(w[1]) == oo[pt, 2]
        assert np.abs(np.array([float(w[2]), float(w[3])]) - op[pt]).max() <= (1e-6 if oo[pt, 2] else 0.0)
    ok0 = oraclebind.OrcKeyFrame().make_lite(f0); ok0.make_rest()
    ok1 = oraclebind.OrcKeyFrame().make_lite(f1)
    eye = np.ascontiguousarray(synth.IDENTITY_POSE, dtype=np.float64).reshape(12); p1 = np.ascontiguousarray(pose1, dtype=np.float64).reshape(12)
    e_lines = [l.split() for l in out if l.startswith("e ")]
    total = 0
    for level in range(4):
        xy, _ = ok0.candidates(level)
        gl = {(int(w[2]), int(w[3])): (float(w[4]), float(w[5])) for w in e_lines if int(w[1]) == level}
        nf = 0
        for k in range(len(xy)):
            o3, o2 = np.zeros(3, dtype=np.int32), np.zeros(2)
            ow.L.orc_epipolar_search(ow.tracker, ok0.h, ok1.h, eye, p1, 1.0, 0.3, 0.1, level, int(xy[k, 0]), int(xy[k, 1]), o3, o2, None)
            key = (int(xy[k, 0]), int(xy[k, 1]))
            assert (key in gl) == bool(o3[0]), (level, k)
            if o3[0]:
                assert np.abs(np.array(gl[key]) - o2).max() <= 1e-6; nf += 1
        hdr = [l.split() for l in out if l.startswith(f"epipolar {level} ")][0]
        assert int(hdr[2]) == nf and int(hdr[4]) == len(xy)
        total += nf
    assert total > 100
    # the new map points of the converged candidates (MapMaker::AddPointEpipolar's tail) and the map file round trip
    cam13 = np.ascontiguousarray(cam.scalars(), dtype=np.float64)
    found_at = {(int(w[1]), int(w[2]), int(w[3])): (float(w[4]), float(w[5])) for w in e_lines}
    p_lines = [l.split() for l in out if l.startswith("p ")]
    assert len(p_lines) == total
    for w in p_lines[::7]:
        level, cx, cy = int(w[1]), int(w[2]), int(w[3])
        vals = np.array([float(v) for v in w[4:]]).reshape(3, 3)
        root = (np.array([cx, cy]) + 0.5) * (1 << level) - 0.5
        world = oraclebind.triangulate(cam13, synth.IDENTITY_POSE, pose1, root, np.array(found_at[(level, cx, cy)]))
        assert np.abs(vals[0] - world).max() <= 1e-9 * max(1.0, np.abs(world).max())
        fields = oraclebind.epipolar_point_fields(cam13, synth.IDENTITY_POSE, level, cx, cy, vals[0])
        assert np.abs(vals[1] - fields[3]).max() <= 1e-12 and np.abs(vals[2] - fields[4]).max() <= 1e-12
    mf = [l.split() for l in out if l.startswith("mapfile")][0]
    assert [int(v) for v in mf[1:]] == [smap.n, 1, 0, smap.n]
    back = common.mapfile_unpack((tmp_path / "shell.vsmap").read_bytes())
    assert np.array_equal(back["keyframes"][0][2], f0) and np.array_equal(back["points"]["world"], smap.world)


def test_shell_reference_shaped_tracker_reads_the_map_like_the_reference(driver, tmp_path):
    """`new Tracker(width, height, *mpCamera, *mpMap, *mpMapMaker)` as jni/jni_part.cpp:27-46 writes it, on reference-shaped types: the adapter
    flattens Map::vpPoints / vpKeyFrames itself, picks up the points a map maker pushes while tracking, and tracks like the oracle."""
    from oracle import oraclebind
    cam, f0, smap = common.scene(n_points=600)
    frames = [synth.render_frame(common.texture(), cam, synth.stream_pose(5 * k, 3)) for k in range(1, 7)]
    scene = str(tmp_path / "reftypes.bin")
    _write_scene(scene, cam, f0, smap, synth.IDENTITY_POSE, frames)
    out = _run(driver, scene, "reftypes")
    poses = [np.array([float(x) for x in l.split()[1:]]).reshape(3, 4) for l in out if l.startswith("pose")]
    msgs = [l[4:] for l in out if l.startswith("msg ")]
    n_first = smap.n - smap.n // 4
    first, rest = common.map_slice(smap, 0, n_first), common.map_slice(smap, n_first, smap.n)
    ow = oraclebind.OrcWorld(cam, f0, first)
    ow.set_pose(synth.IDENTITY_POSE)
    ow.L.orc_tracker_enable_sbi(ow.tracker, synth.Camera(cam.width // 16, cam.height // 16).scalars())
    n_map = n_first
    for k, fr in enumerate(frames):
        if k == len(frames) // 2:
            ow.append_points(rest); n_map = smap.n
        ow.L.orc_tracker_track_frame(ow.tracker, np.ascontiguousarray(fr), cam.width, cam.height, cam.width)
        assert np.abs(poses[k] - ow.get_pose()).max() <= 1e-8, k
        a, f, q, lost, dc = ow.counters()
        exp = "Tracking Map, quality " + {2: "good.", 1: "poor.", 0: "bad."}[q] + " Found:" + "".join(f" {f[l]}/{a[l]}" for l in range(4)) + f" Map: {n_map}P"
        assert msgs[k] == exp, (msgs[k], exp)
    assert out[-1] == "spacebar 1"

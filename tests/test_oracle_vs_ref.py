"""CPU: the oracle restatement against the compiled reference itself (oracle/_ref, built from /root/reference/jni by
oracle/build_ref.sh).  Skipped where that library does not exist; tests/test_oracle_golden.py covers that case."""
import ctypes as C

import numpy as np
import pytest

import common
from oracle import oraclebind, refbind
from visualslam_android_b200 import synth

pytestmark = pytest.mark.skipif(not refbind.available(), reason="oracle/_ref/libvslam_ref.so not built (needs /root/reference)")


def _worlds(P=11, **kw):
    cam, f0, smap = common.scene(**kw)
    rw = refbind.RefWorld(cam.width, cam.height, f0, smap)
    pr, pd = rw.pixel_vectors()
    ow = oraclebind.OrcWorld(cam, f0, smap, P=P, pix_right=pr, pix_down=pd)
    return cam, f0, smap, rw, ow


@pytest.mark.parametrize("kind", ["synthetic", "noise", "flat", "gradient"])
def test_make_keyframe_lite(kind):
    rs = np.random.RandomState(4)
    if kind == "synthetic":
        cam, f0, _ = common.scene()
        im, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST))
    elif kind == "noise":
        im = rs.randint(0, 256, (120, 160)).astype(np.uint8)
    elif kind == "flat":
        im = np.full((64, 96), 77, np.uint8)
    else:
        im = (np.add.outer(np.arange(72) * 3, np.arange(96) * 2) % 256).astype(np.uint8)
    rk, ok = refbind.RefKeyFrame().make_lite(im), oraclebind.OrcKeyFrame().make_lite(im)
    for l in range(4):
        assert np.array_equal(rk.pixels(l), ok.pixels(l))
        assert np.array_equal(rk.corners(l), ok.corners(l))
        assert np.array_equal(rk.row_lut(l), ok.row_lut(l))


@pytest.mark.parametrize("size", [(1920, 1080), (3840, 2160), (640, 360), (352, 272)])
def test_make_keyframe_lite_at_the_other_frame_sizes(size):
    """BASELINE's 1080p and 4K frames (and two sizes whose levels have odd or ragged dimensions) through the reference's own MakeKeyFrame_Lite:
    the GPU parity tests at those sizes compare with the restatement, which is therefore held to the reference there as well."""
    W, H = size
    cam = synth.Camera(W, H)
    im = synth.render_frame(common.texture(4096 if W > 640 else 2048), cam, synth.se3_exp(np.array(synth.CONFIG1_TWIST) * 0.3))
    rk, ok = refbind.RefKeyFrame().make_lite(im), oraclebind.OrcKeyFrame().make_lite(im)
    for l in range(4):
        assert np.array_equal(rk.pixels(l), ok.pixels(l))
        assert np.array_equal(rk.corners(l), ok.corners(l)) and len(rk.corners(l)) > 0
        assert np.array_equal(rk.row_lut(l), ok.row_lut(l))


def test_make_keyframe_rest_nonmax_and_shi_tomasi():
    cam, f0, _ = common.scene()
    im, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.7)
    rgba = np.repeat(im[:, :, None], 4, axis=2).copy()
    rk = refbind.RefKeyFrame().make_lite(im, rgba); rk.make_rest()
    ok = oraclebind.OrcKeyFrame().make_lite(im); ok.make_rest()
    for l in range(4):
        assert np.array_equal(rk.max_corners(l), ok.max_corners(l)), l
        rxy, rs_ = rk.candidates(l); oxy, os_ = ok.candidates(l)
        assert np.array_equal(rxy, oxy) and np.array_equal(rs_, os_), l


def test_camera_project_unproject_derivs():
    cam = synth.Camera(640, 480)
    L, R = oraclebind.lib(), refbind.lib()
    rc = R.ref_cam_create(640.0, 480.0, 1)
    rs = np.random.RandomState(0)
    for _ in range(300):
        p = rs.uniform(-0.9, 0.9, 2)
        ri, oi_, rd, od = np.zeros(2), np.zeros(2), np.zeros(4), np.zeros(4)
        rinv, oinv = C.c_int(), C.c_int()
        R.ref_cam_project(rc, p, ri, C.byref(rinv), rd); L.orc_cam_project(cam.scalars(), p, oi_, C.byref(oinv), od)
        assert np.array_equal(ri, oi_) and np.array_equal(rd, od) and rinv.value == oinv.value
        ru, ou = np.zeros(2), np.zeros(2)
        R.ref_cam_unproject(rc, ri, ru); L.orc_cam_unproject(cam.scalars(), oi_, ou)
        assert np.array_equal(ru, ou)


@pytest.mark.parametrize("P", [11, 8])
def test_stagewise_project_search_pose(P):
    cam, f0, smap, rw, ow = _worlds(P=P) if P == 11 else _worlds_p8()
    f1, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.6)
    start = synth.se3_exp(np.array(synth.CONFIG1_TWIST) * 0.3)
    rw.make_current_kf(f1); ow.make_current_kf(f1)
    rw.set_pose(start); ow.set_pose(start)
    rw.L.ref_tracker_project_all(rw.tracker); ow.L.orc_tracker_project_all(ow.tracker)
    ri, rd = rw.point_states(); oi, od = ow.point_states()
    assert np.array_equal(ri[:, :2], oi[:, :2])
    vis = oi[:, 0] == 1
    cols = [0, 1] + list(range(4, 15))     # (v2Found, columns 2-3, is uninitialised memory in the reference before a search)
    assert np.array_equal(rd[vis][:, cols], od[vis][:, cols]), "projection, derivs, v3Cam, warp: bit for bit"
    idx = np.nonzero(oi[:, 1] >= 0)[0].astype(np.int32)
    for rng, sub in ((10, 0), (30, 8)):
        rw.L.ref_tracker_clear_counters(rw.tracker); ow.L.orc_tracker_clear_counters(ow.tracker)
        nr = rw.L.ref_tracker_search_for_points(rw.tracker, idx, len(idx), rng, sub)
        no = ow.L.orc_tracker_search_for_points(ow.tracker, idx, len(idx), rng, sub)
        assert nr == no
        ri, rd = rw.point_states(); oi, od = ow.point_states()
        assert np.array_equal(ri[idx][:, [2, 3, 5]], oi[idx][:, [2, 3, 5]])
        fnd = idx[oi[idx][:, 3] == 1]
        assert np.array_equal(rd[fnd][:, [2, 3, 15, 30, 31]], od[fnd][:, [2, 3, 15, 30, 31]]), "found / coarse positions: bit for bit"
        assert np.array_equal(rw.counters()[0], ow.counters()[0]) and np.array_equal(rw.counters()[1], ow.counters()[1])
        for k in idx[::37]:
            t = np.zeros(P * P, np.uint8); s, q = C.c_int(), C.c_int()
            rw.L.ref_tracker_point_template(rw.tracker, int(k), t, C.byref(s), C.byref(q))
            ot, os_, oq = ow.point_template(int(k))
            assert np.array_equal(t.reshape(P, P), ot) and (s.value, q.value) == (os_, oq)
    rw.L.ref_tracker_calc_jacobians(rw.tracker, idx, len(idx)); ow.L.orc_tracker_calc_jacobians(ow.tracker, idx, len(idx))
    for sigma, mark in ((0.0, 0), (16.0, 1)):
        ru, ou = np.zeros(6), np.zeros(6)
        rw.L.ref_tracker_calc_pose_update(rw.tracker, idx, len(idx), sigma, mark, 1, ru)
        ow.L.orc_tracker_calc_pose_update(ow.tracker, idx, len(idx), sigma, mark, 1, ou)
        assert np.array_equal(ru, ou), "CalcPoseUpdate 6-vector: bit for bit"
        assert np.array_equal(rw.get_pose(), ow.get_pose())


def _worlds_p8():
    pytest.skip("the reference's Tracker hard-wires PatchFinder(11) (jni/TrackerData.h:44); P=8 is pinned per object in test_patchfinder_p8")


def test_patchfinder_p8_object_level():
    """PatchFinder(8) one object at a time (the 8x8 configuration of north_star) against the restatement's stand-alone entry points."""
    cam, f0, smap, rw, ow = _worlds()
    f1, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.5)
    rk = rw.make_current_kf(f1); okf = ow.make_current_kf(f1)
    rw.set_pose(synth.IDENTITY_POSE); ow.set_pose(synth.IDENTITY_POSE)
    rw.L.ref_tracker_project_all(rw.tracker)
    ri, rd = rw.point_states()
    R, L = rw.L, ow.L
    pf = R.ref_pf_create(8)
    assert R.ref_pf_max_ssd(pf) == 32000
    n_checked = 0
    for k in np.nonzero(ri[:, 1] >= 0)[0][::9]:
        level = int(ri[k, 1]); winv = np.ascontiguousarray(rd[k, 11:15])
        R.ref_pf_set_level_warp(pf, level, winv)
        t = np.zeros(64, np.uint8); s, q = C.c_int(), C.c_int()
        bad = R.ref_pf_make_template(pf, rw.map, int(k), t, C.byref(s), C.byref(q))
        det = winv[0] * winv[3] - winv[1] * winv[2]
        m2 = np.array([winv[3] / det, -winv[1] / det, -winv[2] / det, winv[0] / det])
        invdet = 1.0 / det
        m2 = np.array([(winv[3] * invdet) * (1 << level), (-winv[1] * invdet) * (1 << level), (-winv[2] * invdet) * (1 << level), (winv[0] * invdet) * (1 << level)])
        ot = np.zeros(64, np.uint8); os_, oq = C.c_int(), C.c_int()
        nout = L.orc_make_template(ow.src_kf.h, int(smap.src_level[k]), np.ascontiguousarray(smap.ir_center[k]), 8, m2, ot, C.byref(os_), C.byref(oq))
        assert (nout != 0) == bool(bad)
        assert np.array_equal(t, ot) and (s.value, q.value) == (os_.value, oq.value)
        if bad:
            continue
        pos_r, pos_o = np.zeros(2), np.zeros(2); best = C.c_int(); ev = C.c_long()
        fr = R.ref_pf_find_coarse(pf, rd[k, 0], rd[k, 1], rk.h, 12, pos_r)
        fo = L.orc_find_patch_coarse(okf.h, level, ot, 8, rd[k, 0], rd[k, 1], 12, pos_o, C.byref(best), C.byref(ev))
        assert fr == fo and (not fr or np.array_equal(pos_r, pos_o))
        if fr:
            sp_r, sp_o = np.zeros(2), np.zeros(2)
            cr = R.ref_pf_subpix(pf, rk.h, pos_r, 8, sp_r, None); co = L.orc_subpix(okf.h, level, ot, 8, pos_o, 8, sp_o, None)
            assert cr == co and np.array_equal(sp_r, sp_o)
            cx, cy = int((pos_r[0] + 0.5) / (1 << level)), int((pos_r[1] + 0.5) / (1 << level))
            assert R.ref_pf_zmssd(pf, rk.h, level, cx, cy) == L.orc_zmssd(okf.h, level, ot, 8, cx, cy) == best.value
        n_checked += 1
    assert n_checked > 40
    R.ref_pf_destroy(pf)


@pytest.mark.parametrize("start,frame,vel,n_points", [(0.0, 1.0, None, 1000), (0.0, 1.0, 0.05, 1000), (0.2, 0.5, None, 1000), (0.5, 0.5, 0.02, 1000),
                                                      (0.0, 0.6, None, 2500), (0.0, 0.6, 0.05, 2500),    # 2500: beyond the 1000-patch cap (jni/Tracker.cc:518-527)
                                                      (0.0, 0.6, None, 5000),                            # 5000: the GPU's large-map shuffle test runs against the oracle at this size
                                                      (0.0, 0.6, 0.05, 150), (0.0, 0.6, 0.05, 300), (0.0, 0.6, 0.05, 90)])   # small maps: the coarse-set branches of :425-462, incl. the vNextToSearch overwrite
def test_track_map_whole(start, frame, vel, n_points):
    cam, f0, smap, rw, ow = _worlds(n_points=n_points)
    f1, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * frame)
    sp = synth.se3_exp(np.array(synth.CONFIG1_TWIST) * start)
    rw.make_current_kf(f1); ow.make_current_kf(f1)
    rw.set_pose(sp); ow.set_pose(sp)
    if vel is not None:
        rw.L.ref_tracker_set_velocity(rw.tracker, np.zeros(6), vel); ow.L.orc_tracker_set_velocity(ow.tracker, np.zeros(6), vel)
    rw.L.ref_srand(1)
    rw.L.ref_tracker_track_map(rw.tracker); ow.L.orc_tracker_track_map(ow.tracker)
    assert np.array_equal(rw.get_pose(), ow.get_pose())
    ra, rf, _, _, rdc = rw.counters(); oa, of, _, _, odc = ow.counters()
    assert np.array_equal(ra, oa) and np.array_equal(rf, of) and rdc == odc
    ri, rd = rw.point_states(); oi, od = ow.point_states()
    pvs = oi[:, 1] >= 0
    assert np.array_equal(ri[pvs][:, [0, 1, 2, 3]], oi[pvs][:, [0, 1, 2, 3]])
    srch = oi[:, 2] == 1     # PatchFinder::mbTemplateBad is uninitialised in the reference until the point is searched for the first time
    assert np.array_equal(ri[srch][:, 5], oi[srch][:, 5])
    if n_points > 1000:
        assert srch.sum() == 1000 and pvs.sum() > 1000       # the cap was hit
    fnd = oi[:, 3] == 1
    assert np.array_equal(rd[fnd][:, :4], od[fnd][:, :4])


@pytest.mark.parametrize("vel", [None, 0.05])
def test_track_map_whole_1080p_12000_points(vel):
    """1920 x 1080 with 12000 map points: the level-3 list alone (1200 points) exceeds MaxPatchesPerFrame, so every one of its points is searched and the
    fifth shuffle's result is thrown away whole (jni/Tracker.cc:499-527) -- the branch the GPU's large-map test
    (tests/test_gpu_parity.py::test_track_map_large_maps_shuffle_without_the_swap_chain) holds against the restatement, here held against the reference."""
    cam, f0, smap, rw, ow = _worlds(width=1920, height=1080, n_points=12000, tex_size=4096)
    assert smap.n == 12000
    f1, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.6, 4096)
    rw.make_current_kf(f1); ow.make_current_kf(f1)
    rw.set_pose(synth.IDENTITY_POSE); ow.set_pose(synth.IDENTITY_POSE)
    if vel is not None:
        rw.L.ref_tracker_set_velocity(rw.tracker, np.zeros(6), vel); ow.L.orc_tracker_set_velocity(ow.tracker, np.zeros(6), vel)
    rw.L.ref_srand(1)
    rw.L.ref_tracker_track_map(rw.tracker); ow.L.orc_tracker_track_map(ow.tracker)
    assert np.array_equal(rw.get_pose(), ow.get_pose())
    ra, rf, _, _, rdc = rw.counters(); oa, of, _, _, odc = ow.counters()
    assert np.array_equal(ra, oa) and np.array_equal(rf, of) and rdc == odc and rdc == (1 if vel else 0)
    assert oa[3] >= 1000 and oa[:3].sum() <= (120 if vel else 0)       # level 3 in full; the lower levels only through the coarse set
    ri, rd = rw.point_states(); oi, od = ow.point_states()
    pvs = oi[:, 1] >= 0
    assert np.array_equal(ri[pvs][:, [0, 1, 2, 3]], oi[pvs][:, [0, 1, 2, 3]])
    fnd = oi[:, 3] == 1
    assert np.array_equal(rd[fnd][:, :4], od[fnd][:, :4])


def _two_kf_worlds():
    """Reference and restatement worlds for a map with two source keyframes (MapPoint::pPatchSourceKF differs between points)."""
    cam, kf_frames, kf_poses, smap, src_kf = common.two_keyframe_scene()
    n0 = int((src_kf == 0).sum())
    first = synth.SyntheticMap(**{k: getattr(smap, k)[:n0] for k in ("world", "pix_right_w", "pix_down_w", "ir_center", "src_level", "center_nc", "one_right_nc", "one_down_nc")})
    rw = refbind.RefWorld(cam.width, cam.height, kf_frames[0], first)
    rk1 = refbind.RefKeyFrame().make_lite(kf_frames[1]); rk1.set_pose(kf_poses[1])
    rw.L.ref_map_add_keyframe(rw.map, rk1.h)
    normal = np.array([0.0, 0.0, -1.0])
    for k in range(n0, smap.n):
        rw.L.ref_map_add_point(rw.map, rk1.h, int(smap.src_level[k]), smap.ir_center[k].astype(np.float64), np.ascontiguousarray(smap.world[k]),
                               np.ascontiguousarray(smap.center_nc[k]), np.ascontiguousarray(smap.one_right_nc[k]), np.ascontiguousarray(smap.one_down_nc[k]), normal)
    rw.n = smap.n
    pr, pd = rw.pixel_vectors()
    assert np.abs(pr - smap.pix_right_w).max() < 1e-9 and np.abs(pd - smap.pix_down_w).max() < 1e-9     # RefreshPixelVectors with a keyframe pose
    ow = oraclebind.OrcWorld(cam, kf_frames[0], smap, pix_right=pr, pix_down=pd)
    okf1 = oraclebind.OrcKeyFrame().make_lite(kf_frames[1])
    for k in range(n0, smap.n):
        ow.L.orc_tracker_set_point_source_kf(ow.tracker, k, okf1.h)
    ow._kf1, rw._kf1 = okf1, rk1
    return cam, rw, ow, smap, src_kf, pr, pd


def test_track_map_with_two_source_keyframes():
    """Templates come from the keyframe a point was made in (MapPoint::pPatchSourceKF): a map with points from two keyframes,
    tracked from a third viewpoint, reference against restatement, bit for bit."""
    cam, rw, ow, smap, src_kf, _, _ = _two_kf_worlds()
    tw = np.array([0.05, 0.01, 0.01, 0.005, -0.02, 0.03])
    fr, pose = common.frame_at(cam, tw)
    sp = synth.se3_exp(tw * 0.8)
    rw.make_current_kf(fr); ow.make_current_kf(fr)
    rw.set_pose(sp); ow.set_pose(sp)
    rw.L.ref_srand(1)
    rw.L.ref_tracker_track_map(rw.tracker); ow.L.orc_tracker_track_map(ow.tracker)
    assert np.array_equal(rw.get_pose(), ow.get_pose())
    ra, rf, _, _, _ = rw.counters(); oa, of, _, _, _ = ow.counters()
    assert np.array_equal(ra, oa) and np.array_equal(rf, of)
    ri, rd = rw.point_states(); oi, od = ow.point_states()
    pvs = oi[:, 1] >= 0      # (TrackerData::bFound is uninitialised in the reference for points that never entered the PVS)
    fnd = (oi[:, 3] == 1) & pvs
    assert np.array_equal(ri[pvs][:, [0, 1, 2, 3]], oi[pvs][:, [0, 1, 2, 3]]) and np.array_equal(rd[fnd][:, :4], od[fnd][:, :4])
    assert fnd[src_kf == 0].sum() > 200 and fnd[src_kf == 1].sum() > 200        # points of both keyframes are found
    assert np.abs(ow.get_pose() - pose).max() < 5e-3


def test_track_frame_sequence_no_sbi():
    cam, f0, smap, rw, ow = _worlds()
    rw.L.ref_srand(1)
    rw.L.ref_tracker_set_sbi_rot(rw.tracker, np.zeros(6), 1)
    for k in range(1, 9):
        fr = synth.render_frame(common.texture(), cam, synth.stream_pose(5 * k, 3))
        rw.L.ref_tracker_track_frame_nosbi(rw.tracker, fr, cam.width, cam.height, cam.width)
        ow.L.orc_tracker_track_frame(ow.tracker, fr, cam.width, cam.height, cam.width)
        assert np.array_equal(rw.get_pose(), ow.get_pose()), k
        assert all(np.array_equal(a, b) if isinstance(a, np.ndarray) else a == b for a, b in zip(rw.counters(), ow.counters())), k
        rv, ov = np.zeros(6), np.zeros(6); rm, om = C.c_double(), C.c_double()
        rw.L.ref_tracker_get_velocity(rw.tracker, rv, C.byref(rm)); ow.L.orc_tracker_get_velocity(ow.tracker, ov, C.byref(om))
        assert np.array_equal(rv, ov) and rm.value == om.value


def test_minipatch_find():
    cam, f0, smap, rw, ow = _worlds()
    f1, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.15)
    rk0 = refbind.RefKeyFrame().make_lite(f0); rk1 = refbind.RefKeyFrame().make_lite(f1)
    ok0 = oraclebind.OrcKeyFrame().make_lite(f0); ok1 = oraclebind.OrcKeyFrame().make_lite(f1)
    R, L = rw.L, ow.L
    R.ref_mp_set_max_ssd(100000)
    mp = R.ref_mp_create()
    c = ok0.corners(0)
    c = c[(c[:, 0] > 10) & (c[:, 1] > 10) & (c[:, 0] < 630) & (c[:, 1] < 470)][::41]
    nfound = 0
    for x, y in c:
        R.ref_mp_sample(mp, rk0.h, int(x), int(y), None)
        for use_lut in (0, 1):
            pr, po = np.array([x, y], float), np.array([x, y], float); best = C.c_int()
            fr = R.ref_mp_find(mp, rk1.h, pr, 10, use_lut)
            fo = L.orc_minipatch_find(ok0.h, int(x), int(y), ok1.h, po, 10, use_lut, 100000, C.byref(best))
            assert fr == fo and np.array_equal(pr, po)
            nfound += fr
    assert nfound > 20
    R.ref_mp_destroy(mp)


def test_trail_tracking_start_and_advance():
    """f2 host logic: Tracker::TrailTracking_Start / _Advance (jni/Tracker.cc:264-346) of the compiled reference against the
    restatement, over a short sideways motion: same trails (initial and current positions, list order) after every frame."""
    cam, f0, smap, rw, ow = _worlds()
    R = rw.L
    tw = np.array([0.02, 0.004, 0.0, 0.0, 0.0, 0.003])
    rw.make_current_kf(f0)
    n_ref = R.ref_tracker_trail_start(rw.tracker)
    ot = oraclebind.OrcTrails()
    n_orc = ot.start(oraclebind.OrcKeyFrame().make_lite(f0))
    assert n_ref == n_orc and n_ref > 100

    def ref_trails():
        out = np.zeros((R.ref_tracker_trail_count(rw.tracker), 4))
        if len(out):
            R.ref_tracker_trails(rw.tracker, out)
        return out
    assert np.array_equal(ref_trails(), ot.trails())
    counts = []
    for k in range(1, 5):
        f, _ = common.frame_at(cam, tw * k)
        rw.make_current_kf(f)
        g_ref = R.ref_tracker_trail_advance(rw.tracker, 100000)
        g_orc = ot.advance(oraclebind.OrcKeyFrame().make_lite(f), 100000)
        assert g_ref == g_orc
        tr, to = ref_trails(), ot.trails()
        assert np.array_equal(tr, to)
        counts.append(len(tr))
    assert counts[-1] > 30 and counts[-1] < n_ref      # some trails die, many survive
    assert np.abs(to[:, 2:] - to[:, :2]).max() > 3      # and they moved


@pytest.mark.parametrize("case", ["near", "far_from_keyframe", "very_close"])
def test_refind_common(case):
    """f3: MapMaker::ReFind_Common (jni/MapMaker.cc:967-1036) — the reference's PatchFinder / camera calls in that function's order
    (oracle/ref_harness.cc ref_refind) against the restatement, for every map point in a keyframe whose pose is slightly off.
    `very_close` puts the camera so near the plane that warps are rejected (det > 3 at level 3): the function then still builds the
    template at the level the loop reached, which the restatement reproduces."""
    cam, f0, smap, rw, ow = _worlds()
    tw = {"near": np.array(synth.CONFIG1_TWIST), "far_from_keyframe": np.array([0.05, -0.03, 0.45, 0.02, -0.03, 0.3]),
          "very_close": np.array([0.0, 0.0, -0.935, 0.0, 0.0, 0.0])}[case]
    frame, pose = common.frame_at(cam, tw)
    off = synth.se3_exp(np.array([0.0008, -0.0006, 0.0005, 0.0006, -0.0004, 0.0007]))
    kf_pose = (np.vstack([off, [0, 0, 0, 1]]) @ np.vstack([pose, [0, 0, 0, 1]]))[:3]
    rw.make_current_kf(frame); ow.make_current_kf(frame)
    rw.set_pose(kf_pose); ow.set_pose(kf_pose)
    idx = np.arange(smap.n, dtype=np.int32)
    ro, rp = np.zeros((smap.n, 3), dtype=np.int32), np.zeros((smap.n, 2))
    oo, op = np.zeros((smap.n, 3), dtype=np.int32), np.zeros((smap.n, 2))
    rw.L.ref_refind(rw.tracker, idx, smap.n, 4, 8, ro, rp)
    ow.L.orc_tracker_refind(ow.tracker, idx, smap.n, 4, 8, 0, oo, op)
    assert np.array_equal(ro, oo)
    assert np.array_equal(rp, op)
    if case == "very_close":
        assert (ro[:, 1] == 3).sum() > 5                  # points whose warp was rejected at level 3 still went through the template step
    else:
        assert ro[:, 0].sum() > 300 and (ro[:, 2] == 1).sum() > 50 and ((ro[:, 0] == 1) & (ro[:, 2] == 0)).sum() > 50
    # a cold finder per point gives the same answers when consecutive points differ (the batched GPU semantics)
    ow.L.orc_tracker_refind(ow.tracker, idx, smap.n, 4, 8, 1, oo, op)
    assert np.array_equal(ro, oo) and np.array_equal(rp, op)
    # ... and the reference's OWN MapMaker::ReFind_Common (jni/MapMaker.cc is compiled into oracle/_ref), called point by point: what it
    # returns and the Measurement it files (level, sub-pixel flag, root position) are the restatement's, bit for bit
    mo, mp = np.zeros((smap.n, 4), dtype=np.int32), np.zeros((smap.n, 2))
    rw.L.ref_mm_refind(rw.tracker, idx, smap.n, mo, mp)
    assert np.array_equal(mo[:, 0], oo[:, 0])
    f = oo[:, 0] == 1
    assert np.array_equal(mo[f][:, 1:3], oo[f][:, 1:3]) and np.array_equal(mp[f], op[f])
    assert np.all(mo[~f][:, 3] == 1)                     # everything it did not find is filed under "never retry"


def test_epipolar_search():
    """f3: the search of MapMaker::AddPointEpipolar (jni/MapMaker.cc:525-640) — call sequence on the reference's objects
    (ref_epipolar_search) against the restatement, for the Shi-Tomasi candidates of every level of one keyframe searched in a
    second keyframe taken after a sideways motion."""
    cam, f0, smap, rw, ow = _worlds()
    tw = np.array([0.12, 0.03, 0.02, 0.01, -0.03, 0.02])
    f1, pose1 = common.frame_at(cam, tw)
    rk0 = refbind.RefKeyFrame().make_lite(f0); rk0.make_rest()
    rk1 = refbind.RefKeyFrame().make_lite(f1)
    ok0 = oraclebind.OrcKeyFrame().make_lite(f0); ok0.make_rest()
    ok1 = oraclebind.OrcKeyFrame().make_lite(f1)
    eye = np.ascontiguousarray(synth.IDENTITY_POSE, dtype=np.float64).reshape(12); p1 = np.ascontiguousarray(pose1, dtype=np.float64).reshape(12)
    cam13 = np.ascontiguousarray(cam.scalars(), dtype=np.float64)
    nfound = nbest = n_points = 0
    for level in range(4):
        xy, _ = ok0.candidates(level)
        assert rw.L.ref_kf_num_candidates_l(rk0.h, level) == len(xy)
        for k in range(0, len(xy), max(1, len(xy) // 60)):
            ro, rp = np.zeros(3, dtype=np.int32), np.zeros(2)
            oo, op = np.zeros(3, dtype=np.int32), np.zeros(2)
            for mean, sigma, wig in ((1.0, 0.3, 0.1), (1.4, 0.2, 0.1)):
                rw.L.ref_epipolar_search(rw.tracker, rk0.h, rk1.h, eye, p1, mean, sigma, wig, level, k, ro, rp)
                ow.L.orc_epipolar_search(ow.tracker, ok0.h, ok1.h, eye, p1, mean, sigma, wig, level, int(xy[k, 0]), int(xy[k, 1]), oo, op, None)
                assert np.array_equal(ro, oo) and np.array_equal(rp, op), (level, k, ro, oo, rp, op)
                nfound += int(ro[0]); nbest += int(ro[1] >= 0)
                # the reference's OWN MapMaker::AddPointEpipolar (jni/MapMaker.cc:525-703): succeeds exactly when the search converges, files
                # the refined position as the target measurement, and builds the new MapPoint the restatement builds -- source rays and
                # pixel vectors bit for bit, the triangulated position (MapMaker::ReprojectPoint through the stand-in JacobiSVD) to 1e-9
                out = np.zeros(27)
                ok = rw.L.ref_mm_add_point_epipolar(rw.tracker, rk0.h, rk1.h, eye, p1, mean, sigma, wig, level, k, out)
                assert ok == int(oo[0]), (level, k)
                if ok:
                    assert np.array_equal(out[23:25], op) and np.array_equal(out[18:21], [xy[k, 0], xy[k, 1], level])
                    root = (xy[k] + 0.5) * (1 << level) - 0.5
                    assert np.array_equal(out[21:23], root)
                    world = oraclebind.triangulate(cam13, synth.IDENTITY_POSE, pose1, root, op)
                    assert np.abs(out[0:3] - world).max() <= 1e-9 * max(1.0, np.abs(world).max()), (level, k, out[0:3], world)
                    got = oraclebind.epipolar_point_fields(cam13, synth.IDENTITY_POSE, level, xy[k, 0], xy[k, 1], out[0:3])
                    assert np.array_equal(got.reshape(15), out[3:18]), (level, k)
                    n_points += 1
    assert nfound > 40 and nbest > nfound and n_points > 40


def test_track_frame_recovers_a_lost_tracker():
    """f4: the lost branch of Tracker::TrackFrame (jni/Tracker.cc:134-140) — Relocaliser::AttemptRecovery (ScoreKFs over the map
    keyframes' SmallBlurryImages, ESM alignment to the best one, SE3fromSE2), then TrackMap with the doubled coarse stage and
    AssessTrackingQuality — of the UNMODIFIED reference TrackFrame against the restatement, bit for bit.  Three map keyframes;
    the tracker is blinded for three frames, then shown frames near the second and the third keyframe."""
    cam, f0, smap, rw, ow = _worlds()
    sbi_cam = synth.Camera(cam.width // 16, cam.height // 16).scalars()
    ow.L.orc_tracker_enable_sbi(ow.tracker, sbi_cam)
    kf_twists = [np.zeros(6), np.array([0.10, 0.02, 0.01, 0.01, -0.04, 0.05]), np.array([-0.08, -0.05, 0.02, -0.02, 0.03, -0.06])]
    keep = []
    for k, tw in enumerate(kf_twists):
        fr, pose = common.frame_at(cam, tw)
        if k == 0:
            rk = rw.src_kf
        else:
            rk = refbind.RefKeyFrame().make_lite(fr); rk.set_pose(pose); rw.L.ref_map_add_keyframe(rw.map, rk.h)
        rw.L.ref_kf_make_sbi(rk.h)
        okf = oraclebind.OrcKeyFrame().make_lite(fr)
        ow.L.orc_tracker_add_reloc_keyframe(ow.tracker, okf.h, np.ascontiguousarray(pose, dtype=np.float64).reshape(12))
        keep += [rk, okf]
    rw.set_pose(synth.IDENTITY_POSE); ow.set_pose(synth.IDENTITY_POSE)
    rw.L.ref_srand(1)            # the reference shuffles with the process-wide rand(); the restatement's copy starts at seed 1
    rs = np.random.RandomState(5)
    seq = [common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.2)[0]]
    seq += [rs.randint(0, 255, f0.shape).astype(np.uint8) for _ in range(4)]                                   # noise: quality BAD, then lost
    seq += [common.frame_at(cam, kf_twists[1] + np.array([0.004, -0.003, 0.002, 0.01, 0.008, -0.012]))[0]]     # near keyframe 1
    seq += [common.frame_at(cam, kf_twists[1] + np.array([0.006, -0.002, 0.002, 0.012, 0.006, -0.01]))[0]]
    seq += [rs.randint(0, 255, f0.shape).astype(np.uint8) for _ in range(4)]
    seq += [common.frame_at(cam, kf_twists[2] + np.array([-0.003, 0.004, 0.001, -0.008, 0.01, 0.009]))[0]]     # near keyframe 2
    lost_seen = recovered = 0
    for k, fr in enumerate(seq):
        fr = np.ascontiguousarray(fr)
        rw.L.ref_tracker_track_frame(rw.tracker, fr, cam.width, cam.height, cam.width)
        ow.L.orc_tracker_track_frame(ow.tracker, fr, cam.width, cam.height, cam.width)
        assert np.array_equal(rw.get_pose(), ow.get_pose()), k
        a, f, q, lost, dc = rw.counters(); oa, of, oq, olost, odc = ow.counters()
        assert np.array_equal(a, oa) and np.array_equal(f, of) and (q, lost, dc) == (oq, olost, odc), k
        lost_seen += lost >= 3
        recovered += (k in (5, 11)) and q == 2 and lost == 0
    best, score, nrec = C.c_int(), C.c_double(), C.c_int()
    ow.L.orc_tracker_reloc_info(ow.tracker, C.byref(best), C.byref(score), C.byref(nrec))
    assert lost_seen >= 2 and recovered == 2 and nrec.value >= 2 and best.value == 2


def test_small_blurry_image_pieces():
    """f1: SmallBlurryImage::MakeFromKF, IteratePosRelToTarget and SE3fromSE2 (jni/SmallBlurryImage.cc) — restatement vs compiled reference."""
    cam, f0, smap = common.scene()
    R, L = refbind.lib(), oraclebind.lib()
    R.ref_sbi_reset_size()
    sbi_cam = synth.Camera(cam.width // 16, cam.height // 16)
    rc = R.ref_cam_create(float(cam.width), float(cam.height), 1)
    fa, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.3)
    fb, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.5)
    rka, rkb = refbind.RefKeyFrame().make_lite(fa), refbind.RefKeyFrame().make_lite(fb)
    oka, okb = oraclebind.OrcKeyFrame().make_lite(fa), oraclebind.OrcKeyFrame().make_lite(fb)
    ra, rb = R.ref_sbi_create(rka.h, 0.75), R.ref_sbi_create(rkb.h, 0.75)
    oa, ob = L.orc_sbi_create(oka.h, 0.75), L.orc_sbi_create(okb.h, 0.75)
    w, h = C.c_int(), C.c_int(); L.orc_sbi_dims(oa, C.byref(w), C.byref(h))
    assert (w.value, h.value) == (40, 30)
    for r_, o_ in ((ra, oa), (rb, ob)):
        rt, ot = np.zeros((30, 40), np.float32), np.zeros((30, 40), np.float32)
        rs_, os_ = np.zeros((30, 40), np.uint8), np.zeros((30, 40), np.uint8)
        R.ref_sbi_template(r_, rt); L.orc_sbi_template(o_, ot); R.ref_sbi_small(r_, rs_); L.orc_sbi_small(o_, os_)
        assert np.array_equal(rs_, os_) and np.array_equal(rt, ot), "mimSmall / mimTemplate bit for bit"
    rv, ov, rse2, ose2 = np.zeros(6), np.zeros(6), np.zeros(3), np.zeros(3)
    rscore = R.ref_sbi_rotation(rb, ra, rc, 6, rse2.ctypes.data, rv)
    oscore = L.orc_sbi_rotation(ob, oa, sbi_cam.scalars(), 6, ose2.ctypes.data, ov)
    assert np.array_equal(rse2, ose2) and rscore == oscore, "ESM SE2 alignment: bit for bit"
    assert np.array_equal(rv, ov), "SE3fromSE2(...).ln(): bit for bit"
    assert np.abs(rv[3:]).max() > 1e-4       # a real rotation was estimated


def test_gaussian_blur_stand_in_is_close_to_opencv():
    """The float GaussianBlur of the OpenCV stand-in (third-party arithmetic of the SBI path) against cv2 4.13."""
    cv2 = pytest.importorskip("cv2")
    cam, f0, smap = common.scene()
    L = oraclebind.lib()
    ok = oraclebind.OrcKeyFrame().make_lite(f0)
    o = L.orc_sbi_create(ok.h, 0.75)
    ot, os_ = np.zeros((30, 40), np.float32), np.zeros((30, 40), np.uint8)
    L.orc_sbi_template(o, ot); L.orc_sbi_small(o, os_)
    src = os_.astype(np.float32) - np.float32(os_.sum(dtype=np.uint32)) / np.float32(1200)
    want = cv2.GaussianBlur(src, (9, 9), 0.75, sigmaY=0.75, borderType=cv2.BORDER_REPLICATE)
    assert np.abs(want - ot).max() < 2e-4, np.abs(want - ot).max()


@pytest.mark.parametrize("size", [(640, 480), (640, 360)])
def test_track_frame_with_sbi_is_the_reference_trackframe(size):
    """The unmodified Tracker::TrackFrame (SmallBlurryImage included) against the restatement with its on-board SBI.  640 x 360: level 3 is
    80 x 45, so SmallBlurryImage::MakeFromKF's cv::resize is the general bilinear one (the stand-in OpenCV and the restatement share
    oracle/shim/cv_resize_linear_u8.h, which tests/test_oracle_golden.py pins to the real cv2)."""
    cam, f0, smap, rw, ow = _worlds(width=size[0], height=size[1])
    rw.L.ref_sbi_reset_size()
    rw.L.ref_srand(1)
    ow.L.orc_tracker_enable_sbi(ow.tracker, synth.Camera(cam.width // 16, cam.height // 16).scalars())
    for k in range(1, 9):
        fr = synth.render_frame(common.texture(), cam, synth.stream_pose(5 * k, 2))
        rw.L.ref_tracker_track_frame(rw.tracker, fr, cam.width, cam.height, cam.width)
        ow.L.orc_tracker_track_frame(ow.tracker, fr, cam.width, cam.height, cam.width)
        rv, ov = np.zeros(6), np.zeros(6)
        rw.L.ref_tracker_get_sbi_rot(rw.tracker, rv); ow.L.orc_tracker_get_sbi_rot(ow.tracker, ov)
        assert np.array_equal(rv, ov), k
        assert np.array_equal(rw.get_pose(), ow.get_pose()), k
        assert all(np.array_equal(a, b) if isinstance(a, np.ndarray) else a == b for a, b in zip(rw.counters(), ow.counters())), k


def test_epipolar_new_point_fields_and_triangulation():
    """f3 tail: the patch-source fields + MapPoint::RefreshPixelVectors of a point created by AddPointEpipolar (jni/MapMaker.cc:655-684),
    computed by the reference's MapPoint / ATANCamera objects, against the restatement, bit for bit; and the restated triangulation
    (MapMaker::ReprojectPoint; Eigen's SVD is absent, LAPACK's stands in) recovers the synthetic scene's plane within noise."""
    cam, f0, smap, rw, ow = _worlds()
    tw = np.array([0.12, 0.03, 0.02, 0.01, -0.03, 0.02])
    f1, pose1 = common.frame_at(cam, tw)
    rk0 = refbind.RefKeyFrame().make_lite(f0); rk0.make_rest()
    rk1 = refbind.RefKeyFrame().make_lite(f1)
    ok0 = oraclebind.OrcKeyFrame().make_lite(f0); ok0.make_rest()
    cam13 = np.ascontiguousarray(cam.scalars(), dtype=np.float64)
    src_pose = synth.se3_exp(np.array([0.02, -0.01, 0.03, 0.01, 0.02, -0.01]))       # a non-identity source pose exercises the rotations
    tgt_pose = pose1 @ np.vstack([src_pose, [0, 0, 0, 1]])
    sp = np.ascontiguousarray(src_pose, dtype=np.float64).reshape(12); tp = np.ascontiguousarray(tgt_pose, dtype=np.float64).reshape(12)
    rs = np.random.RandomState(2)
    checked = depth_ok = 0
    for level in range(4):
        xy, _ = ok0.candidates(level)
        for k in range(0, len(xy), max(1, len(xy) // 25)):
            ro, rp = np.zeros(3, dtype=np.int32), np.zeros(2)
            rw.L.ref_epipolar_search(rw.tracker, rk0.h, rk1.h, sp, tp, 1.0, 0.3, 0.1, level, k, ro, rp)
            if not ro[0]:
                continue
            root = (xy[k] + 0.5) * (1 << level) - 0.5
            world = oraclebind.triangulate(cam13, src_pose, tgt_pose, root, rp)
            in_src = src_pose[:, :3] @ world + src_pose[:, 3]
            depth_ok += abs(in_src[2] - 1.0) < 0.05                                     # the synthetic scene is the plane z = 1 of the first camera
            for w in (world, rs.randn(3) + np.array([0, 0, 3.0])):
                ref15 = np.zeros(15)
                rw.L.ref_epipolar_point_fields(rw.tracker, rk0.h, sp, level, k, np.ascontiguousarray(w), ref15)
                got = oraclebind.epipolar_point_fields(cam13, src_pose, level, xy[k, 0], xy[k, 1], w)
                assert np.array_equal(got.reshape(15), ref15), (level, k)
            checked += 1
    assert checked > 30 and depth_ok > 0.9 * checked


def _kf_info(L, fn, tracker):
    v = [C.c_int() for _ in range(4)]
    getattr(L, fn)(tracker, *[C.byref(x) for x in v])
    return [x.value for x in v]


def test_track_frame_keyframe_handoff_heuristics():
    """jni/Tracker.cc:127-132 and :866-872 of the UNMODIFIED Tracker::TrackFrame -- 'add a keyframe when tracking is GOOD, the camera
    moved far enough from the nearest keyframe (relative to scene depth) and 20 frames passed', 'DODGY is BAD when the pose ran away
    from every keyframe', mnFrame / mnLastKeyFrameDropped bookkeeping -- with the MapMaker side (NeedNewKeyFrame, AddKeyFrame,
    IsDistanceToNearestKeyFrameExcessive) driven on the reference's objects, against the restatement: poses, counters, quality, the
    frames at which keyframes are added, and a relocalisation against a keyframe added that way."""
    cam, f0, smap, rw, ow = _worlds()
    rw.L.ref_sbi_reset_size()
    sbi_cam = synth.Camera(cam.width // 16, cam.height // 16).scalars()
    ow.L.orc_tracker_enable_sbi(ow.tracker, sbi_cam)
    rw.L.ref_kf_make_sbi(rw.src_kf.h)
    okf0 = oraclebind.OrcKeyFrame().make_lite(f0)
    ow.L.orc_tracker_add_reloc_keyframe(ow.tracker, okf0.h, np.ascontiguousarray(synth.IDENTITY_POSE, dtype=np.float64).reshape(12))
    rw.set_pose(synth.IDENTITY_POSE); ow.set_pose(synth.IDENTITY_POSE)
    rw.L.ref_srand(1)
    try:
        rw.L.ref_set_keyframe_policy(rw.tracker, 1, 0.1, 0.1, 0.2)
        ow.L.orc_tracker_set_keyframe_policy(ow.tracker, 1, 0.1, 0.1, 0.2, 20)
        base_added = _kf_info(rw.L, "ref_keyframe_info", rw.tracker)[1]
        added_at = []
        step = np.array([0.004, 0.001, 0.0005, 0.0004, -0.0012, 0.0008])
        rs = np.random.RandomState(9)
        frames = [common.frame_at(cam, step * k)[0] for k in range(1, 31)]
        frames += [rs.randint(0, 255, f0.shape).astype(np.uint8) for _ in range(4)]                 # lose tracking ...
        frames += [common.frame_at(cam, step * 27.5)[0], common.frame_at(cam, step * 28.5)[0]]       # ... and come back next to the last keyframe added
        for k, fr in enumerate(frames):
            fr = np.ascontiguousarray(fr)
            rw.L.ref_tracker_track_frame(rw.tracker, fr, cam.width, cam.height, cam.width)
            ow.L.orc_tracker_track_frame(ow.tracker, fr, cam.width, cam.height, cam.width)
            assert np.array_equal(rw.get_pose(), ow.get_pose()), k
            a, f, q, lost, dc = rw.counters(); oa, of, oq, olost, odc = ow.counters()
            assert np.array_equal(a, oa) and np.array_equal(f, of) and (q, lost, dc) == (oq, olost, odc), k
            rn, radd, rframe, rlast = _kf_info(rw.L, "ref_keyframe_info", rw.tracker)
            on, oadd, oframe, olast = _kf_info(ow.L, "orc_tracker_keyframe_info", ow.tracker)
            assert (rn, rframe, rlast) == (on, oframe, olast), (k, rn, on, rframe, oframe, rlast, olast)
            if oadd:
                added_at.append(k + 1)
                assert "Adding key-frame" in rw.message()
        assert added_at == [6, 27], added_at                     # 0.004/frame: first beyond 0.2 * 0.1 * depth at frame 6, then 21 frames later
        assert _kf_info(rw.L, "ref_keyframe_info", rw.tracker)[1] - base_added == 2
        best, score, nrec = C.c_int(), C.c_double(), C.c_int()
        ow.L.orc_tracker_reloc_info(ow.tracker, C.byref(best), C.byref(score), C.byref(nrec))
        assert nrec.value >= 1 and best.value == 2 and ow.counters()[2] == 2        # recovered against the keyframe added at frame 27
        # DODGY -> BAD when the pose is far from every keyframe: the same partly occluded frame with a roomy and with a tiny wiggle scale
        seen = 0
        for frac in (0.55, 0.65, 0.72, 0.8):
            occ = common.frame_at(cam, step * 29.0)[0].copy()
            occ[:int(cam.height * frac)] = rs.randint(0, 255, (int(cam.height * frac), cam.width))
            for wiggle in (0.1, 1e-5):
                rw.L.ref_set_keyframe_policy(rw.tracker, 1, wiggle, 1e30, 0.2)
                ow.L.orc_tracker_set_keyframe_policy(ow.tracker, 1, wiggle, 1e30, 0.2, 20)
                good = common.frame_at(cam, step * 28.8)[0]
                for fr in (good, occ):
                    rw.L.ref_tracker_set_lost(rw.tracker, 0, 2); ow.L.orc_tracker_set_lost(ow.tracker, 0, 2) if fr is good else None
                    fr = np.ascontiguousarray(fr)
                    rw.L.ref_tracker_track_frame(rw.tracker, fr, cam.width, cam.height, cam.width)
                    ow.L.orc_tracker_track_frame(ow.tracker, fr, cam.width, cam.height, cam.width)
                    assert np.array_equal(rw.get_pose(), ow.get_pose()), (frac, wiggle)
                    assert rw.counters()[2:4] == ow.counters()[2:4], (frac, wiggle, rw.counters(), ow.counters())
                a, f, q, lost, dc = ow.counters()
                tot = f.sum() / max(1, a.sum())
                if tot <= 0.3 and wiggle < 1e-3 and q == 0 and lost == 1:
                    seen += 1
        assert seen >= 1, "no DODGY frame in the occlusion sweep"
    finally:
        rw.L.ref_set_keyframe_policy(rw.tracker, 0, 0.1, 0.1, 0.2)


def _submap(smap, sl):
    return synth.SyntheticMap(**{k: getattr(smap, k)[sl] for k in ("world", "pix_right_w", "pix_down_w", "ir_center", "src_level", "center_nc", "one_right_nc", "one_down_nc")})


def test_map_grows_while_tracking():
    """Map::vpPoints.push_back during tracking (what MapMaker::AddPointEpipolar does from its thread): old points keep their
    TrackerData (template cache, M-estimator counters), new ones get theirs on first use (jni/Tracker.cc:372) -- unmodified
    TrackFrame vs restatement across two appends."""
    cam, f0, smap = common.scene()
    parts = [_submap(smap, slice(0, 500)), _submap(smap, slice(500, 800)), _submap(smap, slice(800, 1000))]
    rw = refbind.RefWorld(cam.width, cam.height, f0, parts[0])
    pr, pd = rw.pixel_vectors()
    ow = oraclebind.OrcWorld(cam, f0, parts[0], pix_right=pr, pix_down=pd)
    rw.L.ref_sbi_reset_size()
    rw.L.ref_srand(1)
    ow.L.orc_tracker_enable_sbi(ow.tracker, synth.Camera(cam.width // 16, cam.height // 16).scalars())
    nxt = 1
    for k in range(1, 10):
        if k in (4, 7):
            rw.append_points(parts[nxt])
            pr, pd = rw.pixel_vectors()
            ow.append_points(parts[nxt], pix_right=pr[-parts[nxt].n:], pix_down=pd[-parts[nxt].n:])
            nxt += 1
        fr = synth.render_frame(common.texture(), cam, synth.stream_pose(4 * k, 2))
        rw.L.ref_tracker_track_frame(rw.tracker, fr, cam.width, cam.height, cam.width)
        ow.L.orc_tracker_track_frame(ow.tracker, fr, cam.width, cam.height, cam.width)
        assert np.array_equal(rw.get_pose(), ow.get_pose()), k
        assert all(np.array_equal(a, b) if isinstance(a, np.ndarray) else a == b for a, b in zip(rw.counters(), ow.counters())), k
    assert rw.n == ow.n == 1000 and ow.counters()[0].sum() > 600

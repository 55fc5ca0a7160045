"""CPU: the oracle restatement (oracle/vslam_oracle.cc) against the golden fixture that the COMPILED REFERENCE produced
(tests/golden/ref_small.npz, generator tests/golden/make_golden.py).  Runs without /root/reference and without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oraclebind
from visualslam_android_b200 import synth

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_small.npz"))
W, H = 320, 240


def _smap():
    return synth.SyntheticMap(world=G["world"], pix_right_w=G["pix_right_w"], pix_down_w=G["pix_down_w"], ir_center=G["ir_center"],
                              src_level=G["src_level"], center_nc=G["center_nc"], one_right_nc=G["one_right_nc"], one_down_nc=G["one_down_nc"])


def test_camera_scalars_match_reference_refresh_params():
    cam = synth.Camera(W, H)
    assert np.array_equal(cam.scalars()[:11], G["ref_cam_scalars"][:11])   # fx..maxR, bit for bit (fix_radius variant)
    shipped = synth.Camera(W, H, fix_radius=False)
    assert shipped.largest_radius == 0.0 and shipped.max_r == 0.0


def test_make_keyframe_lite_and_rest_golden():
    kf = oraclebind.OrcKeyFrame().make_lite(G["f1"])
    kf.make_rest()
    for l in range(4):
        assert np.array_equal(kf.pixels(l), G[f"lvl{l}"])
        assert np.array_equal(kf.corners(l), G[f"corners{l}"])
        assert np.array_equal(kf.row_lut(l), G[f"lut{l}"])
        assert np.array_equal(kf.max_corners(l), G[f"max{l}"]), f"non-max suppression level {l}"
        xy, s = kf.candidates(l)
        assert np.array_equal(xy, G[f"cand{l}"]) and np.array_equal(s, G[f"cand_score{l}"]), f"Shi-Tomasi candidates level {l}"


def test_refresh_pixel_vectors_close_to_reference():
    sm = _smap()
    r, d = synth.refresh_pixel_vectors(sm.center_nc, sm.one_right_nc, sm.one_down_nc, sm.world)
    assert np.allclose(r, G["pix_right_w"], rtol=1e-12, atol=1e-18) and np.allclose(d, G["pix_down_w"], rtol=1e-12, atol=1e-18)


def _check_case(ow, tag):
    ints, dbl = ow.point_states()
    gi, gd = G[f"{tag}_ints"], G[f"{tag}_dbl"]
    assert np.array_equal(ints[:, [0, 1]], gi[:, [0, 1]]), "bInImage / nSearchLevel"
    pvs = gi[:, 1] >= 0
    assert np.array_equal(ints[pvs][:, [2, 3, 5]], gi[pvs][:, [2, 3, 4]]), "searched / found / templateBad"
    fnd = pvs & (gi[:, 3] == 1)
    assert np.array_equal(dbl[fnd][:, [0, 1, 2, 3, 15]], gd[fnd]), "v2Image, v2Found, sqrtInvNoise of found points: bit for bit"
    a, f, q, lost, dc = ow.counters()
    assert np.array_equal(np.concatenate([a, f, [dc]]), G[f"{tag}_counters"])
    assert np.array_equal(ow.get_pose(), G[f"{tag}_pose"]), "pose after TrackMap: bit for bit"


def test_track_map_golden_fine_then_coarse():
    cam = synth.Camera(W, H)
    ow = oraclebind.OrcWorld(cam, G["f0"], _smap())
    ow.make_current_kf(G["f1"])
    ow.set_pose(G["start_pose"])
    ow.L.orc_tracker_track_map(ow.tracker)
    _check_case(ow, "B")
    assert G["B_counters"][8] == 0
    ow.make_current_kf(G["f2"])
    ow.L.orc_tracker_set_velocity(ow.tracker, np.zeros(6), 0.05)
    ow.L.orc_tracker_track_map(ow.tracker)
    _check_case(ow, "C")
    assert G["C_counters"][8] == 1, "the fixture's second TrackMap ran the coarse stage"
    assert np.array_equal(ow.point_counts(), G["C_counts"]), "M-estimator inlier / outlier counters"


def test_trail_tracking_golden():
    """Tracker::TrailTracking_Start / _Advance (jni/Tracker.cc:264-346): trail lists of the compiled reference, frame by frame."""
    ot = oraclebind.OrcTrails()
    assert ot.start(oraclebind.OrcKeyFrame().make_lite(G["f0"])) == int(G["E_start_n"][0])
    assert np.array_equal(ot.trails(), G["E_trails0"])
    for k in range(1, 4):
        assert ot.advance(oraclebind.OrcKeyFrame().make_lite(G[f"E_f{k}"]), 100000) == int(G[f"E_good{k}"][0])
        assert np.array_equal(ot.trails(), G[f"E_trails{k}"]), k
    assert 10 < len(G["E_trails3"]) < int(G["E_start_n"][0])


def test_refind_common_golden():
    """MapMaker::ReFind_Common's PatchFinder / camera call sequence on the reference's objects (jni/MapMaker.cc:967-1036)."""
    cam = synth.Camera(W, H)
    smap = _smap()
    ow = oraclebind.OrcWorld(cam, G["f0"], smap)
    ow.make_current_kf(G["f2"]); ow.set_pose(G["F_pose"])
    idx = np.arange(smap.n, dtype=np.int32)
    oo, op = np.zeros((smap.n, 3), dtype=np.int32), np.zeros((smap.n, 2))
    ow.L.orc_tracker_refind(ow.tracker, idx, smap.n, 4, 8, 0, oo, op)
    assert np.array_equal(oo, G["F_flags"]) and np.array_equal(op, G["F_pos"])
    assert oo[:, 0].sum() > 100 and oo[:, 2].sum() > 20


def test_epipolar_search_golden():
    """The search of MapMaker::AddPointEpipolar (jni/MapMaker.cc:525-640) on the reference's objects: found flag, best corner, ZMSSD
    and refined position for candidates of every level."""
    cam = synth.Camera(W, H)
    ow = oraclebind.OrcWorld(cam, G["f0"], _smap())
    k0 = oraclebind.OrcKeyFrame().make_lite(G["f0"]); k3 = oraclebind.OrcKeyFrame().make_lite(G["G_f3"])
    eye = np.ascontiguousarray(synth.IDENTITY_POSE, dtype=np.float64).reshape(12); p3 = np.ascontiguousarray(G["G_pose"], dtype=np.float64).reshape(12)
    nfound = 0
    for row in G["G_rows"]:
        oo, op = np.zeros(3, dtype=np.int32), np.zeros(2)
        ow.L.orc_epipolar_search(ow.tracker, k0.h, k3.h, eye, p3, 1.0, 0.3, 0.1, int(row[0]), int(row[1]), int(row[2]), oo, op, None)
        assert np.array_equal(oo, row[3:6].astype(np.int32)) and np.array_equal(op, row[6:8]), row
        nfound += int(oo[0])
    assert nfound > 20


def test_epipolar_new_point_fields_golden():
    """Patch-source rays and MapPoint::RefreshPixelVectors of points created from converged epipolar matches (tail of
    MapMaker::AddPointEpipolar, jni/MapMaker.cc:655-684; jni/MapPoint.cc:4-29) as the reference's objects computed them."""
    cam13 = np.ascontiguousarray(synth.Camera(W, H).scalars(), dtype=np.float64)
    rows = G["G_fields"]
    assert len(rows) > 40
    for row in rows:
        got = oraclebind.epipolar_point_fields(cam13, row[3:15], int(row[0]), int(row[1]), int(row[2]), row[15:18])
        assert np.array_equal(got.reshape(15), row[18:33]), row[:3]


def test_track_frame_loss_and_relocalisation_golden():
    """The unmodified Tracker::TrackFrame of the reference through a loss of tracking and two relocalisations (Relocaliser over three
    map keyframes): pose and counters after every frame, bit for bit."""
    cam = synth.Camera(W, H)
    smap = _smap()
    ow = oraclebind.OrcWorld(cam, G["f0"], smap)
    ow.L.orc_tracker_enable_sbi(ow.tracker, synth.Camera(W // 16, H // 16).scalars())
    keep = []
    for k in range(3):
        okf = oraclebind.OrcKeyFrame().make_lite(G[f"H_kf{k}"]); keep.append(okf)
        ow.L.orc_tracker_add_reloc_keyframe(ow.tracker, okf.h, np.ascontiguousarray(G[f"H_kfpose{k}"], dtype=np.float64).reshape(12))
    ow.set_pose(synth.IDENTITY_POSE)
    rs5 = np.random.RandomState(5)
    nr = 0
    for k, kind in enumerate(G["H_kinds"]):
        if kind == 0:
            fr = rs5.randint(0, 255, (H, W)).astype(np.uint8)
        else:
            fr = G[f"H_r{nr}"]; nr += 1
        ow.L.orc_tracker_track_frame(ow.tracker, np.ascontiguousarray(fr), W, H, W)
        assert np.array_equal(ow.get_pose(), G["H_poses"][k]), k
        a, f, q, lost, dc = ow.counters()
        assert np.array_equal(np.concatenate([a, f, [q, lost, dc]]), G["H_counters"][k]), k
    assert G["H_counters"][:, 9].max() >= 3 and G["H_counters"][-1, 8] == 2     # tracking was lost, and it ends GOOD


def test_se3_exp_ln_golden():
    L = oraclebind.lib()
    for mu, e, l in zip(G["D_mu"], G["D_exp"], G["D_ln"]):
        oe, ol = np.zeros(12), np.zeros(6)
        L.orc_se3_exp(np.ascontiguousarray(mu), oe); L.orc_se3_ln(oe, ol)
        assert np.array_equal(oe, e) and np.array_equal(ol, l)


def test_tukey_sigma_golden_including_tiny_sets():
    L = oraclebind.lib()
    off = 0
    for n, want in zip(G["D_tukey_n"], G["D_tukey_out"]):
        e = np.ascontiguousarray(G["D_tukey_in"][off:off + n]); off += n
        got = L.orc_tukey_sigma_squared(e, int(n))
        assert got == want or (np.isinf(got) and np.isinf(want)), (n, got, want)   # n=1,2: size_t wrap; n=3 would be inf


def test_glibc_rand_restated():
    L = oraclebind.lib()
    r = L.orc_rand_create(1)
    got = np.array([L.orc_rand_next(r) for _ in range(400)], dtype=np.int64)
    L.orc_rand_destroy(r)
    assert np.array_equal(got, G["D_rand"])
    libc = C.CDLL(None)
    libc.srand(12345)
    want = [libc.rand() for _ in range(1000)]
    r = L.orc_rand_create(12345)
    assert [L.orc_rand_next(r) for _ in range(1000)] == want
    L.orc_rand_destroy(r)


def test_half_sample_is_opencv_resize():
    """cv::resize(prev, lev, size/2) (jni/KeyFrame.cc:22) == (a+b+c+d+2)>>2 for even sizes, probed with cv2."""
    cv2 = pytest.importorskip("cv2")
    kf = oraclebind.OrcKeyFrame().make_lite(G["f1"])
    a = G["f1"]
    for l in range(1, 4):
        a = cv2.resize(a, (a.shape[1] // 2, a.shape[0] // 2))
        assert np.array_equal(a, kf.pixels(l))


def test_fast10_is_the_segment_test():
    """Brute-force 'at least 10 contiguous ring pixels all brighter / all darker' on a noise image."""
    rs = np.random.RandomState(2)
    im = rs.randint(0, 256, (96, 128)).astype(np.uint8)
    im[20:60, 30:90] = (im[20:60, 30:90] // 6 + 100).astype(np.uint8)
    kf = oraclebind.OrcKeyFrame().make_lite(im)
    offs = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]
    I = im.astype(np.int32); Hh, Ww = I.shape
    want = []
    for y in range(3, Hh - 3):
        for x in range(3, Ww - 3):
            c = I[y, x]
            br = [I[y + dy, x + dx] > c + 10 for dx, dy in offs]; dk = [I[y + dy, x + dx] < c - 10 for dx, dy in offs]
            ok = any(all(br[(s + k) % 16] for k in range(10)) or all(dk[(s + k) % 16] for k in range(10)) for s in range(16))
            if ok:
                want.append((x, y))
    assert np.array_equal(kf.corners(0), np.array(want, dtype=np.int32).reshape(-1, 2))
    lut = kf.row_lut(0)
    ys = kf.corners(0)[:, 1]
    assert np.array_equal(lut, np.searchsorted(ys, np.arange(Hh), side="left"))   # LUT[y] = #corners with row < y


def test_restated_cv_resize_linear_matches_opencv_golden_vectors():
    """oracle/shim/cv_resize_linear_u8.h (the cv::resize of SmallBlurryImage::MakeFromKF, jni/SmallBlurryImage.cc:22-30, for level-3 sizes that
    are not even, e.g. 1080p) against vectors made by the real OpenCV (tests/golden/make_resize_golden.py), and live against cv2 when present."""
    import os
    from oracle import oraclebind
    L = oraclebind.lib()
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resize_linear.npz"))
    for k, (sw, sh, dw, dh) in enumerate(g["sizes"]):
        got = np.zeros((dh, dw), dtype=np.uint8)
        L.orc_resize_linear_u8(np.ascontiguousarray(g[f"src{k}"]), int(sw), int(sh), got, int(dw), int(dh))
        assert np.array_equal(got, g[f"dst{k}"]), (sw, sh, dw, dh)
    try:
        import cv2
    except ImportError:
        return
    rng = np.random.default_rng(3)
    for sw, sh, dw, dh in [(240, 135, 120, 67), (17, 9, 8, 4), (30, 16, 45, 24), (240, 135, 120, 68), (160, 120, 80, 60)]:
        src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        got = np.zeros((dh, dw), dtype=np.uint8)
        L.orc_resize_linear_u8(src, sw, sh, got, dw, dh)
        assert np.array_equal(got, cv2.resize(src, (dw, dh))), (sw, sh, dw, dh)

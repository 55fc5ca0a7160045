"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Bit-exact: pyramid pixels, corner lists + row LUT, search levels, template pixels and sums, ZMSSD argmin positions,
found flags, counters.  Tolerance (written at each assert): FP64 geometry 1e-9 relative, pose / update 6-vectors 1e-6
(the contract is 1e-4), sub-pixel positions 1e-6 px.
"""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

import common
from visualslam_android_b200 import synth

pytestmark = pytest.mark.gpu


def _ctx(cam, f0, smap, n_streams=1, **kw):
    from visualslam_android_b200 import api
    ctx = api.Context(cam.width, cam.height, n_streams=n_streams, max_points=smap.n, **kw)
    ctx.set_camera(cam.scalars())
    ctx.upload_source_keyframe(f0)
    ctx.set_map(smap.world, smap.pix_right_w, smap.pix_down_w, smap.ir_center, smap.src_level)
    return ctx


def _orc(cam, f0, smap, **kw):
    from oracle import oraclebind
    return oraclebind.OrcWorld(cam, f0, smap, **kw)


def _check_keyframe(ctx, s, okf):
    for l in range(4):
        assert np.array_equal(ctx.level(s, l), okf.pixels(l)), f"level {l} pixels"
        assert np.array_equal(ctx.corners(s, l), okf.corners(l)), f"level {l} corners"
        assert np.array_equal(ctx.row_lut(s, l), okf.row_lut(l)), f"level {l} row LUT"


@pytest.mark.parametrize("size", [(640, 480), (320, 240), (1920, 1080), (96, 64)])
def test_make_keyframe_lite_bit_exact(size):
    from oracle import oraclebind
    from visualslam_android_b200 import api
    W, H = size
    cam = synth.Camera(W, H)
    tex = common.texture(4096 if W > 640 else 2048)
    ctx = api.Context(W, H, n_streams=3, max_points=8)
    frames = np.stack([synth.render_frame(tex, cam, synth.se3_exp(np.array(synth.CONFIG1_TWIST) * k)) for k in range(3)])
    ctx.make_keyframe_lite(frames)
    for s in range(3):
        _check_keyframe(ctx, s, oraclebind.OrcKeyFrame().make_lite(frames[s]))
    # a second call on the same context (look-back state / tickets are reset per call), different content per stream
    frames2 = frames[::-1].copy()
    ctx.make_keyframe_lite(frames2)
    for s in range(3):
        _check_keyframe(ctx, s, oraclebind.OrcKeyFrame().make_lite(frames2[s]))
    ctx.close()


def test_make_keyframe_lite_edge_images():
    from oracle import oraclebind
    from visualslam_android_b200 import api
    W, H = 160, 120
    rs = np.random.RandomState(3)
    imgs = np.stack([
        np.zeros((H, W), np.uint8),                                   # no corners at all
        rs.randint(0, 256, (H, W)).astype(np.uint8),                  # noise: corners everywhere, including the x=3 / x=W-4 columns
        np.full((H, W), 255, np.uint8),                               # saturated
        ((np.indices((H, W)).sum(0) // 7 % 2) * 255).astype(np.uint8),  # diagonal stripes
    ])
    ctx = api.Context(W, H, n_streams=4, max_points=8, max_corner_frac=1.0)
    ctx.make_keyframe_lite(imgs)
    for s in range(4):
        _check_keyframe(ctx, s, oraclebind.OrcKeyFrame().make_lite(imgs[s]))
    ctx.close()


@pytest.mark.parametrize("size", [(64, 48), (96, 72), (224, 104), (640, 8), (32, 200)])
def test_make_keyframe_lite_threshold_boundaries_and_odd_sizes(size):
    """FAST's comparisons are strict (`>` cb, `<` c_b, jni/vision/cvfast.cpp): images whose pixels sit exactly ON the thresholds of
    their neighbours (values drawn from {c, c +- t, c +- (t+1)} for t = 10 and 15), few-valued noise, and plateaus with single
    outliers, at sizes whose strips, row pairs and 128-pixel chunks are ragged (height 8, widths 32..224, odd level-3 heights)."""
    from oracle import oraclebind
    from visualslam_android_b200 import api
    W, H = size
    rs = np.random.RandomState(W * 1000 + H)
    imgs = []
    for base, t in ((100, 10), (100, 15), (20, 10), (240, 15)):
        vals = np.clip(np.array([base, base + t, base - t, base + t + 1, base - t - 1]), 0, 255)
        imgs.append(vals[rs.randint(0, 5, (H, W))].astype(np.uint8))
    imgs.append(np.where(rs.rand(H, W) < 0.03, 200, 90).astype(np.uint8))          # plateau with isolated bright pixels
    imgs.append((rs.randint(0, 3, (H, W)) * 11 + 60).astype(np.uint8))            # three grey values 11 apart (t = 10 passes, 15 does not)
    imgs = np.stack(imgs)
    ctx = api.Context(W, H, n_streams=len(imgs), max_points=8, max_corner_frac=1.0)
    ctx.make_keyframe_lite(imgs)
    for s in range(len(imgs)):
        _check_keyframe(ctx, s, oraclebind.OrcKeyFrame().make_lite(imgs[s]))
    ctx.close()


def test_corner_capacity_overflow_is_reported():
    from visualslam_android_b200 import api
    W, H = 160, 120
    img = np.random.RandomState(3).randint(0, 256, (1, H, W)).astype(np.uint8)
    ctx = api.Context(W, H, n_streams=1, max_points=8, max_corner_frac=0.01)
    ctx.make_keyframe_lite(img)
    with pytest.raises(api.VslamError) as e:
        ctx.sync()
    assert e.value.code == api.E_CAPACITY
    ctx.close()


def test_make_keyframe_lite_device_input_zero_copy():
    import torch
    from oracle import oraclebind
    from visualslam_android_b200 import api
    W, H = 640, 480
    cam = synth.Camera(W, H)
    frames = np.stack([synth.render_frame(common.texture(), cam, synth.se3_exp(np.array(synth.CONFIG1_TWIST) * k)) for k in range(2)])
    dev = torch.from_numpy(frames).cuda()
    ctx = api.Context(W, H, n_streams=2, max_points=8, cuda_stream=torch.cuda.current_stream().cuda_stream)
    ctx.make_keyframe_lite_ptr(dev.data_ptr(), 2, W, W * H, device=True)
    for s in range(2):
        _check_keyframe(ctx, s, oraclebind.OrcKeyFrame().make_lite(frames[s]))
    ctx.close()


def _compare_states(gi, gd, oi, od, geom_tol=1e-9):
    assert np.array_equal(gi[:, 0], oi[:, 0]), "bInImage"
    assert np.array_equal(gi[:, 1], oi[:, 1]), "nSearchLevel"
    vis = oi[:, 0] == 1
    # projected pixel, derivs, v3Cam, warp: 1e-9 relative (contract 1e-12 on v2Image is checked separately below)
    for cols, name in (((0, 1), "v2Image"), ((4, 5, 6, 7), "derivs"), ((8, 9, 10), "v3Cam"), ((11, 12, 13, 14), "warpInverse")):
        a, b = gd[vis][:, cols], od[vis][:, cols]
        assert np.allclose(a, b, rtol=geom_tol, atol=1e-12), name


def test_atan_cr_matches_host_libm():
    """csrc/atan_dd.cuh against the host's atan (what the reference's ATANCamera calls): bit-equal for >= 99.9 %, never off by more than 1 ulp."""
    import math
    from visualslam_android_b200 import api
    rs = np.random.RandomState(5)
    x = np.concatenate([rs.uniform(0, 0.03, 200000), rs.uniform(0, 1, 100000), rs.uniform(1, 40, 50000), -rs.uniform(0, 2, 20000),
                        np.array([0.0, 1e-12, 1e-9, 0.125, 0.0625, 1.0, 8.0, 1e19, -1e-3])])
    y = api.debug_atan(x)
    ref = np.array([math.atan(v) for v in x])
    same = (y == ref)
    assert same.mean() >= 0.999, same.mean()
    ulp = np.abs(y - ref) / np.maximum(np.spacing(np.abs(ref)), 5e-324)
    assert ulp.max() <= 1.0, ulp.max()


def test_atan_fast_path_is_the_correctly_rounded_value():
    """|x| <= 1/16 takes a ~45-flop path (cubic term in double-double + Ziv's rounding test, csrc/atan_dd.cuh) that must return exactly what
    the double-double routine returns: 2 M arguments in the range the camera model produces, plus the range ends."""
    from visualslam_android_b200 import api
    rs = np.random.RandomState(11)
    x = np.concatenate([rs.uniform(-0.0625, 0.0625, 1200000), rs.uniform(-0.03, 0.03, 800000), 10.0 ** rs.uniform(-9.2, -1.2, 200000),
                        np.array([0.0625, -0.0625, 1e-9, 0.06250000000000001, 0.9999999e-9, 0.0])])
    assert np.array_equal(api.debug_atan(x), api.debug_atan(x, dd_only=True))


def test_project_all_matches_oracle():
    cam, f0, smap = common.scene()
    ctx, ow = _ctx(cam, f0, smap), _orc(cam, f0, smap)
    pose = synth.se3_exp(np.array(synth.CONFIG1_TWIST) * 0.5)
    ctx.set_pose(0, pose); ow.set_pose(pose)
    ctx.project_all(); ow.L.orc_tracker_project_all(ow.tracker)
    gi, gd = ctx.point_states(0); oi, od = ow.point_states()
    _compare_states(gi, gd, oi, od)
    vis = oi[:, 0] == 1
    rel = np.abs(gd[vis][:, :2] - od[vis][:, :2]) / np.maximum(np.abs(od[vis][:, :2]), 1.0)
    assert rel.max() <= 1e-12, rel.max()      # SURVEY.md §8 contract for the projected pixel
    # ... and for the warp matrix (mm2WarpInverse) and the camera derivatives it is built from: 1e-12 relative to the matrix's largest entry
    for cols, name in (((11, 12, 13, 14), "warpInverse"), ((4, 5, 6, 7), "derivs")):
        a, b = gd[vis][:, cols], od[vis][:, cols]
        relm = np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-300)
        assert relm.max() <= 1e-12, (name, relm.max())
    # with the correctly-rounded device atan the whole projection is bit-identical for (almost) every point
    ident = (gd[vis][:, [0, 1, 4, 5, 6, 7, 11, 12, 13, 14]] == od[vis][:, [0, 1, 4, 5, 6, 7, 11, 12, 13, 14]]).all(1)
    assert ident.mean() >= 0.99, ident.mean()
    ctx.close()


@pytest.mark.parametrize("P", [11, 8])
@pytest.mark.parametrize("rng,subpix", [(10, 0), (30, 8), (5, 8)])
def test_search_for_points_bit_exact(P, rng, subpix):
    cam, f0, smap = common.scene()
    ctx, ow = _ctx(cam, f0, smap, patch_size=P), _orc(cam, f0, smap, P=P)
    f1, pose1 = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.6)
    start = synth.se3_exp(np.array(synth.CONFIG1_TWIST) * 0.3)
    ctx.make_keyframe_lite(f1); ow.make_current_kf(f1)
    ctx.set_pose(0, start); ow.set_pose(start)
    ctx.project_all(); ow.L.orc_tracker_project_all(ow.tracker)
    oi, od = ow.point_states()
    # stage isolation (SURVEY.md §8 parity contract): the search is fed the oracle's projected pixels / warp matrices
    ctx.set_point_projection(0, od[:, 0:2], od[:, 11:15], oi[:, 1])
    idx = np.nonzero(oi[:, 1] >= 0)[0].astype(np.int32)
    ctx.set_lists([idx]); ctx.clear_counters(); ow.L.orc_tracker_clear_counters(ow.tracker)
    ctx.search_for_points(rng, subpix)
    nfound = ow.L.orc_tracker_search_for_points(ow.tracker, idx, len(idx), rng, subpix)
    gi, gd = ctx.point_states(0); oi, od = ow.point_states()
    assert np.array_equal(gi[idx][:, [2, 3, 5]], oi[idx][:, [2, 3, 5]]), "searched / found / templateBad flags"
    a, f, *_ = ctx.counters(0); oa, of, *_ = ow.counters()
    assert np.array_equal(a, oa) and np.array_equal(f, of) and f.sum() == nfound
    assert ctx.zmssd_evals() == ow.L.orc_tracker_zmssd_evals(ow.tracker)
    # templates of every searched point: bit-exact pixels and sums
    for k in idx:
        gt, gs, gq = ctx.point_template(0, int(k)); ot, os_, oq = ow.point_template(int(k))
        assert np.array_equal(gt, ot) and (gs, gq) == (os_, oq), f"template of point {k}"
    fnd = idx[oi[idx][:, 3] == 1]
    assert len(fnd) > 100
    if subpix == 0:
        assert np.array_equal(gd[fnd][:, 2:4], od[fnd][:, 2:4]), "coarse positions must be identical"
        assert np.array_equal(gi[fnd][:, 4], oi[fnd][:, 4])
    else:
        assert np.abs(gd[fnd][:, 2:4] - od[fnd][:, 2:4]).max() <= 1e-6, "sub-pixel position (px)"
        assert np.array_equal(gd[fnd][:, 30:32], od[fnd][:, 30:32]), "coarse positions must be identical"
    assert np.array_equal(gd[fnd][:, 15], od[fnd][:, 15])
    ctx.close()


def test_template_cache_and_second_frame():
    """The per-point template reuse cache (jni/PatchFinder.cc:91-102) persists across frames."""
    cam, f0, smap = common.scene()
    ctx, ow = _ctx(cam, f0, smap), _orc(cam, f0, smap)
    for k, sc in enumerate((0.2, 0.21, 0.6)):
        f1, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * sc)
        start = synth.se3_exp(np.array(synth.CONFIG1_TWIST) * (sc - 0.05))
        ctx.make_keyframe_lite(f1); ow.make_current_kf(f1)
        ctx.set_pose(0, start); ow.set_pose(start)
        ctx.project_all(); ow.L.orc_tracker_project_all(ow.tracker)
        oi, _ = ow.point_states()
        idx = np.nonzero(oi[:, 1] >= 0)[0].astype(np.int32)
        ctx.set_lists([idx])
        ctx.search_for_points(10, 0); ow.L.orc_tracker_search_for_points(ow.tracker, idx, len(idx), 10, 0)
        gi, gd = ctx.point_states(0); oi, od = ow.point_states()
        assert np.array_equal(gi[idx][:, [2, 3, 5]], oi[idx][:, [2, 3, 5]]), f"frame {k}"
        fnd = idx[oi[idx][:, 3] == 1]
        assert np.array_equal(gd[fnd][:, 2:4], od[fnd][:, 2:4])
        for j in idx[::97]:
            gt, gs, gq = ctx.point_template(0, int(j)); ot, os_, oq = ow.point_template(int(j))
            assert np.array_equal(gt, ot) and (gs, gq) == (os_, oq)
    ctx.close()


@pytest.mark.parametrize("sigma,mark", [(0.0, False), (16.0, True), (1.0, False)])
def test_calc_pose_update_matches_oracle(sigma, mark):
    cam, f0, smap = common.scene()
    ctx, ow = _ctx(cam, f0, smap), _orc(cam, f0, smap)
    f1, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.6)
    start = synth.se3_exp(np.array(synth.CONFIG1_TWIST) * 0.3)
    ctx.make_keyframe_lite(f1); ow.make_current_kf(f1)
    ctx.set_pose(0, start); ow.set_pose(start)
    ctx.project_all(); ow.L.orc_tracker_project_all(ow.tracker)
    oi, _ = ow.point_states()
    idx = np.nonzero(oi[:, 1] >= 0)[0].astype(np.int32)
    ctx.set_lists([idx])
    ctx.search_for_points(10, 0); ow.L.orc_tracker_search_for_points(ow.tracker, idx, len(idx), 10, 0)
    ctx.calc_jacobians(); ow.L.orc_tracker_calc_jacobians(ow.tracker, idx, len(idx))
    gu = ctx.calc_pose_update(sigma, mark, apply=True)[0]
    ou = np.zeros(6); ow.L.orc_tracker_calc_pose_update(ow.tracker, idx, len(idx), sigma, int(mark), 1, ou)
    assert np.linalg.norm(gu - ou) <= 1e-6 * max(np.linalg.norm(ou), 1e-12), (gu, ou)   # contract: 1e-4 relative
    assert np.abs(ctx.get_pose(0) - ow.get_pose()).max() <= 1e-9
    gi, gd = ctx.point_states(0); oi, od = ow.point_states()
    fnd = idx[oi[idx][:, 3] == 1]
    assert np.allclose(gd[fnd][:, 16:28], od[fnd][:, 16:28], rtol=1e-9, atol=1e-12), "Jacobians"
    if mark:
        assert np.array_equal(ctx.point_counts(0), ow.point_counts()), "inlier / outlier counters"
    _, gs = ctx.updates(0)
    if sigma == 0.0:
        e2 = ((od[fnd][:, 2:4] - od[fnd][:, 0:2]) * od[fnd][:, 15:16]) ** 2
        s2 = ow.L.orc_tukey_sigma_squared(np.ascontiguousarray(e2.sum(1)), len(fnd))
        assert abs(gs[-1] - s2) <= 1e-9 * s2, "Tukey sigma^2"
    ctx.close()


def test_calc_pose_update_degenerate_sets():
    """Zero found points -> zero update (jni/Tracker.cc:712-716); < 4 points exercises the size_t wrap of MEstimator.h:73."""
    cam, f0, smap = common.scene()
    ctx, ow = _ctx(cam, f0, smap), _orc(cam, f0, smap)
    f1, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.2)
    ctx.make_keyframe_lite(f1); ow.make_current_kf(f1)
    ctx.set_pose(0, synth.IDENTITY_POSE); ow.set_pose(synth.IDENTITY_POSE)
    ctx.project_all(); ow.L.orc_tracker_project_all(ow.tracker)
    oi, _ = ow.point_states()
    allidx = np.nonzero(oi[:, 1] >= 0)[0].astype(np.int32)
    for n in (0, 1, 2, 4):
        idx = allidx[:n].copy()
        ctx.set_lists([idx])
        if n:
            ctx.search_for_points(10, 0); ow.L.orc_tracker_search_for_points(ow.tracker, idx, n, 10, 0)
            ctx.calc_jacobians(); ow.L.orc_tracker_calc_jacobians(ow.tracker, idx, n)
        gu = ctx.calc_pose_update(0.0, False, apply=False)[0]
        ou = np.zeros(6); ow.L.orc_tracker_calc_pose_update(ow.tracker, idx if n else np.zeros(1, np.int32), n, 0.0, 0, 0, ou)
        assert np.allclose(gu, ou, rtol=1e-6, atol=1e-15), (n, gu, ou)
    ctx.close()


def _track_map_case(scale_start, scale_frame, velocity=None, P=11, n_points=1000, size=(640, 480), tex=2048):
    cam, f0, smap = common.scene(size[0], size[1], n_points, tex)
    assert smap.n == n_points
    ctx, ow = _ctx(cam, f0, smap, patch_size=P), _orc(cam, f0, smap, P=P)
    f1, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * scale_frame, tex)
    start = synth.se3_exp(np.array(synth.CONFIG1_TWIST) * scale_start)
    ctx.make_keyframe_lite(f1); ow.make_current_kf(f1)
    ctx.set_pose(0, start); ow.set_pose(start)
    if velocity is not None:
        ctx.set_motion(0, np.zeros(6), velocity); ow.L.orc_tracker_set_velocity(ow.tracker, np.zeros(6), velocity)
    ctx.track_map(); ow.L.orc_tracker_track_map(ow.tracker)
    return ctx, ow


def _check_track_map(ctx, ow):
    a, f, q, lost, dc = ctx.counters(0); oa, of, oq, olost, odc = ow.counters()
    assert np.array_equal(a, oa) and np.array_equal(f, of), "manMeasAttempted / manMeasFound"
    assert dc == odc, "mbDidCoarse"
    gi, gd = ctx.point_states(0); oi, od = ow.point_states()
    pvs = oi[:, 1] >= 0
    assert np.array_equal(gi[:, :2], oi[:, :2])
    assert np.array_equal(gi[pvs][:, [2, 3, 5]], oi[pvs][:, [2, 3, 5]]), "searched / found / templateBad per point"
    fnd = oi[:, 3] == 1
    assert np.abs(gd[fnd][:, 2:4] - od[fnd][:, 2:4]).max() <= 1e-6, "found positions (px)"
    gu, gs = ctx.updates(0); ou, os_ = ow.updates()
    assert len(gu) == len(ou)
    for k in range(len(ou)):   # contract: 1e-4 relative on each twist; we hold 1e-6 (plus an absolute floor for ~0 updates)
        assert np.linalg.norm(gu[k] - ou[k]) <= 1e-6 * np.linalg.norm(ou[k]) + 1e-12, (k, gu[k], ou[k])
        assert abs(gs[k] - os_[k]) <= 1e-9 * abs(os_[k])
    assert np.abs(ctx.get_pose(0) - ow.get_pose()).max() <= 1e-9, "final pose"
    assert np.array_equal(ctx.point_counts(0), ow.point_counts()), "inlier / outlier counters"
    # final reprojection residuals of the found points: 1e-4 relative contract, checked at 1e-6 px absolute
    rg = (gd[fnd][:, 2:4] - gd[fnd][:, 0:2]); ro = (od[fnd][:, 2:4] - od[fnd][:, 0:2])
    assert np.abs(rg - ro).max() <= 1e-6


def test_track_map_fine_only_matches_oracle():
    ctx, ow = _track_map_case(0.0, 1.0)     # start pose I, full twist
    _check_track_map(ctx, ow)
    ctx.close()


def test_track_map_with_coarse_stage_matches_oracle():
    ctx, ow = _track_map_case(0.0, 1.0, velocity=0.05)
    assert ow.counters()[4] == 1, "coarse stage must have run in this case"
    _check_track_map(ctx, ow)
    ctx.close()


def test_track_map_more_points_than_the_patch_cap():
    """2500 map points: the potentially-visible set exceeds MaxPatchesPerFrame = 1000, so TrackMap shuffles the remaining fine list
    once more and truncates it (jni/Tracker.cc:518-527) — the fifth std::random_shuffle of the frame."""
    ctx, ow = _track_map_case(0.0, 0.6, n_points=2500)
    a = ow.counters()[0]
    assert a.sum() <= 1000 and a.sum() > 900
    _check_track_map(ctx, ow)
    ctx.close()


@pytest.mark.parametrize("size,n_points,velocity", [((640, 480), 5000, None), ((1920, 1080), 12000, None), ((1920, 1080), 12000, 0.05)])
def test_track_map_large_maps_shuffle_without_the_swap_chain(size, n_points, velocity):
    """Maps of 5000 / 12000 points: the four level shuffles and the fifth one over ~N fine candidates (jni/Tracker.cc:396-397,518-527) run as
    shuffle_parallel -- scratch in shared memory at 5000 points, in the stream's global scratch at 12000 -- and the projection on its own grid
    (k_project_points).  Which points are searched (flags per point, counters per level) is the permutation's fingerprint: a wrong element
    among the first MaxPatchesPerFrame of the shuffled list, or among the coarse set's share of the level-3 / level-2 lists, changes it."""
    ctx, ow = _track_map_case(0.0, 0.6, velocity=velocity, n_points=n_points, size=size, tex=2048 if size[0] == 640 else 4096)
    a = ow.counters()[0]
    assert a.sum() > 900        # (at 12000 points the level-3 list alone exceeds MaxPatchesPerFrame and is searched in full, jni/Tracker.cc:499-527)
    if velocity:
        assert ow.counters()[4] == 1, "coarse stage must have run in this case"
    _check_track_map(ctx, ow)
    ctx.close()


@pytest.mark.parametrize("n_points", [90, 150, 300])
def test_track_map_small_maps_coarse_set_branches(n_points):
    """Coarse-stage list selection when level 3 has fewer points than CoarseMax (jni/Tracker.cc:425-462): level 2 tops the set up, or —
    the reference's quirk — REPLACES it when it fits entirely (150 and 90 points); after the search fewer than CoarseMin points may be found,
    in which case the coarse pose update is skipped."""
    ctx, ow = _track_map_case(0.0, 0.6, velocity=0.05, n_points=n_points)
    _check_track_map(ctx, ow)
    ctx.close()


def test_track_map_patch8():
    ctx, ow = _track_map_case(0.2, 0.5, P=8)
    _check_track_map(ctx, ow)
    ctx.close()


@pytest.mark.parametrize("groups", [1, 3])
def test_track_frame_sequence_multi_stream(groups):
    """4 streams x 6 frames through vslam_track_frame (motion model + TrackMap + quality), each against its own oracle tracker; also
    with the streams split into 3 groups on separate CUDA streams (vslam_params.stream_groups), which must not change any result."""
    from visualslam_android_b200 import api
    cam, f0, smap = common.scene()
    S, K = 4, 6
    ctx = _ctx(cam, f0, smap, n_streams=S)
    ctx.set_params(stream_groups=groups)
    ows = [_orc(cam, f0, smap) for _ in range(S)]
    for k in range(1, K + 1):
        frames = np.stack([synth.render_frame(common.texture(), cam, synth.stream_pose(4 * k, s)) for s in range(S)])
        ctx.track_frame(frames)
        for s, ow in enumerate(ows):
            ow.L.orc_tracker_track_frame(ow.tracker, np.ascontiguousarray(frames[s]), cam.width, cam.height, cam.width)
        poses = ctx.get_poses()
        for s, ow in enumerate(ows):
            assert np.abs(poses[s] - ow.get_pose()).max() <= 1e-8, (k, s)
            a, f, q, lost, dc = ctx.counters(s); oa, of, oq, olost, odc = ow.counters()
            assert np.array_equal(a, oa) and np.array_equal(f, of) and (q, lost, dc) == (oq, olost, odc), (k, s)
    ctx.close()


def test_track_frame_async_pipeline_matches_oracle():
    """vslam_track_frame_async (double-buffered level 0, copy stream, pose write-back) gives the same poses as the oracle."""
    import torch
    cam, f0, smap = common.scene()
    S, K = 2, 5
    ctx = _ctx(cam, f0, smap, n_streams=S)
    ows = [_orc(cam, f0, smap) for _ in range(S)]
    frames = torch.from_numpy(np.stack([np.stack([synth.render_frame(common.texture(), cam, synth.stream_pose(3 * k, s)) for s in range(S)])
                                        for k in range(1, K + 1)])).pin_memory()
    poses = torch.zeros((K, S, 12), dtype=torch.float64).pin_memory()
    ids = [ctx.track_frame_async(frames[k].data_ptr(), cam.width, cam.width * cam.height, poses[k].data_ptr()) for k in range(K)]
    for sid in ids:
        ctx.wait_step(sid)
    for k in range(K):
        for s, ow in enumerate(ows):
            ow.L.orc_tracker_track_frame(ow.tracker, np.ascontiguousarray(frames[k, s].numpy()), cam.width, cam.height, cam.width)
            assert np.abs(poses[k, s].numpy().reshape(3, 4) - ow.get_pose()).max() <= 1e-8, (k, s)
    # the synchronous entry point keeps working on the same context afterwards
    nxt = np.stack([synth.render_frame(common.texture(), cam, synth.stream_pose(3 * (K + 1), s)) for s in range(S)])
    ctx.track_frame(nxt)
    for s, ow in enumerate(ows):
        ow.L.orc_tracker_track_frame(ow.tracker, np.ascontiguousarray(nxt[s]), cam.width, cam.height, cam.width)
        assert np.abs(ctx.get_pose(s) - ow.get_pose()).max() <= 1e-8
    ctx.close()


@pytest.mark.parametrize("kind", ["synthetic", "noise"])
def test_make_keyframe_rest_bit_exact(kind):
    """a13: FAST scores -> non-max -> Shi-Tomasi candidates on the device against the oracle (ints exact, ST score exact: integer sums)."""
    from oracle import oraclebind
    from visualslam_android_b200 import api
    if kind == "synthetic":
        cam = synth.Camera(640, 480)
        im = synth.render_frame(common.texture(), cam, synth.se3_exp(np.array(synth.CONFIG1_TWIST) * 0.7))
    else:
        im = np.random.RandomState(9).randint(0, 256, (240, 320)).astype(np.uint8)
    H, W = im.shape
    ctx = api.Context(W, H, n_streams=2, max_points=8, max_corner_frac=1.0)
    ctx.make_keyframe_lite(np.stack([np.zeros_like(im), im]))
    ctx.make_keyframe_rest(1)
    ok = oraclebind.OrcKeyFrame().make_lite(im); ok.make_rest()
    for l in range(4):
        assert np.array_equal(ctx.max_corners(1, l), ok.max_corners(l)), f"vMaxCorners level {l}"
        gxy, gs = ctx.candidates(1, l); oxy, os_ = ok.candidates(l)
        assert np.array_equal(gxy, oxy), f"vCandidates level {l}"
        assert np.array_equal(gs, os_), f"Shi-Tomasi scores level {l}"
    ctx.make_keyframe_rest(0)       # an image without corners
    assert all(len(ctx.max_corners(0, l)) == 0 for l in range(4))
    ctx.close()


def test_minipatch_trail_tracking_bit_exact():
    """a12: SampleFromImage + FindPatch forward (current frame) and backward (snapshot of the previous frame), as in
    Tracker::TrailTracking_Advance (jni/Tracker.cc:294-346), against the oracle."""
    import ctypes as C
    from oracle import oraclebind
    cam, f0, smap = common.scene()
    f1, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.15)
    ctx = _ctx(cam, f0, smap)
    ok0 = oraclebind.OrcKeyFrame().make_lite(f0); ok1 = oraclebind.OrcKeyFrame().make_lite(f1)
    L = oraclebind.lib()
    ctx.make_keyframe_lite(f0)
    ctx.snapshot_keyframe(0)                     # previous frame
    c = ok0.corners(0)
    c = c[(c[:, 0] > 10) & (c[:, 1] > 10) & (c[:, 0] < 630) & (c[:, 1] < 470)][::9]
    patches = ctx.minipatch_sample(0, c)
    for k in range(0, len(c), 17):
        x, y = c[k]
        assert np.array_equal(patches[k], f0[y - 4:y + 5, x - 4:x + 5])
    ctx.make_keyframe_lite(f1)                   # current frame
    for rng, max_ssd in ((10, 100000), (10, 2000), (3, 100000)):
        pos, found, best = ctx.minipatch_find(0, patches, c.astype(np.float64), rng, max_ssd)
        for k, (x, y) in enumerate(c):
            po = np.array([x, y], float); b = C.c_int()
            fo = L.orc_minipatch_find(ok0.h, int(x), int(y), ok1.h, po, rng, 1, max_ssd, C.byref(b))
            assert fo == found[k] and b.value == best[k] and np.array_equal(po, pos[k]), (rng, max_ssd, k)
        assert found.sum() > 20
    # backward search in the snapshot (which = 1): patches sampled at the forward result in the current frame
    pos, found, _ = ctx.minipatch_find(0, patches, c.astype(np.float64), 10, 100000)
    fw = pos[found == 1].astype(np.int32)
    fw = fw[(fw[:, 0] >= 4) & (fw[:, 1] >= 4) & (fw[:, 0] < 636) & (fw[:, 1] < 476)]
    back_patches = ctx.minipatch_sample(0, fw)
    bpos, bfound, bbest = ctx.minipatch_find(0, back_patches, fw.astype(np.float64), 10, 100000, which=1)
    for k, (x, y) in enumerate(fw):
        po = np.array([x, y], float); b = C.c_int()
        fo = L.orc_minipatch_find(ok1.h, int(x), int(y), ok0.h, po, 10, 1, 100000, C.byref(b))
        assert fo == bfound[k] and b.value == bbest[k] and np.array_equal(po, bpos[k])
    ctx.close()


@pytest.mark.parametrize("size,n_points", [((1920, 1080), 5000), ((3840, 2160), 20000)])
def test_large_configs_keyframe_search_pose(size, n_points):
    """BASELINE configs[2] and [4] at full size: 1080p / 5000 points and 4K / 20000 points.  MakeKeyFrame_Lite is compared with
    the oracle bit for bit; SearchForPoints on every level (8 sub-pixel iterations, beyond the 1000-patch cap, so driven
    directly) and one CalcPoseUpdate are compared with the oracle; plus size-independent properties of the corner list."""
    W, H = size
    cam, f0, smap = common.scene(W, H, n_points, 4096)
    assert smap.n == n_points
    ctx, ow = _ctx(cam, f0, smap), _orc(cam, f0, smap)
    tw = np.array(synth.CONFIG1_TWIST) * 0.3
    f1, _ = common.frame_at(cam, tw, 4096)
    ctx.make_keyframe_lite(f1); okf = ow.make_current_kf(f1)
    _check_keyframe(ctx, 0, okf)
    for l in range(4):   # properties that do not need the oracle: raster order, LUT = exclusive row histogram, x/y ranges
        c = ctx.corners(0, l); w, h = ctx.level_dims(l)
        key = c[:, 1].astype(np.int64) * 65536 + c[:, 0]
        assert np.all(np.diff(key) > 0) and c[:, 0].min() >= 3 and c[:, 0].max() <= w - 4 and c[:, 1].min() >= 3 and c[:, 1].max() <= h - 4
        assert np.array_equal(ctx.row_lut(0, l), np.searchsorted(c[:, 1], np.arange(h), side="left"))
    start = synth.se3_exp(tw * 0.6)
    ctx.set_pose(0, start); ow.set_pose(start)
    ctx.project_all(); ow.L.orc_tracker_project_all(ow.tracker)
    oi, od = ow.point_states()
    gi, gd = ctx.point_states(0)
    assert np.array_equal(gi[:, :2], oi[:, :2])
    ctx.set_point_projection(0, od[:, 0:2], od[:, 11:15], oi[:, 1])
    idx = np.nonzero(oi[:, 1] >= 0)[0].astype(np.int32)
    assert len(idx) > 0.9 * n_points
    ctx.set_lists([idx]); ctx.clear_counters(); ow.L.orc_tracker_clear_counters(ow.tracker)
    ctx.search_for_points(10, 8); ow.L.orc_tracker_search_for_points(ow.tracker, idx, len(idx), 10, 8)
    gi, gd = ctx.point_states(0); oi, od = ow.point_states()
    assert np.array_equal(gi[idx][:, [2, 3, 5]], oi[idx][:, [2, 3, 5]])
    a, f, *_ = ctx.counters(0); oa, of, *_ = ow.counters()
    assert np.array_equal(a, oa) and np.array_equal(f, of)
    assert ctx.zmssd_evals() == ow.L.orc_tracker_zmssd_evals(ow.tracker)
    fnd = idx[oi[idx][:, 3] == 1]
    assert len(fnd) > 0.8 * len(idx)
    assert np.array_equal(gd[fnd][:, 30:32], od[fnd][:, 30:32]) and np.abs(gd[fnd][:, 2:4] - od[fnd][:, 2:4]).max() <= 1e-6
    ctx.calc_jacobians(); ow.L.orc_tracker_calc_jacobians(ow.tracker, idx, len(idx))
    gu = ctx.calc_pose_update(0.0, True, apply=True)[0]
    ou = np.zeros(6); ow.L.orc_tracker_calc_pose_update(ow.tracker, idx, len(idx), 0.0, 1, 1, ou)
    assert np.linalg.norm(gu - ou) <= 1e-6 * np.linalg.norm(ou), (gu, ou)     # contract 1e-4
    assert np.array_equal(ctx.point_counts(0), ow.point_counts())
    ctx.close()


def test_long_sequence_stays_consistent_with_oracle():
    """60 frames of vslam_track_frame.  From identical state one frame agrees to 1e-9 (tests above); over a long sequence the
    device's sin/cos and summation order differ from the host's in the last bit, and the reference's own discontinuities
    (truncated template pixels, the (int) residual cast, Tukey cut-off) amplify that — see DESIGN.md §5.  What must hold:
    both trackers stay locked, find the same number of points within 2 %, and stay as close to each other as to the truth."""
    cam, f0, smap = common.scene()
    ctx, ow = _ctx(cam, f0, smap), _orc(cam, f0, smap)
    worst = 0.0
    for k in range(1, 61):
        fr = synth.render_frame(common.texture(), cam, synth.stream_pose(k, 0))
        ctx.track_frame(fr[None]); ow.L.orc_tracker_track_frame(ow.tracker, fr, cam.width, cam.height, cam.width)
        gp, op, truth = ctx.get_pose(0), ow.get_pose(), synth.stream_pose(k, 0)
        worst = max(worst, np.abs(gp - op).max())
        assert np.abs(gp - truth).max() < 5e-3 and np.abs(op - truth).max() < 5e-3, k
        a, f, q, lost, _ = ctx.counters(0); oa, of, oq, olost, _ = ow.counters()
        assert q == oq == 2 and lost == olost == 0, k
        assert abs(int(f.sum()) - int(of.sum())) <= 0.02 * of.sum(), k
    assert worst < 2e-3, worst
    ctx.close()


@pytest.mark.parametrize("size", [(640, 480), (640, 360), (320, 200)])
def test_track_frame_with_on_device_sbi(size):
    """f1: SmallBlurryImage + CalcSBIRotation on the device: vslam_track_frame is then the reference's whole TrackFrame (good-map
    branch).  The oracle side is the restatement that tests/test_oracle_vs_ref.py pins bit-for-bit to the unmodified TrackFrame.
    640 x 360 and 320 x 200 have odd level-3 heights (45, 25): cv::resize to (cols / 2, rows / 2) is then OpenCV's general fixed-point
    bilinear, not the exact half-sample -- the 1080p case (240 x 135 -> 120 x 67)."""
    cam, f0, smap = common.scene(width=size[0], height=size[1], n_points=1000 if size[0] == 640 else 400)
    S, K = 3, 6
    sbi_cam = synth.Camera(cam.width // 16, cam.height // 16)
    ctx = _ctx(cam, f0, smap, n_streams=S)
    ctx.enable_sbi(sbi_cam.scalars())
    ows = [_orc(cam, f0, smap) for _ in range(S)]
    for ow in ows:
        ow.L.orc_tracker_enable_sbi(ow.tracker, sbi_cam.scalars())
    big = 0.0
    for k in range(1, K + 1):
        frames = np.stack([synth.render_frame(common.texture(), cam, synth.stream_pose(5 * k, s + 1)) for s in range(S)])
        ctx.track_frame(frames)
        for s, ow in enumerate(ows):
            ow.L.orc_tracker_track_frame(ow.tracker, np.ascontiguousarray(frames[s]), cam.width, cam.height, cam.width)
            orot = np.zeros(6); ow.L.orc_tracker_get_sbi_rot(ow.tracker, orot)
            grot = ctx.get_sbi_rotation(s)
            assert np.abs(grot - orot).max() <= 1e-9, (k, s, grot, orot)     # rotation vector (rad)
            big = max(big, np.abs(orot[3:]).max())
            assert np.abs(ctx.get_pose(s) - ow.get_pose()).max() <= 1e-8, (k, s)
            a, f, q, lost, dc = ctx.counters(s); oa, of, oq, olost, odc = ow.counters()
            assert np.array_equal(a, oa) and np.array_equal(f, of) and (q, lost, dc) == (oq, olost, odc), (k, s)
    assert big > 1e-4, "the sequences must contain a measurable inter-frame rotation"
    ctx.close()


@pytest.mark.parametrize("P", [11, 8])
def test_refind_common_batched_over_keyframes(P):
    """f3: MapMaker::ReFind_Common (jni/MapMaker.cc:967-1036) batched — every stream is one keyframe (own image, own pose), the list
    names the map points to re-find; level, found flag, bSubPix and v2RootPos against the oracle (pinned to the reference's calls in
    tests/test_oracle_vs_ref.py::test_refind_common).  The third keyframe is so close to the plane that warps are rejected."""
    cam, f0, smap = common.scene()
    twists = [np.array(synth.CONFIG1_TWIST), np.array([0.05, -0.03, 0.45, 0.02, -0.03, 0.3]), np.array([0.0, 0.0, -0.935, 0.0, 0.0, 0.0])]
    off = synth.se3_exp(np.array([0.0008, -0.0006, 0.0005, 0.0006, -0.0004, 0.0007]))
    S = len(twists)
    ctx = _ctx(cam, f0, smap, n_streams=S, patch_size=P)
    frames, poses = [], []
    for tw in twists:
        fr, pose = common.frame_at(cam, tw)
        frames.append(fr); poses.append((np.vstack([off, [0, 0, 0, 1]]) @ np.vstack([pose, [0, 0, 0, 1]]))[:3])
    ctx.make_keyframe_lite(np.stack(frames))
    idx = np.arange(smap.n, dtype=np.int32)
    for s in range(S):
        ctx.set_pose(s, poses[s])
    ctx.set_lists([idx] * S)
    ctx.refind(4, 8)
    for s in range(S):
        ow = _orc(cam, f0, smap, P=P)
        ow.make_current_kf(frames[s]); ow.set_pose(poses[s])
        oo, op = np.zeros((smap.n, 3), dtype=np.int32), np.zeros((smap.n, 2))
        ow.L.orc_tracker_refind(ow.tracker, idx, smap.n, 4, 8, 1, oo, op)
        fl, pos = ctx.refind_results(s, smap.n)
        searched = oo[:, 1] >= 0                      # (level is only reported for points that reach the template step)
        assert np.array_equal(fl[searched], oo[searched]), s
        assert np.array_equal(fl[~searched, 0], oo[~searched, 0]) and fl[~searched, 0].sum() == 0
        f = oo[:, 0] == 1
        coarse = f & (oo[:, 2] == 0)
        assert np.array_equal(pos[coarse], op[coarse]), s
        sub = f & (oo[:, 2] == 1)
        if sub.any():
            assert np.abs(pos[sub] - op[sub]).max() <= 1e-6, s
        if s < 2:
            assert f.sum() > 300 and sub.sum() > 50 and coarse.sum() > 50
        else:
            assert (oo[:, 1] == 3).sum() > 5
    ctx.close()


@pytest.mark.parametrize("P", [11, 8])
def test_epipolar_search_matches_the_oracle(P):
    """f3: the search of MapMaker::AddPointEpipolar (jni/MapMaker.cc:525-640): every Shi-Tomasi candidate of every level of source
    keyframe 0 searched along its epipolar line in the stream's current keyframe; found flag, best corner and its ZMSSD exact,
    refined position to 1e-6 px, against the oracle (pinned in tests/test_oracle_vs_ref.py::test_epipolar_search)."""
    from oracle import oraclebind
    cam, f0, smap = common.scene()
    tw = np.array([0.12, 0.03, 0.02, 0.01, -0.03, 0.02])
    f1, pose1 = common.frame_at(cam, tw)
    ctx = _ctx(cam, f0, smap, patch_size=P)
    ctx.make_keyframe_lite(f0)
    ctx.make_keyframe_rest(0)
    cands = [ctx.candidates(0, l)[0] for l in range(4)]
    ok0 = oraclebind.OrcKeyFrame().make_lite(f0); ok0.make_rest()
    ok1 = oraclebind.OrcKeyFrame().make_lite(f1)
    ctx.make_keyframe_lite(f1)
    ow = _orc(cam, f0, smap, P=P)
    eye = np.ascontiguousarray(synth.IDENTITY_POSE, dtype=np.float64).reshape(12); p1 = np.ascontiguousarray(pose1, dtype=np.float64).reshape(12)
    total_found = 0
    for level in range(4):
        xy = cands[level]
        assert np.array_equal(xy, ok0.candidates(level)[0])
        for mean, sigma, wig in ((1.0, 0.3, 0.1), (1.4, 0.2, 0.1)):
            found, pos, bi, bs = ctx.epipolar_search(0, 0, level, xy, eye, p1, mean, sigma, wig)
            step = max(1, len(xy) // 150)
            for k in range(0, len(xy), step):
                oo, op = np.zeros(3, dtype=np.int32), np.zeros(2)
                ow.L.orc_epipolar_search(ow.tracker, ok0.h, ok1.h, eye, p1, mean, sigma, wig, level, int(xy[k, 0]), int(xy[k, 1]), oo, op, None)
                assert (found[k], bi[k]) == (oo[0], oo[1]), (level, k, found[k], bi[k], bs[k], oo)
                if oo[1] >= 0:
                    assert bs[k] == oo[2]
                    assert np.abs(pos[k] - op).max() <= 1e-6, (level, k)
            total_found += int(found.sum())
    assert total_found > 200
    ctx.close()


def test_epipolar_new_points_are_triangulated_and_trackable():
    """f3 tail (jni/MapMaker.cc:646-690): candidates found by the epipolar search become map points -- world position by
    MapMaker::ReprojectPoint (restated with LAPACK's SVD in oracle/oraclebind.py; 1e-9 relative, the 4x4 SVD being third-party
    arithmetic), patch-source fields + RefreshPixelVectors against the restatement that tests/test_oracle_vs_ref.py pins bit for bit
    (1e-12: host arithmetic, same operations) -- and a map made ONLY of those new points tracks a third frame."""
    from oracle import oraclebind
    cam, f0, smap = common.scene()
    src_pose = synth.se3_exp(np.array([0.02, -0.01, 0.03, 0.01, 0.02, -0.01]))
    rel = synth.se3_exp(np.array([0.12, 0.03, 0.02, 0.01, -0.03, 0.02]))
    f0s = synth.render_frame(common.texture(), cam, src_pose)
    tgt_pose = rel @ np.vstack([src_pose, [0, 0, 0, 1]])
    f1 = synth.render_frame(common.texture(), cam, tgt_pose)
    ctx = _ctx(cam, f0s, smap)
    ctx.make_keyframe_lite(f0s); ctx.make_keyframe_rest(0)
    cands = [ctx.candidates(0, l)[0] for l in range(4)]
    ctx.make_keyframe_lite(f1)
    cam13 = np.ascontiguousarray(cam.scalars(), dtype=np.float64)
    new = [[] for _ in range(5)]
    for level in range(4):
        xy = cands[level]
        found, pos, _, _ = ctx.epipolar_search(0, 0, level, xy, src_pose, tgt_pose, 1.0, 0.3, 0.1)
        sel = np.nonzero(found)[0]
        world, right, down, irc, lvl = ctx.epipolar_make_points(level, xy[sel], pos[sel], src_pose, tgt_pose)
        assert np.array_equal(irc, xy[sel]) and np.all(lvl == level)
        for j in range(0, len(sel), max(1, len(sel) // 40)):
            k = sel[j]
            root = (xy[k] + 0.5) * (1 << level) - 0.5
            ow_ = oraclebind.triangulate(cam13, src_pose, tgt_pose, root, pos[k])
            assert np.abs(world[j] - ow_).max() <= 1e-9 * max(1.0, np.abs(ow_).max()), (level, k, world[j], ow_)
            fields = oraclebind.epipolar_point_fields(cam13, src_pose, level, xy[k, 0], xy[k, 1], world[j])
            assert np.abs(right[j] - fields[3]).max() <= 1e-12 and np.abs(down[j] - fields[4]).max() <= 1e-12, (level, k)
        in_src = world @ src_pose[:, :3].T + src_pose[:, 3]
        good = np.abs(in_src[:, 2] - np.median(in_src[:, 2])) < 0.1      # drop gross mismatches, as the bundle adjuster's outlier handling would
        for a, v in zip(new, (world[good], right[good], down[good], irc[good], lvl[good])):
            a.append(v)
    world, right, down, irc, lvl = [np.concatenate(a) for a in new]
    assert len(world) > 150
    ctx.close()
    # a fresh tracker whose whole map is the new points (source keyframe = f0s at src_pose) follows the camera to a third view
    from visualslam_android_b200 import api
    ctx = api.Context(cam.width, cam.height, n_streams=1, max_points=len(world))
    ctx.set_camera(cam.scalars()); ctx.upload_source_keyframe(f0s)
    ctx.set_map(world, right, down, irc, lvl)
    third = synth.se3_exp(np.array([0.02, 0.01, 0.005, 0.003, -0.008, 0.005])) @ np.vstack([src_pose, [0, 0, 0, 1]])
    ctx.set_pose(0, src_pose)
    ctx.track_frame(synth.render_frame(common.texture(), cam, third)[None])
    att, fnd, q, lost, dc = ctx.counters(0)
    assert q == 2 and fnd.sum() > 0.6 * att.sum() and att.sum() > 100, (att, fnd, q)
    assert np.abs(ctx.get_pose(0) - third).max() < 0.01, ctx.get_pose(0) - third
    ctx.close()


def test_track_frame_relocalises_lost_streams():
    """f4: the lost branch of Tracker::TrackFrame on the device (k_relocalise: Relocaliser::AttemptRecovery + Tracker::AttemptRecovery,
    then TrackMap with the doubled coarse stage and AssessTrackingQuality in the same frame).  Two streams see different sequences
    (noise frames that make them lose track at different times, then frames near different map keyframes); poses, counters, quality
    and the chosen keyframe against the oracle, which tests/test_oracle_vs_ref.py pins bit-for-bit to the unmodified TrackFrame."""
    from oracle import oraclebind
    cam, f0, smap = common.scene()
    sbi_cam = synth.Camera(cam.width // 16, cam.height // 16).scalars()
    kf_twists = [np.zeros(6), np.array([0.10, 0.02, 0.01, 0.01, -0.04, 0.05]), np.array([-0.08, -0.05, 0.02, -0.02, 0.03, -0.06])]
    kf_frames, kf_poses = zip(*[common.frame_at(cam, tw) for tw in kf_twists])
    S = 2
    ctx = _ctx(cam, f0, smap, n_streams=S, max_source_keyframes=3)
    ctx.enable_sbi(sbi_cam)
    for k in (1, 2):
        ctx.upload_source_keyframe(kf_frames[k], k)
    ctx.set_reloc_keyframes([0, 1, 2], np.stack(kf_poses))
    ows = []
    for s in range(S):
        ow = _orc(cam, f0, smap)
        ow.L.orc_tracker_enable_sbi(ow.tracker, sbi_cam)
        keep = []
        for k in range(3):
            okf = oraclebind.OrcKeyFrame().make_lite(kf_frames[k]); keep.append(okf)
            ow.L.orc_tracker_add_reloc_keyframe(ow.tracker, okf.h, np.ascontiguousarray(kf_poses[k], dtype=np.float64).reshape(12))
        ow._kfs = keep
        ows.append(ow)
    rs = np.random.RandomState(5)
    noise = lambda: rs.randint(0, 255, f0.shape).astype(np.uint8)
    near = lambda k, d: common.frame_at(cam, kf_twists[k] + np.array(d))[0]
    first = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.2)[0]
    seqs = [[first] + [noise() for _ in range(4)] + [near(1, [0.004, -0.003, 0.002, 0.01, 0.008, -0.012]), near(1, [0.006, -0.002, 0.002, 0.012, 0.006, -0.01])]
            + [noise() for _ in range(4)] + [near(2, [-0.003, 0.004, 0.001, -0.008, 0.01, 0.009])],
            [first, near(0, [0.01, 0.0, 0.0, 0.0, 0.0, 0.01])] + [noise() for _ in range(5)] + [near(2, [0.002, 0.001, -0.002, 0.006, -0.004, 0.008])]
            + [near(2, [0.004, 0.002, -0.002, 0.008, -0.002, 0.006]), near(2, [0.006, 0.002, -0.001, 0.009, 0.0, 0.004]), near(2, [0.008, 0.003, 0.0, 0.01, 0.002, 0.002]), near(2, [0.009, 0.003, 0.0, 0.011, 0.003, 0.001])]]
    assert len(seqs[0]) == len(seqs[1])
    recovered = [0, 0]
    for k in range(len(seqs[0])):
        frames = np.stack([np.ascontiguousarray(seqs[s][k]) for s in range(S)])
        ctx.track_frame(frames)
        for s, ow in enumerate(ows):
            ow.L.orc_tracker_track_frame(ow.tracker, frames[s], cam.width, cam.height, cam.width)
            assert np.abs(ctx.get_pose(s) - ow.get_pose()).max() <= 1e-8, (k, s)
            a, f, q, lost, dc = ctx.counters(s); oa, of, oq, olost, odc = ow.counters()
            assert np.array_equal(a, oa) and np.array_equal(f, of) and (q, lost, dc) == (oq, olost, odc), (k, s)
            best, score, nrec = C.c_int(), C.c_double(), C.c_int()
            ow.L.orc_tracker_reloc_info(ow.tracker, C.byref(best), C.byref(score), C.byref(nrec))
            gb, gs, gn, gr = ctx.reloc_info(s)
            assert gn == nrec.value, (k, s)
            if nrec.value:
                assert gb == best.value and abs(gs - score.value) <= 1e-6 * max(1.0, abs(score.value)), (k, s)
            recovered[s] = gn
    assert recovered[0] >= 2 and recovered[1] >= 1
    assert ctx.counters(0)[2] == 2 and ctx.counters(1)[2] == 2       # both streams end with quality GOOD
    ctx.close()


@pytest.mark.parametrize("entry", ["device", "host", "async"])
def test_frame_lookahead_changes_no_result(entry):
    """vslam_params.frame_lookahead (two frame sets; the pyramid / FAST / SmallBlurryImage front end of frame k + 1 on a second CUDA stream
    beside the projection / search / pose back end of frame k) is execution only: the same frames issued back to back WITHOUT any
    synchronisation give bit-identical poses, counters, update twists and per-point state with it and without it -- through all three
    vslam_track_frame* entry points, with streams that lose track (noise frames), relocalise against registered keyframes and carry on,
    and with other calls (getters, a keyframe build) in between, after which the look-ahead has to fall back to a full barrier."""
    import torch
    cam, f0, smap = common.scene()
    sbi_cam = synth.Camera(cam.width // 16, cam.height // 16).scalars()
    kf_twists = [np.zeros(6), np.array([0.10, 0.02, 0.01, 0.01, -0.04, 0.05])]
    kf_frames, kf_poses = zip(*[common.frame_at(cam, tw) for tw in kf_twists])
    S, K = 6, 14
    rs = np.random.RandomState(11)
    seq = np.empty((K, S) + f0.shape, dtype=np.uint8)
    for k in range(K):
        for s in range(S):
            if s in (1, 4) and 2 + (s == 4) <= k < 6 + (s == 4):
                seq[k, s] = rs.randint(0, 255, f0.shape).astype(np.uint8)                 # four frames of noise: the stream gets lost ...
            elif s in (1, 4) and k >= 6 + (s == 4):
                seq[k, s] = common.frame_at(cam, kf_twists[1] + 0.002 * (k - 5) * np.array([1.0, -0.5, 0.5, 2.0, 1.5, -2.0]))[0]   # ... and relocalises near keyframe 1
            else:
                seq[k, s] = synth.render_frame(common.texture(), cam, synth.stream_pose(3 * (k + 1), s))

    def run(lookahead, coarse_chain=0):
        ctx = _ctx(cam, f0, smap, n_streams=S, max_source_keyframes=2)
        ctx.set_params(frame_lookahead=lookahead, coarse_chain=coarse_chain)
        ctx.enable_sbi(sbi_cam)
        ctx.upload_source_keyframe(kf_frames[1], 1)
        ctx.set_reloc_keyframes([0, 1], np.stack(kf_poses))
        W, H = cam.width, cam.height
        mid = {}
        if entry == "device":
            dev = torch.from_numpy(seq).cuda()
            for k in range(K):
                ctx.track_frame_ptr(dev[k].data_ptr(), W, W * H, device=True)
                if k == 8:      # other work between two frames: getters (synchronising) and a keyframe build on the current frame set (kernel launches)
                    mid["poses"] = ctx.get_poses().copy()
                    mid["corners"] = [ctx.corners(0, l).copy() for l in range(4)]
                    mid["level"] = ctx.level(2, 1).copy()
                    ctx.make_keyframe_lite(seq[k])
        elif entry == "host":
            for k in range(K):
                ctx.track_frame(seq[k])
        else:
            pin = torch.from_numpy(seq).pin_memory()
            poses = torch.zeros((K, S, 12), dtype=torch.float64).pin_memory()
            ids = [ctx.track_frame_async(pin[k].data_ptr(), W, W * H, poses[k].data_ptr()) for k in range(K)]
            for sid in ids:
                ctx.wait_step(sid)
            mid["poses_per_step"] = poses.numpy().copy()
        out = dict(poses=ctx.get_poses().copy(), counters=[ctx.counters(s) for s in range(S)], updates=[ctx.updates(s) for s in range(S)],
                   states=[ctx.point_states(s) for s in range(S)], reloc=[ctx.reloc_info(s) for s in range(S)], sbi=[ctx.get_sbi_rotation(s) for s in range(S)],
                   levels=[ctx.level(s, l).copy() for s in (0, S - 1) for l in range(4)], mid=mid)
        ctx.close()
        return out

    a, b = run(0), run(1)

    def same(x, y, what):
        if isinstance(x, dict):
            assert x.keys() == y.keys(), what
            for k in x:
                same(x[k], y[k], f"{what}.{k}")
        elif isinstance(x, (list, tuple)):
            assert len(x) == len(y), what
            for i, (u, v) in enumerate(zip(x, y)):
                same(u, v, f"{what}[{i}]")
        elif isinstance(x, np.ndarray):
            assert np.array_equal(x, y), what
        else:
            assert x == y, (what, x, y)
    same(a, b, "result")
    # vslam_params.coarse_chain (the coarse-stage streams of a frame on a launch chain of their own beside the fine-only streams) is execution
    # only as well: forced on without and with look-ahead, and left to the library (-1: the layout follows an unsynchronised hint)
    same(a, run(0, 1), "coarse_chain=1")
    same(a, run(1, 1), "coarse_chain=1 with look-ahead")
    same(a, run(1, -1), "coarse_chain=-1 with look-ahead")
    assert sum(r[2] for r in a["reloc"]) >= 2, a["reloc"]                  # the two noisy streams did relocalise
    assert all(c[2] == 2 for c in a["counters"]), a["counters"]            # and every stream ends with quality GOOD


def test_coarse_chain_layout_follows_the_hint_and_changes_no_result():
    """48 streams (the smallest context whose default is two launch chains, vslam_params.coarse_chain = -1): half of them speed up for a few frames --
    they enter TrackMap's coarse stage, the layout hint flips, the library falls back to one chain -- and slow down again.  Poses, counters, update twists
    and per-point state are bit-identical to a run with one chain throughout, frame by frame; the fast streams did run the coarse stage, the slow ones
    never did.  (No synchronisation is needed for the hint: it is read after each frame's getters here, and unsynchronised in the bench.)"""
    cam, f0, smap = common.scene()
    S, K, D = 48, 12, 6                                       # D distinct renderings, each tracked by S / D streams
    steps = [[2] * K if d % 2 == 0 else [2, 2, 12, 12, 12, 12, 2, 2, 2, 2, 2, 2] for d in range(D)]
    seq = np.empty((K, S) + f0.shape, dtype=np.uint8)
    for d in range(D):
        p = 0
        for k in range(K):
            p += steps[d][k]
            seq[k, d::D] = synth.render_frame(common.texture(), cam, synth.stream_pose(p, d))

    def run(coarse_chain):
        ctx = _ctx(cam, f0, smap, n_streams=S)
        ctx.set_params(coarse_chain=coarse_chain)
        per_frame = []
        for k in range(K):
            ctx.track_frame(seq[k])
            per_frame.append((ctx.get_poses().copy(), [ctx.counters(s) for s in range(D)]))
        out = dict(per_frame=per_frame, updates=[ctx.updates(s) for s in range(S)], states=[ctx.point_states(s) for s in range(0, S, 7)])
        ctx.close()
        return out

    a, b = run(0), run(-1)
    for k in range(K):
        assert np.array_equal(a["per_frame"][k][0], b["per_frame"][k][0]), k
        for d in range(D):
            for u, v in zip(a["per_frame"][k][1][d], b["per_frame"][k][1][d]):
                assert np.array_equal(u, v), (k, d)
    for u, v in zip(a["updates"], b["updates"]):
        assert np.array_equal(u[0], v[0]) and np.array_equal(u[1], v[1])
    for u, v in zip(a["states"], b["states"]):
        assert np.array_equal(u[0], v[0]) and np.array_equal(u[1], v[1])
    did = np.array([[a["per_frame"][k][1][d][4] for d in range(D)] for k in range(K)])          # mbDidCoarse per frame and rendering
    assert did[:, 0::2].sum() == 0 and np.all(did[:, 1::2].sum(axis=0) >= 2) and np.all(did[-1] == 0), did
    assert all(c[2] == 2 for c in a["per_frame"][-1][1])                                          # every stream ends with quality GOOD


@pytest.mark.parametrize("n_points", [1, 3, 25])
def test_track_frame_tiny_maps_and_blank_frames(n_points):
    """Edge cases of the whole TrackFrame: maps of 1 / 3 / 25 points (Tukey's `n*2-6` wraps or divides by zero, jni/MEstimator.h:73;
    no coarse stage; quality falls to BAD), a blank frame (no FAST corner on any level) and a saturated one, against the oracle."""
    cam, f0, smap_full = common.scene()
    keep = np.arange(smap_full.n)[:: max(1, smap_full.n // n_points)][:n_points]
    smap = synth.SyntheticMap(**{k: getattr(smap_full, k)[keep] for k in ("world", "pix_right_w", "pix_down_w", "ir_center", "src_level", "center_nc", "one_right_nc", "one_down_nc")})
    ctx = _ctx(cam, f0, smap)
    ow = _orc(cam, f0, smap)
    frames = [common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.3)[0], np.zeros_like(f0), common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.5)[0],
              np.full_like(f0, 255), common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.6)[0]]
    for k, fr in enumerate(frames):
        ctx.track_frame(fr[None])
        ow.L.orc_tracker_track_frame(ow.tracker, np.ascontiguousarray(fr), cam.width, cam.height, cam.width)
        gp, op = ctx.get_pose(0), ow.get_pose()
        assert np.array_equal(np.isfinite(gp), np.isfinite(op)), k
        fin = np.isfinite(op)
        assert np.abs(gp[fin] - op[fin]).max() <= 1e-8 if fin.any() else True, (k, gp, op)
        a, f, q, lost, dc = ctx.counters(0); oa, of, oq, olost, odc = ow.counters()
        assert np.array_equal(a, oa) and np.array_equal(f, of) and (q, lost, dc) == (oq, olost, odc), (k, n_points)
        if k in (1, 3):
            assert sum(int(ctx.corners(0, l).shape[0]) for l in range(4)) == 0
    ctx.close()


def test_full_bench_size_256_streams_replicas_and_oracle():
    """BASELINE configs[3] at full size: 256 VGA streams x 1000 points through vslam_track_frame_dev (device-resident frames, SBI on),
    three frames.  Eight distinct camera sequences are replicated 32 times across the streams (interleaved), so that (a) every replica
    must agree with the others BIT FOR BIT — poses, counters, corner lists — whatever SM it ran on and whatever its neighbours did,
    and (b) one replica of each sequence is compared with the oracle."""
    import torch
    cam, f0, smap = common.scene()
    S, D, K = 256, 8, 3
    sbi_cam = synth.Camera(cam.width // 16, cam.height // 16).scalars()
    ctx = _ctx(cam, f0, smap, n_streams=S)
    ctx.enable_sbi(sbi_cam)
    ows = []
    for d in range(D):
        ow = _orc(cam, f0, smap); ow.L.orc_tracker_enable_sbi(ow.tracker, sbi_cam); ows.append(ow)
    for k in range(1, K + 1):
        distinct = np.stack([synth.render_frame(common.texture(), cam, synth.stream_pose(5 * k, d + 1)) for d in range(D)])
        frames = torch.from_numpy(distinct[np.arange(S) % D].copy()).cuda()
        ctx.track_frame_ptr(frames.data_ptr(), cam.width, cam.width * cam.height, device=True)
        poses = ctx.get_poses()
        for d in range(D):
            reps = poses[d::D]
            assert np.array_equal(reps, np.broadcast_to(reps[0], reps.shape)), (k, d)       # bitwise identical replicas
            ows[d].L.orc_tracker_track_frame(ows[d].tracker, np.ascontiguousarray(distinct[d]), cam.width, cam.height, cam.width)
            assert np.abs(poses[d] - ows[d].get_pose()).max() <= 1e-8, (k, d)
            a, f, q, lost, dc = ctx.counters(d); oa, of, oq, olost, odc = ows[d].counters()
            assert np.array_equal(a, oa) and np.array_equal(f, of) and (q, lost, dc) == (oq, olost, odc), (k, d)
            for s in (d + D * 7, d + D * 31):
                assert ctx.counters(s)[0].tolist() == a.tolist() and ctx.counters(s)[1].tolist() == f.tolist()
        if k == K:
            for l in range(4):
                c0 = ctx.corners(3, l)
                assert np.array_equal(c0, ctx.corners(3 + D * 17, l)) and np.array_equal(ctx.row_lut(3, l), ctx.row_lut(3 + D * 30, l))
                order = c0[:, 1].astype(np.int64) * 65536 + c0[:, 0]
                assert np.all(np.diff(order) > 0)                                            # raster order, no duplicates
    ctx.close()


def test_track_map_with_two_source_keyframes():
    """vslam_set_map's src_kf: map points whose patches come from different source keyframes (MapPoint::pPatchSourceKF), templates
    sampled from the right keyframe pyramid; TrackMap against the oracle (pinned to the reference in tests/test_oracle_vs_ref.py)."""
    from oracle import oraclebind
    from visualslam_android_b200 import api
    cam, kf_frames, kf_poses, smap, src_kf = common.two_keyframe_scene()
    ctx = api.Context(cam.width, cam.height, n_streams=1, max_points=smap.n, max_source_keyframes=2)
    ctx.set_camera(cam.scalars())
    ctx.upload_source_keyframe(kf_frames[0], 0); ctx.upload_source_keyframe(kf_frames[1], 1)
    ctx.set_map(smap.world, smap.pix_right_w, smap.pix_down_w, smap.ir_center, smap.src_level, src_kf)
    ow = oraclebind.OrcWorld(cam, kf_frames[0], smap)
    okf1 = oraclebind.OrcKeyFrame().make_lite(kf_frames[1])
    for k in np.nonzero(src_kf == 1)[0]:
        ow.L.orc_tracker_set_point_source_kf(ow.tracker, int(k), okf1.h)
    tw = np.array([0.05, 0.01, 0.01, 0.005, -0.02, 0.03])
    fr, pose = common.frame_at(cam, tw)
    sp = synth.se3_exp(tw * 0.8)
    ctx.make_keyframe_lite(fr); ow.make_current_kf(fr)
    ctx.set_pose(0, sp); ow.set_pose(sp)
    ctx.track_map(); ow.L.orc_tracker_track_map(ow.tracker)
    _check_track_map(ctx, ow)
    gi, _ = ctx.point_states(0)
    assert gi[src_kf == 0, 3].sum() > 200 and gi[src_kf == 1, 3].sum() > 200
    for k in (0, int(np.nonzero(src_kf == 1)[0][5])):          # a template of each keyframe, byte for byte
        if gi[k, 2]:
            gt, gs, gq = ctx.point_template(0, k); ot, os_, oq = ow.point_template(k)
            assert np.array_equal(gt, ot) and (gs, gq) == (os_, oq)
    ctx.close()


def test_api_housekeeping_timing_reset_and_forced_relocalisation():
    """vslam_set_timing / vslam_get_stage_times (per-launch CUDA events), vslam_reset_stream (Tracker::Reset) and vslam_set_lost (a caller
    declaring a stream lost: the next frame must go through the relocaliser and continue tracking)."""
    cam, f0, smap = common.scene()
    ctx = _ctx(cam, f0, smap, n_streams=2)
    ctx.enable_sbi(synth.Camera(cam.width // 16, cam.height // 16).scalars())
    ctx.set_reloc_keyframes([0], synth.IDENTITY_POSE[None])
    fr = [common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * k)[0] for k in (0.1, 0.2, 0.3)]
    ctx.set_timing(True)
    ctx.track_frame(np.stack([fr[0], fr[0]]))
    st = ctx.stage_times()
    for name in ("pyrfast_l0", "pyrfast_l1", "project_lists", "search_fine", "pose_fine", "other"):
        assert st[name][0] > 0 and st[name][1] >= 1, name
    ctx.set_timing(False)
    assert ctx.counters(0)[2] == 2 and ctx.counters(1)[2] == 2
    # stream 1 is declared lost by the caller: the next frame relocalises it against keyframe 0 and tracks on
    ctx.set_lost(1, 3, 0)
    ctx.track_frame(np.stack([fr[1], fr[1]]))
    best, score, nrec, rec = ctx.reloc_info(1)
    assert (best, nrec, rec) == (0, 1, 1) and score < 9e6
    assert ctx.reloc_info(0)[2] == 0
    a, f, q, lost, dc = ctx.counters(1)
    assert q == 2 and lost == 0 and dc == 1            # quality GOOD again; the recovery forces the coarse stage
    assert np.abs(ctx.get_pose(1) - ctx.get_pose(0)).max() < 5e-3
    # Tracker::Reset
    ctx.reset_stream(0)
    a, f, q, lost, dc = ctx.counters(0)
    v, msd, dm, ds = ctx.get_motion(0)
    assert a.sum() == 0 and f.sum() == 0 and (q, lost, dc) == (2, 0, 0) and np.all(v == 0) and (msd, dm, ds) == (0.0, 1.0, 1.0)
    ctx.track_frame(np.stack([fr[2], fr[2]]))
    assert ctx.counters(0)[2] == 2
    ctx.close()


def test_c_abi_rejects_bad_arguments_with_error_codes():
    """Nothing throws or crashes across the C-ABI: bad arguments come back as VSLAM_E_INVALID with a message."""
    from visualslam_android_b200 import api
    cam, f0, smap = common.scene(n_points=200)
    for bad in (dict(width=100, height=480), dict(width=640, height=100), dict(width=640, height=480, patch_size=12), dict(width=640, height=480, n_streams=0)):
        with pytest.raises(api.VslamError) as e:
            api.Context(**{"width": 640, "height": 480, **bad})
        assert e.value.code == api.E_INVALID
    ctx = _ctx(cam, f0, smap, n_streams=2)
    L, h = ctx.L, ctx.h
    assert L.vslam_get_pose(h, 2, np.zeros(12).ctypes.data) == api.E_INVALID                  # stream out of range
    assert L.vslam_get_pose(None, 0, np.zeros(12).ctypes.data) == api.E_INVALID               # null context
    assert L.vslam_upload_source_keyframe(h, 5, f0.ctypes.data, cam.width) == api.E_INVALID    # unknown source keyframe
    assert L.vslam_make_keyframe_lite(h, 1, 2, f0.ctypes.data, cam.width, 0) == api.E_INVALID  # stream range past the end
    assert L.vslam_make_keyframe_lite(h, 0, 1, f0.ctypes.data, cam.width - 1, 0) == api.E_INVALID   # stride smaller than the width
    with pytest.raises(api.VslamError):
        ctx.set_lists([[0, 1, smap.n], []])                                                    # list entry that is not a map point
    with pytest.raises(api.VslamError):
        ctx.set_reloc_keyframes([0], synth.IDENTITY_POSE[None])                               # needs vslam_enable_sbi first
    assert b"enable_sbi" in L.vslam_last_error(h)
    assert L.vslam_minipatch_find(h, 0, 1, 0, None, None, None, None, 10, 100) == api.E_INVALID   # snapshot not taken
    ctx.make_keyframe_lite(np.stack([f0, f0]))                                                 # the context still works afterwards
    ctx.sync()
    assert ctx.corners(1, 0).shape[0] > 100
    ctx.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_track_frame_random_configurations_from_identical_state(seed):
    """Randomised configurations (map size 60..1800, patch 8 / 11, SmallBlurryImage on / off, slow to very fast motion): from identical
    state a frame of vslam_track_frame agrees with the oracle — counters, quality and flags exactly, the pose to 1e-7; the second frame
    too as long as tracking is healthy.  (Later frames inherit last-bit differences that the reference's discontinuities amplify,
    most easily when few points are found: DESIGN.md §5.)"""
    from oracle import oraclebind
    from visualslam_android_b200 import api
    rs = np.random.RandomState(seed)
    for case in range(5):
        n_points = int(rs.choice([60, 200, 700, 1000, 1800])); P = int(rs.choice([11, 11, 8]))
        sbi = bool(rs.rand() < 0.7); speed = float(rs.choice([0.3, 1.0, 3.0, 8.0]))
        tw = rs.uniform(-1, 1, 6) * np.array([0.01, 0.01, 0.005, 0.006, 0.006, 0.01]) * speed
        cam, f0, smap = common.scene(n_points=n_points)
        ctx = api.Context(cam.width, cam.height, n_streams=1, max_points=smap.n, patch_size=P)
        ctx.set_camera(cam.scalars()); ctx.upload_source_keyframe(f0)
        ctx.set_map(smap.world, smap.pix_right_w, smap.pix_down_w, smap.ir_center, smap.src_level)
        ow = oraclebind.OrcWorld(cam, f0, smap, P=P)
        if sbi:
            sc = synth.Camera(cam.width // 16, cam.height // 16).scalars(); ctx.enable_sbi(sc); ow.L.orc_tracker_enable_sbi(ow.tracker, sc)
        for k in (1, 2):
            fr = synth.render_frame(common.texture(), cam, synth.se3_exp(tw * k))
            ctx.track_frame(fr[None]); ow.L.orc_tracker_track_frame(ow.tracker, np.ascontiguousarray(fr), cam.width, cam.height, cam.width)
            a, f, q, lost, dc = ctx.counters(0); oa, of, oq, olost, odc = ow.counters()
            if k == 2 and not (oq == 2 and of.sum() > 0.8 * oa.sum()):
                break      # degraded tracking: the second frame is no longer a from-identical-state comparison
            assert np.array_equal(a, oa) and np.array_equal(f, of) and (q, lost, dc) == (oq, olost, odc), (seed, case, k, n_points, P, sbi, speed)
            assert np.abs(ctx.get_pose(0) - ow.get_pose()).max() <= 1e-7, (seed, case, k, n_points, P, sbi, speed)
        ctx.close()


def test_map_file_round_trip_restart_and_text_export(tmp_path):
    """f4 on-disk format: a context saves its map (camera, two source keyframes, points, relocaliser registration); the file matches the
    documented layout byte for byte (tests/common.py restates it independently, checksum included); a fresh context that loads it
    tracks the same frames to bit-identical poses and relocalises identically; corrupt / truncated / mismatching files are refused
    with the context untouched; a file produced by the independent writer loads too; the text export has the reference's dump layout."""
    from visualslam_android_b200 import api
    cam, kf_frames, kf_poses, smap, src_kf = common.two_keyframe_scene()
    sbi_cam = synth.Camera(cam.width // 16, cam.height // 16).scalars()

    def fresh(**kw):
        c = api.Context(cam.width, cam.height, n_streams=2, max_points=kw.pop("max_points", smap.n), max_source_keyframes=kw.pop("max_source_keyframes", 3), **kw)
        return c
    a = fresh()
    a.set_camera(cam.scalars()); a.enable_sbi(sbi_cam)
    a.upload_source_keyframe(kf_frames[0], 0); a.upload_source_keyframe(kf_frames[1], 2)          # ids need not be dense
    kf_ids = np.where(src_kf == 0, 0, 2).astype(np.int32)
    a.set_map(smap.world, smap.pix_right_w, smap.pix_down_w, smap.ir_center, smap.src_level, kf_ids)
    a.set_reloc_keyframes([2, 0], np.stack([kf_poses[1], kf_poses[0]]))
    path = tmp_path / "scene.vsmap"
    a.save_map_file(path)
    info = api.map_file_info(path)
    assert (info["n_points"], info["n_keyframes"], info["n_reloc_keyframes"]) == (smap.n, 2, 2)
    blob = path.read_bytes()
    got = common.mapfile_unpack(blob)                                                              # asserts the checksum and the exact length
    assert np.array_equal(got["cam13"], np.asarray(cam.scalars(), dtype=np.float64))
    assert [(k, r) for k, r, _ in got["keyframes"]] == [(0, 1), (2, 0)]
    assert np.array_equal(got["keyframes"][0][2], kf_frames[0]) and np.array_equal(got["keyframes"][1][2], kf_frames[1])
    p = got["points"]
    assert np.array_equal(p["world"], smap.world) and np.array_equal(p["right"], smap.pix_right_w) and np.array_equal(p["down"], smap.pix_down_w)
    assert np.array_equal(p["ircenter"], smap.ir_center) and np.array_equal(p["srclevel"], smap.src_level) and np.array_equal(p["srckf"], kf_ids)
    assert np.array_equal(got["reloc"][0], [2, 0]) and np.array_equal(got["reloc"][1], np.stack([kf_poses[1], kf_poses[0]]).reshape(2, 12))
    # the independent writer produces the same bytes from the same content
    again = common.mapfile_pack(cam.width, cam.height, cam.scalars(), got["keyframes"], p, got["reloc"])
    assert again == blob

    # refused files leave the target context untouched: b has no map yet and must still have none afterwards
    b = fresh(); b.enable_sbi(sbi_cam)
    bad = tmp_path / "bad.vsmap"
    for data, code in ((blob[:200000] + bytes([blob[200000] ^ 1]) + blob[200001:], api.E_IO), (blob[:-9], api.E_IO), (blob[:len(blob) // 2], api.E_IO), (blob + b"\0", api.E_IO)):
        bad.write_bytes(data)
        with pytest.raises(api.VslamError) as e:
            b.load_map_file(bad, api.MAP_LOAD_CAMERA | api.MAP_LOAD_RELOC)
        assert e.value.code == code
    small = api.Context(320, 240, n_streams=1, max_points=smap.n, max_source_keyframes=3)
    with pytest.raises(api.VslamError) as e:
        small.load_map_file(path)
    assert e.value.code == api.E_INVALID and "image size" in str(e.value)
    small.close()
    for kw in (dict(max_points=smap.n - 1), dict(max_source_keyframes=2)):                         # keyframe id 2 needs 3 slots
        c = fresh(**kw)
        with pytest.raises(api.VslamError) as e:
            c.load_map_file(path)
        assert e.value.code == api.E_CAPACITY
        c.close()
    nosbi = fresh()
    with pytest.raises(api.VslamError) as e:
        nosbi.load_map_file(path, api.MAP_LOAD_RELOC)
    assert e.value.code == api.E_INVALID
    nosbi.close()
    b.set_camera(cam.scalars())
    first = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.2)[0]
    b.track_frame(np.stack([first, first]))
    assert b.counters(0)[0].sum() == 0                                                             # nothing attempted: still no map

    # restart: load into b (and the independent writer's file into c), then the same frames give the same bits
    b.close(); b = fresh(); b.enable_sbi(sbi_cam)
    b.load_map_file(path, api.MAP_LOAD_CAMERA | api.MAP_LOAD_RELOC)
    other = tmp_path / "other.vsmap"; other.write_bytes(again)
    c = fresh(); c.enable_sbi(sbi_cam); c.load_map_file(other, api.MAP_LOAD_CAMERA | api.MAP_LOAD_RELOC)
    rs = np.random.RandomState(11)
    noise = lambda: rs.randint(0, 255, kf_frames[0].shape).astype(np.uint8)
    tw1 = np.array([0.10, 0.02, 0.01, 0.01, -0.04, 0.05])
    seq = [first, common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.4)[0]] + [noise() for _ in range(4)] + [common.frame_at(cam, tw1 + np.array([0.004, -0.003, 0.002, 0.01, 0.008, -0.012]))[0],
                                                                                                   common.frame_at(cam, tw1 + np.array([0.006, -0.002, 0.002, 0.012, 0.006, -0.01]))[0]]
    for k, fr in enumerate(seq):
        frames = np.stack([fr, seq[0] if k % 2 else fr])
        for x in (a, b, c):
            x.track_frame(frames)
        for s in range(2):
            assert np.array_equal(a.get_pose(s), b.get_pose(s)) and np.array_equal(a.get_pose(s), c.get_pose(s)), (k, s)
            ca, cb = a.counters(s), b.counters(s)
            assert np.array_equal(ca[0], cb[0]) and np.array_equal(ca[1], cb[1]) and ca[2:] == cb[2:], (k, s)
            assert a.reloc_info(s) == b.reloc_info(s) == c.reloc_info(s), (k, s)
    assert a.reloc_info(0)[2] >= 1 and a.reloc_info(0)[0] == 0                                     # recovered against registration index 0 = keyframe id 2
    assert a.counters(0)[1].sum() > 100

    # text export in the layout of the reference's SaveMap dump
    out = tmp_path / "dump"
    b.export_map_text(out)
    lines = (out / "map.dump").read_text().split("\n")
    assert lines[-1] == "" and len(lines) == 3 * smap.n + 1
    for i in (0, 1, smap.n // 2, smap.n - 1):
        x, y, zl = lines[3 * i], lines[3 * i + 1], lines[3 * i + 2]
        z, lvl = zl.rsplit("  ", 1)
        assert len(x) == len(y) == len(z)                                                          # Eigen aligns the column
        assert [float(x), float(y), float(z)] == [float("%g" % v) for v in smap.world[i]] and int(lvl) == smap.src_level[i]
    for k, pose in enumerate([kf_poses[1], kf_poses[0]]):
        rows = (out / "keyframes" / f"{k}.info").read_text().split("\n")
        assert rows[3:] == ["", ""]
        assert np.array_equal(np.array([[float(v) for v in r.split(" ")] for r in rows[:3]]), np.array([[float("%g" % v) for v in r] for r in np.asarray(pose).reshape(3, 4)]))
    for x in (a, b, c):
        x.close()


def test_every_kernel_on_a_ragged_configuration():
    """tests/sanitizer_case.py (352x272, 3 streams: every kernel and entry point once, with plausibility checks) as a plain test."""
    import sanitizer_case
    sanitizer_case.main()


def test_track_frame_keyframe_handoff_on_device():
    """jni/Tracker.cc:127-132,866-872 with the keyframe policy on: requests raised by the k_pose tail at the frames where the
    restatement (pinned to the unmodified TrackFrame in tests/test_oracle_vs_ref.py) adds keyframes; vslam_add_keyframe_from_stream
    (device-to-device copy of the stream's pyramid + relocaliser registration) makes a keyframe the stream later relocalises
    against; DODGY -> BAD far away from every keyframe, DODGY kept with a roomy scale."""
    from oracle import oraclebind
    cam, f0, smap = common.scene()
    sbi_cam = synth.Camera(cam.width // 16, cam.height // 16).scalars()
    okf0 = oraclebind.OrcKeyFrame().make_lite(f0)

    def pair(wiggle, wiggle_dn):
        c = _ctx(cam, f0, smap, n_streams=1, max_source_keyframes=4)
        c.enable_sbi(sbi_cam)
        c.set_reloc_keyframes([0], synth.IDENTITY_POSE[None])
        c.set_keyframe_policy(True, wiggle, wiggle_dn, 0.2, 20)
        o = _orc(cam, f0, smap)
        o.L.orc_tracker_enable_sbi(o.tracker, sbi_cam)
        o.L.orc_tracker_add_reloc_keyframe(o.tracker, okf0.h, np.ascontiguousarray(synth.IDENTITY_POSE, dtype=np.float64).reshape(12))
        o.L.orc_tracker_set_keyframe_policy(o.tracker, 1, wiggle, wiggle_dn, 0.2, 20)
        return c, o
    ctx, ow = pair(0.1, 0.1)
    ctx0, ow0 = ctx, ow

    def info():
        v = [C.c_int() for _ in range(4)]
        ow.L.orc_tracker_keyframe_info(ow.tracker, *[C.byref(x) for x in v])
        return [x.value for x in v]

    def both(fr, tag, noise=False, tol=1e-8, ctx=None, ow=None):
        ctx = ctx or ctx0; ow = ow or ow0
        fr = np.ascontiguousarray(fr)
        ctx.track_frame(fr[None])
        ow.L.orc_tracker_track_frame(ow.tracker, fr, cam.width, cam.height, cam.width)
        a, f, q, lost, dc = ctx.counters(0); oa, of, oq, olost, odc = ow.counters()
        assert (q, lost) == (oq, olost), (tag, q, oq, lost, olost)
        if noise:      # a handful of chance matches on white noise: the (int) casts of the reference amplify 1e-13 differences; only BAD / lost must agree
            return
        assert np.abs(ctx.get_pose(0) - ow.get_pose()).max() <= tol, tag
        assert np.array_equal(a, oa) and np.array_equal(f, of) and dc == odc, tag

    step = np.array([0.004, 0.001, 0.0005, 0.0004, -0.0012, 0.0008])
    rs = np.random.RandomState(9)
    frames = [common.frame_at(cam, step * k)[0] for k in range(1, 31)]
    frames += [rs.randint(0, 255, f0.shape).astype(np.uint8) for _ in range(4)]
    frames += [common.frame_at(cam, step * 27.5)[0], common.frame_at(cam, step * 28.5)[0]]
    added_at, next_id = [], 1
    for k, fr in enumerate(frames):
        both(fr, k, noise=30 <= k < 34)
        req, closest, dist = ctx.keyframe_requests()
        assert np.array_equal(req, ctx.keyframe_requests(flags_only=True))           # the compact per-frame poll agrees with the full read-out
        n_kf, oadd, oframe, olast = info()
        assert bool(req[0]) == bool(oadd), (k, req, oadd, dist)
        if req[0]:
            ctx.add_keyframe_from_stream(0, next_id); next_id += 1; added_at.append(k + 1)
            assert np.array_equal(ctx.keyframe_requests()[0], [0]) and np.array_equal(ctx.keyframe_requests(flags_only=True), [0])
    assert added_at == [6, 27]
    best, score, nrec = C.c_int(), C.c_double(), C.c_int()
    ow.L.orc_tracker_reloc_info(ow.tracker, C.byref(best), C.byref(score), C.byref(nrec))
    gb, gs, gn, gr = ctx.reloc_info(0)
    assert gn == nrec.value >= 1 and gb == best.value == 2 and abs(gs - score.value) <= 1e-6 * max(1.0, abs(score.value))
    # the keyframe copied on the device is the frame the stream saw (frame 27), pyramid included
    okf = oraclebind.OrcKeyFrame().make_lite(frames[26])
    path = os.path.join(tempfile.mkdtemp(), "handoff.vsmap")
    ctx.save_map_file(path)
    saved = common.mapfile_unpack(open(path, "rb").read())
    assert [k for k, _, _ in saved["keyframes"]] == [0, 1, 2] and np.array_equal(saved["keyframes"][2][2], frames[26])
    assert np.array_equal(saved["reloc"][0], [0, 1, 2])
    ctx.close()
    # DODGY -> BAD far from every keyframe (tiny wiggle scale), DODGY stays DODGY with a roomy one; fresh trackers per case (a frame
    # that is mostly noise leaves the two implementations in states that differ in the last bits, see `both`)
    seen = kept = 0
    for frac in (0.55, 0.65, 0.72, 0.8):
        occ = common.frame_at(cam, step * 2.0)[0].copy()
        occ[:int(cam.height * frac)] = rs.randint(0, 255, (int(cam.height * frac), cam.width))
        for wiggle in (0.1, 1e-5):
            c, o = pair(wiggle, 1e30)
            both(common.frame_at(cam, step * 1.0)[0], ("good", frac, wiggle), ctx=c, ow=o)
            both(occ, ("occluded", frac, wiggle), noise=True, ctx=c, ow=o)
            a, f, q, lost, dc = c.counters(0)
            dodgy_by_counts = f.sum() <= 0.3 * a.sum() and (f[2:].sum() >= 0.13 * a[2:].sum() if a[2:].sum() > 10 else f.sum() >= 0.13 * a.sum())
            seen += bool(dodgy_by_counts and wiggle < 1e-3 and q == 0 and lost == 1)
            kept += bool(dodgy_by_counts and wiggle > 1e-3 and q == 1 and lost == 0)
            c.close()
    assert seen >= 1 and kept >= 1


def test_map_grows_while_streams_run():
    """vslam_append_map_points: the map grows twice during a sequence; old points keep their per-stream state (template cache and
    M-estimator counters carry over, which the second frame after an append depends on), new ones start fresh -- against the
    restatement, which tests/test_oracle_vs_ref.py::test_map_grows_while_tracking pins to the unmodified TrackFrame.  Two streams
    with different motions share the growing map."""
    from visualslam_android_b200 import api
    cam, f0, smap = common.scene()
    names = ("world", "pix_right_w", "pix_down_w", "ir_center", "src_level", "center_nc", "one_right_nc", "one_down_nc")
    parts = [synth.SyntheticMap(**{k: getattr(smap, k)[sl] for k in names}) for sl in (slice(0, 500), slice(500, 800), slice(800, 1000))]
    S = 2
    ctx = api.Context(cam.width, cam.height, n_streams=S, max_points=1000)
    ctx.set_camera(cam.scalars()); ctx.upload_source_keyframe(f0)
    ctx.set_map(parts[0].world, parts[0].pix_right_w, parts[0].pix_down_w, parts[0].ir_center, parts[0].src_level)
    sbi_cam = synth.Camera(cam.width // 16, cam.height // 16).scalars()
    ctx.enable_sbi(sbi_cam)
    ows = [_orc(cam, f0, parts[0]) for _ in range(S)]
    for ow in ows:
        ow.L.orc_tracker_enable_sbi(ow.tracker, sbi_cam)
    nxt = 1
    for k in range(1, 10):
        if k in (4, 7):
            p = parts[nxt]; nxt += 1
            ctx.append_map_points(p.world, p.pix_right_w, p.pix_down_w, p.ir_center, p.src_level)
            for ow in ows:
                ow.append_points(p)
        frames = np.stack([synth.render_frame(common.texture(), cam, synth.stream_pose(4 * k, 2 + s)) for s in range(S)])
        ctx.track_frame(frames)
        for s, ow in enumerate(ows):
            ow.L.orc_tracker_track_frame(ow.tracker, np.ascontiguousarray(frames[s]), cam.width, cam.height, cam.width)
            assert np.abs(ctx.get_pose(s) - ow.get_pose()).max() <= 1e-8, (k, s)
            a, f, q, lost, dc = ctx.counters(s); oa, of, oq, olost, odc = ow.counters()
            assert np.array_equal(a, oa) and np.array_equal(f, of) and (q, lost, dc) == (oq, olost, odc), (k, s)
            assert np.array_equal(ctx.point_counts(s), ow.point_counts()), (k, s)
    assert ctx.counters(0)[0].sum() > 600
    with pytest.raises(api.VslamError) as e:
        ctx.append_map_points(parts[1].world, parts[1].pix_right_w, parts[1].pix_down_w, parts[1].ir_center, parts[1].src_level)
    assert e.value.code == api.E_CAPACITY
    ctx.close()


def test_host_mapmaker_loop_grows_the_map_on_the_device():
    """The loop a host MapMaker runs around the library, with no image crossing PCIe after the camera frame itself: the device raises
    a keyframe request -> vslam_add_keyframe_from_stream -> MakeKeyFrame_Rest on the new keyframe -> the closest OLD keyframe becomes
    a scratch stream's keyframe (vslam_make_keyframe_from_source) -> epipolar search of the new keyframe's candidates in it ->
    triangulated points -> vslam_append_map_points.  The grown map is then tracked: the new points are found in later frames and
    the pose stays on the ground truth.  (Functional test: each step has its own parity test above.)"""
    from visualslam_android_b200 import api
    cam, f0, smap = common.scene(n_points=500)
    ctx = api.Context(cam.width, cam.height, n_streams=2, max_points=2000, max_source_keyframes=3)    # stream 0 tracks, stream 1 is MapMaker's scratch
    ctx.set_camera(cam.scalars()); ctx.upload_source_keyframe(f0)
    ctx.set_map(smap.world, smap.pix_right_w, smap.pix_down_w, smap.ir_center, smap.src_level)
    ctx.enable_sbi(synth.Camera(cam.width // 16, cam.height // 16).scalars())
    ctx.set_reloc_keyframes([0], synth.IDENTITY_POSE[None])
    ctx.set_keyframe_policy(True, 0.1, 0.1, 0.2, 20)
    step = np.array([0.008, 0.002, 0.001, 0.0008, -0.0024, 0.0016])
    truth = lambda k: synth.se3_exp(step * k)
    n_before = smap.n
    grown_at = None
    for k in range(1, 16):
        fr = synth.render_frame(common.texture(), cam, truth(k))
        ctx.track_frame(np.stack([fr, fr]))
        assert np.abs(ctx.get_pose(0) - truth(k)).max() < 5e-3, k
        req, closest, dist = ctx.keyframe_requests()
        if req[0] and grown_at is None:
            pose_new = ctx.get_pose(0)
            ctx.add_keyframe_from_stream(0, 1)                                    # Tracker::AddNewKeyFrame
            ctx.make_keyframe_rest(0)                                             # MapMaker::AddKeyFrameFromTopOfQueue: candidates of the new keyframe
            cands = [ctx.candidates(0, l)[0] for l in range(4)]
            _, _, depth_mean, depth_sigma = ctx.get_motion(0)
            ctx.make_keyframe_from_source(1, int(closest[0]))                     # target = the closest old keyframe (keyframe 0, identity pose)
            new = [[] for _ in range(5)]
            for level in range(4):
                found, pos, _, _ = ctx.epipolar_search(1, 1, level, cands[level], pose_new, synth.IDENTITY_POSE, depth_mean, depth_sigma, 0.1)
                sel = np.nonzero(found)[0]
                world, right, down, irc, lvl = ctx.epipolar_make_points(level, cands[level][sel], pos[sel], pose_new, synth.IDENTITY_POSE)
                keep = np.abs(world[:, 2] - 1.0) < 0.1                             # the scene is the plane z = 1; the bundle adjuster would weed out the rest
                for a, v in zip(new, (world[keep], right[keep], down[keep], irc[keep], lvl[keep])):
                    a.append(v)
            world, right, down, irc, lvl = [np.concatenate(a) for a in new]
            assert len(world) > 100, len(world)
            ctx.append_map_points(world, right, down, irc, lvl, np.ones(len(world), dtype=np.int32))
            grown_at = k
    assert grown_at is not None and ctx.n_points > n_before + 100
    ints, dbl = ctx.point_states(0)
    new_pts = ints[n_before:]
    searched = new_pts[:, 2] == 1
    assert searched.sum() > 50 and (new_pts[searched, 3] == 1).mean() > 0.6, (searched.sum(), (new_pts[searched, 3] == 1).mean())
    assert ctx.counters(0)[2] == 2
    ctx.close()


def test_track_map_config_switches_no_truncation_and_other_seed():
    """vslam_config.truncate_error = 0 (the WLS takes the residual as a double instead of the reference's (int) cast) and
    rand_seed != 1 (the per-stream copy of glibc rand() that std::random_shuffle draws from): 2500 points so that every shuffle of
    TrackMap matters, coarse stage on."""
    cam, f0, smap = common.scene(n_points=2500)
    ctx = _ctx(cam, f0, smap, truncate_error=False, rand_seed=7)
    ow = _orc(cam, f0, smap)
    ow.L.orc_tracker_set_truncate(ow.tracker, 0); ow.L.orc_tracker_seed(ow.tracker, 7)
    f1, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.8)
    ctx.make_keyframe_lite(f1); ow.make_current_kf(f1)
    ctx.set_pose(0, synth.IDENTITY_POSE); ow.set_pose(synth.IDENTITY_POSE)
    ctx.set_motion(0, np.zeros(6), 0.05); ow.L.orc_tracker_set_velocity(ow.tracker, np.zeros(6), 0.05)
    ctx.track_map(); ow.L.orc_tracker_track_map(ow.tracker)
    assert ow.counters()[4] == 1
    _check_track_map(ctx, ow)
    # and the default seed really gives another selection of points (the switch is not a no-op)
    ctx1 = _ctx(cam, f0, smap, truncate_error=False)
    ctx1.make_keyframe_lite(f1); ctx1.set_pose(0, synth.IDENTITY_POSE); ctx1.set_motion(0, np.zeros(6), 0.05); ctx1.track_map()
    assert not np.array_equal(ctx1.point_states(0)[0][:, 2], ctx.point_states(0)[0][:, 2])
    ctx.close(); ctx1.close()


@pytest.mark.parametrize("P", [11, 8])
def test_patchfinder_per_object_methods_match_the_oracle(P):
    """jni/PatchFinder.h:45-121 one object at a time (vslam_pf_*): MakeTemplateCoarseCont alone, ZMSSDAtPoint at chosen positions,
    MakeSubPixTemplate + IterateSubPix / IterateSubPixToConvergence from a set start, MakeTemplateCoarseNoWarp."""
    from oracle import oraclebind
    OL = oraclebind.lib()
    cam, f0, smap = common.scene()
    ctx, ow = _ctx(cam, f0, smap, patch_size=P), _orc(cam, f0, smap, P=P)
    f1, _ = common.frame_at(cam, np.array(synth.CONFIG1_TWIST) * 0.6)
    start = synth.se3_exp(np.array(synth.CONFIG1_TWIST) * 0.45)
    okf = ow.make_current_kf(f1); ctx.make_keyframe_lite(f1)
    ctx.set_pose(0, start); ow.set_pose(start)
    ctx.project_all(); ow.L.orc_tracker_project_all(ow.tracker)
    oi, od = ow.point_states()
    ctx.set_point_projection(0, od[:, 0:2], od[:, 11:15], oi[:, 1])
    idx = np.nonzero(oi[:, 1] >= 0)[0].astype(np.int32)
    # the oracle's templates / coarse positions of all points (one batched search on its side only)
    ow.L.orc_tracker_search_for_points(ow.tracker, idx, len(idx), 10, 0)
    oi, od = ow.point_states()
    fnd = idx[oi[idx][:, 3] == 1]
    assert len(fnd) > 100
    rng = np.random.default_rng(5)
    for k in [int(v) for v in rng.choice(fnd, 12, replace=False)]:
        level = int(oi[k, 1])
        # (1) MakeTemplateCoarseCont by itself: template pixels and sums
        assert ctx.pf_make_template(0, k) is False
        gt, gs, gq = ctx.point_template(0, k); ot, os_, oq = ow.point_template(k)
        assert np.array_equal(gt, ot) and (gs, gq) == (os_, oq), f"template of point {k}"
        # (2) ZMSSDAtPoint at positions around the coarse hit, inside and outside the border
        lw, lh = ctx.level_dims(level)
        cx, cy = [int((v + 0.5) / (1 << level) - 0.5 + 0.5) for v in od[k, 30:32]]
        xy = np.array([[cx + dx, cy + dy] for dx in (-3, 0, 2) for dy in (-2, 0, 3)] + [[1, 1], [lw - 1, lh - 1], [P // 2, P // 2], [lw - 1 - P // 2, lh - 1 - P // 2]], dtype=np.int32)
        tm = np.ascontiguousarray(ot.reshape(-1))
        want = np.array([OL.orc_zmssd(okf.h, level, tm, P, int(x), int(y)) for x, y in xy], dtype=np.int32)
        assert np.array_equal(ctx.pf_zmssd_at(0, k, level, xy), want), f"ZMSSD of point {k}"
        assert want[9] == P * P * 500 + 1 and want[10] == P * P * 500 + 1
        # (3) IterateSubPixToConvergence from the coarse position, and the same in single IterateSubPix steps
        coarse = np.ascontiguousarray(od[k, 30:32])
        opos = np.zeros(2)
        ok = OL.orc_subpix(okf.h, level, tm, P, coarse, 8, opos, None)
        gpos, md, conv, last = ctx.pf_subpix(0, k, 8, coarse)
        assert conv == bool(ok)
        assert np.abs(gpos - opos).max() <= 1e-9, (k, gpos, opos)
        pos, md1, n_it = coarse.copy(), 0.0, 0
        for _ in range(8):
            pos, md1, c1, last1 = ctx.pf_subpix(0, k, 1, pos, md1); n_it += 1
            if last1 < 0 or c1:
                break
        assert np.array_equal(pos, gpos) and md1 == md and c1 == conv, "single steps == run to convergence"
    # off the image: IterateSubPix returns a negative value
    _, _, conv, last = ctx.pf_subpix(0, int(fnd[0]), 3, np.array([-50.0, -50.0]))
    assert conv is False and last < 0
    # (4) MakeTemplateCoarseNoWarp: the P x P pixels of the source keyframe level, border rule P / 2 + 1
    k = int(fnd[1])
    for level, (x, y) in ((0, (100, 77)), (2, (31, 40)), (1, (P // 2 + 1, P // 2 + 1)), (1, (P // 2, 50))):
        src = ow.src_kf.pixels(level)
        bad = ctx.pf_make_template_nowarp(0, k, 0, level, x, y)
        h, w = src.shape
        b = P // 2 + 1
        assert bad == (not (x >= b and y >= b and x < w - b and y < h - b))
        if not bad:
            gt, gs, gq = ctx.point_template(0, k)
            want_t = src[y - P // 2:y - P // 2 + P, x - P // 2:x - P // 2 + P]
            assert np.array_equal(gt, want_t) and gs == int(want_t.astype(np.int64).sum()) and gq == int((want_t.astype(np.int64) ** 2).sum())
            gi, _ = ctx.point_states(0)
            assert gi[k, 1] == level
    # user events (jni/jni_part.cpp:49-51)
    assert ctx.take_user_event(0) == 0
    ctx.user_event(0); ctx.user_event(0)
    assert ctx.take_user_event(0) == 1 and ctx.take_user_event(0) == 0
    ctx.close()


def test_corner_lists_are_built_on_demand_after_tracked_frames():
    """A tracked frame leaves only the corner bitmasks behind (k_search_fast reads them); the first call that needs the reference's containers --
    Level::vCorners / vCornerRowLUT (jni/KeyFrame.h:52-58) -- builds them from the bitmasks (vs_ensure_lists).  After every frame the lists of
    every stream and level must be that frame's, bit for bit, also when they were not asked for after the frame before, and the list consumers
    (MakeKeyFrame_Rest here) must see them too."""
    from oracle import oraclebind
    cam, f0, smap = common.scene(n_points=300)
    S = 2
    ctx = _ctx(cam, f0, smap, n_streams=S)
    ctx.enable_sbi(synth.Camera(cam.width // 16, cam.height // 16).scalars())
    for k in range(1, 5):
        frames = np.stack([synth.render_frame(common.texture(), cam, synth.stream_pose(4 * k, s + 1)) for s in range(S)])
        ctx.track_frame(frames)
        if k == 2:
            continue                                   # nobody asks for the lists of frame 2
        for s in range(S):
            okf = oraclebind.OrcKeyFrame().make_lite(frames[s])
            if k == 3 and s == 1:                      # a list consumer comes first: Shi-Tomasi candidates of the frame's corners
                okf.make_rest(); ctx.make_keyframe_rest(s)
                for l in range(4):
                    gxy, gsc = ctx.candidates(s, l); oxy, osc = okf.candidates(l)
                    assert np.array_equal(gxy, oxy) and np.array_equal(gsc, osc), (k, s, l)
            _check_keyframe(ctx, s, okf)
    ctx.close()


@pytest.mark.parametrize("seed", range(6))
def test_make_keyframe_lite_randomised_images(seed):
    """Pyramid + FAST-10 + lists on random content: noise of varying contrast (every rejection / even-ring / full-ring path, dense and sparse
    levels, work items with more than 255 survivors), blocks (long corner-free runs, corners clustered on edges) and a smooth field; several
    streams per launch with different images, sizes whose last 16-pixel chunk is partial at the small levels."""
    from oracle import oraclebind
    rs = np.random.RandomState(100 + seed)
    W, H = [(96, 64), (160, 120), (352, 272), (640, 480), (224, 8 * 13), (1280, 72)][seed]
    S = 3
    imgs = []
    for s in range(S):
        kind = (seed + s) % 3
        if kind == 0:
            amp = [255, 60, 24][s % 3]
            im = rs.randint(0, amp + 1, size=(H, W)) + rs.randint(0, 256 - amp)
        elif kind == 1:
            bs = [5, 9, 17][s % 3]
            blocks = rs.randint(0, 256, size=((H + bs - 1) // bs, (W + bs - 1) // bs))
            im = np.kron(blocks, np.ones((bs, bs), dtype=np.int64))[:H, :W] + rs.randint(0, 12, size=(H, W))
        else:
            yy, xx = np.mgrid[0:H, 0:W]
            im = 128 + 90 * np.sin(xx / 7.0 + s) * np.cos(yy / 5.0) + rs.randint(0, 20, size=(H, W))
        imgs.append(np.clip(im, 0, 255).astype(np.uint8))
    frames = np.stack(imgs)
    from visualslam_android_b200 import api
    ctx = api.Context(W, H, n_streams=S, max_points=8, max_corner_frac=1.0)
    ctx.make_keyframe_lite(frames)
    for s in range(S):
        _check_keyframe(ctx, s, oraclebind.OrcKeyFrame().make_lite(frames[s]))
    ctx.close()

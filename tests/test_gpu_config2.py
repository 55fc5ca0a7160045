"""BASELINE config 2 on the GPU: the whole Tracker::TrackFrame (jni/Tracker.cc:76-146 — MakeKeyFrame_Lite, SmallBlurryImage,
motion model, TrackMap coarse+fine with 10 Tukey-WLS iterations each, quality) over a 300-frame synthetic VGA sequence.

Two modes, and why there are two: the reference's tracker AMPLIFIES any pose perturbation by about x1.3 per frame
(motion-model feedback; tests/test_host_logic_cpu.py::test_reference_algorithm_amplifies_a_one_ulp_perturbation shows it with the
oracle against itself), so two implementations that are not bit-identical in every libm call drift apart until a discrete event
(another corner wins, a truncated template pixel or an (int)-cast residual flips) after 50-90 frames, whatever their per-frame
accuracy.  The contract tolerance (1e-4 relative on every update twist and on the final residuals, counters exact)
is therefore asserted per frame from the reference's state (teacher-forced); the free-running run is reported and bounded.
"""
import json
import os

import numpy as np
import pytest

import common
from visualslam_android_b200 import synth

pytestmark = pytest.mark.gpu

N_FRAMES = 300
TOL = 1e-4            # SURVEY.md §8 parity contract: SE3 update twist and final reprojection residuals, relative


def _setup(n_streams):
    from oracle import oraclebind
    from visualslam_android_b200 import api
    cam, f0, smap = common.scene()
    sbi_cam = synth.Camera(cam.width // 16, cam.height // 16).scalars()
    ctx = api.Context(cam.width, cam.height, n_streams=n_streams, max_points=smap.n)
    ctx.set_camera(cam.scalars()); ctx.upload_source_keyframe(f0)
    ctx.set_map(smap.world, smap.pix_right_w, smap.pix_down_w, smap.ir_center, smap.src_level)
    ctx.enable_sbi(sbi_cam)
    ows = [oraclebind.OrcWorld(cam, f0, smap) for _ in range(n_streams)]
    for ow in ows:
        ow.L.orc_tracker_enable_sbi(ow.tracker, sbi_cam)
    frames = np.stack([common.render_sequence(cam, [synth.stream_pose(k, s) for k in range(1, N_FRAMES + 1)]) for s in range(n_streams)], axis=1)
    return cam, smap, ctx, ows, frames          # frames: (N_FRAMES, S, H, W)


def _oracle_motion(ow):
    import ctypes as C
    v = np.zeros(6); m, dm, ds = C.c_double(), C.c_double(), C.c_double()
    ow.L.orc_tracker_get_velocity(ow.tracker, v, C.byref(m))
    ow.L.orc_tracker_get_scene_depth(ow.tracker, C.byref(dm), C.byref(ds))
    return v, m.value, dm.value, ds.value


def _report(name, obj):
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, name), "w") as f:
            json.dump(obj, f, indent=1)
    print(json.dumps(obj))


def test_config2_300_frames_teacher_forced_meets_the_contract():
    """Every frame starts from the ORACLE's tracker state (pose, velocity, scene depth; the per-point template cache, the rand() state and
    the SmallBlurryImage of the last frame are the device's own) and must reproduce the oracle's frame: counters, quality, coarse flag and
    number of iterations exactly; every one of the 10 (+10) update twists to 1e-4 relative (norm-wise, the contract); the final pose; the
    found set and the final reprojection residuals of every found point to 1e-4 relative."""
    S = 2
    cam, smap, ctx, ows, frames = _setup(S)
    worst_upd, worst_pose, worst_res, n_upd, mism_pix, n_border, n_subpix, n_flipped, n_found, n_tmpl_px = 0.0, 0.0, 0.0, 0, 0, 0, 0, 0, 0, 0
    for k in range(N_FRAMES):
        if k > 0:
            for s, ow in enumerate(ows):
                v, m, dm, ds = _oracle_motion(ow)
                ctx.set_pose(s, ow.get_pose()); ctx.set_motion(s, v, m, dm, ds)
        ctx.track_frame(frames[k])
        for s, ow in enumerate(ows):
            ow.L.orc_tracker_track_frame(ow.tracker, np.ascontiguousarray(frames[k, s]), cam.width, cam.height, cam.width)
            a, f, q, lost, dc = ctx.counters(s); oa, of, oq, olost, odc = ow.counters()
            assert np.array_equal(a, oa) and np.array_equal(f, of) and (q, lost, dc) == (oq, olost, odc), (k, s, a, oa, f, of)
            gu, gs = ctx.updates(s); ou, os_ = ow.updates()
            assert len(gu) == len(ou) == (20 if dc else 10), (k, s)
            for it in range(len(ou)):
                err = np.linalg.norm(gu[it] - ou[it])
                assert err <= TOL * np.linalg.norm(ou[it]) + 1e-10, (k, s, it, gu[it], ou[it])       # 1e-10: absolute floor for the last, vanishing updates
                if np.linalg.norm(ou[it]) > 1e-7:
                    worst_upd = max(worst_upd, err / np.linalg.norm(ou[it])); n_upd += 1
            assert np.allclose(gs, os_, rtol=TOL, atol=0), (k, s)                                        # Tukey sigma^2 per iteration
            d = np.abs(ctx.get_pose(s) - ow.get_pose()).max()
            assert d <= 1e-7, (k, s, d)
            worst_pose = max(worst_pose, d)
            if k % 25 == 24 or k == 0:     # the per-point end state (sampled: 1000 ctypes calls per stream on the oracle side)
                gi, gd = ctx.point_states(s); oi, od = ow.point_states()
                assert np.array_equal(gi[:, :6], oi[:, :6]), (k, s)                                      # in-image, level, searched, found, sub-pixel, bad-template flags
                fnd = oi[:, 3] == 1
                # coarse (integer-corner) positions: exact -- except for a point whose TEMPLATE differs: the template pixels are truncated doubles
                # (jni/vision/ImageHandler.cpp:12-19) sampled at positions that carry the device's own last-bit history (the per-point template
                # cache is not teacher-forced), so once in ~10^5 pixels a truncation falls on the other side and another corner can win the ZMSSD.
                dcoarse = np.abs(gd[:, 30:32] - od[:, 30:32]).max(axis=1)
                flipped = fnd & (dcoarse > 0)
                for i in np.nonzero(flipped)[0]:
                    gt, *_ = ctx.point_template(s, int(i)); ot, *_ = ow.point_template(int(i))
                    assert (gt != ot).any(), (k, s, i, "coarse position differs although the templates are identical")
                n_flipped += int(flipped.sum()); n_found += int(fnd.sum())
                # found positions: exact for the points without sub-pixel refinement, 1e-6 px for the refined ones -- except where the
                # float-blended inverse-compositional iteration (jni/PatchFinder.cc:272-350) sits on its convergence threshold and the
                # two sides stop one iteration apart ("identical except documented borderline"): counted, not compared
                dfound = np.abs(gd[:, 2:4] - od[:, 2:4]).max(axis=1)
                border = fnd & (dfound > 1e-6) & ~flipped
                assert np.all(oi[border, 4] == 1) and dfound[border].max(initial=0.0) < 0.25, (k, s, np.nonzero(border)[0], dfound[border])
                n_border += int(border.sum()); n_subpix += int((fnd & (oi[:, 4] == 1)).sum())
                assert dfound[fnd & (oi[:, 4] == 0) & ~flipped].max(initial=0.0) == 0.0, (k, s)
                ok = fnd & ~border & ~flipped
                res_g, res_o = gd[ok][:, 28:30], od[ok][:, 28:30]                                        # final (found - projected) / 2^level
                # residuals are differences of ~100 px quantities and tend to zero at the optimum: 1e-4 relative to max(|r|, 1 level-px), i.e. 1e-6 of the projections
                rel = (np.abs(res_g - res_o) / np.maximum(np.abs(res_o), 1.0)).max()
                assert rel <= TOL, (k, s, rel, int(np.argmax(np.abs(res_g - res_o).max(axis=1))))
                worst_res = max(worst_res, rel)
                for i in np.nonzero(oi[:, 2] == 1)[0][::7]:
                    gt, gsum, gsq = ctx.point_template(s, int(i)); ot, osum, osq = ow.point_template(int(i))
                    mism_pix += int((gt != ot).sum()); n_tmpl_px += gt.size
    _report("r02_config2_teacher_forced.json", {"frames": N_FRAMES, "streams": S, "worst_update_relerr": worst_upd, "updates_compared": n_upd,
                                               "worst_pose_absdiff": worst_pose, "worst_final_residual_relerr": worst_res,
                                               "template_pixel_mismatches_sampled": mism_pix, "template_pixels_sampled": n_tmpl_px, "tolerance": TOL,
                                               "subpix_points_sampled": n_subpix, "subpix_borderline_points": n_border,
                                               "found_points_sampled": n_found, "points_with_flipped_template_pixel_and_other_corner": n_flipped})
    # discrete events of the un-forced per-point state (template cache): rare, counted, bounded
    assert mism_pix <= 1e-3 * n_tmpl_px and n_flipped <= 1e-3 * n_found + 1, (mism_pix, n_tmpl_px, n_flipped, n_found)
    assert n_border <= 0.05 * n_subpix + 2, (n_border, n_subpix)
    ctx.close()


def test_config2_300_frames_free_running_report():
    """No teacher: both trackers run on their own state for 300 frames.  Asserted: both stay locked (quality GOOD, nothing lost), track the
    true camera equally well, find the same number of points within 2 %, agree to 1e-9 for the first frames and never differ by more than the
    trackers' own error against the truth.  Reported (gpurun_out/r02_config2_free_running.json): the first frame at which the pose difference
    exceeds 1e-9 / 1e-6 / 1e-4, the per-frame growth factor before the first discrete event and the number of frames with unequal counters."""
    S = 2
    cam, smap, ctx, ows, frames = _setup(S)
    first = {1e-9: None, 1e-6: None, 1e-4: None}
    hist = np.zeros((N_FRAMES, S)); cnt_mismatch = 0; err_g = err_o = 0.0
    for k in range(N_FRAMES):
        ctx.track_frame(frames[k])
        poses = ctx.get_poses()
        for s, ow in enumerate(ows):
            ow.L.orc_tracker_track_frame(ow.tracker, np.ascontiguousarray(frames[k, s]), cam.width, cam.height, cam.width)
            op, truth = ow.get_pose(), synth.stream_pose(k + 1, s)
            d = hist[k, s] = np.abs(poses[s] - op).max()
            err_g, err_o = max(err_g, np.abs(poses[s] - truth).max()), max(err_o, np.abs(op - truth).max())
            a, f, q, lost, dc = ctx.counters(s); oa, of, oq, olost, odc = ow.counters()
            assert q == oq == 2 and lost == olost == 0, (k, s)
            assert abs(int(f.sum()) - int(of.sum())) <= 0.02 * of.sum(), (k, s)
            cnt_mismatch += not (np.array_equal(a, oa) and np.array_equal(f, of))
            for thr in first:
                if first[thr] is None and d > thr:
                    first[thr] = (k, s, float(d))
    assert hist[:10].max() <= 1e-9, hist[:10].max()
    assert err_g < 5e-3 and err_o < 5e-3, (err_g, err_o)
    assert hist.max() <= 2.0 * max(err_g, err_o), (hist.max(), err_g, err_o)         # as close to each other as either is to the truth
    k9 = first[1e-9][0] if first[1e-9] else N_FRAMES - 1
    growth = float((hist[k9].max() / max(hist[2].max(), 1e-17)) ** (1.0 / max(k9 - 2, 1))) if k9 > 2 else None
    _report("r02_config2_free_running.json", {"frames": N_FRAMES, "streams": S, "first_frame_exceeding": {str(t): v for t, v in first.items()},
                                             "growth_factor_per_frame_until_1e-9": growth, "worst_pose_absdiff": float(hist.max()),
                                             "frames_with_unequal_counters": int(cnt_mismatch), "tracker_error_vs_truth": {"gpu": err_g, "oracle": err_o},
                                             "pose_absdiff_every_10_frames": [float(x) for x in hist.max(axis=1)[::10]]})
    ctx.close()


def test_track_frame_against_the_compiled_reference_directly():
    """GPU against oracle/_ref (the reference's own Tracker::TrackFrame, compiled by oracle/build_ref.sh) without the oracle port in between:
    12 frames with SmallBlurryImage, counters / quality exact, pose 1e-8."""
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref/libvslam_ref.so was not built (no /root/reference at build time)")
    from visualslam_android_b200 import api
    cam, f0, smap = common.scene()
    W, H = cam.width, cam.height
    ctx = api.Context(W, H, n_streams=1, max_points=smap.n)
    ctx.set_camera(cam.scalars()); ctx.upload_source_keyframe(f0)
    ctx.set_map(smap.world, smap.pix_right_w, smap.pix_down_w, smap.ir_center, smap.src_level)
    ctx.enable_sbi(synth.Camera(W // 16, H // 16).scalars())
    rw = refbind.RefWorld(W, H, f0, smap)
    rw.L.ref_srand(1); rw.L.ref_sbi_reset_size()
    frames = common.render_sequence(cam, [synth.stream_pose(3 * k, 1) for k in range(1, 13)])
    for k, fr in enumerate(frames):
        ctx.track_frame(fr[None])
        rw.L.ref_tracker_track_frame(rw.tracker, np.ascontiguousarray(fr), W, H, W)
        a, f, q, lost, dc = ctx.counters(0); ra, rf, rq, rlost, rdc = rw.counters()
        assert np.array_equal(a, ra) and np.array_equal(f, rf) and (q, lost, dc) == (rq, rlost, rdc), (k, a, ra, f, rf)
        assert np.abs(ctx.get_pose(0) - rw.get_pose()).max() <= 1e-8, k
    gi, gd = ctx.point_states(0); ri, rd = rw.point_states()
    pvs = gi[:, 1] >= 0          # (TrackerData::bFound and PatchFinder::mbTemplateBad are uninitialised in the reference until a point enters the PVS / is searched)
    assert np.array_equal(gi[pvs][:, :4], ri[pvs][:, :4])
    assert np.array_equal(gi[ri[:, 3] == 1][:, 4], ri[ri[:, 3] == 1][:, 4])      # bDidSubPix is only assigned when a point is found (jni/Tracker.cc:657-672)
    srch = gi[:, 2] == 1
    assert np.array_equal(gi[srch][:, 5], ri[srch][:, 5])
    fnd = (ri[:, 3] == 1) & pvs
    assert np.array_equal(gd[fnd][:, 30:32], rd[fnd][:, 30:32]) and np.abs(gd[fnd][:, 2:4] - rd[fnd][:, 2:4]).max() <= 1e-6
    ctx.close()

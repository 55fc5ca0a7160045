"""One compact pass over EVERY kernel of the library, for `compute-sanitizer` (memcheck / racecheck / initcheck / synccheck):

    compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitizer_case.py

Its own script because the sanitizer slows kernels 10-100x (the parity suite would take hours); tests/test_gpu_parity.py also runs it
as a plain GPU test.  (On the pool this round was developed on, compute-sanitizer is closed by the operators, so only the plain run
has been made there.)  No oracle here: this run only has to
touch every code path with ragged sizes; results are checked for plausibility so that a silently skipped path is noticed."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visualslam_android_b200 import api, synth  # noqa: E402


def main():
    W, H, S = 352, 272, 3           # ragged: 11 x 32 columns, height not a multiple of the 16-row strips; 3 streams
    cam = synth.Camera(W, H)
    tex = synth.make_texture(1024)
    f0 = synth.render_frame(tex, cam, synth.IDENTITY_POSE)
    ctx = api.Context(W, H, n_streams=S, max_points=400, max_source_keyframes=2)
    ctx.set_camera(cam.scalars())
    # keyframe stage: pyramid + FAST, MakeKeyFrame_Rest, MiniPatch
    ctx.make_keyframe_lite(np.stack([f0] * S))
    corners = [ctx.corners(0, l) for l in range(4)]
    dims = [ctx.level_dims(l) for l in range(4)]
    assert len(corners[0]) > 300, len(corners[0])
    ctx.make_keyframe_rest(0)
    cands = [ctx.candidates(0, l)[0] for l in range(4)]
    assert sum(len(c) for c in cands) > 50
    ctx.snapshot_keyframe(0)
    xy = cands[0][:40]
    patches = ctx.minipatch_sample(0, xy)
    pos, found, _ = ctx.minipatch_find(0, patches, xy.astype(np.float64), which=1)
    assert found.sum() == len(xy)
    # map from those corners; tracking with SmallBlurryImage and the relocaliser
    smap = synth.build_map(cam, corners, dims, 400)
    ctx.upload_source_keyframe(f0, 0)
    pose1 = synth.se3_exp(np.array([0.10, 0.02, 0.01, 0.01, -0.04, 0.05]))
    f1 = synth.render_frame(tex, cam, pose1)
    ctx.upload_source_keyframe(f1, 1)
    ctx.set_map(smap.world, smap.pix_right_w, smap.pix_down_w, smap.ir_center, smap.src_level)
    ctx.enable_sbi(synth.Camera(W // 16, H // 16).scalars())
    ctx.set_reloc_keyframes([0, 1], np.stack([synth.IDENTITY_POSE, pose1]))
    tw = np.array(synth.CONFIG1_TWIST)
    for k in range(1, 4):
        frames = np.stack([synth.render_frame(tex, cam, synth.se3_exp(tw * 0.2 * k * (1 + 0.3 * s))) for s in range(S)])
        if k == 3:
            ctx.set_lost(1, 3)                                  # stream 1 goes through the relocaliser on this frame
        ctx.track_frame(frames)
    for s in range(S):
        a, f, q, lost, dc = ctx.counters(s)
        assert f.sum() > 0.5 * a.sum() > 50, (s, a, f)
    assert ctx.reloc_info(1)[2] == 1
    # the pipelined entry point
    bufs = [np.ascontiguousarray(frames) for _ in range(2)]
    for i in [ctx.track_frame_async(b.ctypes.data, W, W * H) for b in bufs]:
        ctx.wait_step(i)
    # MapMaker's searches + new points, and the stage-wise entry points
    lists = np.tile(np.arange(smap.n, dtype=np.int32), (S, 1))
    ctx.set_lists(lists)
    ctx.refind()
    flags, _ = ctx.refind_results(0, smap.n)
    assert flags[:, 0].sum() > 100
    ctx.make_keyframe_lite(f1)
    nf = 0
    for level in range(4):
        found, pos, _, _ = ctx.epipolar_search(0, 0, level, cands[level], synth.IDENTITY_POSE, pose1, 1.0, 0.3, 0.1)
        sel = np.nonzero(found)[0]
        world, *_ = ctx.epipolar_make_points(level, cands[level][sel], pos[sel], synth.IDENTITY_POSE, pose1)
        nf += len(sel)
    assert nf > 20, nf
    ctx.project_all(); ctx.clear_counters(); ctx.search_for_points(8, 4); ctx.project_and_derivs(); ctx.calc_jacobians(); ctx.calc_pose_update(0.0, True, True)
    # PatchFinder one object at a time (patchfinder_ops.cu) and a search range wide enough for the word-by-word window walk of k_search_fast
    ints, dbl = ctx.point_states(0)
    pt = int(np.nonzero(ints[:, 3] == 1)[0][0])
    assert ctx.pf_make_template(0, pt) is False
    lvl = int(ints[pt, 1]); lw, lh = ctx.level_dims(lvl)
    z = ctx.pf_zmssd_at(0, pt, lvl, np.array([[lw // 2, lh // 2], [0, 0], [lw - 1, lh - 1]], dtype=np.int32))
    assert z[1] == z[2] == 11 * 11 * 500 + 1 and 0 <= z[0]
    ctx.pf_subpix(0, pt, 8, dbl[pt, 30:32])
    ctx.pf_make_template_nowarp(0, pt, -1, 0, 0, 0); ctx.pf_make_template_nowarp(0, pt, 1, 3, 1, 1)
    ctx.project_all(); ctx.search_for_points(40, 2)
    ctx.track_map()
    with tempfile.TemporaryDirectory() as d:
        ctx.save_map_file(os.path.join(d, "m.vsmap"))
        ctx.load_map_file(os.path.join(d, "m.vsmap"), api.MAP_LOAD_CAMERA | api.MAP_LOAD_RELOC)
    ctx.sync()
    print("sanitizer case ok: launches", ctx.kernel_launches())
    ctx.close()


if __name__ == "__main__":
    main()

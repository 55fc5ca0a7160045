#!/usr/bin/env python
"""Benchmark of the B200 tracking front-end.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[3]): S = 256 independent synthetic VGA camera streams per GPU, 1000 map points,
one `step` = one frame of every stream through Tracker::TrackFrame's good-map branch (MakeKeyFrame_Lite: 4-level pyramid +
FAST-10 + row LUT; SmallBlurryImage + CalcSBIRotation; ApplyMotionModel; TrackMap coarse+fine with 10 Tukey-WLS iterations
each; UpdateMotionModel; AssessTrackingQuality).  Streams never exchange data: N GPUs = N independent contexts, no collective on the data path.

  value  tracked frames/s with the frames already resident in HBM (vslam_track_frame_dev), CUDA-event timed
  e2e    the same through the host-buffer C-ABI call (vslam_track_frame: pinned host frames copied in every step) plus a
         device->host read of every stream's pose, every step
  roofline      KeyFrame::MakeKeyFrame_Lite (k_pyramid_fast + k_fast_levels + k_corner_lists): algorithmic bytes / event-timed duration
  cpu_baseline  the reference's own sources (oracle/_ref) — or the oracle port when that library is absent — tracking the
                same kind of sequence on the host cores, one process per core

`--impl reference` times only the CPU arm and prints the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, N_POINTS = 640, 480, 1000
TEX_SIZE = 2048
# --frame: the other frame sizes north_star asks throughput for: (W, H, map points, streams per GPU)
FRAME_CONFIGS = {"1080p": (1920, 1080, 5000, 148), "4k": (3840, 2160, 20000, 148)}   # one CTA of the per-stream kernels per SM
STREAMS_PER_GPU = 256   # --scaling weak: streams per GPU; --scaling strong (default, BASELINE configs[3]): streams in TOTAL, stream s on GPU s mod G
POOL = 24               # distinct frame sets kept resident per leg (78.6 MB each at 256 VGA streams)
FRAME_STEP = 2          # synthetic sequence index advance per step (≈ 1 px of image motion per frame)
METRIC = "tracked_frames_per_sec"
UNIT = "frames/s"
CPU_FRAMES_PER_STEP = 25   # frames per process per step in the CPU arms (a bounded sample of the same workload)
SCALING = "strong"
WORKLOAD = ("configs[3]: 256 independent synthetic VGA (640x480) camera streams, 1000 map points, full TrackFrame-equivalent per frame (4-level pyramid + "
            "FAST-10 + row LUT, SmallBlurryImage rotation estimate, motion model, fine PatchFinder search + 10 Tukey-WLS iterations; the coarse stage with its "
            "10 further iterations runs when the motion model asks for it -- not at this camera speed, see the `fast_motion` leg), P=11")


# ------------------------------------------------------------------------------------------------ synthetic data
def stream_poses(streams, n_frames, frame_step=None):
    """(len(streams), n_frames, 3, 4) camera poses; `streams`: a count (streams 0..n-1) or the list of stream ids (= sequence seeds)."""
    from visualslam_android_b200 import synth
    ids = range(streams) if isinstance(streams, int) else streams
    step = FRAME_STEP if frame_step is None else frame_step
    return np.stack([np.stack([synth.stream_pose(step * k, s) for k in range(n_frames)]) for s in ids])


def render_frames_torch(tex_t, cam, poses, device):
    """Torch version of synth.render_frame for a batch of poses (B,3,4) -> uint8 (B,H,W).  Input generation only."""
    import torch
    B = poses.shape[0]
    size = tex_t.shape[0]
    v, u = torch.meshgrid(torch.arange(cam.height, device=device, dtype=torch.float64), torch.arange(cam.width, device=device, dtype=torch.float64), indexing="ij")
    dx = (u - cam.cx) / cam.fx
    dy = (v - cam.cy) / cam.fy
    rd = torch.sqrt(dx * dx + dy * dy)
    r = torch.tan(rd * cam.w) * cam.one_over_two_tan
    f = torch.where(rd > 0.01, r / rd.clamp_min(1e-300), torch.ones_like(rd))
    ray = torch.stack([dx * f, dy * f, torch.ones_like(dx)], dim=-1).to(torch.float32)          # (H,W,3)
    P = torch.as_tensor(poses, device=device, dtype=torch.float64)
    R, t = P[:, :, :3], P[:, :, 3]
    o = -(R.transpose(1, 2) @ t[:, :, None])[:, :, 0]                                              # (B,3)
    d = torch.einsum("hwc,bcd->bhwd", ray, R.to(torch.float32))                                    # R^T ray
    s = ((1.0 - o[:, 2]).to(torch.float32))[:, None, None] / d[..., 2]
    X = o[:, 0].to(torch.float32)[:, None, None] + s * d[..., 0]
    Y = o[:, 1].to(torch.float32)[:, None, None] + s * d[..., 1]
    tu = torch.remainder(X * float(cam.fx) + size / 2.0, size - 1.0)
    tv = torch.remainder(Y * float(cam.fx) + size / 2.0, size - 1.0)
    iu = tu.floor().long().clamp_(0, size - 2)
    iv = tv.floor().long().clamp_(0, size - 2)
    fu = tu - iu
    fv = tv - iv
    flat = tex_t.reshape(-1)
    i00 = iv * size + iu
    val = (1 - fv) * ((1 - fu) * flat[i00] + fu * flat[i00 + 1]) + fv * ((1 - fu) * flat[i00 + size] + fu * flat[i00 + size + 1])
    return (val + 0.5).floor().clamp_(0, 255).to(torch.uint8)


def keyframe_corners(f0, on_gpu, device=0):
    """FAST corners and level sizes of one keyframe, for choosing the synthetic map's points (input preparation, outside every timed
    region).  The GPU arm uses the library itself; only the CPU arms (cpu_baseline, --impl reference, --cpu-stages) use the oracle —
    both give the same corners bit for bit (tests/test_gpu_parity.py), so every arm tracks the same map."""
    h, w = f0.shape
    if on_gpu:
        from visualslam_android_b200 import api
        ctx = api.Context(w, h, n_streams=1, max_points=8, device=device)
        ctx.make_keyframe_lite(f0[None])
        out = [ctx.corners(0, l) for l in range(4)], [ctx.level_dims(l) for l in range(4)]
        ctx.close()
        return out
    from oracle import oraclebind
    kf = oraclebind.OrcKeyFrame().make_lite(f0)
    return [kf.corners(l) for l in range(4)], [kf.dims(l) for l in range(4)]


def build_scene(on_gpu=False, device=0, size=None):
    """KF0 + map of N_POINTS points chosen among KF0's FAST corners.  size: (W, H, map points, texture size), default = the module's workload."""
    from visualslam_android_b200 import synth
    w, h, n_points, tex_size = size or (W, H, N_POINTS, TEX_SIZE)
    cam = synth.Camera(w, h)
    tex = synth.make_texture(tex_size)
    f0 = synth.render_frame(tex, cam, synth.IDENTITY_POSE)
    corners, dims = keyframe_corners(f0, on_gpu, device)
    smap = synth.build_map(cam, corners, dims, n_points)
    return cam, tex, f0, smap


def short_leg(name, size, stream_ids, K, Wm, device, stream, frame_step=None, pool=6, barrier=None, scene=None):
    """An extra, shorter leg of the GPU arm on its own context: whole TrackFrame of len(stream_ids) streams of `size` = (W, H, map points, texture
    size), frames resident in HBM, CUDA-event timed over K steps after Wm warm-up steps.  Returns (ms per step of THIS rank, info dict)."""
    import torch
    from visualslam_android_b200 import api, synth
    w, h, n_points, tex_size = size
    dev = torch.device("cuda", device)
    cam, tex, f0, smap = scene or build_scene(on_gpu=True, device=device, size=size)
    tex_t = torch.from_numpy(tex.astype(np.float32)).to(dev)
    S = len(stream_ids)
    M = int(max(3, min(pool, 5e9 // (S * w * h))))
    poses = stream_poses(stream_ids, M + 1, frame_step)
    frames = torch.empty((M, S, h, w), dtype=torch.uint8, device=dev)
    rb = 64 if w * h <= 640 * 480 else 8
    for k in range(1, M + 1):
        for s0 in range(0, S, rb):
            frames[k - 1, s0:s0 + rb] = render_frames_torch(tex_t, cam, poses[s0:s0 + rb, k], dev)
    ctx = api.Context(w, h, n_streams=S, max_points=smap.n, device=device, cuda_stream=stream.cuda_stream)
    ctx.set_camera(cam.scalars())
    sbi = True                   # (heights that are not a multiple of 16 -- 1080p -- take cv::resize's general bilinear path for the thumbnail)
    ctx.enable_sbi(synth.Camera(w // 16, h // 16).scalars())
    ctx.upload_source_keyframe(f0)
    ctx.set_map(smap.world, smap.pix_right_w, smap.pix_down_w, smap.ir_center, smap.src_level)
    step = 0
    for k in range(Wm):
        ctx.track_frame_ptr(frames[tri(step, M)].data_ptr(), w, h * w, device=True); step += 1
    ctx.sync()
    if barrier:
        barrier()
    stats0 = ctx.search_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(K):
        ctx.track_frame_ptr(frames[tri(step, M)].data_ptr(), w, h * w, device=True); step += 1
    e1.record(stream)
    ctx.sync()
    ms = e0.elapsed_time(e1) / K
    stats1 = ctx.search_stats()
    probe = list(range(0, S, max(1, S // 8)))
    info = {"leg": name, "frame": [w, h], "map_points": n_points, "streams_this_gpu": S, "pool_sets": M, "frame_step": FRAME_STEP if frame_step is None else frame_step,
            "small_blurry_image": sbi,
            "found_per_frame_mean": float(np.mean([ctx.counters(s_)[1].sum() for s_ in probe])),
            "quality_good_frac": float(np.mean([ctx.counters(s_)[2] == 2 for s_ in probe])),
            "coarse_stage_frac": float(np.mean([ctx.counters(s_)[4] for s_ in probe])),
            "wls_iterations_per_frame": float(np.mean([len(ctx.updates(s_)[0]) for s_ in probe])),
            "templates_generated_per_step": (stats1["templates_generated"] - stats0["templates_generated"]) / K,
            "subpix_refinements_per_step": (stats1["subpix_refinements"] - stats0["subpix_refinements"]) / K}
    ctx.close()
    del frames
    torch.cuda.empty_cache()
    return ms, info


# ------------------------------------------------------------------------------------------------ CPU arm
def tri(j, M):
    """j-th step of a run -> index into a pool of M consecutive frame sets, traversed as a triangle wave (0,1,..,M-1,M-2,..,0,1,..): any
    number of steps sees a continuous camera motion that stays over the mapped region.  The GPU arm and both CPU arms use this schedule."""
    if M == 1:
        return 0
    p = j % (2 * M - 2)
    return p if p < M else 2 * M - 2 - p


def pool_size(K, Wm):
    """Frame sets kept per stream; the same on every arm for the same --steps / --warmup."""
    Wm = max(Wm, 3)
    return min(POOL, 2 * (Wm + K) + Wm)


def cpu_worker(path, shape, stream, n_warm, n_steps, frames_per_step, use_ref):
    """One process = one camera stream tracked by the reference (oracle/_ref) or the oracle port.  Prints per-step seconds."""
    from visualslam_android_b200 import synth
    n_seq, M = shape
    frames = np.load(path + ".frames.npy", mmap_mode="r")[stream]           # (M, H, W): only this stream's pages are touched
    d = np.load(path + ".scene.npz", allow_pickle=False)
    f0 = d["f0"]
    smap = synth.SyntheticMap(world=d["world"], pix_right_w=d["right"], pix_down_w=d["down"], ir_center=d["irc"], src_level=d["lvl"],
                              center_nc=d["cnc"], one_right_nc=d["rnc"], one_down_nc=d["dnc"])
    cam = synth.Camera(W, H)
    a4, f4 = np.zeros(4, dtype=np.int32), np.zeros(4, dtype=np.int32)
    if use_ref:
        import ctypes as C
        from oracle import refbind
        rw = refbind.RefWorld(W, H, f0, smap)
        rw.L.ref_srand(1)
        rw.L.ref_sbi_reset_size()
        # keyframe 0 joins the relocaliser (KeyFrame::MakeKeyFrame_Rest leaves pSBI behind, jni/KeyFrame.cc:98): a tracker that gets lost
        # then recovers through Relocaliser::AttemptRecovery (jni/Relocaliser.cc:17-42) instead of dereferencing a null pSBI
        rw.L.ref_kf_make_sbi(rw.src_kf.h)
        step = lambda fr: rw.L.ref_tracker_track_frame(rw.tracker, fr, W, H, W)      # the unmodified Tracker::TrackFrame, SmallBlurryImage included
        q, lost, dc = C.c_int(), C.c_int(), C.c_int()
        counters = lambda: rw.L.ref_tracker_counters(rw.tracker, a4, f4, C.byref(q), C.byref(lost), C.byref(dc))
    else:
        from oracle import oraclebind
        ow = oraclebind.OrcWorld(cam, f0, smap)
        ow.L.orc_tracker_enable_sbi(ow.tracker, synth.Camera(W // 16, H // 16).scalars())
        step = lambda fr: ow.L.orc_tracker_track_frame(ow.tracker, fr, W, H, W)
        def counters():
            a, f, *_ = ow.counters()
            a4[:] = a; f4[:] = f
    k = 0
    times, found, attempted = [], 0, 0
    for s in range(n_warm + n_steps):
        t0 = time.perf_counter()
        for _ in range(frames_per_step):
            step(np.ascontiguousarray(frames[tri(k, M)]))
            k += 1
        times.append(time.perf_counter() - t0)
        if s >= n_warm:                      # counters of the step's last frame (outside the timed part)
            counters()
            found += int(f4.sum()); attempted += int(a4.sum())
    print(json.dumps({"steps": times[n_warm:], "found_per_frame": found / max(n_steps, 1), "attempted_per_frame": attempted / max(n_steps, 1)}))


def run_cpu_arm(cam, tex, f0, smap, n_procs, n_warm, n_steps, frames_per_step, M):
    """Launch n_procs tracker processes concurrently, process i on the GPU arm's stream i (same seeds, same triangle-wave pool of M frame
    sets, same FRAME_STEP); returns (frames/s aggregate, seconds per step, kind, tracking stats)."""
    from oracle import refbind
    use_ref = refbind.available()
    n_seq = n_procs
    frames = render_cpu_or_gpu(tex, cam, stream_poses(n_seq, M + 1)[:, 1:])          # (n_seq, M, H, W): what GPU rank 0 feeds streams 0..n_seq-1
    fd, path = tempfile.mkstemp(prefix="vslam_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    os.close(fd)
    np.save(path + ".frames.npy", frames)
    np.savez(path + ".scene.npz", f0=f0, world=smap.world, right=smap.pix_right_w, down=smap.pix_down_w, irc=smap.ir_center,
             lvl=smap.src_level, cnc=smap.center_nc, rnc=smap.one_right_nc, dnc=smap.one_down_nc)
    try:
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cpu-worker", path, f"{n_seq},{M}", str(i % n_seq), str(n_warm),
                                   str(n_steps), str(frames_per_step), str(int(use_ref))], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                                  env={**os.environ, "CUDA_VISIBLE_DEVICES": "", "OMP_NUM_THREADS": "1"}) for i in range(n_procs)]
        outs, failed = [], []
        for i, p in enumerate(procs):
            try:
                o, e = p.communicate(timeout=900)
            except subprocess.TimeoutExpired:
                p.kill(); o, e = p.communicate()
                failed.append(f"worker {i}: timed out"); continue
            if p.returncode != 0:
                why = f"signal {-p.returncode}" if p.returncode < 0 else f"exit code {p.returncode}"
                failed.append(f"worker {i} (stream {i % n_seq}): {why}; stderr tail: {e[-1500:]!r}")
                continue
            outs.append(json.loads(o.strip().splitlines()[-1]))
        if failed:
            raise RuntimeError("cpu worker(s) failed: " + " | ".join(failed))
    finally:
        for suffix in ("", ".frames.npy", ".scene.npz"):
            if os.path.exists(path + suffix):
                os.unlink(path + suffix)
    per_step = np.array([o["steps"] for o in outs])          # (procs, steps) seconds
    step_s = per_step.max(axis=0)                             # a step ends when the slowest worker finished it
    total = float(step_s.sum())
    fps = n_procs * n_steps * frames_per_step / total
    stats = {"found_per_frame_mean": float(np.mean([o["found_per_frame"] for o in outs])),
             "attempted_per_frame_mean": float(np.mean([o["attempted_per_frame"] for o in outs]))}
    return fps, total / n_steps, ("reference" if use_ref else "port"), stats


def sweep_config5(device=0):
    """SURVEY.md §8(d) config 5 on one GPU: 3840x2160 frames and a 20000-point map.
    (1) pyramid + FAST-10 + row LUT for batches of 1..64 frames per launch pair: ms, frames/s, algorithmic GB/s and the fraction of the
        measured HBM peak.  Every timed launch reads a frame set that was not touched by the previous launches worth > 126 MB (L2).
    (2) SearchForPoints (no sub-pixel) for N points x search range: ms, candidates scored (identical on CPU and GPU), integer TMAC/s."""
    import torch
    from visualslam_android_b200 import api, synth
    W4, H4 = 3840, 2160
    dev = torch.device("cuda", device)
    torch.cuda.set_device(dev)
    cam = synth.Camera(W4, H4)
    tex = synth.make_texture(4096)
    tex_t = torch.from_numpy(tex.astype(np.float32)).to(dev)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
    out = {"what": "config 5 sweep (3840x2160)", "hbm_peak_gbs": peak, "pyramid_fast": [], "search": []}
    n_pool = 20                                              # 20 x 8.3 MB = 166 MB > L2
    poses = np.stack([synth.stream_pose(3 * k, 1) for k in range(1, 65 + n_pool)])
    frames = torch.empty((64 + n_pool, H4, W4), dtype=torch.uint8, device=dev)
    for k in range(0, len(frames), 4):
        frames[k:k + 4] = render_frames_torch(tex_t, cam, poses[k:k + 4], dev)
    torch.cuda.synchronize()
    for S in (1, 2, 4, 8, 16, 32, 64):
        ctx = api.Context(W4, H4, n_streams=S, max_points=16, cuda_stream=stream.cuda_stream)
        reps, t = 12, []
        for r in range(3 + reps):
            off = (r * max(S, n_pool)) % (len(frames) - S + 1) if S < n_pool else (r % 2) * (len(frames) - S)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.make_keyframe_lite_ptr(frames[off].data_ptr(), S, W4, H4 * W4, device=True)
            e1.record(stream); ctx.sync()
            if r >= 3:
                t.append(e0.elapsed_time(e1))
        ms = float(np.median(t))
        corners = sum(int(ctx.corners(0, l).shape[0]) for l in range(4))
        alg = S * (1.328125 * W4 * H4 + 4 * sum((H4 >> l) for l in range(4)) + 4.0 * corners)
        out["pyramid_fast"].append({"frames_per_launch": S, "ms": ms, "frames_per_s": S / (ms * 1e-3), "algorithmic_gbs": alg / (ms * 1e-3) / 1e9,
                                    "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak, "corners_per_frame": corners})
        ctx.close()
    # (2) patch search: one 4K stream, N map points, range sweep
    f0 = synth.render_frame(tex, cam, synth.IDENTITY_POSE)
    kf_corners, kf_dims = keyframe_corners(f0, True, device)
    f1 = synth.render_frame(tex, cam, synth.se3_exp(np.array(synth.CONFIG1_TWIST)))
    dp4a_peak = api.dp4a_peak_tmacs()
    for N in (1000, 5000, 20000):
        smap = synth.build_map(cam, kf_corners, kf_dims, N)
        ctx = api.Context(W4, H4, n_streams=1, max_points=smap.n, cuda_stream=stream.cuda_stream)
        ctx.set_camera(cam.scalars()); ctx.upload_source_keyframe(f0)
        ctx.set_map(smap.world, smap.pix_right_w, smap.pix_down_w, smap.ir_center, smap.src_level)
        ctx.set_pose(0, synth.IDENTITY_POSE)
        ctx.make_keyframe_lite(f1[None])
        ctx.project_all()
        ints, _ = ctx.point_states(0)
        lst = np.nonzero((ints[:, 0] == 1) & (ints[:, 1] >= 0))[0].astype(np.int32)
        ctx.set_lists([lst])
        for rng in (5, 10, 20, 40):
            t = []; ev = 0
            for r in range(2 + 6):
                ev0 = ctx.zmssd_evals()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); ctx.search_for_points(rng, 0); e1.record(stream); ctx.sync()
                ev = ctx.zmssd_evals() - ev0
                if r >= 2:
                    t.append(e0.elapsed_time(e1))
            ms = float(np.median(t))
            out["search"].append({"points": int(len(lst)), "range": rng, "ms": ms, "points_per_s": len(lst) / (ms * 1e-3), "candidates_scored": int(ev),
                                  "tmacs": 3.0 * 121 * ev / (ms * 1e-3) / 1e12, "dp4a_peak_tmacs": dp4a_peak})
        ctx.close()
    return out


def cpu_stage_times(reps=20, warm=5):
    """per-stage times of BASELINE config 1 (one VGA frame, 1000 map points, start pose I, frame at CONFIG1_TWIST) on
    one host core: the reference's own sources (oracle/_ref, P = 11 as shipped) and the oracle port at P = 11 and P = 8.
    Median of `reps` repetitions after `warm` warm-ups, milliseconds."""
    from oracle import oraclebind, refbind
    from visualslam_android_b200 import synth
    cam, tex, f0, smap = build_scene()
    f1 = synth.render_frame(tex, cam, synth.se3_exp(np.array(synth.CONFIG1_TWIST)))
    eye = synth.IDENTITY_POSE

    def med(fn, setup=None):
        ts = []
        for r in range(warm + reps):
            if setup:
                setup()
            t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
        return float(np.median(ts[warm:]))

    def stages(kind, w, L, pre):
        t = w.tracker
        g = lambda name: getattr(L, pre + name)
        out = {}
        kf_make = (lambda: refbind.RefKeyFrame().make_lite(f1)) if kind == "reference" else (lambda: oraclebind.OrcKeyFrame().make_lite(f1))
        out["make_keyframe_lite"] = med(kf_make)
        w.make_current_kf(f1)
        reset = lambda: (w.set_pose(eye), g("tracker_set_velocity")(t, np.zeros(6), 0.0))
        out["track_map_fine_only"] = med(lambda: g("tracker_track_map")(t), reset)
        reset_c = lambda: (w.set_pose(eye), g("tracker_set_velocity")(t, np.zeros(6), 0.05))
        out["track_map_coarse_and_fine"] = med(lambda: g("tracker_track_map")(t), reset_c)
        w.set_pose(eye)
        out["project_all"] = med(lambda: g("tracker_project_all")(t))
        ints, _ = w.point_states()
        lst = np.ascontiguousarray(np.nonzero((ints[:, 0] == 1) & (ints[:, 1] >= 0))[0].astype(np.int32))
        out["search_for_points_range10"] = med(lambda: g("tracker_search_for_points")(t, lst, len(lst), 10, 0))
        out["search_for_points_range10_subpix8"] = med(lambda: g("tracker_search_for_points")(t, lst, len(lst), 10, 8))
        g("tracker_calc_jacobians")(t, lst, len(lst))
        mu = np.zeros(6)
        out["calc_pose_update"] = med(lambda: g("tracker_calc_pose_update")(t, lst, len(lst), 0.0, 0, 0, mu))
        out["points_searched"] = int(len(lst))
        return out

    def sequence_fps(kind, w, L, pre, n=40):
        """whole TrackFrame (SBI included) over a short config-2 style sequence on ONE core."""
        frames = [np.ascontiguousarray(synth.render_frame(tex, cam, synth.stream_pose(FRAME_STEP * k, 1))) for k in range(1, n + 6)]
        w.set_pose(eye)
        getattr(L, pre + "tracker_set_velocity")(w.tracker, np.zeros(6), 0.0)
        for fr in frames[:5]:
            getattr(L, pre + "tracker_track_frame")(w.tracker, fr, W, H, W)
        t0 = time.perf_counter()
        for fr in frames[5:]:
            getattr(L, pre + "tracker_track_frame")(w.tracker, fr, W, H, W)
        return n / (time.perf_counter() - t0)

    res = {}
    if refbind.available():
        rw = refbind.RefWorld(W, H, f0, smap)
        res["reference_P11"] = stages("reference", rw, rw.L, "ref_")
        rw2 = refbind.RefWorld(W, H, f0, smap)
        rw2.L.ref_srand(1); rw2.L.ref_sbi_reset_size()
        res["reference_P11"]["track_frame_sequence_fps_one_core"] = sequence_fps("reference", rw2, rw2.L, "ref_")
    for P in (11, 8):
        ow = oraclebind.OrcWorld(cam, f0, smap, P=P)
        res[f"port_P{P}"] = stages("port", ow, ow.L, "orc_")
        ow2 = oraclebind.OrcWorld(cam, f0, smap, P=P)
        ow2.L.orc_tracker_enable_sbi(ow2.tracker, synth.Camera(W // 16, H // 16).scalars())
        res[f"port_P{P}"]["track_frame_sequence_fps_one_core"] = sequence_fps("port", ow2, ow2.L, "orc_")
    return {"what": "per-stage CPU times, BASELINE config 1, one core, ms (median of %d after %d warm-ups); the stand-in cv::Mat / cv::resize of the "
                    "reference build are plain scalar C++, not OpenCV's SIMD paths" % (reps, warm), "cores": 1, "stages_ms": res}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region by a background thread through NVML (pynvml)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = False
        self._thr = None

    def _run(self):
        import pynvml as nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                r = get_reasons(self.handle)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import threading

            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[0].isdigit() else self.gpu
            self.handle = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        except Exception:
            self._thr = None

    def stop(self):
        self._stop = True
        if self._thr is not None:
            self._thr.join(timeout=2)
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.samples:
            out["sm_mhz"] = float(np.median(self.samples))
        return out


# ------------------------------------------------------------------------------------------------ main
def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--cpu-worker":
        a = sys.argv[2:]
        cpu_worker(a[0], tuple(int(x) for x in a[1].split(",")), int(a[2]), int(a[3]), int(a[4]), int(a[5]), bool(int(a[6])))
        return
    # stdout carries exactly ONE JSON line: whatever a library writes to fd 1 (NCCL prints its version banner there) goes to stderr
    sys.stdout.flush()
    real_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_out.write(json.dumps(obj) + "\n"); real_out.flush()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=STREAMS_PER_GPU, help="camera streams: in total (--scaling strong) or per GPU (--scaling weak)")
    ap.add_argument("--scaling", default=SCALING, choices=["strong", "weak"], help="strong (default, BASELINE configs[3]): --streams in total, stream s tracked by GPU s mod G; "
                    "weak: --streams per GPU")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the fast-motion, weak-scaling, 1080p and 4K legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--host-alloc", default="pinned", choices=["pinned", "wc"], help="how the e2e leg's host frame pool is allocated")
    ap.add_argument("--groups", type=int, default=0, help="vslam_params.stream_groups (0 = library default)")
    ap.add_argument("--lookahead", type=int, default=-1, choices=[-1, 0, 1], help="vslam_params.frame_lookahead (-1 = library default: on up to 160 streams per GPU)")
    ap.add_argument("--frame", default="vga", choices=["vga"] + sorted(FRAME_CONFIGS), help="frame size of the whole-TrackFrame workload: vga (the headline "
                    "config), 1080p (5000 map points, 148 streams per GPU) or 4k (20000 points, 148 streams)")
    ap.add_argument("--sweep", action="store_true", help="SURVEY §8(d) config 5: 4K frames, batch-size sweep of pyramid+FAST and points x range sweep of the patch search (one GPU)")
    ap.add_argument("--cpu-stages", action="store_true", help="SURVEY §8(d)(i): per-stage CPU times of config 1 (reference build and oracle port), no GPU work")
    args = ap.parse_args()
    K, Wm = args.steps, max(args.warmup, 3 if args.impl == "ours" else 1)
    if args.frame != "vga":
        global W, H, N_POINTS, TEX_SIZE, WORKLOAD
        W, H, N_POINTS, s_default = FRAME_CONFIGS[args.frame]
        TEX_SIZE = 4096
        if args.streams == STREAMS_PER_GPU:
            args.streams = s_default
        args.no_cpu_baseline = True
        global POOL
        POOL = int(max(4, min(24, 6e9 // (args.streams * W * H))))      # keep the resident pool (and its pinned host copy) under ~6 GB
        args.scaling = "weak"; args.no_extra_legs = True
        WORKLOAD = (f"{args.frame}: {args.streams} independent synthetic {W}x{H} camera streams per GPU, {N_POINTS} map points, full TrackFrame-equivalent per frame "
                    f"(MaxPatchesPerFrame = 1000 as in the reference), P=11")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from visualslam_android_b200 import sharding
    stream_ids = sharding.streams_for_rank(args.streams, rank, world) if args.scaling == "strong" else sharding.weak_streams(args.streams, rank)
    streams_total = args.streams if args.scaling == "strong" else args.streams * world
    spg = len(sharding.streams_for_rank(args.streams, 0, world)) if args.scaling == "strong" else args.streams
    config = {"workload": WORKLOAD, "streams_total": streams_total, "streams_per_gpu": spg, "frame": [W, H], "map_points": N_POINTS, "patch": 11,
              "parallelism": (f"{streams_total} streams in total, stream s on GPU s mod {world}, no collective" if args.scaling == "strong" else f"{args.streams} streams on each of {world} GPU(s), no collective"),
              "l2": f"inputs larger than L2: each step reads a different {spg * W * H / 1e6:.1f} MB frame set out of a pool of {POOL} sets = {POOL * spg * W * H / 1e6:.0f} MB per GPU (> 126 MB L2; "
                    "triangle-wave order, a set comes round again after up to 46 steps); no explicit flush",
              "e2e_host_frames": "pinned (cudaHostAlloc default)" if args.host_alloc == "pinned" else "pinned, write-combined"}

    if args.impl == "reference":
        if rank != 0:
            return
        cam, tex, f0, smap = build_scene()
        n_procs = int(os.environ.get("VSLAM_BENCH_CPU_PROCS", 0)) or os.cpu_count() or 1
        fps_step = CPU_FRAMES_PER_STEP
        M = pool_size(K, Wm)
        fps, step_s, kind, stats = run_cpu_arm(cam, tex, f0, smap, n_procs, Wm, K, fps_step, M)
        sample = (f"{n_procs} processes = the GPU arm's streams 0..{n_procs - 1} (one tracker each, same seeds, same triangle-wave pool of {M} frame sets, "
                  f"FRAME_STEP {FRAME_STEP}), {fps_step} frames per process per step; unmodified Tracker::TrackFrame (SmallBlurryImage included), "
                  "keyframe 0 registered with the relocaliser")
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": Wm, "ms_per_step": step_s * 1e3,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u8/i32 + f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": fps, "unit": UNIT, "cores": n_procs, "kind": kind, "sample": sample},
                "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0, "tracking": stats}
        emit(line)
        return

    if args.sweep:
        emit(sweep_config5(local_rank))
        return
    if args.cpu_stages:
        emit(cpu_stage_times())
        return

    import torch
    import torch.distributed as dist
    from visualslam_android_b200 import api, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if os.environ.get("VSLAM_BENCH_AFFINITY", "1") == "1":
        # bind this rank to the CPUs NVML reports as local to its GPU, BEFORE the pinned frame pool is allocated (first touch => local NUMA node)
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[local_rank]) if vis and vis.split(",")[0].isdigit() else local_rank
            nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(idx))
        except Exception:
            pass
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    S = len(stream_ids)
    cam, tex, f0, smap = build_scene(on_gpu=True, device=local_rank)
    tex_t = torch.from_numpy(tex.astype(np.float32)).to(dev)
    # A pool of M consecutive frame sets per stream, traversed as a triangle wave (1,2,..,M,M-1,..,1,2,..) so that any number of
    # steps sees a continuous camera motion.  A set is re-read at the earliest two steps later, after > 126 MB of other traffic.
    M = pool_size(K, Wm)
    poses = stream_poses(stream_ids, M + 1)                                     # (S, M+1, 3, 4); index 0 = identity (the source keyframe)
    frames_dev = torch.empty((M, S, H, W), dtype=torch.uint8, device=dev)
    for k in range(1, M + 1):
        rb = 64 if W * H <= 640 * 480 else 8          # render batch (float64 temporaries of B x H x W)
        for s0 in range(0, S, rb):
            frames_dev[k - 1, s0:s0 + rb] = render_frames_torch(tex_t, cam, poses[s0:s0 + rb, k], dev)
    torch.cuda.synchronize()
    if args.host_alloc == "wc":
        # write-combined pinned memory (cudaHostAllocWriteCombined | Portable): the CPU never reads the frame pool, and on a multi-socket
        # host the PCIe reads of WC memory are not snooped
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        ptr = ctypes.c_void_p()
        nbytes = M * S * H * W
        rc = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04 | 0x01))
        if rc != 0:
            raise SystemExit(f"cudaHostAlloc(write-combined) failed: {rc}")
        buf = (ctypes.c_uint8 * nbytes).from_address(ptr.value)
        frames_host = torch.frombuffer(buf, dtype=torch.uint8).view(M, S, H, W)
        rt.cudaMemcpy(ctypes.c_void_p(ptr.value), ctypes.c_void_p(frames_dev.data_ptr()), ctypes.c_size_t(nbytes), ctypes.c_int(2))   # device -> host
        _keep_alive = (rt, buf)
    else:
        frames_host = torch.empty((M, S, H, W), dtype=torch.uint8).pin_memory()
        frames_host.copy_(frames_dev)
    torch.cuda.synchronize()

    _tri = globals()["tri"]
    def tri(j):      # j-th step of the whole run -> index into the pool (module-level tri: the schedule shared with the CPU arms)
        return _tri(j, M)

    stream = torch.cuda.Stream(device=dev)       # an explicit stream: the library launches on it and the CUDA events below are recorded on it
    torch.cuda.set_stream(stream)
    ctx = api.Context(W, H, n_streams=S, max_points=smap.n, device=local_rank, cuda_stream=stream.cuda_stream)
    prm = {}
    if args.groups:
        prm["stream_groups"] = args.groups
    if args.lookahead >= 0:
        prm["frame_lookahead"] = args.lookahead
    if prm:
        ctx.set_params(**prm)
    ctx.set_camera(cam.scalars())
    ctx.enable_sbi(synth.Camera(W // 16, H // 16).scalars())       # SmallBlurryImage + CalcSBIRotation on the device, like the reference's TrackFrame
    ctx.upload_source_keyframe(f0)
    ctx.set_map(smap.world, smap.pix_right_w, smap.pix_down_w, smap.ir_center, smap.src_level)
    config["frame_lookahead"] = ("on: pyramid + FAST + SmallBlurryImage of frame k+1 on a second stream beside projection / search / pose of frame k (two frame sets)"
                                 if ctx.frame_lookahead_active() else "off")
    config["coarse_chain"] = ("library default: while no stream tried TrackMap's coarse stage in its latest frame (and the map has at most 2040 points) the "
                              "coarse stage's kernels run as a launch chain of their own beside the fine stage instead of in front of it")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    fs = H * W
    # ---- leg 1: frames resident in HBM -------------------------------------------------------------------------
    step_no = 0
    for k in range(Wm):
        ctx.track_frame_ptr(frames_dev[tri(step_no)].data_ptr(), W, fs, device=True); step_no += 1
    ctx.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    def value_leg(first_step):
        step = first_step
        launches0 = ctx.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(K):
            ctx.track_frame_ptr(frames_dev[tri(step)].data_ptr(), W, fs, device=True); step += 1
        e1.record(stream)
        barrier()
        return step, ctx.kernel_launches() - launches0, max_over_ranks(e0.elapsed_time(e1))

    step_no, launches, dev_ms = value_leg(step_no)
    remeasured = None
    # per-stage pass (not part of `value`): CUDA events around every launch; the library serialises the side-stream branch
    # (SmallBlurryImage + projection) behind the level-0 launch in this mode, so every stage is timed alone
    Ks = min(K, 25)
    ctx.set_timing(True)
    evals0 = ctx.zmssd_evals()
    stats0 = ctx.search_stats()
    for k in range(Ks):
        ctx.track_frame_ptr(frames_dev[tri(step_no)].data_ptr(), W, fs, device=True); step_no += 1
    barrier()
    evals1 = ctx.zmssd_evals()
    stats1 = ctx.search_stats()
    stage = ctx.stage_times()
    ctx.set_timing(False)
    # a timed region several times longer than the sum of its own kernels means the box stalled (seen once on a cold box: 40 s inside
    # a 2-step region); such a run is re-measured once, and the line says so
    if max_over_ranks(1.0 if dev_ms / K > 3.0 * sum(v[0] for v in stage.values()) / Ks else 0.0) > 0:   # same decision on every rank
        remeasured = dev_ms / K
        step_no, launches, dev_ms = value_leg(step_no)
    ctx.sync()     # raises on corner-capacity overflow
    probe = list(range(0, S, max(1, S // 8)))
    corners_per_step = sum(int(ctx.corners(s, l).shape[0]) for l in range(4) for s in probe) * (S / len(probe))
    found = np.array([ctx.counters(s)[1].sum() for s in range(0, S, max(1, S // 16))])
    quality = np.array([ctx.counters(s)[2] for s in range(0, S, max(1, S // 16))])
    value = streams_total * K / (dev_ms * 1e-3)

    # ---- leg 2: end to end through the host-buffer C-ABI call, pose read-back every step ------------------------------
    # vslam_track_frame_async: pinned host frames -> (copy stream) -> kernels -> every stream's pose copied back to pinned host
    # memory, each step; the copy of step k overlaps the kernels of step k-1 (two level-0 buffers).
    poses_pin = torch.empty((2, S, 12), dtype=torch.float64).pin_memory()
    prev = None
    for k in range(Wm):
        sid = ctx.track_frame_async(frames_host[tri(step_no)].data_ptr(), W, fs, poses_pin[k & 1].data_ptr()); step_no += 1
        if prev is not None:
            ctx.wait_step(prev)
        prev = sid
    ctx.wait_step(prev)
    barrier()

    def e2e_leg(first_step):
        step, acc, prev = first_step, 0.0, None
        t0 = time.perf_counter()
        for k in range(K):
            sid = ctx.track_frame_async(frames_host[tri(step)].data_ptr(), W, fs, poses_pin[k & 1].data_ptr()); step += 1
            if prev is not None:
                ctx.wait_step(prev)
                acc += float(poses_pin[(k - 1) & 1, :, 3].sum())      # the result of step k-1 is consumed on the host
            prev = sid
        ctx.wait_step(prev)
        acc += float(poses_pin[(K - 1) & 1, :, 3].sum())
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        assert np.isfinite(acc)
        return step, max_over_ranks(wall_ms)

    step_no, e2e_ms = e2e_leg(step_no)
    clocks = sampler.stop()

    # ---- what bounds e2e: the host-to-device copy of a step's frames.  Peak of this box, measured here: pinned cudaMemcpyAsync of the same
    # frame sets on the library's copy path alone (no kernels), all ranks copying at the same time.
    h2d_bytes = S * fs
    stage_buf = torch.empty((S, H, W), dtype=torch.uint8, device=dev)
    for k in range(2):
        stage_buf.copy_(frames_host[tri(k)], non_blocking=True)
    barrier()
    nh = max(8, min(K, 40))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(nh):
        stage_buf.copy_(frames_host[tri(step_no + k)], non_blocking=True)
    e1.record(stream)
    barrier()
    h2d_ms = max_over_ranks(e0.elapsed_time(e1)) / nh
    h2d_peak = h2d_bytes / (h2d_ms * 1e-3) / 1e9                       # GB/s per GPU with all ranks copying
    # the e2e leg is timed on the host's clock over K short steps: a step time more than twice both of its bounds (the kernels of the `value`
    # leg, the copy just measured) means the host stalled inside the region (seen once: 0.76 ms per step where 0.29 is the rule, 32 streams);
    # such a leg is re-measured once, and the line says so -- the policy of the `value` leg
    e2e_remeasured = None
    if max_over_ranks(1.0 if e2e_ms / K > 2.0 * max(dev_ms / K, h2d_ms) else 0.0) > 0:      # same decision on every rank
        e2e_remeasured = e2e_ms / K
        step_no, e2e_ms = e2e_leg(step_no)
    e2e_value = streams_total * K / (e2e_ms * 1e-3)
    h2d_achieved = h2d_bytes / (e2e_ms / K * 1e-3) / 1e9
    h2d_roofline = {"bound": "host-to-device copy (PCIe / host memory)", "achieved": h2d_achieved, "peak": h2d_peak, "unit": "GB/s per GPU", "frac": h2d_achieved / h2d_peak,
                    "aggregate_peak_gbs": h2d_peak * world, "bytes_per_step_per_gpu": h2d_bytes,
                    "how": f"pinned cudaMemcpyAsync of {nh} frame sets per rank, all {world} rank(s) at once, CUDA events, max over ranks; achieved = the e2e leg's bytes over its step time"}
    del stage_buf

    # ---- roofline of the pyramid+FAST stage ----------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    stage_sum_ms = sum(v[0] for v in stage.values()) / Ks
    lut_bytes = 4 * sum((H >> l) for l in range(4))
    alg_bytes_step = S * (1.328125 * W * H + lut_bytes) + 4.0 * corners_per_step            # whole stage, per step
    # (a) the COMPLETE KeyFrame::MakeKeyFrame_Lite -- pyramid, FAST-10 of four levels, raster-ordered corner lists and row LUTs: the three
    #     launches of vslam_make_keyframe_lite_dev -- timed live here with CUDA events over the resident frame pool (a different frame set
    #     every launch, > L2 in between), stage timing on so that every launch is also timed alone
    Kr = max(8, min(K, 30))
    for k in range(3):
        ctx.make_keyframe_lite_ptr(frames_dev[tri(step_no + k)].data_ptr(), S, W, fs, device=True)
    ctx.sync()
    ctx.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(Kr):
        ctx.make_keyframe_lite_ptr(frames_dev[tri(step_no + 3 + k)].data_ptr(), S, W, fs, device=True)
    e1.record(stream)
    barrier()
    mk_ms = max_over_ranks(e0.elapsed_time(e1)) / Kr
    mk_stage = ctx.stage_times()
    ctx.set_timing(False)
    mk_l = [mk_stage[f"pyrfast_l{l}"][0] / Kr for l in range(3)]       # level-0 launch, levels 1-3 launch, corner-list launch
    pyr_ms_sum = sum(mk_l)
    achieved = alg_bytes_step / (pyr_ms_sum * 1e-3) / 1e9
    # (b) what a tracked frame runs of it: the two FAST launches (the corner bitmasks are what k_search_fast reads; lists are built on demand)
    in_step_ms = (stage["pyrfast_l0"][0] + stage["pyrfast_l1"][0]) / Ks
    in_step_bytes = S * 1.328125 * W * H * (1.0 + 1.0 / 8.0)                               # frames in, level 1-3 images + one corner bit per pixel out
    # level-0 launch alone (75 % of the FAST pixels): reads level 0, writes the images of levels 1-3 + level 0's corner bits
    l0_bytes = S * (1.328125 * W * H + W * H / 8.0)
    l0_ms = mk_l[0]
    # DRAM bytes and executed warp-instructions of that launch: read from the committed ncu --set full capture of the CURRENT kernels
    # (profiles/r02_ncu_counters.json, written by profiles/tools/ncu_extract.py from the .ncu-rep of `bench.py --steps 2`; a property of the
    # workload, not of the run).  null when the capture does not cover this configuration.
    l0_traffic = l0_inst = None; counters_src = None; traffic_stage = None
    cpath = os.path.join(ROOT, "profiles", "r02_ncu_counters.json")
    if os.path.exists(cpath):
        cj = json.load(open(cpath))
        kk = cj.get("kernels", {})
        if "k_pyramid_fast" in kk and cj.get("streams") == S and cj.get("frame") == [W, H]:
            dram = lambda n: (kk[n].get("dram_bytes_read") or 0.0) + (kk[n].get("dram_bytes_write") or 0.0)
            l0_traffic = dram("k_pyramid_fast"); l0_inst = kk["k_pyramid_fast"]["warp_instructions"]; counters_src = "profiles/r02_ncu_counters.json"
            if all(n in kk for n in ("k_fast_levels", "k_corner_lists")):
                traffic_stage = l0_traffic + dram("k_fast_levels") + dram("k_corner_lists")
    sm_clock_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
    issue = None
    if l0_inst:
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        issue = {"bound": "instruction issue", "warp_instructions_per_launch": l0_inst, "achieved": l0_inst / (l0_ms * 1e-3) / 1e9,
                 "peak": sms * 4 * sm_clock_hz / 1e9, "unit": "G warp-instr/s", "frac": l0_inst / (l0_ms * 1e-3) / (sms * 4 * sm_clock_hz)}
    roofline = {"kernel": "k_pyramid_fast + k_fast_levels + k_corner_lists = KeyFrame::MakeKeyFrame_Lite (pyramid + FAST-10 + raster-ordered corner lists + row LUT; 3 launches: "
                          "level 0 (+ level 1-3 images), levels 1-3, lists)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic_stage if traffic_stage is not None else l0_traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_step": alg_bytes_step, "ms_per_step": pyr_ms_sum, "ms_per_launch": {"level0": mk_l[0], "levels1_3": mk_l[1], "corner_lists": mk_l[2]},
                "ms_per_step_back_to_back": mk_ms,
                "how": f"{Kr} calls of vslam_make_keyframe_lite_dev on {S} resident frames each (a different set of the pool per call), CUDA events around every launch on the library's stream; "
                       "ms_per_step = sum of the three launches' own times, ms_per_step_back_to_back = one event pair around all calls",
                "in_tracked_frame": {"what": "the two FAST launches a tracked frame runs (corner bitmasks out; k_search_fast reads them, lists are built on demand)",
                                     "ms": in_step_ms, "algorithmic_bytes": in_step_bytes, "achieved": in_step_bytes / (in_step_ms * 1e-3) / 1e9,
                                     "frac": in_step_bytes / (in_step_ms * 1e-3) / 1e9 / peak, "share_of_step": in_step_ms / stage_sum_ms},
                "share_of_step": in_step_ms / stage_sum_ms,
                "level0_launch": {"algorithmic_bytes": l0_bytes, "ms": l0_ms, "achieved": l0_bytes / (l0_ms * 1e-3) / 1e9,
                                  "frac": l0_bytes / (l0_ms * 1e-3) / 1e9 / peak, "traffic": l0_traffic, "issue_roofline": issue},
                "counters_source": counters_src,
                "note": "traffic = dram read+write bytes of the three launches (level0_launch.traffic: of the level-0 launch) from the committed ncu --set full capture; "
                        "the stage is instruction-issue bound, not HBM bound: DESIGN.md section 4.1 and profiles/"}
    # ZMSSD: 3*P^2 integer MACs per scored candidate over the time of the two search kernels, against a measured dp4a peak
    evals_timed = evals1 - evals0
    search_ms = stage["search_fine"][0] + stage["search_coarse"][0]
    dp4a_peak = api.dp4a_peak_tmacs()
    zm_achieved = 3.0 * 11 * 11 * evals_timed / (search_ms * 1e-3) / 1e12
    zmssd = {"kernel": "k_search (template generation + FindPatchCoarse/ZMSSD + sub-pixel refinement)", "bound": "integer pipe", "achieved": zm_achieved,
             "peak": dp4a_peak, "unit": "TMAC/s", "frac": zm_achieved / dp4a_peak, "candidates_scored_per_step": evals_timed / Ks,
             "peak_source": "dp4a micro-benchmark of this run (vslam_debug_dp4a_peak)",
             "note": "ZMSSD is a small part of k_search (about 6 candidates per point); the kernel is latency / issue bound, see profiles/"}
    # SURVEY.md §8(d) per-stage rates: templates/s, sub-pixel points/s (counted on a sample of streams from the per-point flags the last
    # step left behind, scaled to all streams) over the search kernels' time; Gauss-Newton iterations/s per stream over the pose kernels' time
    sample_streams = list(range(0, S, max(1, S // 8)))
    attempted_per_frame = float(np.mean([ctx.counters(s_)[0].sum() for s_ in sample_streams]))
    n_tmpl = (stats1["templates_generated"] - stats0["templates_generated"]) / Ks       # device counters over the Ks steps of the stage pass
    n_subpix = (stats1["subpix_refinements"] - stats0["subpix_refinements"]) / Ks
    n_searched = attempted_per_frame * S
    pose_ms = (stage["pose_fine"][0] + stage["pose_coarse"][0]) / Ks
    its_per_frame = float(np.mean([len(ctx.updates(s_)[0]) for s_ in sample_streams]))
    stage_rates = {"templates_generated_per_sec": n_tmpl / (search_ms / Ks * 1e-3), "subpix_points_per_sec": n_subpix / (search_ms / Ks * 1e-3),
                   "points_searched_per_sec": n_searched / (search_ms / Ks * 1e-3),
                   "wls_iterations_per_sec_per_stream": its_per_frame / (pose_ms * 1e-3), "wls_iterations_per_frame": its_per_frame,
                   "per_step": {"points_searched": n_searched, "templates_generated": n_tmpl, "subpix_points": n_subpix},
                   "note": f"templates / sub-pixel refinements: device counters over the {Ks} steps of the stage pass (template regeneration is bursty: the 0.07 "
                           "reuse test of MakeTemplateCoarseCont trips for many points in the same frame); points searched: attempted counters of the last "
                           f"frame on {len(sample_streams)} sampled streams; times from the serialised stage pass"}
    stages_ms = {k: round(v[0] / Ks, 4) for k, v in stage.items() if v[1]}
    la_on = config["frame_lookahead"] != "off"
    stages_ms["note"] = ("timed in a separate serialised pass; in the `value` leg " +
                         ("the front end of a frame (`pyrfast_l0`, `pyrfast_l1` = levels 1-3, SmallBlurryImage = `other`) runs on a second stream beside the previous "
                          "frame's back end (`project_lists`, searches, pose iterations): frame look-ahead, so the step time approaches the back end's sum"
                          if la_on else
                          "SmallBlurryImage + projection (`other`, `project_lists`) run on a side stream beside `pyrfast_l1` (= levels 1-3), so the stages sum to more than ms_per_step"))

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": dev_ms / K, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "u8/i32 (pyramid, FAST, ZMSSD) + f64 (projection, WLS)", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": S * H * W * world, "d2h_bytes_per_step": S * 12 * 8 * world, "ms_per_step": e2e_ms / K,
                    "limiter": "host-to-device copy" if h2d_achieved > 0.8 * h2d_peak else "kernels (copy overlapped)",
                    **({"remeasured_after_stall_ms_per_step": e2e_remeasured} if e2e_remeasured else {})},
            "h2d_roofline": h2d_roofline,
            "gpu_launches": int(launches), "clocks": clocks, **({"remeasured_after_stall_ms_per_step": remeasured} if remeasured else {}), "roofline": roofline, "zmssd_roofline": zmssd, "stages_ms_per_step": stages_ms, "stage_rates": stage_rates,
            "tracking": {"found_per_frame_mean": float(found.mean()), "quality_good_frac": float((quality == 2).mean()),
                         "zmssd_evals_total": int(ctx.zmssd_evals())}}

    top_stage = max(((k, v) for k, v in stages_ms.items() if k != "note"), key=lambda kv: kv[1])[0]
    back_ms = sum(stages_ms.get(k, 0.0) for k in ("project_lists", "search_coarse", "pose_coarse", "search_fine", "pose_fine"))
    line["limiter"] = {"value": (f"per-stream latency chain of a frame's back end (projection -> searches -> pose iterations: {back_ms:.3f} ms of one-CTA-per-stream / "
                                 f"dependent kernels whose length does not shrink with the stream count), largest stage: {top_stage}" if la_on
                                 else f"kernel time, largest stage: {top_stage}"), "e2e": line["e2e"]["limiter"]}
    ctx.close()
    del frames_dev, frames_host
    torch.cuda.empty_cache()

    # ---- extra legs (own contexts, frames resident, CUDA events, max over ranks): every rank takes part
    if not args.no_extra_legs:
        Ke, We = max(5, min(K, 20)), 3
        vga = (W, H, N_POINTS, TEX_SIZE)
        scene_vga = (cam, tex, f0, smap)
        # (1) fast camera: the motion model's speed gate (jni/Tracker.cc:815-819,425) switches the coarse stage on in every frame -- 10 + 10
        #     Gauss-Newton iterations -- and templates are regenerated as the warp moves past the 0.07 re-use test
        ms, info = short_leg("fast_motion", vga, stream_ids, Ke, We, local_rank, stream, frame_step=6 * FRAME_STEP, pool=10, barrier=barrier, scene=scene_vga)
        ms = max_over_ranks(ms)
        line["fast_motion"] = {"value": streams_total / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, **info}
        # (2) the other scaling mode, so that one line carries both
        if world > 1:
            other = "weak" if args.scaling == "strong" else "strong"
            ids2 = sharding.weak_streams(args.streams, rank) if other == "weak" else sharding.streams_for_rank(args.streams, rank, world)
            ms, info = short_leg(other, vga, ids2, Ke, We, local_rank, stream, pool=12, barrier=barrier, scene=scene_vga)
            ms = max_over_ranks(ms)
            tot2 = args.streams * world if other == "weak" else args.streams
            line[other + "_scaling"] = {"value": tot2 / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "streams_total": tot2, **info}
        # (3) the other frame sizes of north_star, 148 streams per GPU = one CTA of the per-stream kernels per SM
        line["other_sizes"] = {}
        for nm, (w2, h2, np2, s2) in FRAME_CONFIGS.items():
            ms, info = short_leg(nm, (w2, h2, np2, 4096), sharding.weak_streams(s2, rank), max(4, Ke // 2), 3, local_rank, stream, pool=4, barrier=barrier)
            ms = max_over_ranks(ms)
            line["other_sizes"][nm] = {"value": s2 * world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "scaling": "weak", **info}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_procs = os.cpu_count() or 1
        fps_step, cpu_steps = CPU_FRAMES_PER_STEP, 6
        fps, step_s, kind, stats = run_cpu_arm(cam, tex, f0, smap, n_procs, 1, cpu_steps, fps_step, M)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": n_procs, "kind": kind,
                                "sample": f"{n_procs} processes = streams 0..{n_procs - 1} of this run x {cpu_steps * fps_step} frames of the same triangle-wave pool "
                                          f"({M} frame sets, 1000 map points), "
                                          f"{'reference jni/ sources compiled by oracle/build_ref.sh' if kind == 'reference' else 'oracle port'}, "
                                          "whole Tracker::TrackFrame (SmallBlurryImage included) on both sides",
                                "tracking": stats}
        gf, cf = line["tracking"]["found_per_frame_mean"], stats["found_per_frame_mean"]
        line["cpu_baseline"]["same_workload"] = bool(abs(gf - cf) <= 0.02 * max(gf, cf))      # found points per frame agree within 2 %
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


def render_cpu_or_gpu(tex, cam, poses):
    """(n_seq, n_frames, 3, 4) poses -> uint8 frames; GPU renderer if a device is visible, numpy otherwise."""
    try:
        import torch
        if torch.cuda.is_available():
            dev = torch.device("cuda", 0)
            tex_t = torch.from_numpy(tex.astype(np.float32)).to(dev)
            out = [torch.cat([render_frames_torch(tex_t, cam, poses[s, k0:k0 + 32], dev) for k0 in range(0, poses.shape[1], 32)]).cpu().numpy()
                   for s in range(poses.shape[0])]
            return np.stack(out)
    except Exception:
        pass
    from visualslam_android_b200 import synth
    return np.stack([np.stack([synth.render_frame(tex, cam, poses[s, k]) for k in range(poses.shape[1])]) for s in range(poses.shape[0])])


if __name__ == "__main__":
    main()

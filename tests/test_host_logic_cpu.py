"""CPU: synthetic-input determinism and the multi-GPU (stream-sharding) host logic on 2 gloo ranks."""
import os
import subprocess
import sys

import numpy as np

from visualslam_android_b200 import sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synthetic_inputs_are_deterministic():
    a, b = synth.make_texture(256), synth.make_texture(256)
    assert np.array_equal(a, b) and a.std() > 20
    cam = synth.Camera(160, 120)
    p = synth.stream_pose(7, 3)
    assert np.array_equal(synth.render_frame(a, cam, p), synth.render_frame(b, cam, p))
    assert np.allclose(synth.stream_pose(0, 5), synth.IDENTITY_POSE)
    assert np.allclose(synth.se3_exp(np.zeros(6)), synth.IDENTITY_POSE)
    R = synth.se3_exp([0.1, -0.2, 0.3, 0.4, -0.5, 0.6])[:, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-14)


def test_shard_assignment_partitions_streams():
    for total, world in ((256, 1), (256, 8), (10, 4), (3, 8)):
        seen = []
        for r in range(world):
            ids = sharding.streams_for_rank(total, r, world)
            assert all(i % world == r for i in ids)          # stream s -> GPU (s mod G)
            seen += ids
        assert sorted(seen) == list(range(total))
    assert sharding.weak_streams(256, 3) == list(range(768, 1024))


WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from visualslam_android_b200 import sharding
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%d" %% int(sys.argv[2]), rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
ids = sharding.streams_for_rank(10, rank, 2)
elapsed_ms = 5.0 if rank == 0 else 7.5          # the slower rank defines the step
frames = len(ids) * 4
agg = sharding.aggregate_throughput(frames, elapsed_ms, dist)
print("RESULT", rank, len(ids), agg)
dist.destroy_process_group()
"""


def test_two_rank_gloo_aggregation():
    port = 29500 + os.getpid() % 2000
    code = WORKER % ROOT
    procs = [subprocess.Popen([sys.executable, "-c", code, str(r), str(port)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    vals = []
    for o, _ in outs:
        line = [l for l in o.splitlines() if l.startswith("RESULT")][0].split()
        vals.append((int(line[2]), float(line[3])))
    assert [v[0] for v in vals] == [5, 5]
    # 10 streams x 4 frames over max(5, 7.5) ms on both ranks
    assert all(abs(v[1] - 40 / 7.5e-3) < 1e-6 for v in vals)


def test_reference_algorithm_amplifies_a_one_ulp_perturbation():
    """Why long free-running sequences cannot be held to a fixed tolerance (tests/test_gpu_config2.py): the ORACLE against
    ITSELF, one copy with t_x of its pose nudged by 1e-15 after frame 10.  Same code, same libm, same frames — and the two runs drift
    apart by roughly x1.3 per frame (the motion model feeds the pose difference back), i.e. by many orders of magnitude within 60 frames.
    Any implementation that is not bit-identical in every floating-point operation meets the same fate."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common
    from oracle import oraclebind
    cam, f0, smap = common.scene()
    W, H = cam.width, cam.height
    sbi = synth.Camera(W // 16, H // 16).scalars()
    a, b = oraclebind.OrcWorld(cam, f0, smap), oraclebind.OrcWorld(cam, f0, smap)
    for o in (a, b):
        o.L.orc_tracker_enable_sbi(o.tracker, sbi)
    diff = []
    for k in range(1, 71):
        fr = synth.render_frame(common.texture(), cam, synth.stream_pose(k, 0))
        for o in (a, b):
            o.L.orc_tracker_track_frame(o.tracker, fr, W, H, W)
        if k == 10:
            assert np.array_equal(a.get_pose(), b.get_pose())          # identical until the nudge
            p = b.get_pose().copy(); p[0, 3] += 1e-15; b.set_pose(p)
        diff.append(np.abs(a.get_pose() - b.get_pose()).max())
    assert diff[11] < 1e-14
    assert diff[69] > 1e4 * diff[11], (diff[11], diff[69])                # > 4 orders of magnitude in 58 frames
    growth = (diff[59] / diff[11]) ** (1.0 / 48)
    assert 1.1 < growth < 1.8, growth


def test_reference_arm_runs_on_the_gpu_arms_workload():
    """`bench.py --impl reference` (the driver's reference arm) on a tiny budget: exits 0, one JSON line, tracks the triangle-wave pool the
    GPU arm uses (about 950 of the 1000 map points found per frame) and registers keyframe 0 with the relocaliser."""
    import json
    from oracle import refbind
    env = {**os.environ, "VSLAM_BENCH_CPU_PROCS": "2"}
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] == ("reference" if refbind.available() else "port")
    assert line["tracking"]["found_per_frame_mean"] > 900

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


def _shuffle_without_the_chain(v0, tgt, m):
    """The rules of shuffle_parallel (visualslam_android_b200/csrc/track.cu) restated with numpy: buckets of steps by target, every step either
    names its source element or links to an earlier step, every output position follows the links.  No sequential state anywhere."""
    n = len(v0)
    buckets = [[] for _ in range(n)]
    for k in range(n):
        if tgt[k] != k:
            buckets[tgt[k]].append(k)                       # (unordered in the kernel: filled by atomics)
    res = np.empty(n, dtype=np.int64)
    last = np.full(n, -1, dtype=np.int64)
    for k in range(n):
        q = tgt[k]
        if q == k:
            res[k] = k
        else:
            earlier = [x for x in buckets[q] if x < k]
            res[k] = max(earlier) if earlier else -(q + 1)
        if buckets[k]:
            last[k] = max(buckets[k])
    out = np.empty(m, dtype=v0.dtype)
    for p in range(m):
        r = last[p]
        if r < 0:
            r = res[p]
            while r < 0:
                r = res[-(r + 1)]
        out[p] = v0[r]
    return out


def test_shuffle_without_the_swap_chain_equals_std_random_shuffle():
    """std::random_shuffle is `for k in 1..n-1: swap(v[k], v[rand() % (k+1)])` (libstdc++ bits/stl_algo.h:4581-4597).  The kernel that builds the
    tracker's search lists replaces that chain by a data-parallel evaluation; this is its rule set against the chain, on
    single shuffles, on truncated ones (TrackMap keeps the first MaxPatchesPerFrame elements of the fifth shuffle) and on several independent
    segments handled as one problem (the four level lists).  The CUDA code itself is held to the oracle's shuffle by the GPU parity tests."""
    rs = np.random.RandomState(5)
    for n in [1, 2, 3, 7, 64, 257, 1500]:
        for _ in range(6):
            v0 = rs.permutation(10 * n + 3)[:n]
            tgt = np.array([0] + [rs.randint(0, k + 1) for k in range(1, n)], dtype=np.int64)
            if n > 3 and rs.rand() < 0.5:
                tgt[rs.randint(1, n, size=n // 3)] = rs.randint(0, 2)           # crowd the low positions: long buckets, links to links
                tgt = np.minimum(tgt, np.arange(n))
            v = v0.copy()
            for k in range(1, n):
                j = tgt[k]
                v[k], v[j] = v[j], v[k]
            for m in sorted({n, min(n, 5), n // 2}):
                assert np.array_equal(_shuffle_without_the_chain(v0, tgt, m), v[:m]), (n, m)
    # four segments at once: a segment's first element has no step (target = itself), targets are positions of the packed array
    sizes = [300, 0, 1, 450]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    v0 = rs.permutation(offs[-1])
    tgt = np.arange(offs[-1])
    v = v0.copy()
    for q, nseg in enumerate(sizes):
        for k in range(1, nseg):
            j = rs.randint(0, k + 1)
            tgt[offs[q] + k] = offs[q] + j
            a, b = offs[q] + k, offs[q] + j
            v[a], v[b] = v[b], v[a]
    assert np.array_equal(_shuffle_without_the_chain(v0, tgt, len(v0)), v)


def test_glibc_rand_jump_matrix_reproduces_the_serial_generator():
    """glibc's TYPE_3 rand() is r[i] = r[i-3] + r[i-31] (mod 2^32), output r[i] >> 1: linear, so 31 * B draws are one 31 x 31 matrix applied to the
    ring.  k_project_lists draws a frame's random numbers in chunks of 248 started from M^c * ring (glibc_rand_fill_parallel, M = T^8 computed by
    upload_rand_jump the way it is computed here); this checks chunk starts and draws against the generator run serially."""
    rs = np.random.RandomState(9)
    B = 8

    def block(r, out=None):
        for q in range(31):
            r[(q + 3) % 31] = (r[(q + 3) % 31] + r[q]) & 0xffffffff
            if out is not None:
                out.append(r[(q + 3) % 31] >> 1)

    M = np.zeros((31, 31), dtype=object)
    for j in range(31):
        r = [0] * 31
        r[j] = 1
        for _ in range(B):
            block(r)
        for i in range(31):
            M[i, j] = r[i]
    ring0 = [int(x) for x in rs.randint(0, 2 ** 32, 31, dtype=np.uint64)]
    n = 1000                                                  # a VGA frame's draws: 4 full chunks + 8 draws
    serial, r = [], list(ring0)
    while len(serial) < n + 31:
        block(r, serial)
    serial = serial[:n]
    chunked, start = [], list(ring0)
    for c in range((n + 31 * B - 1) // (31 * B)):
        r, out = list(start), []
        for _ in range(B):
            block(r, out)
        chunked += out
        start = [int(sum(M[i, j] * start[j] for j in range(31)) & 0xffffffff) for i in range(31)]
        assert start == r                                      # the jump lands where the serial generator is
    assert chunked[:n] == serial

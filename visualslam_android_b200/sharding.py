"""Host logic for multi-GPU runs: camera streams are independent, so they are sharded across GPUs with no collective on the
data path.  The only communication is the final max-over-ranks of the timed region."""
from __future__ import annotations


def streams_for_rank(total_streams: int, rank: int, world: int) -> list[int]:
    """Strong-scaling assignment: stream s lives on GPU (s mod world)."""
    return [s for s in range(total_streams) if s % world == rank]


def weak_streams(streams_per_gpu: int, rank: int) -> list[int]:
    """Weak-scaling assignment used by bench.py: every GPU tracks its own block of `streams_per_gpu` streams."""
    return list(range(rank * streams_per_gpu, (rank + 1) * streams_per_gpu))


def aggregate_throughput(frames_this_rank: int, elapsed_ms: float, dist=None) -> float:
    """Whole-job frames/s: total frames of all ranks over the slowest rank's device time."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return frames_this_rank / (elapsed_ms * 1e-3)
    import torch
    t = torch.tensor([elapsed_ms], dtype=torch.float64)
    f = torch.tensor([float(frames_this_rank)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t, f = t.cuda(), f.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    return float(f.item()) / (float(t.item()) * 1e-3)

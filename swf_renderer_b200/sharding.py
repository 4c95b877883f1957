"""Frame sharding across GPUs.

Frames (and morph-ratio steps) share no mutable state, so they shard over ranks with NO collective on the data
path: frame ``f`` goes to rank ``f mod world``; every rank owns one renderer handle on its own device and
registers the asset store locally (definitions are replicated, exactly like a per-process reference renderer).
The only communication is control-plane: a barrier around timed regions and a MAX reduction of the per-rank
device time (torch.distributed, NCCL on GPUs / gloo on CPU for the tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def frames_of_rank(n_frames: int, rank: int, world: int) -> List[int]:
    """Global frame indices rendered by ``rank`` (round-robin, paint order inside a frame is never split)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_frames, world))


def owner_of_frame(frame: int, world: int) -> Tuple[int, int]:
    """(rank, local slot) of a global frame index."""
    return frame % world, frame // world


def gather_order(n_frames: int, world: int) -> List[Tuple[int, int]]:
    """For each global frame, where a gather (optional, off the hot path) finds it: (rank, local slot)."""
    return [owner_of_frame(f, world) for f in range(n_frames)]


def reduce_max_time(ms: float, group=None) -> float:
    """MAX over ranks of a device time, as the benchmark contract requires (never a wall clock)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(ms)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def total_throughput(units_per_rank: Sequence[float], ms_max: float) -> float:
    """Whole-job units per second = everything all ranks processed / the slowest rank's time."""
    return float(sum(units_per_rank)) / (ms_max / 1e3)

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


def gather_frames(local_frames, n_frames: int, dst: int = 0, group=None):
    """Optional gather of finished frames onto rank ``dst`` (off the hot path, its own line in the bench).

    ``local_frames``: this rank's frames as a tensor [n_local, H, W, 4] in local-slot order - on the GPU the
    renderer's own frame store (`HeadlessRenderer.device_frames()`, no copy), so with the NCCL backend the pixels travel
    GPU to GPU over NVLink / NVSwitch; on CPU (gloo) a host tensor.  ``n_frames`` is the global frame count.  Returns
    [n_frames, H, W, 4] in global frame order on ``dst`` and None elsewhere.  Frame f sits on rank f mod world in slot
    f div world, so the ranks' tensors interleave; ranks with one frame fewer are padded for the collective."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        if local_frames.shape[0] != n_frames:
            raise ValueError("rank holds %d frames, expected %d" % (local_frames.shape[0], n_frames))
        return local_frames
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = len(range(rank, n_frames, world))
    if local_frames.shape[0] != mine:
        raise ValueError("rank %d holds %d frames, expected %d" % (rank, local_frames.shape[0], mine))
    slots = (n_frames + world - 1) // world
    send = local_frames
    if mine < slots:  # pad to the common slot count
        pad = torch.zeros((slots - mine,) + tuple(local_frames.shape[1:]), dtype=local_frames.dtype, device=local_frames.device)
        send = torch.cat([local_frames, pad], dim=0)
    send = send.contiguous()
    parts = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, parts, dst=dst, group=group)
    if rank != dst:
        return None
    # parts[r][s] is global frame s * world + r: stack along a new rank axis and flatten (slot, rank) -> frame
    out = torch.stack(parts, dim=1).reshape((slots * world,) + tuple(send.shape[1:]))
    return out[:n_frames]


def peer_gather_frames(renderer, dst: int = 0, group=None):
    """The same gather inside the library (`swfr_gather_frames`): every rank exports its renderer's frame store (a CUDA
    IPC handle, 128 bytes over the control plane), rank ``dst`` copies the stores GPU to GPU with one strided
    asynchronous peer copy per rank (NVLink / NVSwitch; no NCCL, no staging tensor) into its gather buffer, in global
    frame order.  Returns (frames tensor [n_frames, H, W, 4], device ms of the copies) on ``dst`` and (None, None)
    elsewhere.  The sources must not render again before the barrier at the end."""
    import torch.distributed as dist

    exp = renderer.export_frames()  # waits for this rank's render
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return renderer.gather_frames([exp])
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    exports = [None] * world
    dist.all_gather_object(exports, exp, group=group)  # control plane only: 128 bytes per rank
    out = renderer.gather_frames(exports) if rank == dst else (None, None)
    dist.barrier(group=group)  # the sources' stores are free again
    return out


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

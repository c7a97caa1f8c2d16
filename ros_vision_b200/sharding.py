"""Frame / camera-stream sharding across the GPUs of one box.

The reference scales by running one detector process per camera on one GPU
(src/ros_vision_launch/launch/launch_vision.py:231-310).  On an 8-GPU box the unit of work stays the
frame (or the camera stream) and there is no exchange step: rank r owns its frames, runs its own
detector on its own GPU and only the tiny detection records are gathered.  torch.distributed is used
for the barrier / max-over-ranks timing and the optional result gather -- never for frame data.
"""
from __future__ import annotations

from typing import Dict, List, Sequence


def frames_for_rank(num_frames: int, world_size: int, rank: int) -> List[int]:
    """Global batch -> ranks: frame f goes to rank f mod G (config 5)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return list(range(rank, num_frames, world_size))


def streams_for_rank(num_streams: int, world_size: int, rank: int) -> List[int]:
    """Camera streams -> GPUs: stream s goes to GPU s mod G (config 4: one 1600x1200 stream per GPU)."""
    return frames_for_rank(num_streams, world_size, rank)


def partition_is_valid(num_items: int, world_size: int) -> bool:
    seen = sorted(i for r in range(world_size) for i in frames_for_rank(num_items, world_size, r))
    return seen == list(range(num_items))


def gather_detections(local: Dict[int, Sequence], group=None) -> Dict[int, Sequence] | None:
    """Collects {frame index: detections} from every rank on rank 0 (a few hundred bytes per frame).

    Uses torch.distributed's object gather (gloo or nccl); returns the merged dict on rank 0, None elsewhere.
    """
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(local)
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    out = [None] * world if rank == 0 else None
    dist.gather_object(dict(local), out, dst=0, group=group)
    if rank != 0:
        return None
    merged: Dict[int, Sequence] = {}
    for part in out:
        for k, v in part.items():
            if k in merged:
                raise RuntimeError(f"frame {k} was processed by two ranks")
            merged[k] = v
    return merged


def max_over_ranks(seconds: float, device=None, group=None) -> float:
    """Timed-region length of the whole job = the slowest rank's."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())

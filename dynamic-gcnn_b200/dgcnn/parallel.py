"""Data-parallel plumbing (device-agnostic host logic).

The reference's only parallelism is synchronous data parallel over in-process towers with a CPU-side mean of the
per-tower gradients (/root/reference/dgcnn/trainval.py:26-29,59-69) and the batch cut into one MINIBATCH_SIZE
slice per GPU (main_funcs.py:145-152).  Here: one process per GPU, towers mapped onto ranks, every tower adds
grad/len(GPUS) into one flat buffer, and ONE all-reduce(sum) of that buffer per optimizer step reproduces the
tower mean.  EdgeConv itself is per cloud and never shards.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def tower_assignment(num_towers: int, world: int, rank: int) -> List[int]:
    """Towers (entries of flags.GPUS) executed by `rank`.  world == towers: one each; fewer ranks than towers:
    a contiguous block each (single process: all of them); more ranks than towers: the surplus ranks idle."""
    if num_towers <= 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("bad tower assignment request towers=%d world=%d rank=%d" % (num_towers, world, rank))
    if num_towers % world and world % num_towers:
        raise ValueError("len(--gpus)=%d is incompatible with world size %d" % (num_towers, world))
    if world >= num_towers:
        stride = world // num_towers
        return [rank // stride] if rank % stride == 0 else []
    per = num_towers // world
    return list(range(rank * per, (rank + 1) * per))


def allreduce_flat_(flat: torch.Tensor) -> torch.Tensor:
    """The single collective of a training step: in-place SUM over ranks of the flat gradient bucket
    (each tower pre-divided its contribution by len(GPUS), so the sum is trainval.py:64-69's mean)."""
    world, _ = world_info()
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat

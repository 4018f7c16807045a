"""Data-parallel plumbing (device-agnostic host logic).

The reference's only parallelism is synchronous data parallel over in-process towers with a CPU-side mean of the
per-tower gradients (/root/reference/dgcnn/trainval.py:26-29,59-69) and the batch cut into one MINIBATCH_SIZE
slice per GPU (main_funcs.py:145-152).  Here: one process per GPU, towers mapped onto ranks, every tower adds
grad/len(GPUS) into one flat buffer, and ONE all-reduce(sum) of that buffer per optimizer step reproduces the
tower mean.  EdgeConv itself is per cloud and never shards.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

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


HEAD_SCOPES = ("MergedEdgeConv", "FC", "Final")     # model.py:65-101: everything after the EdgeConv stack


def head_split_offset(names: Sequence[str], numels: Sequence[int]) -> Optional[int]:
    """Offset (in elements) inside the flat buffer where the head's variables start, if they form its tail.

    Backward produces the head's gradients (Final, FC*, MergedEdgeConv: ~96 % of the bytes) first and the EdgeConv
    stack's last; the flat buffer is laid out in declaration order (EdgeConv*, then the head), so "the head" is a
    suffix of the buffer and can be reduced while the EdgeConv backward is still running.  None: no such suffix."""
    def is_head(n: str) -> bool:
        parts = n.split("/")
        return any(p == s or (s == "FC" and p.startswith("FC") and p[2:].isdigit()) for p in parts for s in HEAD_SCOPES)
    flags = [is_head(n) for n in names]
    if True not in flags:
        return None
    first = flags.index(True)
    if first == 0 or not all(flags[first:]):
        return None
    return int(sum(numels[:first]))


class GradBuckets(object):
    """The flat gradient buffer reduced as TWO in-place SUM all-reduces instead of one, so that the first overlaps the
    rest of the backward pass (the reference averages tower gradients only after every tower is done,
    trainval.py:59-69; the mean is linear, so reducing a finished slice early changes nothing):

        head = flat[split:]   head variables + the trailing loss / accuracy slots -- complete as soon as the head's
                              backward is; reduce_head_async() right there (asynchronous: NCCL's stream runs it
                              while the caller's stream carries on with the EdgeConv backward)
        tail = flat[:split]   EdgeConv variables; reduce_tail() after backward, then wait()

    Device-agnostic (gloo on CPU in the tests, NCCL on the GPUs; inside a CUDA-graph capture the two collectives and
    their fork / join become graph nodes)."""

    def __init__(self, flat: torch.Tensor, split: int):
        if not (0 < int(split) < flat.numel()):
            raise ValueError("split %d outside the flat buffer of %d elements" % (int(split), flat.numel()))
        self.flat, self.split = flat, int(split)
        self._works = []

    def reduce_head_async(self) -> None:
        world, _ = world_info()
        if world >= 1 and dist.is_available() and dist.is_initialized():
            self._works.append(dist.all_reduce(self.flat[self.split:], op=dist.ReduceOp.SUM, async_op=True))

    def reduce_tail(self) -> None:
        world, _ = world_info()
        if world >= 1 and dist.is_available() and dist.is_initialized():
            self._works.append(dist.all_reduce(self.flat[:self.split], op=dist.ReduceOp.SUM, async_op=True))

    def wait(self) -> None:
        """the caller's stream (CUDA) or thread (CPU) continues only after both buckets hold the sums"""
        works, self._works = self._works, []
        for w in works:
            if w is not None:
                w.wait()

"""Multi-GPU plumbing of the RoI path: images shard across ranks, there is NO data-path collective.

The reference shards the same way: detectron2's launch() starts one process per GPU and the unlabeled
loader hands every rank ``IMG_PER_BATCH_UNLABEL / world_size`` images (coin/data/build.py:153-159); every
op on the RoI path is per image (trainer.py:465, clip_roi_heads.py:297,351, rpn.py:149,209). The only
collective of the reference on this path's step is DDP's gradient all-reduce, which is outside it.
torch.distributed is used here for exactly two things: a barrier around the timed region and the
max-over-ranks reduction of the measured times (backend nccl on GPUs, gloo in the CPU tests).
"""
from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_images(n_images: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment of a global batch's image indices to ranks (weak scaling: the bench gives
    every rank the same number of images; a global batch is split like this)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} out of [0, {world})")
    return list(range(rank, n_images, world))


def max_over_ranks(values: Sequence[float], device=None) -> List[float]:
    """Element-wise maximum over all ranks (identity when torch.distributed is not initialised)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def whole_job_rate(images_per_rank: int, world: int, steps: int, ms_max_over_ranks: float) -> float:
    """Aggregate images/s of the job: all ranks' images over the slowest rank's time."""
    return world * images_per_rank * steps / (ms_max_over_ranks / 1e3)

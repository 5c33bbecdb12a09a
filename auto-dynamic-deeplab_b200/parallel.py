"""Batch sharding across the GPUs of one box.

Independent units = images: one process per GPU, contiguous batch shards, replicated weights and
**no data-path collective** — inference is embarrassingly batch-parallel (the reference's own
multi-GPU inference is `nn.DataParallel` scatter/gather, train.py:177).  The only exchange is the
optional final all-reduce of the [nc, nc] int64 confusion matrix when a global mIoU is wanted
(`Evaluator.all_reduce`).  `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is plumbing."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of `n_items` for `rank` (first `n_items % world` ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; (0, 1, 0) when not launched by it."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def all_reduce_confusion(cm_int64: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Sum an int64 confusion matrix over all ranks (in place); no-op when not distributed."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(cm_int64, op=dist.ReduceOp.SUM, group=group)
    return cm_int64

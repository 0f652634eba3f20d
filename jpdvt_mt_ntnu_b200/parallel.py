"""Multi-GPU plumbing for the JPDVT hot path: one process per GPU, torch.distributed over NCCL/NVLink.

Sampling shards naturally - puzzles are independent (the reference strides its file list by rank,
image_model/inference_ddp.py:325) - so ranks exchange nothing during the 250-step loop; the only collectives are the
final gather of the int32 placements and the reference's closing statistics (SUM of [puzzles correct, pieces correct,
count], MAX of the wall time: inference_ddp.py:485-495).  Everything here also runs on the gloo backend (CPU tests).
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: str = "nccl") -> Tuple[int, int, int]:
    rank, size, local = world()
    if size > 1 and not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, size, local


def strided_shard(items: Sequence, rank: int, world_size: int) -> List:
    """The reference's partition: image_paths[rank::world_size] (inference_ddp.py:325)."""
    return list(items[rank::world_size])


def block_shard(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) puzzle range of a rank; sizes differ by at most one."""
    base, extra = divmod(total, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_placements(pred: torch.Tensor, counts: Sequence[int] | None = None, group=None) -> torch.Tensor:
    """All ranks' int32 placements [n_r, G*G] -> [sum n_r, G*G] in rank order (ragged shards are padded for the
    collective and trimmed afterwards)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return pred
    size = dist.get_world_size(group)
    if counts is None:
        n = torch.tensor([pred.shape[0]], device=pred.device, dtype=torch.int64)
        all_n = [torch.zeros_like(n) for _ in range(size)]
        dist.all_gather(all_n, n, group=group)
        counts = [int(v.item()) for v in all_n]
    cap = max(counts)
    padded = pred
    if pred.shape[0] < cap:
        padded = torch.cat([pred, pred.new_zeros(cap - pred.shape[0], pred.shape[1])])
    bufs = [torch.empty_like(padded) for _ in range(size)]
    dist.all_gather(bufs, padded.contiguous(), group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])


def broadcast_state(buffers: Sequence[torch.Tensor], group=None, src: int = 0) -> None:
    """Every rank adopts rank `src`'s copy of each buffer, in place - what DistributedDataParallel's constructor does for the
    module's parameters and buffers (train_JPDVT.py:231; the ranks seed differently at :115-116, so without it the replicas
    start from different weights)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    root = dist.get_global_rank(group, src) if group is not None else src
    for b in buffers:
        dist.broadcast(b, src=root, group=group)


def sum_gradients(flat: torch.Tensor, group=None, async_op: bool = False):
    """SUM all-reduce of a flat gradient buffer over the data-parallel group; returns (work or None, scale) where
    `scale` = 1 / world_size is what turns the sum into DDP's average (train_JPDVT.py:231,370) - the B200 trainer folds it
    into the optimizer kernel instead of spending a pass on it."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None, 1.0
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return (work if async_op else None), 1.0 / dist.get_world_size(group)


def reduce_stats(puzzle_correct: float, piece_correct: float, count: float, seconds: float, device, group=None):
    """inference_ddp.py:485-495: SUM of the three counters, MAX of the elapsed time."""
    stats = torch.tensor([puzzle_correct, piece_correct, count], device=device, dtype=torch.float32)
    tmax = torch.tensor([seconds], device=device, dtype=torch.float32)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX, group=group)
    return stats.tolist(), float(tmax.item())

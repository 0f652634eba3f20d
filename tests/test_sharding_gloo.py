"""world_size-2 gloo test of the multi-GPU sampling plumbing: shard, gather placements, reduce stats."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from jpdvt_mt_ntnu_b200 import parallel
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        begin, end = parallel.block_shard(total, rank, world)
        # each rank "solves" its shard: placement row b is the puzzle id repeated (checkable after the gather)
        pred = torch.arange(begin, end, dtype=torch.int32)[:, None].repeat(1, 9)
        allp = parallel.gather_placements(pred)
        assert allp.shape == (total, 9)
        assert torch.equal(allp[:, 0], torch.arange(total, dtype=torch.int32))
        stats, tmax = parallel.reduce_stats(float(end - begin), 9.0 * (end - begin), float(end - begin), 1.0 + rank, torch.device("cpu"))
        assert stats == [float(total), 9.0 * total, float(total)] and tmax == float(world)
        # strided shard (the reference's own partition) covers every item exactly once
        mine = torch.tensor(parallel.strided_shard(list(range(total)), rank, world), dtype=torch.int64)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([mine.numel()]))
        assert sum(int(s) for s in sizes) == total
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 11), nprocs=2, join=True)

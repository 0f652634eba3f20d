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


def _train_worker(rank, world, port):
    """Host logic of the data-parallel trainer (jpdvt_mt_ntnu_b200/trainer.py) on gloo: replicas seeded differently agree
    after broadcast_state; sum_gradients + the 1/world scale equals the gradient of the concatenated batch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from jpdvt_mt_ntnu_b200 import parallel
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                      # train_JPDVT.py:115-116: a different seed on every rank
        p, ema, m = torch.randn(1000), torch.randn(1000), torch.randn(1000)
        steps = torch.tensor([7 * (rank + 1)])
        parallel.broadcast_state((p, ema, m, steps))
        torch.manual_seed(100)
        want = torch.randn(1000)
        assert torch.equal(p, want) and int(steps) == 7
        # a linear model y = w.x: mean-loss gradient over the whole batch == average of the two half-batch gradients
        g = torch.Generator().manual_seed(5)
        x, y, w = torch.randn(8, 16, generator=g), torch.randn(8, generator=g), torch.randn(16, generator=g)
        full = (2 * (x @ w - y)[:, None] * x).mean(0)
        half = slice(rank * 4, rank * 4 + 4)
        mine = (2 * (x[half] @ w - y[half])[:, None] * x[half]).mean(0)
        work, scale = parallel.sum_gradients(mine)
        assert work is None and scale == 0.5
        assert torch.allclose(mine * scale, full, atol=1e-6)
        work, scale = parallel.sum_gradients(torch.ones(4), async_op=True)
        work.wait()
    finally:
        dist.destroy_process_group()


def test_two_rank_trainer_host_logic():
    port = _free_port()
    mp.spawn(_train_worker, args=(2, port), nprocs=2, join=True)


def test_two_rank_sharding_and_gather():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 11), nprocs=2, join=True)

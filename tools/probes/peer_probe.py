#!/usr/bin/env python3
"""Probe (run under torchrun, >= 2 GPUs): which ways of mapping a peer GPU's buffer work on this box?
 (a) CUDA IPC through torch's storage sharing (cudaIpcGetMemHandle / cudaIpcOpenMemHandle underneath)
 (b) torch.distributed._symmetric_memory (CUDA VMM + fabric / fd handles), incl. the NVSwitch multicast pointer."""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
dev = torch.device("cuda", torch.cuda.current_device())
n = 64 << 20

# (a) IPC
try:
    buf = torch.full((n,), float(rank + 1), device=dev)
    h = buf.untyped_storage()._share_cuda_()
    handles = [None] * world
    dist.all_gather_object(handles, h)
    peers = []
    for r, hh in enumerate(handles):
        if r == rank:
            peers.append(buf)
        else:
            st = torch.UntypedStorage._new_shared_cuda(*hh)
            peers.append(torch.tensor([], dtype=torch.float32, device=dev).set_(st, 0, (n,)))
    torch.cuda.synchronize(); dist.barrier()
    s = [float(p[:1024].sum()) for p in peers]
    out = torch.empty(n, device=dev)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        out.copy_(peers[(rank + 1) % world])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"[rank {rank}] IPC ok: peer sums {s}; peer->local copy {n * 4 / dt / 1e9:.0f} GB/s", flush=True)
    acc = torch.zeros(n, device=dev)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        torch.add(peers[(rank + 1) % world], peers[rank], out=acc)       # SM loads from peer memory
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"[rank {rank}] SM read of peer memory: {n * 4 / dt / 1e9:.0f} GB/s (one peer)", flush=True)
    dist.barrier()
except Exception as e:  # noqa: BLE001
    print(f"[rank {rank}] IPC FAILED: {type(e).__name__}: {e}", flush=True)

# (b) symmetric memory
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(n, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    t.fill_(rank + 1)
    hdl.barrier()
    mc = getattr(hdl, "multicast_ptr", None)
    print(f"[rank {rank}] symm ok: buffer_ptrs {[hex(p) for p in hdl.buffer_ptrs][:4]} signal_pads {len(hdl.signal_pad_ptrs)} "
          f"multicast_ptr {hex(mc) if mc else mc}", flush=True)
    pt = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
    print(f"[rank {rank}] symm peer value {float(pt[0])}", flush=True)
    hdl.barrier()
except Exception as e:  # noqa: BLE001
    print(f"[rank {rank}] symm FAILED: {type(e).__name__}: {e}", flush=True)
dist.destroy_process_group()

#!/usr/bin/env python3
"""Time the data-parallel optimizer step alone, under torchrun (>= 2 GPUs): the fused peer-memory kernel
(`jpdvt_adamw_ema_peer`, multimem and plain peer variants) against NCCL all-reduce + the full `jpdvt_adamw_ema` pass, on the
JPDVT parameter count (130.7 M).  CUDA events on the launching stream, max over ranks."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jpdvt_mt_ntnu_b200 import _lib, peer                      # noqa: E402
from jpdvt_mt_ntnu_b200._lib import check, ptr                 # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
total = int(os.environ.get("PARAMS", "130747208"))
reps = int(os.environ.get("REPS", "20"))
st = lambda: torch.cuda.current_stream(dev).cuda_stream


def timed(fn, label, nbytes_link):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{label:58s} {ms.item():7.3f} ms   ({nbytes_link / ms.item() / 1e6:6.0f} GB/s per GPU over the links)", flush=True)


step = [0]
for mc in ((True, False) if os.environ.get("JPDVT_PEER_VARIANT", "")[:1] != "t" else (False,)):
    px = peer.PeerExchange(total, dev, None, multicast=mc)
    px.grads.normal_()
    px.p.normal_()
    px.ema.copy_(px.p)
    torch.cuda.synchronize()
    dist.barrier()

    def fused():
        step[0] += 1
        check(lib.jpdvt_adamw_ema_peer(C.byref(px.next_epoch()), ptr(px.p), ptr(px.m), ptr(px.v), ptr(px.ema), step[0], 1.0 / world,
                                       1e-4, 0.9, 0.999, 1e-8, 0.0, 0.9999, st()), "peer")
    link = (world - 1) / world * total * (4 + 2)             # gradients in + bf16 operands out, per GPU
    how = "multimem" if px.multicast else ("peer ld/st per thread" if os.environ.get("JPDVT_PEER_VARIANT", "")[:1] == "t" else "bulk async copies")
    timed(fused, f"fused peer step, {how} ({world} GPUs)", link)
    px.check()
    del px

g = torch.randn(total, device=dev)
p, m, v = torch.randn(total, device=dev), torch.zeros(total, device=dev), torch.zeros(total, device=dev)
ema, pb = p.clone(), torch.empty(total, device=dev, dtype=torch.bfloat16)


def nccl():
    step[0] += 1
    dist.all_reduce(g)
    check(lib.jpdvt_adamw_ema(ptr(p), ptr(g), ptr(m), ptr(v), ptr(ema), ptr(pb), total, step[0], 1.0 / world, 1e-4, 0.9, 0.999, 1e-8,
                              0.0, 0.9999, st()), "adamw")


timed(nccl, f"NCCL all-reduce + full AdamW/EMA pass ({world} GPUs)", 2 * (world - 1) / world * total * 4)
timed(lambda: dist.all_reduce(g), "  of which the NCCL all-reduce", 2 * (world - 1) / world * total * 4)
dist.barrier()
dist.destroy_process_group()

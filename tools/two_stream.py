#!/usr/bin/env python3
"""Experiment: does splitting the C2 batch into NSTREAMS independent shards on as many CUDA streams
let the memory-bound launches of one shard (LayerNorm, proj) overlap the tensor-bound launches of another?  Each shard has
its own engine workspace, shares the packed weights, and is driven by its own host thread (the C loop enqueues a whole
sample_loop call per shard).  Prints ms per diffusion step of the whole batch for 1 stream and for NSTREAMS streams."""
import os
import sys
import threading

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jpdvt_mt_ntnu_b200.diffusion import create_diffusion      # noqa: E402
from jpdvt_mt_ntnu_b200.engine import DenoiserEngine           # noqa: E402
from jpdvt_mt_ntnu_b200.models import DiT_models               # noqa: E402
from jpdvt_mt_ntnu_b200.weights import seeded_state            # noqa: E402

size, batch = int(os.environ.get("SIZE", "192")), int(os.environ.get("BATCH", "256"))
steps, reps, ns = int(os.environ.get("STEPS", "125")), int(os.environ.get("REPS", "3")), int(os.environ.get("NSTREAMS", "2"))
T = (size // 16) ** 2
model = DiT_models["JPDVT"](input_size=size)
model.load_state_dict(seeded_state(model.state_dict()))
model.cuda()
d = create_diffusion("250")
g = torch.Generator().manual_seed(0)
cond = (torch.rand(batch, 3, size, size, generator=g) * 2 - 1).cuda()
noise = torch.randn(1, T, 8, generator=g).repeat(batch, 1, 1).cuda()
step_noise = torch.randn(1, batch, T, 8, device="cuda")
eng = model.engine()
tabs = d.device_tables(cond.device)
dev = cond.device


def timed(fn):
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / steps)
    return out


with torch.no_grad():
    eng.sample_loop(tabs, cond, noise, step_noise, first_step=0, last_step=10)
    one = timed(lambda: eng.sample_loop(tabs, cond, noise, step_noise, first_step=0, last_step=steps))
    print(f"1 stream  x batch {batch}: ms/step", " ".join(f"{v:.3f}" for v in one), f"-> {batch / (min(one) * 0.25):.1f} puzzles/s")

    per = batch // ns
    engines, streams = [], [torch.cuda.Stream(device=dev) for _ in range(ns)]
    for i in range(ns):
        e = DenoiserEngine(model.depth, size, dev)
        e.weights = eng.weights
        engines.append(e)
    shards = [(cond[i * per:(i + 1) * per].contiguous(), noise[i * per:(i + 1) * per].contiguous(),
               step_noise[:, i * per:(i + 1) * per].contiguous()) for i in range(ns)]

    def shard(i, last):
        torch.cuda.set_device(dev)
        with torch.no_grad(), torch.cuda.stream(streams[i]):
            engines[i].sample_loop(tabs, *shards[i], first_step=0, last_step=last)

    def fan_out(last):
        cur = torch.cuda.current_stream(dev)
        for s in streams:
            s.wait_stream(cur)
        th = [threading.Thread(target=shard, args=(i, last)) for i in range(ns)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for s in streams:
            cur.wait_stream(s)

    fan_out(10)
    many = timed(lambda: fan_out(steps))
    print(f"{ns} streams x batch {per}: ms/step", " ".join(f"{v:.3f}" for v in many), f"-> {batch / (min(many) * 0.25):.1f} puzzles/s")

#!/usr/bin/env python3
"""Where does the end-to-end training loop lose time against the device-resident one?  Variants of the host side."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from jpdvt_mt_ntnu_b200.diffusion import create_diffusion
from jpdvt_mt_ntnu_b200.models import DiT_models, get_2d_sincos_pos_embed
from jpdvt_mt_ntnu_b200.trainer import Trainer, BatchPrefetcher, LossLog
from jpdvt_mt_ntnu_b200.weights import seeded_state
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
S, G, batch, steps = 192, 3, 128, 20
model = DiT_models["JPDVT"](input_size=S); model.load_state_dict(seeded_state(model.state_dict(), seed=1234)); model.to(dev)
d = create_diffusion(""); tr = Trainer(model, d)
x_pin = (torch.rand(batch, 3, S, S) * 2 - 1).pin_memory(); x = x_pin.to(dev)
piece = torch.tensor(get_2d_sincos_pos_embed(8, G)).unsqueeze(0).float().to(dev)
kw = dict(block_size=S // G, patch_size=16, add_mask=False, grid_size=G)
def one(xin):
    t = torch.randint(0, d.num_timesteps, (batch,), device=dev)
    return tr.step(xin, t, piece, **kw)
def timed(fn, label):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); fn(); host = time.perf_counter() - t0; e1.record(); torch.cuda.synchronize()
    print(f"{label:46s} {e0.elapsed_time(e1) / steps:7.3f} ms/step   host enqueue {host * 1e3 / steps:6.3f} ms/step", flush=True)
for _ in range(5): one(x)
log = LossLog()
for rep in range(2):
    timed(lambda: [one(x) for _ in range(steps)], "device-resident input")
    timed(lambda: [log.push(one(x)) for _ in range(steps)], "  + LossLog.push")
    timed(lambda: [one(x).cpu() for _ in range(steps)], "  + loss.cpu() every step")
    timed(lambda: [one(x_pin.to(dev, non_blocking=True)) for _ in range(steps)], "H2D on the compute stream")
    timed(lambda: [one(xin) for xin in BatchPrefetcher((x_pin for _ in range(steps)), dev)], "BatchPrefetcher")
    timed(lambda: [log.push(one(xin)) for xin in BatchPrefetcher((x_pin for _ in range(steps)), dev)], "BatchPrefetcher + LossLog")

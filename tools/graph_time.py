#!/usr/bin/env python3
"""250-step sampling latency at small batch: plain launches vs CUDA-graph replay (BATCH from the environment)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jpdvt_mt_ntnu_b200.diffusion import create_diffusion
from jpdvt_mt_ntnu_b200.models import DiT_models
from jpdvt_mt_ntnu_b200.weights import seeded_state
B = int(os.environ.get("BATCH", "1"))
model = DiT_models["JPDVT"](input_size=192); model.load_state_dict(seeded_state(model.state_dict())); model.cuda()
d = create_diffusion("250"); T = 144
g = torch.Generator().manual_seed(0)
cond = (torch.rand(B, 3, 192, 192, generator=g) * 2 - 1).cuda(); noise = torch.randn(1, T, 8, generator=g).repeat(B, 1, 1).cuda()
sn = torch.randn(250, B, T, 8, device="cuda"); eng = model.engine(); tabs = d.device_tables(cond.device)
def timed(fn, n=3):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): out = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3, out
with torch.no_grad():
    a, ra = timed(lambda: eng.sample_loop(tabs, cond, noise, sn))
    b, rb = timed(lambda: eng.sample_loop_graphed(tabs, cond, noise, sn))
print(f"B={B}: plain {a:.1f} ms / loop, graph {b:.1f} ms / loop, identical={torch.equal(ra['sample'], rb['sample'])}")

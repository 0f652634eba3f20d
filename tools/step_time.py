#!/usr/bin/env python3
"""ms per diffusion step of the C2 workload (JPDVT 3x3 @192, batch 256) through jpdvt_sample_loop, CUDA events,
sustained (a few hundred steps so the power cap has settled).  STEPS / REPS / BATCH / SIZE from the environment."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jpdvt_mt_ntnu_b200.diffusion import create_diffusion      # noqa: E402
from jpdvt_mt_ntnu_b200.models import DiT_models               # noqa: E402
from jpdvt_mt_ntnu_b200.weights import seeded_state            # noqa: E402

size, batch = int(os.environ.get("SIZE", "192")), int(os.environ.get("BATCH", "256"))
steps, reps = int(os.environ.get("STEPS", "125")), int(os.environ.get("REPS", "4"))
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
out = []
with torch.no_grad():
    eng.sample_loop(tabs, cond, noise, step_noise, first_step=0, last_step=20)
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        eng.sample_loop(tabs, cond, noise, step_noise, first_step=0, last_step=steps)
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / steps)
print("ms/step:", " ".join(f"{v:.3f}" for v in out), f"  -> {batch / (min(out[1:] or out) * 250e-3):.1f} puzzles/s")

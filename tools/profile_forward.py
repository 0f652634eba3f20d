#!/usr/bin/env python3
"""A few denoiser forwards (+ posterior updates) at the bench shape, for ncu captures: 3 diffusion steps of the C2
workload (JPDVT 3x3 @192, batch 256) through jpdvt_sample_loop.  92 kernel launches per step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jpdvt_mt_ntnu_b200.diffusion import create_diffusion      # noqa: E402
from jpdvt_mt_ntnu_b200.models import DiT_models               # noqa: E402
from jpdvt_mt_ntnu_b200.weights import seeded_state            # noqa: E402

size = int(os.environ.get("SIZE", "192"))
batch = int(os.environ.get("BATCH", "256"))
steps = int(os.environ.get("STEPS", "3"))
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
with torch.no_grad():
    eng.sample_loop(d.device_tables(cond.device), cond, noise, step_noise, first_step=0, last_step=steps)
torch.cuda.synchronize()
print("ok")

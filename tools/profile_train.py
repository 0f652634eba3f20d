#!/usr/bin/env python3
"""A few training steps of the C3 workload (JPDVT 3x3 @192, batch 128) through Trainer.step, for ncu captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jpdvt_mt_ntnu_b200.diffusion import create_diffusion                    # noqa: E402
from jpdvt_mt_ntnu_b200.models import DiT_models, get_2d_sincos_pos_embed    # noqa: E402
from jpdvt_mt_ntnu_b200.trainer import Trainer                               # noqa: E402
from jpdvt_mt_ntnu_b200.weights import seeded_state                          # noqa: E402

size, batch, steps = int(os.environ.get("SIZE", "192")), int(os.environ.get("BATCH", "128")), int(os.environ.get("STEPS", "2"))
grid = int(os.environ.get("GRID", "3"))
model = DiT_models["JPDVT"](input_size=size)
model.load_state_dict(seeded_state(model.state_dict()))
model.cuda()
d = create_diffusion("")
tr = Trainer(model, d, lr=1e-4, weight_decay=0.0, ema_decay=0.9999)
g = torch.Generator().manual_seed(0)
x = (torch.rand(batch, 3, size, size, generator=g) * 2 - 1).cuda()
piece = torch.tensor(get_2d_sincos_pos_embed(8, grid)).unsqueeze(0).float().cuda()
torch.manual_seed(0)
for _ in range(steps):
    t = torch.randint(0, d.num_timesteps, (batch,), device="cuda")
    loss = tr.step(x, t, piece, block_size=size // grid, patch_size=16, add_mask=False, grid_size=grid)
torch.cuda.synchronize()
print("ok", float(loss))

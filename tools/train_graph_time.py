#!/usr/bin/env python3
"""ms per training step (BASELINE configs[2]: JPDVT 3x3 @192, batch 128 by default) through Trainer.step: host-issued launches
vs the whole step replayed from a CUDA graph.  BATCH / SIZE / GRID / STEPS / MASK from the environment."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jpdvt_mt_ntnu_b200.diffusion import create_diffusion                        # noqa: E402
from jpdvt_mt_ntnu_b200.models import DiT_models, get_2d_sincos_pos_embed        # noqa: E402
from jpdvt_mt_ntnu_b200.trainer import Trainer                                   # noqa: E402
from jpdvt_mt_ntnu_b200.weights import seeded_state                              # noqa: E402

S, G, B = int(os.environ.get("SIZE", "192")), int(os.environ.get("GRID", "3")), int(os.environ.get("BATCH", "128"))
steps, mask = int(os.environ.get("STEPS", "40")), os.environ.get("MASK", "0") == "1"
dev = torch.device("cuda", 0)
model = DiT_models["JPDVT"](input_size=S)
model.load_state_dict(seeded_state(model.state_dict(), seed=1234))
d = create_diffusion("")
tr = Trainer(model.to(dev), d, lr=1e-4)
x = (torch.rand(B, 3, S, S) * 2 - 1).to(dev)
piece = torch.tensor(get_2d_sincos_pos_embed(8, G)).unsqueeze(0).float().to(dev)
kw = dict(block_size=S // G, patch_size=16, add_mask=mask, grid_size=G)


def run(graph):
    def one():
        t = torch.randint(0, d.num_timesteps, (B,), device=dev)
        return tr.step(x, t, piece, graph=graph, **kw)
    for _ in range(5):
        one()
    out = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            loss = one()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / steps)
    return out, float(loss)


for graph in (False, True, False, True):
    ms, loss = run(graph)
    print(f"graph={graph}: ms/step " + " ".join(f"{v:.3f}" for v in ms) + f"  -> {B / (min(ms) * 1e-3):.0f} img/s   loss {loss:.4f}")

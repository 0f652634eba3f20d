#!/usr/bin/env python3
"""cProfile of the host side of Trainer.step (C3 shape): where do the ~15 ms of enqueue time per step go?"""
import cProfile, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from jpdvt_mt_ntnu_b200.diffusion import create_diffusion
from jpdvt_mt_ntnu_b200.models import DiT_models, get_2d_sincos_pos_embed
from jpdvt_mt_ntnu_b200.trainer import Trainer
from jpdvt_mt_ntnu_b200.weights import seeded_state
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
S, G, batch, steps = 192, 3, int(os.environ.get("BATCH", "128")), 20
model = DiT_models["JPDVT"](input_size=S); model.load_state_dict(seeded_state(model.state_dict(), seed=1234)); model.to(dev)
d = create_diffusion(""); tr = Trainer(model, d)
x = (torch.rand(batch, 3, S, S) * 2 - 1).to(dev)
piece = torch.tensor(get_2d_sincos_pos_embed(8, G)).unsqueeze(0).float().to(dev)
kw = dict(block_size=S // G, patch_size=16, add_mask=False, grid_size=G)
def one():
    t = torch.randint(0, d.num_timesteps, (batch,), device=dev)
    return tr.step(x, t, piece, **kw)
for _ in range(5): one()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
for _ in range(steps): one()
host = time.perf_counter() - t0
pr.disable(); torch.cuda.synchronize()
print(f"host enqueue {host * 1e3 / steps:.3f} ms/step")
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(45)

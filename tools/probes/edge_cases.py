#!/usr/bin/env python3
"""Edge cases end to end: empty and single-puzzle batches through p_sample_loop / solve_puzzles / training_losses."""
import os, sys, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from jpdvt_mt_ntnu_b200 import assignment
from jpdvt_mt_ntnu_b200.diffusion import create_diffusion
from jpdvt_mt_ntnu_b200.models import DiT, get_2d_sincos_pos_embed
from jpdvt_mt_ntnu_b200.weights import seeded_state

size, grid = 192, 3
T = (size // 16) ** 2
model = DiT(input_size=size, depth=2, hidden_size=768, patch_size=16, num_heads=12)
model.load_state_dict(seeded_state(model.state_dict(), seed=7)); model.cuda()
d = create_diffusion("5")
g = torch.Generator().manual_seed(0)
full = (torch.rand(5, 3, size, size, generator=g) * 2 - 1).cuda()
noise1 = torch.randn(1, T, 8, generator=g).cuda()
with torch.no_grad():
    ref = d.p_sample_loop(model.forward, full, (5, T, 8), noise1.repeat(5, 1, 1), clip_denoised=False)
for B in (0, 1, 2, 5):
    for graph in (False, True):
        try:
            with torch.no_grad():
                out = d.p_sample_loop(model.forward, full[:B], (B, T, 8), noise1.repeat(B, 1, 1), clip_denoised=False, graph=graph)
                order, pred = assignment.solve_puzzles(out, grid)
            torch.cuda.synchronize()
            err = float((out - ref[:B]).abs().max()) if B else 0.0
            print(f"sampling B={B} graph={graph}: ok, shape {tuple(out.shape)}, pred {tuple(pred.shape)}, max |diff| vs rows of the batch-5 run {err:.2e}")
        except Exception as e:  # noqa: BLE001
            print(f"sampling B={B} graph={graph}: {type(e).__name__}: {e}")
piece = torch.tensor(get_2d_sincos_pos_embed(8, grid)).unsqueeze(0).float().cuda()
td = create_diffusion("")
for B in (0, 1):
    for mask in (False, True):
        try:
            t = torch.randint(0, 1000, (B,), device="cuda")
            terms = td.training_losses(model, full[:B], t, piece, None, block_size=size // grid, patch_size=16, add_mask=mask, grid_size=grid)
            if B:
                terms["loss"].mean().backward()
            torch.cuda.synchronize()
            print(f"training B={B} mask={mask}: ok, loss {terms['loss'].tolist()}")
        except Exception as e:  # noqa: BLE001
            print(f"training B={B} mask={mask}: {type(e).__name__}: {e}")

"""Seeded parity cases shared by oracle/make_golden.py and tests/.  TEST INFRASTRUCTURE ONLY.

Inputs are never stored: they are regenerated from these seeds with the CPU generators
(deterministic across machines), so the committed fixtures hold reference OUTPUTS only.
Input recipes follow SURVEY.md 8(d): images rand*2-1, one randn(1,T,8) noise row repeated over the
batch (inferencetexmet.py:313), np.random permutations, weights randn*0.02 on every tensor but pos_embed.
"""
import random

import numpy as np
import torch

from . import jpdvt_oracle as orc

TAP_TOKEN_STRIDE = 5
TAP_CHANNEL_STRIDE = 37

FORWARD_CASES = {
    "tiny48":  dict(size=48,  depth=2,  batch=3, wseed=11,   seed=101),
    "d2_192":  dict(size=192, depth=2,  batch=2, wseed=12,   seed=102),
    "d2_256":  dict(size=256, depth=2,  batch=2, wseed=13,   seed=103),
    "d2_288":  dict(size=288, depth=2,  batch=2, wseed=14,   seed=104),
    "full192": dict(size=192, depth=12, batch=2, wseed=1234, seed=105),
    # std 0.04 makes gates/modulations O(0.1..1) so every block contributes visibly to the output
    "hot192":  dict(size=192, depth=4,  batch=2, wseed=15,   seed=106, wstd=0.04),
}

SAMPLING_CASES = {
    "tiny48_s10":   dict(size=48,  depth=2,  batch=3, grid=3, wseed=21,   seed=201, respacing="10",  loop_seed=5),
    "d2_192_s250":  dict(size=192, depth=2,  batch=2, grid=3, wseed=22,   seed=202, respacing="250", loop_seed=6),
    "d2_256g4_s25": dict(size=256, depth=2,  batch=2, grid=4, wseed=23,   seed=203, respacing="25",  loop_seed=7),
    "full192_s250": dict(size=192, depth=12, batch=2, grid=3, wseed=1234, seed=204, respacing="250", loop_seed=8),
    # BASELINE configs[3] (C4): 4x4 @256 px, all 12 blocks, all 250 steps
    "c4_256g4_s250": dict(size=256, depth=12, batch=2, grid=4, wseed=1234, seed=205, respacing="250", loop_seed=12),
    # BASELINE configs[4] (C5): 3x3 @288 px with 1-2 missing pieces per puzzle = slots of the scrambled condition image set to
    # zero (inference_visualize_missing_patches.ipynb cell 9), all 12 blocks, all 250 steps
    "c5_288_miss_s250": dict(size=288, depth=12, batch=2, grid=3, wseed=1234, seed=206, respacing="250", loop_seed=13, missing=(1, 2)),
}

# DDIM: the reference's ddim_sample calls p_mean_variance without `condition` (gaussian_diffusion.py:546-553, TypeError);
# the goldens run the reference's OWN ddim_sample / ddim_sample_loop_progressive code (lines 531-698, unmodified) with
# the missing argument supplied at that single call site (oracle/make_golden.py: golden_ddim)
DDIM_CASES = {
    "tiny48_ddim10_eta0":  dict(size=48,  depth=2, batch=3, grid=3, wseed=21, seed=201, respacing="10",     loop_seed=9,  eta=0.0),
    "tiny48_ddim10_eta05": dict(size=48,  depth=2, batch=3, grid=3, wseed=21, seed=201, respacing="10",     loop_seed=10, eta=0.5),
    "d2_192_ddim25_eta1":  dict(size=192, depth=2, batch=2, grid=3, wseed=22, seed=202, respacing="ddim25", loop_seed=11, eta=1.0),
}

TRAINING_CASES = {
    "tiny96":      dict(size=96,  depth=2, batch=3, grid=3, wseed=31, seed=301, add_mask=False),
    "tiny96_mask": dict(size=96,  depth=2, batch=3, grid=3, wseed=32, seed=302, add_mask=True),
    "g4_128_mask": dict(size=128, depth=2, batch=2, grid=4, wseed=33, seed=303, add_mask=True),
    "d2_192":      dict(size=192, depth=2, batch=2, grid=3, wseed=34, seed=304, add_mask=False),
    # 4x4 @256 px (T = 256, BASELINE configs[3]): the tiled tcgen05 attention backward end to end
    "g4_256":      dict(size=256, depth=2, batch=2, grid=4, wseed=35, seed=305, add_mask=False),
    # 3x3 @288 px (T = 324, the reference trainer's default --image-size) with masked pieces: the key-block-item backward
    "d2_288_mask": dict(size=288, depth=2, batch=2, grid=3, wseed=36, seed=306, add_mask=True),
}

GRAD_KEYS = [
    "time_emb_out2.weight", "time_emb_out1.bias", "time_emb_in.weight", "x_embedder.proj.weight",
    "t_embedder.mlp.0.weight", "t_embedder.mlp.2.bias",
    "blocks.0.attn.qkv.weight", "blocks.0.attn.proj.bias", "blocks.0.adaLN_modulation.1.weight",
    "blocks.1.mlp.fc1.weight", "blocks.1.mlp.fc2.bias", "blocks.1.adaLN_modulation.1.bias",
    "final_layer.linear.weight", "final_layer.adaLN_modulation.1.weight",
]


def tokens(case):
    return (case["size"] // 16) ** 2


def state_for(case):
    """Synthetic weights of a case (regenerated, never stored)."""
    return orc.seeded_state(orc.blank_state(case["size"], case["depth"]), seed=case["wseed"], std=case.get("wstd", 0.02))


def forward_inputs(case):
    g = torch.Generator().manual_seed(case["seed"])
    B, S, T = case["batch"], case["size"], tokens(case)
    img = torch.rand(B, 3, S, S, generator=g) * 2 - 1
    t = torch.randint(0, 1000, (B,), generator=g)
    x_t = torch.randn(B, T, 8, generator=g)
    return img, t, x_t


def sampling_inputs(case):
    g = torch.Generator().manual_seed(case["seed"])
    B, S, T, G = case["batch"], case["size"], tokens(case), case["grid"]
    img = torch.rand(B, 3, S, S, generator=g) * 2 - 1
    rs = np.random.RandomState(case["seed"])
    perms = [rs.permutation(G * G) for _ in range(B)]
    cond = torch.cat([orc.scramble(img[b:b + 1], perms[b], G) for b in range(B)], 0)
    if case.get("missing"):
        cond = zero_slots(cond, missing_slots(case), G)
    noise = torch.randn(1, T, 8, generator=g).repeat(B, 1, 1)
    return cond, noise


def missing_slots(case):
    """Per puzzle: the slots of the scrambled image that are blanked (r in [lo, hi] of them, seeded)."""
    lo, hi = case["missing"]
    rs = np.random.RandomState(case["seed"] + 7919)
    n = case["grid"] ** 2
    return [sorted(rs.choice(n, size=int(rs.randint(lo, hi + 1)), replace=False).tolist()) for _ in range(case["batch"])]


def zero_slots(cond, slots, grid):
    """Masked-puzzle inference: the selected slots of the (already scrambled) condition image become zeros."""
    cond = cond.clone()
    p = cond.shape[-1] // grid
    for b, ss in enumerate(slots):
        for s in ss:
            r, c = divmod(int(s), grid)
            cond[b, :, r * p:(r + 1) * p, c * p:(c + 1) * p] = 0
    return cond


def sampling_perms(case):
    rs = np.random.RandomState(case["seed"])
    return [rs.permutation(case["grid"] ** 2) for _ in range(case["batch"])]


def kept_steps(n):
    return sorted({0, 1, n // 2, n - 1})


def training_inputs(case):
    g = torch.Generator().manual_seed(case["seed"])
    B, S, G = case["batch"], case["size"], case["grid"]
    x = torch.rand(B, 3, S, S, generator=g) * 2 - 1
    t = torch.randint(0, 1000, (B,), generator=g)
    piece = torch.from_numpy(orc.sincos_2d(8, G)).float().unsqueeze(0)
    return x, t, piece


def seed_training_rngs(case):
    torch.manual_seed(case["seed"]); np.random.seed(case["seed"]); random.seed(case["seed"])


def training_draws(case):
    """Replay gaussian_diffusion.py:751-795's draws in order: randn_like(x) [torch], permutation [numpy],
    (randint [numpy], sample [random]) per image if add_mask, randn_like(te) [torch]."""
    seed_training_rngs(case)
    B, S, G, T = case["batch"], case["size"], case["grid"], tokens(case)
    noise_x = torch.randn(B, 3, S, S)
    perm = np.random.permutation(G * G)
    masks = None
    if case["add_mask"]:
        masks = torch.ones(B, G * G)
        for i in range(B):
            r = np.random.randint(0, G)
            masks[i, random.sample(range(G * G), r)] = 0
    noise_te = torch.randn(B, T, 8)
    return dict(noise_x=noise_x, perm=perm, masks=masks, noise_te=noise_te)

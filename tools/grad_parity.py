#!/usr/bin/env python3
"""Per-tensor gradient parity table: every parameter's gradient from the B200 training path against fp32 autograd through
the CPU oracle on the same draws (the oracle is pinned to the unmodified reference: tests/golden/training_*.npz).
Writes one JSON object per case: name -> {cos, rel_l2, ref_norm, numel}.  Run on the GPU box:

    python tools/grad_parity.py gpurun_out/grad_parity.json [case ...]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch                                                     # noqa: E402

from oracle import cases                                         # noqa: E402
from oracle import jpdvt_oracle as orc                           # noqa: E402


def run(name):
    from jpdvt_mt_ntnu_b200.diffusion import create_diffusion
    from jpdvt_mt_ntnu_b200.models import DiT
    case = cases.TRAINING_CASES[name]
    m = DiT(input_size=case["size"], depth=case["depth"], hidden_size=768, patch_size=16, num_heads=12)
    m.load_state_dict(cases.state_for(case))
    m.cuda()
    d = create_diffusion("")
    x, t, piece = cases.training_inputs(case)
    d._draws = cases.training_draws(case)
    kw = dict(block_size=case["size"] // case["grid"], patch_size=16, add_mask=case["add_mask"], grid_size=case["grid"])
    terms = d.training_losses(m, x.cuda(), t.cuda(), piece.cuda(), None, **kw)
    terms["loss"].mean().backward()
    st = {k: v.clone().requires_grad_(k != "pos_embed") for k, v in cases.state_for(case).items()}
    model = orc.OracleDenoiser.__new__(orc.OracleDenoiser)
    model.w, model.depth, model.heads, model.patch = st, case["depth"], 12, 16
    dr = cases.training_draws(case)
    o = orc.training_losses(orc.Schedule(""), model, x, t, piece, dr["perm"], dr["noise_x"], dr["noise_te"],
                            block_size=kw["block_size"], grid=case["grid"], masks=dr["masks"])
    o["loss"].mean().backward()
    table = {}
    for pname, p in m.named_parameters():
        if not p.requires_grad:
            continue
        ref, got = st[pname].grad.double().flatten(), p.grad.cpu().double().flatten()
        table[pname] = {"cos": torch.nn.functional.cosine_similarity(got, ref, dim=0).item(),
                        "rel_l2": ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item(),
                        "ref_norm": ref.norm().item(), "numel": ref.numel()}
    return {"mse_rel_err": ((terms["mse"].detach().cpu() - o["mse"].detach()).abs() / o["mse"].detach().abs()).max().item(),
            "tensors": table}


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "grad_parity.json")
    names = sys.argv[2:] or list(cases.TRAINING_CASES)
    res = {n: run(n) for n in names}
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(res, open(out, "w"), indent=1)
    for n, r in res.items():
        worst = sorted(r["tensors"].items(), key=lambda kv: kv[1]["cos"])[:6]
        print(n, "mse rel err %.2e" % r["mse_rel_err"], "worst cos:", [(k, round(v["cos"], 5), "%.1e" % v["ref_norm"]) for k, v in worst])

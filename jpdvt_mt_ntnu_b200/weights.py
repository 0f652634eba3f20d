"""Weight utilities for the B200 denoiser: synthetic (benchmark) weights and reference-checkpoint loading.

`seeded_state` implements the synthetic-weight recipe of SURVEY.md 8(d): a freshly initialised JPDVT returns exact
zeros (adaLN-Zero + zero final layer, image_model/models.py:216-225), so benchmarks and parity tests overwrite every
tensor except the fixed sin-cos `pos_embed` with randn(seed) * std, drawn in state-dict order on the CPU generator.
"""
from __future__ import annotations

from typing import Dict

import torch


def seeded_state(state: Dict[str, torch.Tensor], seed: int = 1234, std: float = 0.02) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in state.items():
        out[k] = v.detach().clone() if k == "pos_embed" else torch.randn(tuple(v.shape), generator=g, dtype=torch.float32) * std
    return out


def load_reference_checkpoint(model, path: str, use_ema: bool = False, map_location="cpu") -> dict:
    """Load a checkpoint written by the reference trainer (`{"model","ema","opt","args","train_steps"}`,
    image_model/train_JPDVT.py:410-416) the way its inference scripts do (image_model/inference.py:207-211):
    keys that exist in the model are loaded, strict=False.  Returns a report of what matched."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    src = ckpt["ema" if use_ema else "model"] if isinstance(ckpt, dict) and "model" in ckpt else ckpt
    own = model.state_dict()
    usable = {k: v for k, v in src.items() if k in own and tuple(v.shape) == tuple(own[k].shape)}
    skipped = sorted(k for k in src if k not in usable)
    missing = sorted(k for k in own if k not in usable)
    model.load_state_dict(usable, strict=False)
    return {"loaded": len(usable), "skipped": skipped, "missing": missing,
            "train_steps": ckpt.get("train_steps") if isinstance(ckpt, dict) else None}

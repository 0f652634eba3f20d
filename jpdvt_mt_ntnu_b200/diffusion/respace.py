"""Host-side mirror of image_model/diffusion/respace.py: timestep respacing.

`space_timesteps` picks the retained steps, `SpacedDiffusion` re-derives the betas of the shortened chain and maps a
respaced step index back to the original timestep before the network sees it.  In the reference that mapping is a
per-call `th.tensor(timestep_map)` upload + gather inside `_WrappedModel.__call__` (respace.py:124-129); here the map
is a device-resident int32 table read directly by the timestep-embedding kernel.
"""
import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """respace.py:12-62.  section_counts: "250", "10,15,20", [10, 15, 20] or "ddimN".  Returns a set of kept steps."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            wanted = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                picked = range(0, num_timesteps, stride)
                if len(picked) == wanted:
                    return set(picked)
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(part) for part in section_counts.split(",")]
    per_section, remainder = divmod(num_timesteps, len(section_counts))
    kept, offset = [], 0
    for idx, count in enumerate(section_counts):
        length = per_section + (1 if idx < remainder else 0)
        if length < count:
            raise ValueError(f"cannot divide section of {length} steps into {count}")
        step = 1 if count <= 1 else (length - 1) / (count - 1)
        kept.extend(offset + round(pos) for pos in _arith(count, step))
        offset += length
    return set(kept)


def _arith(count, step):
    """0, step, 2*step, ... accumulated by repeated addition, as the reference does (rounding matters)."""
    pos, out = 0.0, []
    for _ in range(count):
        out.append(pos)
        pos += step
    return out


class SpacedDiffusion(GaussianDiffusion):
    """A diffusion process that only visits `use_timesteps` of a base process (respace.py:65-114)."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(kwargs["betas"])
        base = GaussianDiffusion(**kwargs)
        self.timestep_map = []
        new_betas, prev = [], 1.0
        for i, abar in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - abar / prev)
                prev = abar
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)

    def timestep_map_list(self):
        return list(self.timestep_map)

    def _model_timesteps(self, t):
        return self.device_tables(t.device)["timestep_map64"][t] if t.is_cuda else th.tensor(self.timestep_map, dtype=t.dtype)[t]

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.original_num_steps)

    def _scale_timesteps(self, t):
        return t


class _WrappedModel:
    """respace.py:117-129: callable that maps respaced indices to original timesteps before calling the model.
    Kept for callers that build it explicitly; SpacedDiffusion itself maps timesteps on the device."""

    def __init__(self, model, timestep_map, original_num_steps):
        self.model = model
        self.timestep_map = timestep_map
        self.original_num_steps = original_num_steps
        self._cache = {}

    def __call__(self, x, ts, time_emb, **kwargs):
        key = (str(ts.device), ts.dtype)
        if key not in self._cache:
            self._cache[key] = th.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)
        return self.model(x, self._cache[key][ts], time_emb, **kwargs)

"""Host-side mirror of the reference's `models` module (image_model/models.py) for the B200 path.

Same public surface - `DiT`, `DiT_models`, `JPDVT*` factories, `get_2d_sincos_pos_embed` - same constructor
arguments, same state-dict keys/shapes (so reference checkpoints load, SURVEY.md 5 "Checkpoint / resume"), same
`forward(x, t, time_emb, y=None) -> (image, time_emb_out)`.  The arithmetic is NOT here: forward hands raw pointers
to libjpdvt_sm100.so (hand-written sm_100a kernels).  There is no CPU or eager-PyTorch fallback; calling forward
without a B200 raises.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib

__all__ = ["DiT", "DiT_models", "get_2d_sincos_pos_embed", "get_2d_sincos_pos_embed_from_grid",
           "get_1d_sincos_pos_embed_from_grid", "TimestepEmbedder", "DiTBlock", "FinalLayer", "modulate"]


# --------------------------------------------------------------------------------------------- positional tables
def get_1d_sincos_pos_embed_from_grid(embed_dim: int, pos: np.ndarray) -> np.ndarray:
    """models.py:349-366: [sin | cos] of pos * 10000^(-k/(embed_dim/2)), fp64, shape (M, embed_dim)."""
    if embed_dim % 2:
        raise AssertionError("embed_dim must be even")
    omega = 1.0 / 10000 ** (np.arange(embed_dim // 2, dtype=np.float64) / (embed_dim / 2.0))
    ang = np.einsum("m,d->md", np.asarray(pos).reshape(-1), omega)
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def get_2d_sincos_pos_embed_from_grid(embed_dim: int, grid: np.ndarray) -> np.ndarray:
    """models.py:337-346: first half encodes grid[0], second half grid[1]."""
    if embed_dim % 2:
        raise AssertionError("embed_dim must be even")
    return np.concatenate([get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[0]),
                           get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[1])], axis=1)


def get_2d_sincos_pos_embed(embed_dim: int, grid_size: int, cls_token: bool = False, extra_tokens: int = 0) -> np.ndarray:
    """models.py:319-334: [grid_size**2, embed_dim]; grid[0] is the COLUMN coordinate ("w goes first")."""
    coords = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(coords, coords), axis=0).reshape(2, 1, grid_size, grid_size)
    emb = get_2d_sincos_pos_embed_from_grid(embed_dim, grid)
    if cls_token and extra_tokens > 0:
        emb = np.concatenate([np.zeros([extra_tokens, embed_dim]), emb], axis=0)
    return emb


def modulate(x: torch.Tensor, shift: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """models.py:19-20.  Kept for API completeness; the CUDA path fuses this with LayerNorm (jpdvt_ln_modulate_fwd)."""
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


# --------------------------------------------------------------------------------------------- parameter containers
# The sub-modules below only own parameters under the reference's names; none of their forward()s is on the hot path.

class _PatchEmbed(nn.Module):
    """timm PatchEmbed's parameter layout: `.proj` = Conv2d(k = s = patch) (models.py:169)."""

    def __init__(self, img_size, patch_size, in_chans, embed_dim, bias=True):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)


class TimestepEmbedder(nn.Module):
    """models.py:27-64 parameter layout: mlp.0 = Linear(256, hidden), mlp.2 = Linear(hidden, hidden)."""

    def __init__(self, hidden_size, frequency_embedding_size=256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size, bias=True), nn.SiLU(),
                                 nn.Linear(hidden_size, hidden_size, bias=True))
        self.frequency_embedding_size = frequency_embedding_size


class _Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class DiTBlock(nn.Module):
    """models.py:101-122 parameter layout (norm1/norm2 carry no parameters: elementwise_affine=False)."""

    def __init__(self, hidden_size, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.attn = _Attention(hidden_size, num_heads)
        self.mlp = _Mlp(hidden_size, int(hidden_size * mlp_ratio))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 6 * hidden_size, bias=True))


class FinalLayer(nn.Module):
    """models.py:125-142 parameter layout."""

    def __init__(self, hidden_size, patch_size, out_channels):
        super().__init__()
        self.linear = nn.Linear(hidden_size, patch_size * patch_size * out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 2 * hidden_size, bias=True))


# --------------------------------------------------------------------------------------------- the denoiser
class DiT(nn.Module):
    """JPDVT denoiser (models.py:145-293) running on hand-written sm_100a kernels.

    Supported on the CUDA path: hidden_size 768, 12 heads, patch 16, 3 channels, any depth, input_size % 16 == 0 -
    i.e. every configuration for which the reference's own forward type-checks (SURVEY.md headline 6).
    """

    def __init__(self, input_size=255, patch_size=2, in_channels=3, hidden_size=1152, depth=28, num_heads=16,
                 mlp_ratio=4.0, class_dropout_prob=0.1, num_classes=0, learn_sigma=False):
        super().__init__()
        self.learn_sigma = learn_sigma
        self.in_channels = in_channels
        self.out_channels = in_channels * 2 if learn_sigma else in_channels
        self.patch_size = patch_size
        self.num_heads = num_heads
        self.hidden_size = hidden_size
        self.depth = depth
        self.input_size = input_size
        self.mlp_ratio = mlp_ratio

        self.x_embedder = _PatchEmbed(input_size, patch_size, in_channels, hidden_size, bias=True)
        self.t_embedder = TimestepEmbedder(hidden_size)
        self.pos_embed = nn.Parameter(torch.zeros(1, self.x_embedder.num_patches, hidden_size), requires_grad=False)
        self.time_emb_in = nn.Linear(8, 768)        # hard-wired widths, as in models.py:176-179
        self.time_emb_out1 = nn.Linear(768, 64)
        self.time_emb_out2 = nn.Linear(64, 8)
        self.blocks = nn.ModuleList([DiTBlock(hidden_size, num_heads, mlp_ratio=mlp_ratio) for _ in range(depth)])
        self.final_layer = FinalLayer(hidden_size, patch_size, self.out_channels)
        self.initialize_weights()
        self.__dict__["_engine"] = None
        self.__dict__["_engine_key"] = None

    # -- initialisation scheme of models.py:187-225 ---------------------------------------------------------------
    def initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        grid = int(self.x_embedder.num_patches ** 0.5)
        self.pos_embed.data.copy_(torch.from_numpy(get_2d_sincos_pos_embed(self.pos_embed.shape[-1], grid)).float().unsqueeze(0))
        w = self.x_embedder.proj.weight.data
        nn.init.xavier_uniform_(w.view(w.shape[0], -1))
        nn.init.zeros_(self.x_embedder.proj.bias)
        for lin in (self.t_embedder.mlp[0], self.t_embedder.mlp[2], self.time_emb_in, self.time_emb_out1, self.time_emb_out2):
            nn.init.normal_(lin.weight, std=0.02)
        for blk in self.blocks:                       # adaLN-Zero: a fresh model outputs exact zeros
            nn.init.zeros_(blk.adaLN_modulation[-1].weight)
            nn.init.zeros_(blk.adaLN_modulation[-1].bias)
        nn.init.zeros_(self.final_layer.adaLN_modulation[-1].weight)
        nn.init.zeros_(self.final_layer.adaLN_modulation[-1].bias)
        nn.init.zeros_(self.final_layer.linear.weight)
        nn.init.zeros_(self.final_layer.linear.bias)

    # -- engine management ---------------------------------------------------------------------------------------
    def _check_supported(self):
        if (self.hidden_size, self.num_heads, self.patch_size, self.in_channels, self.learn_sigma) != (768, 12, 16, 3, False) \
                or int(self.hidden_size * self.mlp_ratio) != 3072 or self.input_size % 16 != 0:
            raise NotImplementedError(
                f"DiT(hidden={self.hidden_size}, heads={self.num_heads}, patch={self.patch_size}) is outside the JPDVT hot "
                "path: the committed reference forward only type-checks for hidden 768 / patch 16 (time_emb_in is "
                "Linear(8,768) and time_emb_out1 is applied to the p*p*3-wide final layer, models.py:176-177,287-288)")

    def _weights_key(self):
        # `_epoch` is bumped by the B200 Trainer, whose optimizer kernel updates parameters behind autograd's back
        return (self.__dict__.get("_epoch", 0),) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self, device: Optional[torch.device] = None):
        """The packed-weight engine for the parameters as they are now (re-packed when any parameter changed)."""
        from .engine import DenoiserEngine
        self._check_supported()
        dev = next(self.parameters()).device if device is None else device
        if dev.type != "cuda":
            raise _lib.JpdvtError("JPDVT parameters are on the CPU: move the model to a B200 (`.cuda()`); no CPU path exists")
        eng = self.__dict__.get("_engine")
        key = self._weights_key()
        if eng is None or eng.device != dev:
            eng = DenoiserEngine(self.depth, self.input_size, dev)
            self.__dict__["_engine"], self.__dict__["_engine_key"] = eng, None
        if self.__dict__.get("_engine_key") != key:
            eng.load_state({k: v for k, v in self.state_dict().items()})
            self.__dict__["_engine_key"] = key
        return eng

    def train_engine(self, device: Optional[torch.device] = None):
        """Training twin of `engine()`: bf16 operand copies in both orientations, activation tape, gradient scratch."""
        from .training import TrainEngine
        self._check_supported()
        dev = next(self.parameters()).device if device is None else device
        if dev.type != "cuda":
            raise _lib.JpdvtError("JPDVT parameters are on the CPU: move the model to a B200 (`.cuda()`); no CPU path exists")
        eng = self.__dict__.get("_train_engine")
        if eng is not None and self.__dict__.get("_adopted_by_trainer"):
            return eng          # operand buffers are owned and refreshed by jpdvt_mt_ntnu_b200.trainer.Trainer
        key = self._weights_key()
        if eng is None or eng.device != dev:
            eng = TrainEngine(self.depth, self.input_size, dev)
            self.__dict__["_train_engine"], self.__dict__["_train_engine_key"] = eng, None
        if self.__dict__.get("_train_engine_key") != key:
            eng.load_state({k: v for k, v in self.state_dict().items()})
            self.__dict__["_train_engine_key"] = key
        return eng

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k in ("_engine", "_engine_key", "_train_engine", "_train_engine_key", "_stage_hook", "_adopted_by_trainer") else copy.deepcopy(v, memo)
        return new

    def unpatchify(self, x: torch.Tensor) -> torch.Tensor:
        """models.py:227-240: (N, T, p*p*C) -> (N, C, H, W)."""
        from . import ops
        n, t, _ = x.shape
        side = int(t ** 0.5)
        if side * side != t:
            raise AssertionError("token count is not a square")
        self._check_supported()
        return ops.unpatchify(x.reshape(n * t, -1).float().contiguous(), n, side * 16)

    # -- forward --------------------------------------------------------------------------------------------------
    def forward(self, x, t, time_emb, y=None):
        """models.py:273-293: x (N,3,S,S) image condition, t (N,) timesteps, time_emb (N,T,8) noisy position latents.

        Returns (image head (N,3,S,S), time_emb_out (N,T,8)).
        """
        needs_grad = torch.is_grad_enabled() and (
            any(p.requires_grad for p in self.parameters()) or x.requires_grad or time_emb.requires_grad)
        if needs_grad:
            from .training import denoiser_forward_with_grad
            return denoiser_forward_with_grad(self, x, t, time_emb)
        img, te = self.engine(x.device).forward(x, t, time_emb, need_image=True)
        return img, te

    def forward_latents(self, x, t, time_emb):
        """Sampling fast path: only the position latents (the image head is discarded by p_mean_variance,
        diffusion/gaussian_diffusion.py:281)."""
        return self.engine(x.device).forward(x, t, time_emb, need_image=False)[1]

    def forward_with_cfg(self, x, t, y, cfg_scale):
        raise NotImplementedError("forward_with_cfg is dead code in the reference (models.py:295-311 calls forward with "
                                  "the wrong arity) and is not part of the JPDVT hot path")


# --------------------------------------------------------------------------------------------- config zoo (models.py:373-424)
def _cfg(depth, hidden_size, patch_size, num_heads):
    def make(**kwargs):
        return DiT(depth=depth, hidden_size=hidden_size, patch_size=patch_size, num_heads=num_heads, **kwargs)
    return make


DiT_models = {
    "DiT-XL/2": _cfg(28, 1152, 2, 16), "DiT-XL/4": _cfg(28, 1152, 4, 16), "DiT-XL/8": _cfg(28, 1152, 8, 16),
    "DiT-L/2": _cfg(24, 1024, 2, 16), "DiT-L/4": _cfg(24, 1024, 4, 16), "DiT-L/8": _cfg(24, 1024, 8, 16),
    "DiT-B/2": _cfg(12, 768, 2, 12), "DiT-B/4": _cfg(12, 768, 4, 12), "DiT-B/8": _cfg(12, 768, 8, 12),
    "DiT-S/2": _cfg(12, 384, 2, 6), "DiT-S/4": _cfg(12, 384, 4, 6), "DiT-S/8": _cfg(12, 384, 8, 6),
    "JPDVT": _cfg(12, 768, 16, 12), "JPDVT-S": _cfg(12, 768, 32, 12), "JPDVT-T": _cfg(12, 768, 64, 12),
}
JPDVT = DiT_models["JPDVT"]
JPDVT_S = DiT_models["JPDVT-S"]
JPDVT_T = DiT_models["JPDVT-T"]

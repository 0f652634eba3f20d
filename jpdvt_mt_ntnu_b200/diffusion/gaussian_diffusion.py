"""Host-side mirror of the reference's Gaussian diffusion (image_model/diffusion/gaussian_diffusion.py).

Same public names and call signatures (`GaussianDiffusion`, `q_sample`, `p_mean_variance`, `p_sample`,
`p_sample_loop[_progressive]`, `ddim_sample[_loop]`, `training_losses`, `get_named_beta_schedule`, the enums,
`mean_flat`, `_extract_into_tensor`), so the reference's train/inference scripts run unchanged.  The host part keeps
the fp64 numpy schedule tables exactly as the reference builds them; everything per-element runs in sm_100a kernels
(libjpdvt_sm100.so) with the tables resident on the device, so a sampling step makes no host<->device copies
(the reference makes >= 8 per step: gaussian_diffusion.py:516,926; respace.py:125; models.py:52-54).

Only the branches that are live in the reference are implemented (START_X or EPSILON mean, FIXED_SMALL variance,
MSE loss - SURVEY.md 8a rows 21/27 list the reference's dead or crashing branches).
"""
from __future__ import annotations

import enum
import math
import os
import random
from typing import Dict, Iterator, Optional

import numpy as np
import torch as th

from .. import _lib, ops


def mean_flat(tensor: th.Tensor) -> th.Tensor:
    """Mean over all non-batch dimensions (gaussian_diffusion.py:18-22).  Utility for callers; `training_losses` computes
    its MSE terms (and their gradient) in the fused loss kernels below instead."""
    return tensor.mean(dim=list(range(1, tensor.dim())))


class _MseLossFn(th.autograd.Function):
    """terms["mse"] of training_losses (gaussian_diffusion.py:835-838) as one kernel pair: per-sample
    mean((te_tgt - te_out)^2) [+ mean((x_tgt - x_out)^2 * (1 - masks))], gradient w.r.t. the two model outputs."""

    @staticmethod
    def forward(ctx, te_out, x_out, te_tgt, x_tgt, keep_slots, grid):
        te_out, te_tgt = te_out.detach().float().contiguous(), te_tgt.detach().float().contiguous()
        if x_out is not None:
            x_out, x_tgt = x_out.detach().float().contiguous(), x_tgt.detach().float().contiguous()
            keep_slots = keep_slots.detach().float().contiguous()
        ctx.saved = (te_out, te_tgt, x_out, x_tgt, keep_slots, grid)
        return ops.mse_loss_fwd(te_out, te_tgt, x_out, x_tgt, keep_slots, grid)

    @staticmethod
    def backward(ctx, dloss):
        te_out, te_tgt, x_out, x_tgt, keep_slots, grid = ctx.saved
        d_te, d_img = ops.mse_loss_bwd(dloss.float().contiguous(), te_out, te_tgt, x_out, x_tgt, keep_slots, grid)
        return d_te, d_img, None, None, None, None


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()

    def is_vb(self):
        return self in (LossType.KL, LossType.RESCALED_KL)


# ------------------------------------------------------------------------------------------------- schedules
def get_beta_schedule(beta_schedule, *, beta_start, beta_end, num_diffusion_timesteps):
    """gaussian_diffusion.py:67-97 (fp64)."""
    n = num_diffusion_timesteps
    if beta_schedule == "linear":
        betas = np.linspace(beta_start, beta_end, n, dtype=np.float64)
    elif beta_schedule == "quad":
        betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=np.float64) ** 2
    elif beta_schedule in ("warmup10", "warmup50"):
        frac = 0.1 if beta_schedule == "warmup10" else 0.5
        betas = beta_end * np.ones(n, dtype=np.float64)
        k = int(n * frac)
        betas[:k] = np.linspace(beta_start, beta_end, k, dtype=np.float64)
    elif beta_schedule == "const":
        betas = beta_end * np.ones(n, dtype=np.float64)
    elif beta_schedule == "jsd":
        betas = 1.0 / np.linspace(n, 1, n, dtype=np.float64)
    else:
        raise NotImplementedError(beta_schedule)
    assert betas.shape == (n,)
    return betas


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """gaussian_diffusion.py:126-143."""
    n = num_diffusion_timesteps
    return np.array([min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)])


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """gaussian_diffusion.py:100-123: "linear" is Ho et al. rescaled to the step count."""
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return get_beta_schedule("linear", beta_start=scale * 0.0001, beta_end=scale * 0.02,
                                 num_diffusion_timesteps=num_diffusion_timesteps)
    if schedule_name == "squaredcos_cap_v2":
        return betas_for_alpha_bar(num_diffusion_timesteps, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def _extract_into_tensor(arr, timesteps, broadcast_shape):
    """gaussian_diffusion.py:917-929: fp64 gather, cast to fp32, broadcast.  Utility for callers; the sampling and
    training paths below read device-resident fp32 tables inside the kernels instead."""
    res = th.from_numpy(np.asarray(arr)).to(device=timesteps.device)[timesteps].float()
    while res.dim() < len(broadcast_shape):
        res = res[..., None]
    return res + th.zeros(broadcast_shape, device=timesteps.device)


def _resolve_denoiser(model):
    """Our DiT behind `model`, `model.forward`, DDP / DataParallel wrappers or a _WrappedModel - else None."""
    from ..models import DiT
    seen = 0
    while seen < 4:
        seen += 1
        if isinstance(model, DiT):
            return model
        if hasattr(model, "__self__") and getattr(model, "__name__", "") == "forward":
            model = model.__self__
        elif hasattr(model, "module"):
            model = model.module
        elif hasattr(model, "model") and hasattr(model, "timestep_map"):
            model = model.model
        else:
            return None
    return None


class GaussianDiffusion:
    """Schedule tables + sampling / training entry points (gaussian_diffusion.py:146-843)."""

    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type):
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        betas = np.array(betas, dtype=np.float64)
        if betas.ndim != 1:
            raise AssertionError("betas must be 1-D")
        if not ((betas > 0).all() and (betas <= 1).all()):
            raise AssertionError("betas must lie in (0, 1]")
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        self.alphas_cumprod = ac
        self.alphas_cumprod_prev = np.append(1.0, ac[:-1])
        self.alphas_cumprod_next = np.append(ac[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(ac)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - ac)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - ac)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / ac)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / ac - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - ac)
        self.posterior_log_variance_clipped = (
            np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
            if len(self.posterior_variance) > 1 else np.array([]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - ac)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - ac)
        self._dev_tables: Dict[str, dict] = {}

    # ------------------------------------------------------------------ device-resident tables
    def timestep_map_list(self):
        return list(range(self.num_timesteps))

    def device_tables(self, device) -> dict:
        """fp32 casts of the fp64 tables, uploaded once per device (the reference re-uploads on every gather)."""
        key = str(device)
        tabs = self._dev_tables.get(key)
        if tabs is None:
            f32 = lambda a: th.tensor(np.asarray(a, dtype=np.float64), dtype=th.float64).to(th.float32).to(device)
            n = self.num_timesteps
            tabs = {
                "num_steps": n,
                "coef1": f32(self.posterior_mean_coef1), "coef2": f32(self.posterior_mean_coef2),
                "logvar": f32(self.posterior_log_variance_clipped), "var": f32(self.posterior_variance),
                "sqrt_ac": f32(self.sqrt_alphas_cumprod), "sqrt_1mac": f32(self.sqrt_one_minus_alphas_cumprod),
                "recip": f32(self.sqrt_recip_alphas_cumprod), "recipm1": f32(self.sqrt_recipm1_alphas_cumprod),
                "step_ids": th.arange(n - 1, -1, -1, dtype=th.int32, device=device),
                "timestep_map": th.tensor(self.timestep_map_list(), dtype=th.int32, device=device),
                "timestep_map64": th.tensor(self.timestep_map_list(), dtype=th.int64, device=device),
            }
            self._dev_tables[key] = tabs
        return tabs

    def _ddim_tables(self, device, eta: float) -> dict:
        tabs = self.device_tables(device)
        key = f"ddim_{eta!r}"
        if key not in tabs:
            ab, abp = self.alphas_cumprod, self.alphas_cumprod_prev
            sigma = eta * np.sqrt((1 - abp) / (1 - ab)) * np.sqrt(1 - ab / abp)
            f32 = lambda a: th.tensor(a, dtype=th.float64).to(th.float32).to(device)
            tabs[key] = {"recip": tabs["recip"], "recipm1": tabs["recipm1"], "sqrt_abp": f32(np.sqrt(abp)),
                         "dir": f32(np.sqrt(np.maximum(1 - abp - sigma ** 2, 0.0))), "sigma": f32(sigma)}
        return tabs[key]

    def _model_timesteps(self, t: th.Tensor) -> th.Tensor:
        """Respaced index -> timestep the network sees (respace.py:124-129); identity for the base process."""
        return t

    # ------------------------------------------------------------------ forward process
    def q_mean_variance(self, x_start, t):
        shape = x_start.shape
        return (_extract_into_tensor(self.sqrt_alphas_cumprod, t, shape) * x_start,
                _extract_into_tensor(1.0 - self.alphas_cumprod, t, shape),
                _extract_into_tensor(self.log_one_minus_alphas_cumprod, t, shape))

    def q_sample(self, x_start, t, noise=None, keep_mask=None):
        """sqrt(abar_t) x0 + sqrt(1 - abar_t) noise (gaussian_diffusion.py:217-232); optional masked blend (:800)."""
        if noise is None:
            noise = th.randn_like(x_start)
        if noise.shape != x_start.shape:
            raise AssertionError("noise must have the shape of x_start")
        tabs = self.device_tables(x_start.device)
        return ops.q_sample(x_start.float(), noise.float(), tabs["sqrt_ac"], tabs["sqrt_1mac"], t.to(th.int64), keep_mask)

    def q_posterior_mean_variance(self, x_start, x_t, t):
        """gaussian_diffusion.py:234-254."""
        if x_start.shape != x_t.shape:
            raise AssertionError("x_start and x_t must have the same shape")
        tabs = self.device_tables(x_t.device)
        mean, _ = ops.posterior_step(x_start.float(), x_t.float(), x_t.float(), tabs["coef1"], tabs["coef2"], tabs["logvar"],
                                     t.to(th.int64))
        bshape = [-1] + [1] * (x_t.dim() - 1)
        var = tabs["var"][t].reshape(bshape).expand_as(x_t)
        logvar = tabs["logvar"][t].reshape(bshape).expand_as(x_t)
        return mean, var, logvar

    # ------------------------------------------------------------------ reverse process, one step
    def _call_model(self, model, condition, t, x, model_kwargs):
        """`_, out = model(condition, map[t], x)` (gaussian_diffusion.py:281 via respace.py:124-129); skips the unused
        image head when `model` is the B200 denoiser."""
        ts = self._model_timesteps(t)
        dit = _resolve_denoiser(model)
        if dit is not None and not model_kwargs and not th.is_grad_enabled():
            return dit.forward_latents(condition, ts, x)
        inner = model.model if (hasattr(model, "model") and hasattr(model, "timestep_map")) else model
        out = inner(condition, ts, x, **(model_kwargs or {}))
        return out[1]

    def p_mean_variance(self, model, condition, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None):
        """gaussian_diffusion.py:256-344.  Variance is FIXED_SMALL (the reference forces it at :288)."""
        B = x.shape[0]
        if t.shape != (B,):
            raise AssertionError("t must have one entry per batch element")
        out = self._call_model(model, condition, t, x, model_kwargs)
        extra = None
        if isinstance(out, tuple):
            out, extra = out
        if self.model_mean_type == ModelMeanType.START_X:
            pred = out
        elif self.model_mean_type == ModelMeanType.EPSILON:
            pred = self._predict_xstart_from_eps(x, t, out)
        else:
            raise NotImplementedError(self.model_mean_type)
        if denoised_fn is not None:
            pred = denoised_fn(pred)
        if clip_denoised:
            pred = pred.clamp(-1, 1)
        mean, var, logvar = self.q_posterior_mean_variance(pred, x, t)
        return {"mean": mean, "variance": var, "log_variance": logvar, "pred_xstart": pred, "extra": extra}

    def _predict_xstart_from_eps(self, x_t, t, eps):
        s = x_t.shape
        return (_extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, s) * x_t
                - _extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, s) * eps)

    def _predict_eps_from_xstart(self, x_t, t, pred_xstart):
        s = x_t.shape
        return ((_extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, s) * x_t - pred_xstart)
                / _extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, s))

    def p_sample(self, model, condition, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                 noise=None):
        """gaussian_diffusion.py:388-431.  `noise` (extra keyword) replaces the internal randn_like for parity tests."""
        out = self.p_mean_variance(model, condition, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                   model_kwargs=model_kwargs)
        if noise is None:
            noise = th.randn_like(x)
        if cond_fn is not None:
            raise NotImplementedError("cond_fn guidance is never used by the reference callers (always None)")
        tabs = self.device_tables(x.device)
        _, sample = ops.posterior_step(out["pred_xstart"].float(), x.float(), noise.float(), tabs["coef1"], tabs["coef2"],
                                       tabs["logvar"], t.to(th.int64))
        return {"sample": sample, "pred_xstart": out["pred_xstart"]}

    # ------------------------------------------------------------------ reverse process, whole loop
    def _fast_path(self, model, clip_denoised, denoised_fn, cond_fn, model_kwargs):
        if clip_denoised or denoised_fn is not None or cond_fn is not None or model_kwargs:
            return None
        if self.model_mean_type != ModelMeanType.START_X:
            return None
        return _resolve_denoiser(model)

    def _draw_step_noise(self, noise: th.Tensor) -> Optional[th.Tensor]:
        """The per-step noise of p_sample (gaussian_diffusion.py:424 `th.randn_like(x)`).  Default: None - the posterior
        kernel draws it in place (Philox4x32-10 keyed by a seed taken from torch's default generator, so `torch.manual_seed`
        still makes runs repeatable; no noise buffer, no generator launches).  JPDVT_TORCH_NOISE=1: one torch `normal_()`
        per step in loop order, which consumes the device generator exactly like the reference's loop does."""
        if os.environ.get("JPDVT_TORCH_NOISE", "")[:1] != "1":
            return None
        buf = th.empty((self.num_timesteps,) + tuple(noise.shape), device=noise.device, dtype=th.float32)
        for k in range(self.num_timesteps):
            buf[k].normal_()
        return buf

    def p_sample_loop(self, model, condition, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, progress=False, chain=False, step_noise=None, graph=False):
        """gaussian_diffusion.py:433-478.  Default behaviour reproduces the reference exactly, including its loop quirk
        (every step is fed the initial `noise`, :518-529); `chain=True` (extra keyword) feeds the running sample."""
        dit = self._fast_path(model, clip_denoised, denoised_fn, cond_fn, model_kwargs)
        if dit is not None and not progress:
            dev = condition.device
            if noise is None:
                noise = th.randn(*shape, device=dev)
            if step_noise is None:
                step_noise = self._draw_step_noise(noise)
            with th.no_grad():
                eng = dit.engine(dev)
                # graph=True (extra keyword): replay the whole loop from a CUDA graph - worth it for small, repeated batches
                run = eng.sample_loop_graphed if graph else eng.sample_loop
                state = run(self.device_tables(dev), condition, noise, step_noise, chain=chain)
            return state["sample"]
        final = None
        for final in self.p_sample_loop_progressive(model, condition, shape, noise=noise, clip_denoised=clip_denoised,
                                                    denoised_fn=denoised_fn, cond_fn=cond_fn, model_kwargs=model_kwargs,
                                                    device=device, progress=progress, chain=chain, step_noise=step_noise):
            pass
        return final["sample"]

    def p_sample_loop_progressive(self, model, condition, shape, noise=None, clip_denoised=True, denoised_fn=None,
                                  cond_fn=None, model_kwargs=None, device=None, progress=False, chain=False,
                                  step_noise=None) -> Iterator[dict]:
        """gaussian_diffusion.py:480-529: yields {"sample", "pred_xstart"} per step."""
        if device is None:
            device = condition.device
        if not isinstance(shape, (tuple, list)):
            raise AssertionError("shape must be a tuple or list")
        if noise is None:
            noise = th.randn(*shape, device=device)
        steps = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            steps = tqdm(steps)
        dit = self._fast_path(model, clip_denoised, denoised_fn, cond_fn, model_kwargs)
        if dit is not None:
            eng, tabs = dit.engine(condition.device), self.device_tables(condition.device)
            if step_noise is None:
                step_noise = self._draw_step_noise(noise)
            state = None
            for k, _ in enumerate(steps):
                with th.no_grad():
                    state = eng.sample_loop(tabs, condition, noise, step_noise, chain=chain, first_step=k, last_step=k + 1,
                                            state=state)
                yield {"sample": state["sample"].clone(), "pred_xstart": state["x0"].clone()}
            return
        x = noise
        for k, i in enumerate(steps):
            t = th.full((shape[0],), i, device=device, dtype=th.long)
            with th.no_grad():
                eps = step_noise[k] if step_noise is not None else None
                out = self.p_sample(model, condition, x if chain else noise, t, clip_denoised=clip_denoised,
                                    denoised_fn=denoised_fn, cond_fn=cond_fn, model_kwargs=model_kwargs, noise=eps)
            yield out
            x = out["sample"]

    # ------------------------------------------------------------------ DDIM (reference: broken)
    def ddim_sample(self, model, condition, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                    eta=0.0, noise=None):
        """gaussian_diffusion.py:531-578 with the `condition` argument threaded through (the reference omits it at
        :547 and raises TypeError).  Pinned against the reference's own DDIM code run with that one argument supplied
        (tests/golden/ddim_*.npz, oracle/make_golden.py golden_ddim)."""
        if cond_fn is not None:
            raise NotImplementedError("cond_fn guidance is never used by the reference callers")
        out = self.p_mean_variance(model, condition, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                   model_kwargs=model_kwargs)
        if noise is None:
            noise = th.randn_like(x)
        sample = ops.ddim_step(out["pred_xstart"].float(), x.float(), noise.float(), self._ddim_tables(x.device, float(eta)),
                               t.to(th.int64))
        return {"sample": sample, "pred_xstart": out["pred_xstart"]}

    def ddim_sample_loop(self, model, condition, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0, step_noise=None):
        final = None
        for final in self.ddim_sample_loop_progressive(model, condition, shape, noise=noise, clip_denoised=clip_denoised,
                                                       denoised_fn=denoised_fn, cond_fn=cond_fn, model_kwargs=model_kwargs,
                                                       device=device, progress=progress, eta=eta, step_noise=step_noise):
            pass
        return final["sample"]

    def ddim_sample_loop_progressive(self, model, condition, shape, noise=None, clip_denoised=True, denoised_fn=None,
                                     cond_fn=None, model_kwargs=None, device=None, progress=False, eta=0.0,
                                     step_noise=None):
        """gaussian_diffusion.py:652-698 (DDIM feeds the running sample, unlike p_sample_loop)."""
        if device is None:
            device = condition.device
        img = noise if noise is not None else th.randn(*shape, device=device)
        steps = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            steps = tqdm(steps)
        for k, i in enumerate(steps):
            t = th.full((shape[0],), i, device=device, dtype=th.long)
            with th.no_grad():
                out = self.ddim_sample(model, condition, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                       cond_fn=cond_fn, model_kwargs=model_kwargs, eta=eta,
                                       noise=step_noise[k] if step_noise is not None else None)
            yield out
            img = out["sample"]

    # ------------------------------------------------------------------ training
    @staticmethod
    def draw_scramble(batch: int, grid: int, add_mask: bool):
        """The host-side draws of one training step in the reference's order (gaussian_diffusion.py:757, 777-783): the piece
        permutation (numpy) and, with add_mask, per sample r = randint(0, G) zeroed slots (numpy + `random`).
        -> (perm int32 [G*G], keep float32 [B, G*G] or None)."""
        n = grid * grid
        perm = th.as_tensor(np.random.permutation(n), dtype=th.int32)
        keep = None
        if add_mask:
            keep = th.ones(batch, n)
            for i in range(batch):
                r = np.random.randint(0, grid)
                keep[i, random.sample(range(n), r)] = 0
        return perm, keep

    def _to_device_async(self, host: "th.Tensor", device) -> "th.Tensor":
        """Small host-drawn tables (the step's permutation, the mask slots) -> device WITHOUT a stream sync: a pageable
        host-to-device copy blocks the host until the stream has drained, i.e. once per training step (measured: 10 of the
        15 ms of host time per C3 step sat in two 36-byte `.to(device)` calls).  Staged through a small ring of pinned
        buffers; a slot is rewritten only after the copy that read it has run."""
        ring = self.__dict__.setdefault("_pin_ring", {})
        key = (host.dtype, tuple(host.shape))
        ent = ring.get(key)
        if ent is None:
            ent = ring[key] = {"pin": th.empty((16,) + tuple(host.shape), dtype=host.dtype).pin_memory(), "ev": [None] * 16, "i": 0}
        i = ent["i"]
        ent["i"] = (i + 1) % 16
        if ent["ev"][i] is not None:
            ent["ev"][i].synchronize()
        ent["pin"][i].copy_(host)
        out = ent["pin"][i].to(device, non_blocking=True)
        ev = th.cuda.Event()
        ev.record(th.cuda.current_stream(device))
        ent["ev"][i] = ev
        return out

    @staticmethod
    def _scramble(x, perm, grid, block):
        """[B,C,(g h),(g w)] -> pieces permuted so that slot i holds original piece perm[i] (:757-775): the rearrange /
        index / rearrange of the reference as one device gather (jpdvt_gather_pieces).  `perm`: host array or a device
        int32 tensor."""
        B = x.shape[0]
        if x.shape[-1] != grid * block or x.shape[-2] != grid * block:
            raise AssertionError(f"{tuple(x.shape)} images do not tile a {grid}x{grid} puzzle of {block}-pixel pieces")
        if not (isinstance(perm, th.Tensor) and perm.device == x.device):
            perm = th.as_tensor(np.asarray(perm), dtype=th.int32).to(x.device)
        idx = perm.to(th.int32).unsqueeze(0).expand(B, grid * grid).contiguous()
        return ops.gather_pieces(x.float().contiguous(), idx, grid)

    def training_losses(self, model, x_start, t, time_emb_start, model_kwargs=None, noise=None, block_size=96,
                        patch_size=16, add_mask=False, grid_size=3):
        """gaussian_diffusion.py:736-843: one training step's loss terms {"mse", "loss"}, each [N], differentiable
        w.r.t. the model parameters.  Random draws are consumed in the reference's order (torch randn for the image,
        numpy permutation, numpy/`random` mask draws, torch randn for the latents); the reference's per-step
        `plt.imsave` side effect (:796) is dropped on purpose."""
        if self.loss_type not in (LossType.MSE, LossType.RESCALED_MSE):
            raise NotImplementedError("the KL loss path of the reference is broken (SURVEY.md 8a row 27); only MSE is live")
        if model_kwargs is None:
            model_kwargs = {}
        B = x_start.shape[0]
        if B == 0:
            raise ValueError("training_losses: empty batch (the reference's loader drops incomplete batches, drop_last=True, train_JPDVT.py:311-319)")
        G, n = grid_size, grid_size * grid_size
        draws = getattr(self, "_draws", None)      # parity tests inject the reference's CPU draws here
        dev_draws = getattr(self, "_device_draws", None)   # Trainer's CUDA-graph step: permutation / mask slots already on the device
        noise_x = th.randn_like(x_start) if draws is None else draws["noise_x"].to(x_start.device)
        perm = None if dev_draws is not None else (np.random.permutation(n) if draws is None else np.asarray(draws["perm"]))
        keep_slots = None
        if dev_draws is not None:
            keep_slots = dev_draws.get("keep") if add_mask else None
        elif add_mask:
            if draws is not None:
                keep_slots = draws["masks"].clone()
            else:
                keep_slots = th.ones(B, n)
                for i in range(B):
                    r = np.random.randint(0, G)
                    keep_slots[i, random.sample(range(n), r)] = 0
        on_gpu = x_start.device.type == "cuda"
        if dev_draws is not None:
            perm_dev = dev_draws["perm"]
        else:
            perm_dev = (self._to_device_async(th.as_tensor(np.asarray(perm), dtype=th.int32), x_start.device) if on_gpu
                        else th.as_tensor(np.asarray(perm), dtype=th.int32))
            if keep_slots is not None and on_gpu:
                keep_slots = self._to_device_async(keep_slots.to(th.float32), x_start.device)
        x0 = self._scramble(x_start, perm_dev, G, block_size)
        tok = block_size // patch_size
        te = time_emb_start.to(x_start.device).float().expand(B, -1, -1)[:, perm_dev.long()]
        te0 = te.reshape(B, G, 1, G, 1, -1).expand(B, G, tok, G, tok, te.shape[-1]).reshape(B, n * tok * tok, -1).contiguous()
        noise_te = th.randn_like(te0) if draws is None else draws["noise_te"].to(te0.device)
        t = t.to(th.int64)
        if keep_slots is None:
            # masks == 1 everywhere: x_t * 0 + 1 * x_start == x_start exactly, the image condition stays clean (:800);
            # noise_x was still drawn above so the generator advances like the reference's
            keep, x_t = None, x0
        else:
            k = keep_slots.to(x0.device).reshape(B, 1, G, 1, G, 1).expand(B, x0.shape[1], G, block_size, G, block_size)
            keep = k.reshape_as(x0).contiguous()
            x_t = self.q_sample(x0, t, noise=noise_x, keep_mask=keep)
        te_t = self.q_sample(te0, t, noise=noise_te)
        inner = model.model if (hasattr(model, "model") and hasattr(model, "timestep_map")) else model
        x_out, te_out = inner(x_t, self._model_timesteps(t), te_t, **model_kwargs)
        if self.model_mean_type == ModelMeanType.START_X:
            target_x, target_te = x0, te0
        elif self.model_mean_type == ModelMeanType.EPSILON:
            target_x, target_te = noise_x, noise_te
        else:
            raise NotImplementedError(self.model_mean_type)
        if x_out.shape != x0.shape:
            raise AssertionError("model image output must match x_start")
        # mean_flat((target_te - te_out)**2) [+ mean_flat((target_x - x_out)**2 * (1 - masks))] (:835-838), fused
        if add_mask:
            mse = _MseLossFn.apply(te_out, x_out, target_te, target_x, keep_slots.to(x0.device), G)
        else:
            mse = _MseLossFn.apply(te_out, None, target_te, None, None, G)
        return {"mse": mse, "loss": mse}

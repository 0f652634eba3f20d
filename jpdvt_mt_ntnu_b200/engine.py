"""Device-side state of one JPDVT denoiser: packed bf16 weights, workspaces and the launch entry points.

This is host plumbing around libjpdvt_sm100.so: PyTorch owns every allocation (weights, workspaces, outputs); the C
side only sees raw pointers and the current CUDA stream (SURVEY.md 8b "Ownership").
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import Sampler, Weights, Workspace, check, ptr

HIDDEN = 768
LATENT = 8


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.bfloat16).contiguous()


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


class PackedWeights:
    """Reference-keyed fp32 state dict -> the layouts the kernels read (include/jpdvt_b200.h: jpdvt_weights)."""

    _generation = 0

    def __init__(self, state: Dict[str, torch.Tensor], depth: int, image_size: int, device: torch.device):
        self.depth, self.image_size = depth, image_size
        self.tokens = (image_size // 16) ** 2
        g = lambda k: state[k].to(device)
        w = g("x_embedder.proj.weight")
        if tuple(w.shape[1:]) != (3, 16, 16) or w.shape[0] != HIDDEN:
            raise _lib.JpdvtError(
                f"unsupported patch embedding {tuple(w.shape)}: the B200 path covers hidden 768 / patch 16 "
                "(the only wiring the reference forward type-checks for, models.py:177,287-288)")
        t: Dict[str, torch.Tensor] = {}
        t["w_patch"] = _bf16(w.reshape(HIDDEN, -1))
        t["b_embed"] = _f32(g("x_embedder.proj.bias") + g("time_emb_in.bias"))
        t["w_in_t"] = _f32(g("time_emb_in.weight").t())
        pos = g("pos_embed")
        if pos.shape[1] != self.tokens:
            raise _lib.JpdvtError(f"pos_embed has {pos.shape[1]} tokens, image size {image_size} needs {self.tokens}")
        t["pos"] = _f32(pos[0])
        t["t_w0"], t["t_b0"] = _f32(g("t_embedder.mlp.0.weight")), _f32(g("t_embedder.mlp.0.bias"))
        t["t_w2"], t["t_b2"] = _f32(g("t_embedder.mlp.2.weight")), _f32(g("t_embedder.mlp.2.bias"))
        blk = lambda i, k: g(f"blocks.{i}.{k}")
        t["w_ada"] = _bf16(torch.cat([blk(i, "adaLN_modulation.1.weight") for i in range(depth)]
                                     + [g("final_layer.adaLN_modulation.1.weight")], 0))
        t["b_ada"] = _f32(torch.cat([blk(i, "adaLN_modulation.1.bias") for i in range(depth)]
                                    + [g("final_layer.adaLN_modulation.1.bias")], 0))
        for name, key in (("qkv", "attn.qkv"), ("proj", "attn.proj"), ("fc1", "mlp.fc1"), ("fc2", "mlp.fc2")):
            t["w_" + name] = _bf16(torch.stack([blk(i, key + ".weight") for i in range(depth)], 0))
            t["b_" + name] = _f32(torch.stack([blk(i, key + ".bias") for i in range(depth)], 0))
        t["w_final"], t["b_final"] = _bf16(g("final_layer.linear.weight")), _f32(g("final_layer.linear.bias"))
        if t["w_final"].shape != (HIDDEN, HIDDEN):
            raise _lib.JpdvtError("final_layer.linear must be 768x768 (patch 16, 3 channels)")
        t["w_head1"], t["b_head1"] = _bf16(g("time_emb_out1.weight")), _f32(g("time_emb_out1.bias"))
        t["w_head2"], t["b_head2"] = _f32(g("time_emb_out2.weight")), _f32(g("time_emb_out2.bias"))
        self.tensors = t
        PackedWeights._generation += 1
        self.generation = PackedWeights._generation        # monotonically increasing: CUDA-graph cache key (ids get reused)
        self.n_mod = depth * 6 * HIDDEN + 2 * HIDDEN
        s = Weights()
        s.depth, s.tokens, s.image_size, s.reserved = depth, self.tokens, image_size, 0
        for k, v in t.items():
            setattr(s, k, ptr(v))
        self.struct = s


class DenoiserEngine:
    def __init__(self, depth: int, image_size: int, device: torch.device):
        _lib.require_device()
        self.lib = _lib.load()
        self.depth, self.image_size, self.device = depth, image_size, device
        self.tokens = (image_size // 16) ** 2
        self.weights: Optional[PackedWeights] = None
        self._ws_key = None
        self._ws_tensors: Dict[str, torch.Tensor] = {}
        self._ws = None
        self._noise_key: Optional[torch.Tensor] = None       # device int64[2] {seed, call counter} of the in-kernel step noise
        self._noise_calls = 0

    # ------------------------------------------------------------------ weights / workspace
    def load_state(self, state: Dict[str, torch.Tensor]) -> None:
        self.weights = PackedWeights(state, self.depth, self.image_size, self.device)

    def workspace(self, batch: int, cond_rows: int, need_image: bool, step_rows: int = 0) -> Workspace:
        rows = batch * self.tokens
        key = self._ws_key
        if key is None or key[0] < rows or key[1] < cond_rows or (need_image and not key[2]) or key[3] < step_rows:
            rows_cap = max(rows, key[0] if key else 0)
            cond_cap = max(cond_rows, key[1] if key else 0)
            step_cap = max(step_rows, key[3] if key else 0)
            img = need_image or bool(key and key[2])
            dev, bf, f32 = self.device, torch.bfloat16, torch.float32
            n_mod = self.depth * 6 * HIDDEN + 2 * HIDDEN
            fold = os.environ.get("JPDVT_LN_FOLD", "")[:1] == "1"      # opt-in path: its 100 MB of buffers only when asked for
            t = {
                "x": torch.empty(rows_cap, HIDDEN, device=dev, dtype=f32),
                "xn": torch.empty(rows_cap, HIDDEN, device=dev, dtype=bf),
                "qkv": torch.empty(rows_cap, 3 * HIDDEN, device=dev, dtype=bf),
                "attn": torch.empty(rows_cap, HIDDEN, device=dev, dtype=bf),
                "hid": torch.empty(rows_cap, 4 * HIDDEN, device=dev, dtype=bf),
                "y": torch.empty(rows_cap, HIDDEN, device=dev, dtype=bf),
                "y32": torch.empty(rows_cap, HIDDEN, device=dev, dtype=f32) if img else None,
                "c": torch.empty(cond_cap, HIDDEN, device=dev, dtype=f32),
                "silu_c": torch.empty(cond_cap, HIDDEN, device=dev, dtype=f32),
                "silu_c_bf16": torch.empty(cond_cap, HIDDEN, device=dev, dtype=bf),
                "mod": torch.empty(cond_cap, n_mod, device=dev, dtype=f32),
                # LayerNorm folded into qkv / fc1 (uniform-timestep forwards; csrc/fold.cu)
                "w_fold": torch.empty(self.depth * 7 * HIDDEN, HIDDEN, device=dev, dtype=bf) if fold else None,
                "fold_u": torch.empty(self.depth * 7 * HIDDEN, device=dev, dtype=f32) if fold else None,
                "fold_v": torch.empty(self.depth * 7 * HIDDEN, device=dev, dtype=f32) if fold else None,
                "row_stats": torch.empty(rows_cap, 2 * (HIDDEN // 256), 2, device=dev, dtype=f32) if fold else None,
                # conditioning of every step of one sample_loop call, computed ahead of the loop (api.cu: jpdvt_sample_loop)
                "c_steps": torch.empty(step_cap, HIDDEN, device=dev, dtype=f32) if step_cap else None,
                "silu_c_steps": torch.empty(step_cap, HIDDEN, device=dev, dtype=f32) if step_cap else None,
                "mod_steps": torch.empty(step_cap, n_mod, device=dev, dtype=f32) if step_cap else None,
                # hidden activations of the timestep MLP between its two phases (a buffer of its own: csrc/elementwise.cu)
                "te_hid": torch.empty(max(cond_cap, step_cap, 1), HIDDEN, device=dev, dtype=f32),
                # loop-invariant embedding of one sample_loop call (api.cu: jpdvt_sample_loop; JPDVT_HOIST_EMBED=0 = per step)
                "x_embed": torch.empty(rows_cap, HIDDEN, device=dev, dtype=f32) if step_cap else None,
            }
            ws = Workspace()
            ws.rows, ws.cond_rows, ws.step_rows = rows_cap, cond_cap, step_cap
            for k, v in t.items():
                setattr(ws, k, ptr(v))
            self.__dict__.get("_graphs", {}).clear()           # captured graphs hold the old buffers' addresses
            self._ws_tensors, self._ws, self._ws_key = t, ws, (rows_cap, cond_cap, img, step_cap)
        return self._ws

    # ------------------------------------------------------------------ one forward
    def forward(self, img: torch.Tensor, t: Optional[torch.Tensor], x_t: torch.Tensor, need_image: bool = True,
                step_ptr: Optional[torch.Tensor] = None, tmap: Optional[torch.Tensor] = None):
        """DiT.forward (image_model/models.py:273-293) -> (image or None, time_emb_out)."""
        assert self.weights is not None, "load_state() first"
        B = img.shape[0]
        if tuple(img.shape[1:]) != (3, self.image_size, self.image_size):
            raise _lib.JpdvtError(f"image batch {tuple(img.shape)} does not match input_size {self.image_size}")
        if tuple(x_t.shape) != (B, self.tokens, LATENT):
            raise _lib.JpdvtError(f"time_emb {tuple(x_t.shape)} must be [{B}, {self.tokens}, {LATENT}]")
        img = img.to(torch.float32).contiguous()
        x_t = x_t.to(torch.float32).contiguous()
        if t is not None:
            t = t.to(device=img.device, dtype=torch.int64).contiguous()
            if t.shape != (B,):
                raise _lib.JpdvtError(f"t {tuple(t.shape)} must be [{B}]")
        with _lib.on_device(self.device):
            ws = self.workspace(B, B if t is not None else 1, need_image)
            te = torch.empty(B, self.tokens, LATENT, device=img.device, dtype=torch.float32)
            out_img = torch.empty(B, 3, self.image_size, self.image_size, device=img.device, dtype=torch.float32) if need_image else None
            check(self.lib.jpdvt_denoiser_forward(C.byref(self.weights.struct), C.byref(ws), ptr(img), ptr(t), ptr(step_ptr),
                                                  ptr(tmap), ptr(x_t), ptr(te), ptr(out_img), B, _lib.stream_ptr(self.device)),
                  "jpdvt_denoiser_forward")
        return out_img, te

    # ------------------------------------------------------------------ per-step noise drawn inside the posterior kernel
    def next_noise_key(self) -> torch.Tensor:
        """Device int64[2] {seed, call counter} for one sample_loop call.  The seed is drawn once per engine from torch's
        default CPU generator (so `torch.manual_seed` makes runs repeatable); the counter advances on the device, in stream
        order, so a replayed CUDA graph draws fresh noise on every replay too."""
        if self._noise_key is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            self._noise_key = torch.tensor([seed, 0], dtype=torch.int64, device=self.device)
            self._noise_inc = torch.tensor([0, 1], dtype=torch.int64, device=self.device)
        else:
            self._noise_key.add_(self._noise_inc)
        self._noise_calls += 1
        return self._noise_key

    # ------------------------------------------------------------------ whole reverse loop
    def sample_loop_graphed(self, tables: dict, condition: torch.Tensor, noise: torch.Tensor,
                            step_noise: Optional[torch.Tensor], chain: bool = False) -> dict:
        """`sample_loop` replayed from a CUDA graph (captured once per batch size / step count / weights): for small
        batches the ~23,000 launches of a 250-step loop are launch- and gap-bound, the graph removes the host from the loop.
        Inputs are copied into graph-owned buffers; the returned tensors are copies."""
        n_steps = tables["num_steps"]
        if condition.shape[0] == 0:                                  # an empty shard: nothing to capture
            return self.sample_loop(tables, condition, noise, step_noise, chain=chain)
        strided = step_noise is not None and step_noise.dim() == noise.dim() + 1 and step_noise.shape[0] > 1
        # The graph bakes in raw pointers to the packed weights and the schedule tables: key on the weights' generation
        # counter (ids are reused after a re-pack) and on the tables' storage, and pin both objects in the entry.
        key = (tuple(condition.shape), n_steps, bool(chain), bool(strided), step_noise is None, self.weights.generation,
               tables["coef1"].data_ptr(), tables["timestep_map"].data_ptr())
        cache = self.__dict__.setdefault("_graphs", {})
        ent = cache.get(key)
        inputs = [("cond", condition), ("noise", noise)] + ([("step_noise", step_noise)] if step_noise is not None else [])
        with _lib.on_device(self.device):
            if ent is None:
                cache.clear()                                        # one graph at a time: they pin workspaces and inputs
                ent = {k: torch.empty_like(src, dtype=torch.float32).contiguous() for k, src in inputs}
                ent.update(step_noise=ent.get("step_noise"), state=None, weights=self.weights, tables=tables)
                for k, src in inputs:
                    ent[k].copy_(src)
                if step_noise is None:
                    self.next_noise_key()
                side = torch.cuda.Stream(device=condition.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):                        # warm-up outside capture (function attributes, workspaces)
                    ent["state"] = self.sample_loop(tables, ent["cond"], ent["noise"], ent["step_noise"], chain=chain, _advance_key=False)
                torch.cuda.current_stream(self.device).wait_stream(side)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.sample_loop(tables, ent["cond"], ent["noise"], ent["step_noise"], chain=chain, state=ent["state"], _advance_key=False)
                ent["graph"] = g
                cache[key] = ent
            for k, src in inputs:
                ent[k].copy_(src)
            if step_noise is None:
                self.next_noise_key()                                # advances the device-side call counter before the replay
            ent["graph"].replay()
            return {k: (v.clone() if v is not None else None) for k, v in ent["state"].items()}

    def sample_loop(self, tables: dict, condition: torch.Tensor, noise: torch.Tensor, step_noise: Optional[torch.Tensor],
                    chain: bool = False, record: bool = False, first_step: int = 0, last_step: Optional[int] = None,
                    state: Optional[dict] = None, _advance_key: bool = True) -> dict:
        """SpacedDiffusion.p_sample_loop on the device (diffusion/gaussian_diffusion.py:433-529).  step_noise: the per-step
        noise tensors ([steps, B, T, 8] or one [B, T, 8] reused), or None = drawn inside the posterior kernel (Philox)."""
        assert self.weights is not None
        with _lib.on_device(self.device):
            return self._sample_loop(tables, condition, noise, step_noise, chain, record, first_step, last_step, state, _advance_key)

    def _sample_loop(self, tables, condition, noise, step_noise, chain, record, first_step, last_step, state, advance_key):
        B = condition.shape[0]
        n_steps = tables["num_steps"]
        last_step = n_steps if last_step is None else last_step
        condition = condition.to(device=self.device, dtype=torch.float32).contiguous()
        noise = noise.to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(condition.shape[1:]) != (3, self.image_size, self.image_size) or tuple(noise.shape) != (B, self.tokens, LATENT):
            raise _lib.JpdvtError(f"sample_loop: condition {tuple(condition.shape)} / noise {tuple(noise.shape)} do not fit "
                                  f"input_size {self.image_size}")
        noise_key = None
        if step_noise is None:
            noise_key = self.next_noise_key() if (advance_key or self._noise_key is None) else self._noise_key
        else:
            step_noise = step_noise.to(device=self.device, dtype=torch.float32).contiguous()
            ok = tuple(step_noise.shape) == tuple(noise.shape) or (step_noise.dim() == noise.dim() + 1
                                                                  and tuple(step_noise.shape[1:]) == tuple(noise.shape)
                                                                  and step_noise.shape[0] in (1, n_steps))
            if not ok:
                raise _lib.JpdvtError(f"sample_loop: step_noise {tuple(step_noise.shape)} must be {tuple(noise.shape)} or "
                                      f"[{n_steps}, ...] of it")
        ws = self.workspace(B, 1, False, step_rows=max(0, last_step - first_step))
        dev = condition.device
        if state is None:
            state = {
                "x0": torch.empty(B, self.tokens, LATENT, device=dev, dtype=torch.float32),
                "sample": torch.empty(B, self.tokens, LATENT, device=dev, dtype=torch.float32),
                "traj_x0": torch.empty(n_steps, B, self.tokens, LATENT, device=dev, dtype=torch.float32) if record else None,
                "traj_sample": torch.empty(n_steps, B, self.tokens, LATENT, device=dev, dtype=torch.float32) if record else None,
            }
        s = Sampler()
        s.num_steps, s.chain = n_steps, int(bool(chain))
        s.step_ids, s.timestep_map = ptr(tables["step_ids"]), ptr(tables["timestep_map"])
        s.coef1, s.coef2, s.logvar = ptr(tables["coef1"]), ptr(tables["coef2"]), ptr(tables["logvar"])
        s.step_noise = ptr(step_noise)
        s.noise_key = ptr(noise_key)
        s.step_noise_stride = step_noise.stride(0) if (step_noise is not None and step_noise.dim() == noise.dim() + 1
                                                       and step_noise.shape[0] > 1) else 0
        s.x0, s.sample = ptr(state["x0"]), ptr(state["sample"])
        s.traj_x0, s.traj_sample = ptr(state["traj_x0"]), ptr(state["traj_sample"])
        check(self.lib.jpdvt_sample_loop(C.byref(self.weights.struct), C.byref(ws), C.byref(s), ptr(condition), ptr(noise),
                                         B, first_step, last_step, _lib.stream_ptr(self.device)), "jpdvt_sample_loop")
        return state

"""Training-side host plumbing: the differentiable denoiser call used by `training_losses`
(image_model/diffusion/gaussian_diffusion.py:817 inside image_model/train_JPDVT.py:357-370).

`DiT.forward` under autograd routes here.  One `torch.autograd.Function` wraps the whole denoiser: its forward runs
`jpdvt_train_forward` (sm_100a kernels, activations kept on a tape), its backward runs the three C stages
`jpdvt_train_backward_{head,block,embed}` and hands every parameter its fp32 gradient, so `loss.backward()`,
`torch.optim.AdamW`, DDP's gradient hooks and EMA all work unchanged on top of it.  No PyTorch op computes any part of
the denoiser or of its gradient.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Optional

import torch

from . import _lib
from ._lib import BwdScratch, Grads, Tape, WeightsT, check, ptr
from .engine import HIDDEN, LATENT, PackedWeights


def _grad_layout(depth: int):
    """(field, shape) of jpdvt_grads in order, plus the parameter names each field (or slice of it) feeds."""
    n_mod = depth * 6 * HIDDEN + 2 * HIDDEN
    return [
        ("w_patch", (HIDDEN, HIDDEN)), ("b_patch", (HIDDEN,)), ("w_in", (HIDDEN, LATENT)), ("b_in", (HIDDEN,)),
        ("t_w0", (HIDDEN, 256)), ("t_b0", (HIDDEN,)), ("t_w2", (HIDDEN, HIDDEN)), ("t_b2", (HIDDEN,)),
        ("w_ada", (n_mod, HIDDEN)), ("b_ada", (n_mod,)),
        ("w_qkv", (depth, 3 * HIDDEN, HIDDEN)), ("b_qkv", (depth, 3 * HIDDEN)),
        ("w_proj", (depth, HIDDEN, HIDDEN)), ("b_proj", (depth, HIDDEN)),
        ("w_fc1", (depth, 4 * HIDDEN, HIDDEN)), ("b_fc1", (depth, 4 * HIDDEN)),
        ("w_fc2", (depth, HIDDEN, 4 * HIDDEN)), ("b_fc2", (depth, HIDDEN)),
        ("w_final", (HIDDEN, HIDDEN)), ("b_final", (HIDDEN,)), ("w_head1", (64, HIDDEN)), ("b_head1", (64,)),
        ("w_head2", (LATENT, 64)), ("b_head2", (LATENT,)),
    ]


class TrainEngine:
    """Packed weights (+ transposed copies), tape, scratch and gradient buffers of one denoiser on one device."""

    def __init__(self, depth: int, image_size: int, device: torch.device):
        _lib.require_device()
        self.lib = _lib.load()
        self.depth, self.image_size, self.device = depth, image_size, device
        self.tokens = (image_size // 16) ** 2
        self.n_mod = depth * 6 * HIDDEN + 2 * HIDDEN
        self.weights: Optional[PackedWeights] = None
        self.wt_tensors: Dict[str, torch.Tensor] = {}
        self.wt = None
        self._batch = None
        self.tape_t: Dict[str, torch.Tensor] = {}
        self.scr_t: Dict[str, torch.Tensor] = {}
        self.tape = None
        self.scratch = None
        self.zeros = torch.zeros(max(self.n_mod, 4 * HIDDEN), device=device, dtype=torch.float32)
        self.w_struct = None
        self.last_flat = None
        self.grad_buffer: Optional[torch.Tensor] = None    # persistent flat gradient buffer owned by the Trainer (peer-mapped)
        self._keepalive = None
        self.ticket = 0      # forward counter: a backward must match the forward whose activations are on the tape

    def __deepcopy__(self, memo):
        return None

    # ------------------------------------------------------------------ weights
    def load_state(self, state: Dict[str, torch.Tensor]) -> None:
        """bf16 operand copies in both orientations ([out,in] for the forward / wgrad, [in,out] for the data gradients)."""
        self.weights = PackedWeights(state, self.depth, self.image_size, self.device)
        t = self.weights.tensors
        tr = lambda x: x.transpose(-1, -2).contiguous()
        w = {
            "w_qkv_t": tr(t["w_qkv"]), "w_proj_t": tr(t["w_proj"]), "w_fc1_t": tr(t["w_fc1"]), "w_fc2_t": tr(t["w_fc2"]),
            "w_final_t": tr(t["w_final"]), "w_head1_t": tr(t["w_head1"]), "w_ada_t": tr(t["w_ada"]),
            "t_w2_t": tr(t["t_w2"].to(torch.bfloat16)),
        }
        s = WeightsT()
        for k, v in w.items():
            setattr(s, k, ptr(v))
        self.wt_tensors, self.wt = w, s
        self.w_struct = self.weights.struct

    def adopt(self, w_struct, wt_struct, keepalive=None) -> None:
        """Use operand buffers owned by someone else (the Trainer's flat bf16 shadow + transposed copies)."""
        self.w_struct, self.wt, self._keepalive = w_struct, wt_struct, keepalive

    # ------------------------------------------------------------------ buffers
    def _ensure(self, batch: int) -> None:
        if self._batch == batch:
            return
        dev, bf, f32, d = self.device, torch.bfloat16, torch.float32, self.depth
        M, T = batch * self.tokens, self.tokens
        e = lambda *shape, dtype=bf: torch.empty(*shape, device=dev, dtype=dtype)
        tp = {
            "cols": e(M, HIDDEN), "x": e(2 * d + 1, M, HIDDEN, dtype=f32), "xn1": e(d, M, HIDDEN), "qkv": e(d, M, 3 * HIDDEN),
            "lse2": e(d, batch, 12, T, dtype=f32), "att": e(d, M, HIDDEN), "y1": e(d, M, HIDDEN), "xn2": e(d, M, HIDDEN),
            "hpre": e(d, M, 4 * HIDDEN), "h": e(d, M, 4 * HIDDEN), "y2": e(d, M, HIDDEN), "xnf": e(M, HIDDEN),
            "yfin": e(M, HIDDEN), "yfin32": e(M, HIDDEN, dtype=f32), "headpre": e(M, 64, dtype=f32),
            "feat": e(batch, 256, dtype=f32), "tpre": e(batch, HIDDEN, dtype=f32), "c": e(batch, HIDDEN, dtype=f32),
            "silu_c": e(batch, HIDDEN, dtype=f32), "silu_c_bf16": e(batch, HIDDEN), "mod": e(batch, self.n_mod, dtype=f32),
            "thid": e(batch, HIDDEN, dtype=f32),
        }
        tape = Tape()
        tape.rows, tape.batch, tape.reserved = M, batch, 0
        for k, v in tp.items():
            setattr(tape, k, ptr(v))
        need = int(self.lib.jpdvt_train_wgrad_scratch_floats(d, batch, T))
        sc = {
            "dx": e(M, HIDDEN, dtype=f32), "dxn": e(M, HIDDEN, dtype=f32), "dy": e(M, HIDDEN), "dh": e(M, 4 * HIDDEN),
            "dqkv": e(M, 3 * HIDDEN), "datt": e(M, HIDDEN), "dpre": e(M, 64), "dmod": e(batch, self.n_mod, dtype=f32),
            "dmod_bf16": e(batch, self.n_mod), "small_f32": e(4, batch, HIDDEN, dtype=f32), "small_bf16": e(4, batch, HIDDEN),
            "wgrad_scratch": e(max(need, 4), dtype=f32),
            "part": e(max(int(self.lib.jpdvt_bwd_part_floats(batch, T)), 4), dtype=f32),
        }
        scratch = BwdScratch()
        for k, v in sc.items():
            setattr(scratch, k, ptr(v))
        scratch.zeros = ptr(self.zeros)
        self.tape_t, self.tape, self.scr_t, self.scratch, self._batch = tp, tape, sc, scratch, batch

    def new_grads(self):
        """Fresh zero-filled flat fp32 gradient buffer + the struct of pointers into it + per-field views."""
        layout = _grad_layout(self.depth)
        sizes = [int(torch.Size(shape).numel()) for _, shape in layout]
        if self.grad_buffer is not None:            # the peers read this very buffer: same address every step
            flat = self.grad_buffer[:sum(sizes)]
            self.grad_buffer.zero_()
        else:
            flat = torch.zeros(sum(sizes), device=self.device, dtype=torch.float32)
        g, views, off = Grads(), {}, 0
        for (name, shape), n in zip(layout, sizes):
            v = flat[off:off + n].view(shape)
            views[name] = v
            setattr(g, name, v.data_ptr())
            off += n
        return flat, g, views

    # ------------------------------------------------------------------ forward / backward
    def forward(self, img: torch.Tensor, t: torch.Tensor, x_t: torch.Tensor):
        B = img.shape[0]
        if tuple(img.shape[1:]) != (3, self.image_size, self.image_size) or tuple(x_t.shape) != (B, self.tokens, LATENT) or t.shape != (B,):
            raise _lib.JpdvtError(f"bad training input shapes: img {tuple(img.shape)}, t {tuple(t.shape)}, time_emb {tuple(x_t.shape)}")
        with _lib.on_device(self.device):
            self._ensure(B)
            self.ticket += 1
            te = torch.empty(B, self.tokens, LATENT, device=self.device, dtype=torch.float32)
            out_img = torch.empty(B, 3, self.image_size, self.image_size, device=self.device, dtype=torch.float32)
            check(self.lib.jpdvt_train_forward(C.byref(self.w_struct), C.byref(self.tape), ptr(img), ptr(t), ptr(x_t), ptr(te),
                                               ptr(out_img), B, _lib.stream_ptr(self.device)), "jpdvt_train_forward")
        return out_img, te

    def backward(self, d_te: torch.Tensor, d_img: Optional[torch.Tensor], x_t: torch.Tensor,
                 stage_done: Optional[Callable[[str, Dict[str, torch.Tensor]], None]] = None):
        """Runs the backward stages; `stage_done(name, views)` is called after each stage (head, block<i>, embed) so a
        data-parallel trainer can start reducing that stage's gradients while the next stage computes."""
        with _lib.on_device(self.device):
            return self._backward(d_te, d_img, x_t, stage_done)

    def _backward(self, d_te, d_img, x_t, stage_done):
        flat, g, views = self.new_grads()
        self.scr_t["dmod"].zero_()
        st = _lib.stream_ptr(self.device)
        args = (C.byref(self.w_struct), C.byref(self.wt), C.byref(self.tape), C.byref(self.scratch), C.byref(g))
        check(self.lib.jpdvt_train_backward_head(*args, ptr(d_te), ptr(d_img) if d_img is not None else None, st),
              "jpdvt_train_backward_head")
        if stage_done:
            stage_done("head", views)
        for i in range(self.depth - 1, -1, -1):
            check(self.lib.jpdvt_train_backward_block(*args, i, st), "jpdvt_train_backward_block")
            if stage_done:
                stage_done(f"block{i}", views)
        check(self.lib.jpdvt_train_backward_embed(*args, ptr(x_t), st), "jpdvt_train_backward_embed")
        if stage_done:
            stage_done("embed", views)
        self.last_flat = flat
        return flat, views


def param_grad_map(module, views: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """jpdvt_grads fields -> reference parameter names (state-dict keys)."""
    d = module.depth
    out = {
        "x_embedder.proj.weight": views["w_patch"].view(HIDDEN, 3, 16, 16), "x_embedder.proj.bias": views["b_patch"],
        "time_emb_in.weight": views["w_in"], "time_emb_in.bias": views["b_in"],
        "t_embedder.mlp.0.weight": views["t_w0"], "t_embedder.mlp.0.bias": views["t_b0"],
        "t_embedder.mlp.2.weight": views["t_w2"], "t_embedder.mlp.2.bias": views["t_b2"],
        "final_layer.linear.weight": views["w_final"], "final_layer.linear.bias": views["b_final"],
        "time_emb_out1.weight": views["w_head1"], "time_emb_out1.bias": views["b_head1"],
        "time_emb_out2.weight": views["w_head2"], "time_emb_out2.bias": views["b_head2"],
        "final_layer.adaLN_modulation.1.weight": views["w_ada"][d * 6 * HIDDEN:], "final_layer.adaLN_modulation.1.bias": views["b_ada"][d * 6 * HIDDEN:],
    }
    for i in range(d):
        b = f"blocks.{i}."
        out[b + "adaLN_modulation.1.weight"] = views["w_ada"][i * 6 * HIDDEN:(i + 1) * 6 * HIDDEN]
        out[b + "adaLN_modulation.1.bias"] = views["b_ada"][i * 6 * HIDDEN:(i + 1) * 6 * HIDDEN]
        for short, key in (("qkv", "attn.qkv"), ("proj", "attn.proj"), ("fc1", "mlp.fc1"), ("fc2", "mlp.fc2")):
            out[b + key + ".weight"] = views["w_" + short][i]
            out[b + key + ".bias"] = views["b_" + short][i]
    return out


class _DenoiserFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, img, t, x_t, *params):
        eng = module.train_engine()
        img = img.detach().to(torch.float32).contiguous()
        x_t = x_t.detach().to(torch.float32).contiguous()
        t = t.detach().to(device=img.device, dtype=torch.int64).contiguous()
        out_img, te = eng.forward(img, t, x_t)
        ctx.module, ctx.eng, ctx.x_t, ctx.ticket = module, eng, x_t, eng.ticket
        ctx.set_materialize_grads(False)
        return out_img, te

    @staticmethod
    def backward(ctx, d_img, d_te):
        module, eng = ctx.module, ctx.eng
        if eng.ticket != ctx.ticket:
            raise _lib.JpdvtError("backward() of a denoiser call whose activations were overwritten by a later forward; "
                                  "run forward and backward of one micro-batch before the next forward")
        if d_te is None:
            d_te = torch.zeros(ctx.x_t.shape, device=ctx.x_t.device, dtype=torch.float32)
        hook = getattr(module, "_stage_hook", None)
        flat, views = eng.backward(d_te.to(torch.float32).contiguous(),
                                   d_img.to(torch.float32).contiguous() if d_img is not None else None, ctx.x_t, hook)
        by_name = param_grad_map(module, views)
        grads: List[Optional[torch.Tensor]] = []
        for name, p in module.named_parameters():
            grads.append(by_name.get(name) if p.requires_grad else None)
        return (None, None, None, None, *grads)


def denoiser_forward_with_grad(module, x, t, time_emb):
    """DiT.forward under autograd: (image head, time_emb_out), differentiable w.r.t. every trainable parameter."""
    params = [p for _, p in module.named_parameters()]
    return _DenoiserFn.apply(module, x, t, time_emb, *params)

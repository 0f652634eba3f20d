"""One Python function per C-ABI kernel entry point (torch tensors in, torch tensors out).

Used by the model/diffusion mirrors and by the parity tests; each wrapper only validates shapes, allocates the
output with torch and forwards raw pointers + the current stream to libjpdvt_sm100.so.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

HIDDEN, LATENT = 768, 8


def _lib_dev():
    return _lib.require_device()


def _need(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if t.dtype != dtype:
        raise _lib.JpdvtError(f"{name}: expected {dtype}, got {t.dtype}")
    return t.contiguous()


def ln_modulate(x: torch.Tensor, shift: torch.Tensor, scale: torch.Tensor, tokens: int,
                delta: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None,
                out_x: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [rows,768] fp32; shift/scale(/gate) [n_cond,768] fp32 with n_cond == rows/tokens or 1 -> bf16 [rows,768].
    With `delta` (bf16 [rows,768]) the residual stream is updated first: out_x = x + gate * delta (gate None = 1);
    out_x defaults to x itself (in place)."""
    lib = _lib_dev()
    shift, scale = _need(shift, torch.float32, "shift"), _need(scale, torch.float32, "scale")
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise _lib.JpdvtError("x must be contiguous fp32")
    rows = x.shape[0]
    stride = 0 if shift.shape[0] == 1 else HIDDEN
    y = torch.empty(rows, HIDDEN, device=x.device, dtype=torch.bfloat16)
    if delta is not None and out_x is None:
        out_x = x
    check(lib.jpdvt_ln_modulate_fwd(ptr(x), ptr(out_x) if out_x is not None else None,
                                    ptr(_need(delta, torch.bfloat16, "delta")) if delta is not None else None,
                                    ptr(_need(gate, torch.float32, "gate")) if gate is not None else None,
                                    ptr(shift), ptr(scale), stride, ptr(y), rows, tokens, stream_ptr()), "ln_modulate")
    return y


def gemm_bias(a, w, bias, want_f32_copy: bool = False):
    lib = _lib_dev()
    a, w, bias = _need(a, torch.bfloat16, "a"), _need(w, torch.bfloat16, "w"), _need(bias, torch.float32, "bias")
    m, k = a.shape
    n = w.shape[0]
    out = torch.empty(m, n, device=a.device, dtype=torch.bfloat16)
    out32 = torch.empty(m, n, device=a.device, dtype=torch.float32) if want_f32_copy else None
    check(lib.jpdvt_gemm_bias(ptr(a), ptr(w), ptr(bias), ptr(out), ptr(out32), m, n, k, stream_ptr()), "gemm_bias")
    return (out, out32) if want_f32_copy else out


def gemm_bias_f32(a, w, bias):
    lib = _lib_dev()
    a, w, bias = _need(a, torch.bfloat16, "a"), _need(w, torch.bfloat16, "w"), _need(bias, torch.float32, "bias")
    m, k = a.shape
    n = w.shape[0]
    out = torch.empty(m, n, device=a.device, dtype=torch.float32)
    check(lib.jpdvt_gemm_bias_f32(ptr(a), ptr(w), ptr(bias), ptr(out), m, n, k, stream_ptr()), "gemm_bias_f32")
    return out


def gemm_bias_gelu(a, w, bias):
    lib = _lib_dev()
    a, w, bias = _need(a, torch.bfloat16, "a"), _need(w, torch.bfloat16, "w"), _need(bias, torch.float32, "bias")
    m, k = a.shape
    n = w.shape[0]
    out = torch.empty(m, n, device=a.device, dtype=torch.bfloat16)
    check(lib.jpdvt_gemm_bias_gelu(ptr(a), ptr(w), ptr(bias), ptr(out), m, n, k, stream_ptr()), "gemm_bias_gelu")
    return out


def gemm_bias_gate(a, w, bias, gate, tokens: int) -> torch.Tensor:
    """gate[row // tokens] * (a @ w.T + bias) -> bf16.  gate [n_cond,N] with n_cond == rows/tokens or 1."""
    lib = _lib_dev()
    a, w = _need(a, torch.bfloat16, "a"), _need(w, torch.bfloat16, "w")
    bias, gate = _need(bias, torch.float32, "bias"), _need(gate, torch.float32, "gate")
    m, k = a.shape
    n = w.shape[0]
    stride = 0 if gate.shape[0] == 1 else n
    out = torch.empty(m, n, device=a.device, dtype=torch.bfloat16)
    check(lib.jpdvt_gemm_bias_gate(ptr(a), ptr(w), ptr(bias), ptr(gate), stride, ptr(out), m, n, k, tokens, stream_ptr()),
          "gemm_bias_gate")
    return out


def gemm_bias_gate_residual(x, a, w, bias, gate, tokens: int) -> torch.Tensor:
    """x += gate[row // tokens] * (a @ w.T + bias), in place on the fp32 residual stream x [M,N]; returns x."""
    lib = _lib_dev()
    a, w = _need(a, torch.bfloat16, "a"), _need(w, torch.bfloat16, "w")
    bias, gate = _need(bias, torch.float32, "bias"), _need(gate, torch.float32, "gate")
    m, k = a.shape
    n = w.shape[0]
    if x.dtype != torch.float32 or not x.is_contiguous() or tuple(x.shape) != (m, n) or not x.is_cuda:
        raise _lib.JpdvtError("gemm_bias_gate_residual: x must be a contiguous fp32 CUDA tensor of shape [M, N]")
    stride = 0 if gate.shape[0] == 1 else n
    check(lib.jpdvt_gemm_bias_gate_residual(ptr(a), ptr(w), ptr(bias), ptr(gate), stride, ptr(x), m, n, k, tokens,
                                            stream_ptr()), "gemm_bias_gate_residual")
    return x


def gemm_bias_gate_residual_ln(x, a, w, bias, gate, shift, scale, tokens: int):
    """x += gate[row // tokens] * (a @ w.T + bias) in place (fp32 [M,768]), then xn = LN(x) * (1 + scale) + shift in the same
    kernel; returns (x, xn bf16 [M,768]).  gate / shift / scale: [n_cond,768] with n_cond == rows/tokens or 1."""
    lib = _lib_dev()
    a, w = _need(a, torch.bfloat16, "a"), _need(w, torch.bfloat16, "w")
    bias, gate = _need(bias, torch.float32, "bias"), _need(gate, torch.float32, "gate")
    shift, scale = _need(shift, torch.float32, "shift"), _need(scale, torch.float32, "scale")
    m, k = a.shape
    n = w.shape[0]
    if x.dtype != torch.float32 or not x.is_contiguous() or tuple(x.shape) != (m, n) or not x.is_cuda:
        raise _lib.JpdvtError("gemm_bias_gate_residual_ln: x must be a contiguous fp32 CUDA tensor of shape [M, N]")
    if shift.shape != scale.shape or gate.shape[0] != shift.shape[0]:
        raise _lib.JpdvtError("gemm_bias_gate_residual_ln: gate / shift / scale must have the same number of conditioning rows")
    g_stride = 0 if gate.shape[0] == 1 else n
    xn = torch.empty(m, n, device=a.device, dtype=torch.bfloat16)
    check(lib.jpdvt_gemm_bias_gate_residual_ln(ptr(a), ptr(w), ptr(bias), ptr(gate), g_stride, ptr(x), ptr(shift), ptr(scale),
                                               g_stride, ptr(xn), m, n, k, tokens, stream_ptr()), "gemm_bias_gate_residual_ln")
    return x, xn


def gemm_bias_gate_residual_copy(x, a, w, bias, gate, tokens: int):
    """x += gate[row // tokens] * (a @ w.T + bias) in place (fp32 [M,N], N % 256 == 0); also returns bf16(x) and the rows'
    (sum, sum of squares) partials [M, 2N/256, 2] - the producer half of the folded LayerNorm."""
    lib = _lib_dev()
    a, w = _need(a, torch.bfloat16, "a"), _need(w, torch.bfloat16, "w")
    bias, gate = _need(bias, torch.float32, "bias"), _need(gate, torch.float32, "gate")
    m, k = a.shape
    n = w.shape[0]
    if x.dtype != torch.float32 or not x.is_contiguous() or tuple(x.shape) != (m, n) or not x.is_cuda:
        raise _lib.JpdvtError("gemm_bias_gate_residual_copy: x must be a contiguous fp32 CUDA tensor of shape [M, N]")
    xb = torch.empty(m, n, device=a.device, dtype=torch.bfloat16)
    stats = torch.empty(m, 2 * (n // 256), 2, device=a.device, dtype=torch.float32)
    check(lib.jpdvt_gemm_bias_gate_residual_copy(ptr(a), ptr(w), ptr(bias), ptr(gate), 0 if gate.shape[0] == 1 else n, ptr(x),
                                                 ptr(xb), ptr(stats), m, n, k, tokens, stream_ptr()), "gemm_bias_gate_residual_copy")
    return x, xb, stats


def fold_ln_weights(w_qkv, w_fc1, b_qkv, b_fc1, mod):
    """W' = W (1 + scale), u = rowsum(W'), v = b + W @ shift for the qkv / fc1 matrices of every block.  w_qkv [depth,2304,768],
    w_fc1 [depth,3072,768] bf16; mod: row 0 of the adaLN table.  Returns (w_fold [depth,5376,768] bf16, u, v [depth,5376])."""
    lib = _lib_dev()
    w_qkv, w_fc1 = _need(w_qkv, torch.bfloat16, "w_qkv"), _need(w_fc1, torch.bfloat16, "w_fc1")
    b_qkv, b_fc1, mod = _need(b_qkv, torch.float32, "b_qkv"), _need(b_fc1, torch.float32, "b_fc1"), _need(mod, torch.float32, "mod")
    depth = w_qkv.shape[0]
    wf = torch.empty(depth, 5376, 768, device=w_qkv.device, dtype=torch.bfloat16)
    u = torch.empty(depth, 5376, device=w_qkv.device, dtype=torch.float32)
    v = torch.empty_like(u)
    check(lib.jpdvt_fold_ln_weights(ptr(w_qkv), ptr(w_fc1), ptr(b_qkv), ptr(b_fc1), ptr(mod), ptr(wf), ptr(u), ptr(v), depth,
                                    stream_ptr()), "fold_ln_weights")
    return wf, u, v


def gemm_ln_folded(xb, stats, w_fold, u, v, gelu: bool = False):
    """[gelu_tanh](rstd * (xb @ w_fold.T) - rstd * mean * u + v) -> bf16 [M,N]; mean / rstd per row from `stats`."""
    lib = _lib_dev()
    xb, w_fold = _need(xb, torch.bfloat16, "xb"), _need(w_fold, torch.bfloat16, "w_fold")
    stats, u, v = _need(stats, torch.float32, "stats"), _need(u, torch.float32, "u"), _need(v, torch.float32, "v")
    m, k = xb.shape
    n = w_fold.shape[0]
    out = torch.empty(m, n, device=xb.device, dtype=torch.bfloat16)
    check(lib.jpdvt_gemm_ln_folded(1 if gelu else 0, ptr(xb), ptr(stats), stats.shape[1], ptr(w_fold), ptr(u), ptr(v), ptr(out),
                                   m, n, k, stream_ptr()), "gemm_ln_folded")
    return out


def patchify(img: torch.Tensor) -> torch.Tensor:
    lib = _lib_dev()
    img = _need(img, torch.float32, "img")
    b, _, s, _ = img.shape
    cols = torch.empty(b * (s // 16) ** 2, HIDDEN, device=img.device, dtype=torch.bfloat16)
    check(lib.jpdvt_patchify(ptr(img), ptr(cols), b, s, stream_ptr()), "patchify")
    return cols


def unpatchify(y: torch.Tensor, batch: int, size: int) -> torch.Tensor:
    lib = _lib_dev()
    y = _need(y, torch.float32, "y")
    img = torch.empty(batch, 3, size, size, device=y.device, dtype=torch.float32)
    check(lib.jpdvt_unpatchify(ptr(y), ptr(img), batch, size, stream_ptr()), "unpatchify")
    return img


def gemm_patch_embed(cols, w_patch, bias, x_t, w_in_t, pos, tokens: int) -> torch.Tensor:
    lib = _lib_dev()
    m = cols.shape[0]
    x = torch.empty(m, HIDDEN, device=cols.device, dtype=torch.float32)
    check(lib.jpdvt_gemm_patch_embed(ptr(_need(cols, torch.bfloat16, "cols")), ptr(_need(w_patch, torch.bfloat16, "w")),
                                     ptr(_need(bias, torch.float32, "bias")), ptr(_need(x_t, torch.float32, "x_t")),
                                     ptr(_need(w_in_t, torch.float32, "w_in_t")), ptr(_need(pos, torch.float32, "pos")),
                                     ptr(x), m, tokens, stream_ptr()), "gemm_patch_embed")
    return x


def final_head(y, w1, b1, w2, b2) -> torch.Tensor:
    lib = _lib_dev()
    m = y.shape[0]
    out = torch.empty(m, LATENT, device=y.device, dtype=torch.float32)
    check(lib.jpdvt_final_head_fwd(ptr(_need(y, torch.bfloat16, "y")), ptr(_need(w1, torch.bfloat16, "w1")),
                                   ptr(_need(b1, torch.float32, "b1")), ptr(_need(w2, torch.float32, "w2")),
                                   ptr(_need(b2, torch.float32, "b2")), ptr(out), m, stream_ptr()), "final_head")
    return out


def attention(qkv: torch.Tensor, batch: int, tokens: int, return_lse: bool = False):
    lib = _lib_dev()
    qkv = _need(qkv, torch.bfloat16, "qkv")
    out = torch.empty(batch * tokens, HIDDEN, device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty(batch, 12, tokens, device=qkv.device, dtype=torch.float32) if return_lse else None
    check(lib.jpdvt_attention_fwd(ptr(qkv), ptr(out), ptr(lse), batch, tokens, stream_ptr()), "attention")
    return (out, lse) if return_lse else out


def timestep_embed(t: torch.Tensor, w0, b0, w2, b2) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = _lib_dev()
    t = _need(t, torch.int64, "t")
    n = t.shape[0]
    c = torch.empty(n, HIDDEN, device=t.device, dtype=torch.float32)
    sc = torch.empty_like(c)
    hid = torch.empty_like(c)          # scratch of its own between the two grid-wide phases (must not alias c / sc)
    check(lib.jpdvt_timestep_embed(ptr(t), n, None, None, ptr(_need(w0, torch.float32, "w0")), ptr(_need(b0, torch.float32, "b0")),
                                   ptr(_need(w2, torch.float32, "w2")), ptr(_need(b2, torch.float32, "b2")), ptr(c), ptr(sc),
                                   ptr(hid), stream_ptr()), "timestep_embed")
    return c, sc


def adaln_table(silu_c, w_all, b_all) -> torch.Tensor:
    lib = _lib_dev()
    rows, n_out = silu_c.shape[0], w_all.shape[0]
    out = torch.empty(rows, n_out, device=silu_c.device, dtype=torch.float32)
    check(lib.jpdvt_adaln_table(ptr(_need(silu_c, torch.float32, "silu_c")), rows, ptr(_need(w_all, torch.bfloat16, "w_all")),
                                ptr(_need(b_all, torch.float32, "b_all")), ptr(out), n_out, stream_ptr()), "adaln_table")
    return out


def posterior_step(x0, x_t, noise, coef1, coef2, logvar, t: torch.Tensor):
    """-> (mean, sample); t int64 [B] respaced step indices."""
    lib = _lib_dev()
    x0, x_t, noise = _need(x0, torch.float32, "x0"), _need(x_t, torch.float32, "x_t"), _need(noise, torch.float32, "noise")
    mean, sample = torch.empty_like(x0), torch.empty_like(x0)
    n, per = x0.numel(), x0.numel() // x0.shape[0]
    check(lib.jpdvt_posterior_step(ptr(x0), ptr(x_t), ptr(noise), ptr(coef1), ptr(coef2), ptr(logvar),
                                   ptr(_need(t, torch.int64, "t")), None, ptr(mean), ptr(sample), n, per, stream_ptr()),
          "posterior_step")
    return mean, sample


def philox_key(seed: int, call: int, device) -> torch.Tensor:
    """Device int64[2] {seed, call counter}: the key of the in-kernel per-step noise (jpdvt_sampler.noise_key)."""
    return torch.tensor([int(seed) & 0x7fffffffffffffff, int(call)], dtype=torch.int64, device=device)


def philox_normal(key: torch.Tensor, n: int, step: int = 0, return_raw: bool = False):
    """The normals posterior_step_philox draws for loop position `step` (n % 4 == 0); optionally the raw Philox4x32-10 words."""
    lib = _lib_dev()
    key = _need(key, torch.int64, "key")
    out = torch.empty(n, device=key.device, dtype=torch.float32)
    raw = torch.empty(n, device=key.device, dtype=torch.int32) if return_raw else None
    check(lib.jpdvt_philox_normal(ptr(out), ptr(raw), n, int(step), ptr(key), stream_ptr()), "philox_normal")
    return (out, raw) if return_raw else out


def posterior_step_philox(x0, x_t, key: torch.Tensor, step: int, coef1, coef2, logvar, t: torch.Tensor) -> torch.Tensor:
    """posterior_step whose noise is drawn inside the kernel from Philox (key, loop position `step`) -> sample."""
    lib = _lib_dev()
    x0, x_t = _need(x0, torch.float32, "x0"), _need(x_t, torch.float32, "x_t")
    sample = torch.empty_like(x0)
    n, per = x0.numel(), x0.numel() // x0.shape[0]
    check(lib.jpdvt_posterior_step_philox(ptr(x0), ptr(x_t), ptr(_need(key, torch.int64, "key")), int(step), ptr(coef1), ptr(coef2),
                                          ptr(logvar), ptr(_need(t, torch.int64, "t")), None, ptr(sample), n, per, stream_ptr()),
          "posterior_step_philox")
    return sample


def mse_loss_fwd(te_out, te_tgt, img_out=None, img_tgt=None, keep_slots=None, grid: int = 0) -> torch.Tensor:
    """Per-sample loss terms of training_losses (gaussian_diffusion.py:835-838): mean((te_tgt - te_out)^2)
    [+ mean((img_tgt - img_out)^2 * (1 - keep))]; keep_slots fp32 [B, grid*grid]."""
    lib = _lib_dev()
    te_out, te_tgt = _need(te_out, torch.float32, "te_out"), _need(te_tgt, torch.float32, "te_tgt")
    b = te_out.shape[0]
    per = te_out.numel() // max(b, 1)
    loss = torch.empty(b, device=te_out.device, dtype=torch.float32)
    part = torch.empty(max(int(lib.jpdvt_mse_part_floats(b)), 1), device=te_out.device, dtype=torch.float32)
    size = 0
    if img_out is not None:
        img_out, img_tgt = _need(img_out, torch.float32, "img_out"), _need(img_tgt, torch.float32, "img_tgt")
        keep_slots = _need(keep_slots, torch.float32, "keep_slots")
        size = img_out.shape[-1]
        if img_out.shape != img_tgt.shape or tuple(keep_slots.shape) != (b, grid * grid) or img_out.shape[1] != 3:
            raise _lib.JpdvtError("mse_loss: image / mask shapes do not fit")
    check(lib.jpdvt_mse_loss_fwd(ptr(te_out), ptr(te_tgt), per, ptr(img_out), ptr(img_tgt), ptr(keep_slots), size, grid, ptr(part),
                                 ptr(loss), b, stream_ptr()), "mse_loss_fwd")
    return loss


def mse_loss_bwd(dloss, te_out, te_tgt, img_out=None, img_tgt=None, keep_slots=None, grid: int = 0):
    """-> (d_te, d_img or None) for upstream per-sample gradients dloss [B]."""
    lib = _lib_dev()
    dloss = _need(dloss, torch.float32, "dloss")
    b = te_out.shape[0]
    per = te_out.numel() // max(b, 1)
    d_te = torch.empty_like(te_out)
    d_img = torch.empty_like(img_out) if img_out is not None else None
    size = img_out.shape[-1] if img_out is not None else 0
    check(lib.jpdvt_mse_loss_bwd(ptr(te_out), ptr(te_tgt), per, ptr(img_out), ptr(img_tgt), ptr(keep_slots), size, grid, ptr(dloss),
                                 ptr(d_te), ptr(d_img), b, stream_ptr()), "mse_loss_bwd")
    return d_te, d_img


def ddim_step(x0, x_t, noise, tabs: dict, t: torch.Tensor) -> torch.Tensor:
    """tabs: fp32 device tables recip/recipm1/sqrt_abp/dir/sigma indexed by respaced step."""
    lib = _lib_dev()
    x0, x_t, noise = _need(x0, torch.float32, "x0"), _need(x_t, torch.float32, "x_t"), _need(noise, torch.float32, "noise")
    out = torch.empty_like(x0)
    n, per = x0.numel(), x0.numel() // x0.shape[0]
    check(lib.jpdvt_ddim_step(ptr(x0), ptr(x_t), ptr(noise), ptr(tabs["recip"]), ptr(tabs["recipm1"]), ptr(tabs["sqrt_abp"]),
                              ptr(tabs["dir"]), ptr(tabs["sigma"]), ptr(_need(t, torch.int64, "t")), None, ptr(out), n, per,
                              stream_ptr()), "ddim_step")
    return out


def q_sample(x0, noise, sqrt_ac, sqrt_1mac, t, keep: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib_dev()
    x0, noise = _need(x0, torch.float32, "x0"), _need(noise, torch.float32, "noise")
    out = torch.empty_like(x0)
    n, per = x0.numel(), x0.numel() // x0.shape[0]
    check(lib.jpdvt_q_sample(ptr(x0), ptr(noise), ptr(sqrt_ac), ptr(sqrt_1mac), ptr(_need(t, torch.int64, "t")),
                             ptr(keep.contiguous()) if keep is not None else None, ptr(out), n, per, stream_ptr()), "q_sample")
    return out


def assign_from_scores(scores: torch.Tensor, sentinel: float = 1e9):
    """scores fp64 [B,n,n] (rows = slots, cols = grid cells) -> (order, pred) int32 [B,n]  (inference.py:113-125,306)."""
    lib = _lib_dev()
    scores = _need(scores, torch.float64, "scores")
    b, n, _ = scores.shape
    order = torch.empty(b, n, device=scores.device, dtype=torch.int32)
    pred = torch.empty_like(order)
    check(lib.jpdvt_assign_from_scores(ptr(scores), b, n, float(sentinel), ptr(order), ptr(pred), stream_ptr()),
          "assign_from_scores")
    return order, pred


def assign_greedy_l1(latents: torch.Tensor, canon: torch.Tensor, grid: int, sentinel: float = 1e9,
                     return_scores: bool = False):
    """latents fp32 [B,T,8] -> (order, pred[, scores])  (inference.py:294-306)."""
    lib = _lib_dev()
    latents, canon = _need(latents, torch.float32, "latents"), _need(canon, torch.float32, "canon")
    b, t, _ = latents.shape
    n = grid * grid
    side = int(round((t // n) ** 0.5))
    if n * side * side != t:
        raise _lib.JpdvtError(f"{t} tokens do not tile a {grid}x{grid} puzzle (S/(16*G) must be integral, inference.py:295)")
    order = torch.empty(b, n, device=latents.device, dtype=torch.int32)
    pred = torch.empty_like(order)
    scores = torch.empty(b, n, n, device=latents.device, dtype=torch.float64) if return_scores else None
    check(lib.jpdvt_assign_greedy_l1(ptr(latents), ptr(canon), b, grid, side, float(sentinel), ptr(order), ptr(pred),
                                     ptr(scores), stream_ptr()), "assign_greedy_l1")
    return (order, pred, scores) if return_scores else (order, pred)


def gather_pieces(images: torch.Tensor, perm: torch.Tensor, grid: int, keep: Optional[torch.Tensor] = None,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out slot i = images piece perm[b, i]; slots with keep[b, i] == 0 are zeroed.  images fp32 [B,C,S,S], perm int32
    [B,G*G], keep uint8 [B,G*G] or None.  perm = scramble indices -> the scramble of inference_ddp.py:382-395;
    perm = `order` of the assignment -> the reconstruction of inference_ddp.py:449-455."""
    lib = _lib_dev()
    images, perm = _need(images, torch.float32, "images"), _need(perm, torch.int32, "perm")
    b, c, s, s2 = images.shape
    if s != s2 or tuple(perm.shape) != (b, grid * grid):
        raise _lib.JpdvtError(f"gather_pieces: images {tuple(images.shape)} / perm {tuple(perm.shape)} do not fit a {grid}x{grid} puzzle")
    if keep is not None:
        keep = _need(keep, torch.uint8, "keep")
        if tuple(keep.shape) != tuple(perm.shape):
            raise _lib.JpdvtError("gather_pieces: keep must have the shape of perm")
    if out is None:
        out = torch.empty_like(images)
    check(lib.jpdvt_gather_pieces(ptr(images), ptr(out), ptr(perm), ptr(keep), b, c, s, grid, stream_ptr()), "gather_pieces")
    return out


def crop_pieces(images: torch.Tensor, grid: int, out_piece: int) -> torch.Tensor:
    """Crop-gap erosion (train_JPDVT.py:345-349): centre-crop every piece of the grid x grid puzzle to out_piece pixels and
    re-tile.  images fp32 [B,C,S,S] with S % grid == 0 -> [B,C,grid*out_piece,grid*out_piece]."""
    lib = _lib_dev()
    images = _need(images, torch.float32, "images")
    b, c, s, s2 = images.shape
    if s != s2 or s % grid != 0:
        raise _lib.JpdvtError(f"crop_pieces: {s}x{s2} images do not tile a {grid}x{grid} puzzle")
    in_piece = s // grid
    if not 0 < out_piece <= in_piece:
        raise _lib.JpdvtError(f"crop_pieces: cannot crop {in_piece}-pixel pieces to {out_piece}")
    off = int(round((in_piece - out_piece) / 2.0))            # torchvision.transforms.functional.center_crop
    out = torch.empty(b, c, grid * out_piece, grid * out_piece, device=images.device, dtype=torch.float32)
    check(lib.jpdvt_crop_pieces(ptr(images), ptr(out), b, c, grid, in_piece, out_piece, off, stream_ptr()), "crop_pieces")
    return out


def score_placements(pred: torch.Tensor, truth: torch.Tensor, totals: Optional[torch.Tensor] = None):
    """(correct int32 [B], matches int32 [B]) for pred/truth int32 [B,n] (inference_ddp.py:431-447); `totals` (int64 [3],
    optional) accumulates (puzzles correct, pieces correct, puzzles)."""
    lib = _lib_dev()
    pred, truth = _need(pred, torch.int32, "pred"), _need(truth, torch.int32, "truth")
    if pred.shape != truth.shape or pred.dim() != 2:
        raise _lib.JpdvtError("score_placements: pred and truth must both be [B, n]")
    if totals is not None and (totals.dtype != torch.int64 or totals.numel() != 3 or not totals.is_contiguous()):
        raise _lib.JpdvtError("score_placements: totals must be a contiguous int64 tensor of 3 elements")
    b, n = pred.shape
    correct = torch.empty(b, device=pred.device, dtype=torch.int32)
    matches = torch.empty_like(correct)
    check(lib.jpdvt_score_placements(ptr(pred), ptr(truth), b, n, ptr(correct), ptr(matches), ptr(totals), stream_ptr()),
          "score_placements")
    return correct, matches


# ------------------------------------------------------------------------------------------------- training kernels
def gemm_wgrad(p: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """dW [out_rows, n_cols] fp32 = p[m, out_rows].T @ q[m, n_cols]  (bf16 operands, tcgen05 MN-major GEMM)."""
    lib = _lib_dev()
    p, q = _need(p, torch.bfloat16, "p"), _need(q, torch.bfloat16, "q")
    m, out_rows = p.shape
    n_cols = q.shape[1]
    dw = torch.empty(out_rows, n_cols, device=p.device, dtype=torch.float32)
    need = lib.jpdvt_wgrad_scratch_floats(m, out_rows, n_cols)
    scratch = torch.empty(max(need, 1), device=p.device, dtype=torch.float32)
    check(lib.jpdvt_gemm_wgrad(ptr(p), ptr(q), ptr(dw), ptr(scratch), m, out_rows, n_cols, stream_ptr()), "gemm_wgrad")
    return dw


def gemm_dgelu(a, w, pre) -> torch.Tensor:
    """(a @ w.T) * pre -> bf16, where `pre` holds gelu_tanh'(fc1 pre-activation) (kept by the training forward)."""
    lib = _lib_dev()
    a, w, pre = _need(a, torch.bfloat16, "a"), _need(w, torch.bfloat16, "w"), _need(pre, torch.bfloat16, "pre")
    m, k = a.shape
    n = w.shape[0]
    out = torch.empty(m, n, device=a.device, dtype=torch.bfloat16)
    check(lib.jpdvt_gemm_dgelu(ptr(a), ptr(w), ptr(pre), ptr(out), m, n, k, stream_ptr()), "gemm_dgelu")
    return out


def gemm_dgrad(dy, w, gprime=None, out_dtype=torch.float32, with_colsum: bool = False):
    """dX = dY . W from the layer's own [out, in] weight (MN-major B operand); `gprime`: multiply by gelu' (bf16 output);
    `with_colsum` (dGELU form): also the column sums of the output -> (out, colsum [n_in])."""
    lib = _lib_dev()
    dy, w = _need(dy, torch.bfloat16, "dy"), _need(w, torch.bfloat16, "w")
    m, k_out = dy.shape
    n_in = w.shape[1]
    if gprime is not None:
        out_dtype = torch.bfloat16
    out = torch.empty(m, n_in, device=dy.device, dtype=out_dtype)
    cs = torch.zeros(n_in, device=dy.device, dtype=torch.float32) if with_colsum else None
    check(lib.jpdvt_gemm_dgrad(ptr(dy), ptr(w), ptr(_need(gprime, torch.bfloat16, "gprime")) if gprime is not None else None,
                               ptr(out) if out_dtype == torch.bfloat16 else None, ptr(out) if out_dtype == torch.float32 else None,
                               ptr(cs), m, n_in, k_out, stream_ptr()), "gemm_dgrad")
    return (out, cs) if with_colsum else out


def attention_bwd(qkv, o, d_o, lse2, batch: int, tokens: int, with_bias_grad: bool = False):
    """-> dqkv, or (dqkv, dbias [2304] = column sums of dqkv) with `with_bias_grad`."""
    lib = _lib_dev()
    dqkv = torch.empty_like(qkv)
    dbias = torch.zeros(3 * HIDDEN, device=qkv.device, dtype=torch.float32) if with_bias_grad else None
    check(lib.jpdvt_attention_bwd(ptr(_need(qkv, torch.bfloat16, "qkv")), ptr(_need(o, torch.bfloat16, "o")),
                                  ptr(_need(d_o, torch.bfloat16, "d_o")), ptr(_need(lse2, torch.float32, "lse2")), ptr(dqkv),
                                  ptr(dbias), batch, tokens, stream_ptr()), "attention_bwd")
    return (dqkv, dbias) if with_bias_grad else dqkv


def gate_bwd(dx, y, gate, tokens: int):
    """-> (dy bf16, dgate [B,768], dbias [768])."""
    lib = _lib_dev()
    batch = dx.shape[0] // tokens
    dy = torch.empty_like(y)
    dgate = torch.zeros(batch, HIDDEN, device=dx.device, dtype=torch.float32)
    dbias = torch.zeros(HIDDEN, device=dx.device, dtype=torch.float32)
    part = torch.empty(max(int(lib.jpdvt_bwd_part_floats(batch, tokens)), 4), device=dx.device, dtype=torch.float32)
    check(lib.jpdvt_gate_bwd(ptr(_need(dx, torch.float32, "dx")), ptr(_need(y, torch.bfloat16, "y")),
                             ptr(_need(gate, torch.float32, "gate")), HIDDEN, ptr(dy), ptr(dgate), HIDDEN, ptr(dbias), ptr(part),
                             batch, tokens, stream_ptr()), "gate_bwd")
    return dy, dgate, dbias


def ln_modulate_bwd(x, dxn, scale, tokens: int, dx: Optional[torch.Tensor] = None):
    """-> (dx fp32 (accumulated into `dx` when given), dshift [B,768], dscale [B,768], dx_bf16)."""
    lib = _lib_dev()
    batch = x.shape[0] // tokens
    acc = dx is not None
    if dx is None:
        dx = torch.empty_like(x)
    dshift = torch.zeros(batch, HIDDEN, device=x.device, dtype=torch.float32)
    dscale = torch.zeros_like(dshift)
    dxb = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    part = torch.empty(max(int(lib.jpdvt_bwd_part_floats(batch, tokens)), 4), device=x.device, dtype=torch.float32)
    check(lib.jpdvt_ln_modulate_bwd(ptr(_need(x, torch.float32, "x")), ptr(_need(dxn, torch.float32, "dxn")),
                                    ptr(_need(scale, torch.float32, "scale")), HIDDEN, ptr(dx), int(acc), ptr(dshift), ptr(dscale),
                                    HIDDEN, ptr(dxb), ptr(part), batch, tokens, stream_ptr()), "ln_modulate_bwd")
    return dx, dshift, dscale, dxb


def ln_gate_bwd(x, dxn, scale, tokens: int, dx: Optional[torch.Tensor] = None, y=None, gate=None):
    """Fused LayerNorm-modulate backward + the gate backward that consumes its dx.
    -> (dx fp32, dshift, dscale, dx_bf16, dy bf16 | None, dgate | None, dbias | None)."""
    lib = _lib_dev()
    batch = x.shape[0] // tokens
    acc = dx is not None
    if dx is None:
        dx = torch.empty_like(x)
    dshift = torch.zeros(batch, HIDDEN, device=x.device, dtype=torch.float32)
    dscale = torch.zeros_like(dshift)
    dxb = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    dy = dgate = dbias = None
    if y is not None:
        dy = torch.empty_like(_need(y, torch.bfloat16, "y"))
        dgate, dbias = torch.zeros_like(dshift), torch.zeros(HIDDEN, device=x.device, dtype=torch.float32)
    check(lib.jpdvt_ln_gate_bwd(ptr(_need(x, torch.float32, "x")), ptr(_need(dxn, torch.float32, "dxn")),
                                ptr(_need(scale, torch.float32, "scale")), HIDDEN, ptr(dx), int(acc), ptr(dshift), ptr(dscale), HIDDEN,
                                ptr(dxb), ptr(y), ptr(_need(gate, torch.float32, "gate")) if gate is not None else None, HIDDEN,
                                ptr(dy), ptr(dgate), HIDDEN, ptr(dbias), batch, tokens, stream_ptr()), "ln_gate_bwd")
    return dx, dshift, dscale, dxb, dy, dgate, dbias


def colsum(src: torch.Tensor) -> torch.Tensor:
    lib = _lib_dev()
    rows, cols = src.shape
    out = torch.zeros(cols, device=src.device, dtype=torch.float32)
    fn = lib.jpdvt_colsum_bf16 if src.dtype == torch.bfloat16 else lib.jpdvt_colsum_f32
    check(fn(ptr(src.contiguous()), rows, cols, ptr(out), stream_ptr()), "colsum")
    return out


def _wrap_all():
    """Every public wrapper runs with its first CUDA tensor argument's device current (see _lib.on_tensor_device)."""
    import types
    for name, fn in list(globals().items()):
        if isinstance(fn, types.FunctionType) and not name.startswith("_") and fn.__module__ == __name__:
            globals()[name] = _lib.on_tensor_device(fn)


_wrap_all()

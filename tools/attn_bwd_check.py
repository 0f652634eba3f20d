#!/usr/bin/env python3
"""Developer check of the tcgen05 attention backward (T = 144): per-output error against fp32 autograd of torch SDPA and
against the mma.sync kernel (JPDVT_ATTN_BWD_LEGACY=1 in a second process), plus timing at the C3 shape (batch 128).

    python tools/attn_bwd_check.py [batch ...]
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jpdvt_mt_ntnu_b200 import ops  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


T = int(os.environ.get("T", "144"))
for B in [int(v) for v in sys.argv[1:]] or [1, 2, 13]:
    torch.manual_seed(B)
    qkv = (torch.randn(B * T, 2304, device="cuda") * 1.2).bfloat16()
    d_o = torch.randn(B * T, 768, device="cuda").bfloat16()
    x = qkv.float().requires_grad_(True)
    q, k, v = x.reshape(B, T, 3, 12, 64).permute(2, 0, 3, 1, 4)
    F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, 768).backward(d_o.float())
    o, lse = ops.attention(qkv, B, T, return_lse=True)
    dqkv = ops.attention_bwd(qkv, o, d_o, lse, B, T)
    torch.cuda.synchronize()
    g = x.grad
    print(f"B={B}: dQ {rel(dqkv[:, :768].float(), g[:, :768]):.3e}  dK {rel(dqkv[:, 768:1536].float(), g[:, 768:1536]):.3e}  "
          f"dV {rel(dqkv[:, 1536:].float(), g[:, 1536:]):.3e}")
    # where are the errors: per token-row block of one sample / head
    d = (dqkv.float() - g).reshape(B, T, 3, 12, 64)
    gg = g.reshape(B, T, 3, 12, 64)
    for part, name in enumerate(("dQ", "dK", "dV")):
        main = rel(d[:, :128, part] + gg[:, :128, part], gg[:, :128, part])
        tail = rel(d[:, 128:, part] + gg[:, 128:, part], gg[:, 128:, part])
        print(f"    {name}: rows 0..127 {main:.3e}   rows 128..{T - 1} {tail:.3e}")

B = int(os.environ.get("TIME_BATCH", "128"))
qkv = (torch.randn(B * T, 2304, device="cuda") * 1.2).bfloat16()
d_o = torch.randn(B * T, 768, device="cuda").bfloat16()
o, lse = ops.attention(qkv, B, T, return_lse=True)
for _ in range(3):
    ops.attention_bwd(qkv, o, d_o, lse, B, T)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    ops.attention_bwd(qkv, o, d_o, lse, B, T)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print(f"attention_bwd B={B} T={T}: {us:.1f} us per launch  ({5 * 2 * B * 12 * T * T * 64 / us / 1e6:.0f} TFLOP/s algorithmic)")

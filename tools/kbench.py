#!/usr/bin/env python3
"""Per-kernel A/B timing at the bench shape (C2: batch 256 x 144 tokens), CUDA events on the launching stream.

    python tools/kbench.py [--kernels a,b,...] [--iters 20] [--rounds 5] [--batch 256] [--size 192] [--once]

Kernels are timed round-robin (`rounds` passes of `iters` back-to-back launches each) so clock drift under the power
cap hits every candidate alike; the table reports the median and the minimum pass.  `--once` launches every selected
kernel exactly once after one warm-up launch (the command to wrap in ncu).
"""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jpdvt_mt_ntnu_b200 import ops                                # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kernels", default="")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--rounds", type=int, default=5)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--size", type=int, default=192)
ap.add_argument("--once", action="store_true")
ap.add_argument("--json", default="")
args = ap.parse_args()

T = (args.size // 16) ** 2
B = args.batch
M = B * T
dev = torch.device("cuda")
torch.manual_seed(0)
bf = lambda *s, k=1.0: (torch.randn(*s, device=dev) * k).bfloat16()
xn, att, hid, qkv = bf(M, 768), bf(M, 768), bf(M, 3072, k=0.1), bf(M, 2304)
x = torch.randn(M, 768, device=dev)
w_qkv, w_proj, w_fc1, w_fc2 = bf(2304, 768, k=0.02), bf(768, 768, k=0.02), bf(3072, 768, k=0.02), bf(768, 3072, k=0.02)
b_qkv, b_proj, b_fc1, b_fc2 = (torch.randn(n, device=dev) * 0.02 for n in (2304, 768, 3072, 768))
gate = torch.randn(1, 768, device=dev) * 0.01
shift, scale = torch.randn(1, 768, device=dev), torch.randn(1, 768, device=dev)

# folded-LayerNorm operands (csrc/fold.cu): one block's W', u, v and the row sums of x
_wq, _w1 = w_qkv[None].contiguous(), w_fc1[None].contiguous()
_mod = torch.randn(6 * 768 + 2 * 768, device=dev) * 0.3
wf, fu, fv = ops.fold_ln_weights(_wq, _w1, b_qkv[None].contiguous(), b_fc1[None].contiguous(), _mod)
wf_qkv, wf_fc1 = wf[0, :2304].contiguous(), wf[0, 2304:].contiguous()
xb = x.bfloat16()
stats = torch.zeros(M, 6, 2, device=dev)
stats[:, 0, 0], stats[:, 0, 1] = x.sum(1), (x * x).sum(1)
w12q = torch.randn(12, 2304, 768, device=dev).bfloat16() * 0.02
w12f = torch.randn(12, 3072, 768, device=dev).bfloat16() * 0.02
b12q, b12f, mod12 = torch.randn(12, 2304, device=dev), torch.randn(12, 3072, device=dev), torch.randn(12 * 4608 + 1536, device=dev)

GF, MB = 1e9, 1e6
KERNELS = {
    "qkv_fold": (lambda: ops.gemm_ln_folded(xb, stats, wf_qkv, fu[0, :2304], fv[0, :2304]), 2.0 * M * 768 * 2304, "flop"),
    "fc1_fold": (lambda: ops.gemm_ln_folded(xb, stats, wf_fc1, fu[0, 2304:], fv[0, 2304:], gelu=True), 2.0 * M * 768 * 3072, "flop"),
    "fc2_copy": (lambda: ops.gemm_bias_gate_residual_copy(x, hid, w_fc2, b_fc2, gate, T), 2.0 * M * 3072 * 768, "flop"),
    "proj_copy": (lambda: ops.gemm_bias_gate_residual_copy(x, att, w_proj, b_proj, gate, T), 2.0 * M * 768 * 768, "flop"),
    "fold_w": (lambda: ops.fold_ln_weights(w12q, w12f, b12q, b12f, mod12), 12 * 5376 * 768 * 4.0, "byte"),
    "qkv": (lambda: ops.gemm_bias(xn, w_qkv, b_qkv), 2.0 * M * 768 * 2304, "flop"),
    "fc1": (lambda: ops.gemm_bias_gelu(xn, w_fc1, b_fc1), 2.0 * M * 768 * 3072, "flop"),
    "fc2_bf16": (lambda: ops.gemm_bias_gate(hid, w_fc2, b_fc2, gate, T), 2.0 * M * 3072 * 768, "flop"),
    "proj_bf16": (lambda: ops.gemm_bias_gate(att, w_proj, b_proj, gate, T), 2.0 * M * 768 * 768, "flop"),
    "fc2_resid": (lambda: ops.gemm_bias_gate_residual(x, hid, w_fc2, b_fc2, gate, T), 2.0 * M * 3072 * 768, "flop"),
    "proj_resid": (lambda: ops.gemm_bias_gate_residual(x, att, w_proj, b_proj, gate, T), 2.0 * M * 768 * 768, "flop"),
    "fc2_resid_ln": (lambda: ops.gemm_bias_gate_residual_ln(x, hid, w_fc2, b_fc2, gate, shift, scale, T), 2.0 * M * 3072 * 768, "flop"),
    "proj_resid_ln": (lambda: ops.gemm_bias_gate_residual_ln(x, att, w_proj, b_proj, gate, shift, scale, T), 2.0 * M * 768 * 768, "flop"),
    "attn": (lambda: ops.attention(qkv, B, T), 4.0 * B * 12 * T * T * 64, "flop"),
    "ln": (lambda: ops.ln_modulate(x, shift, scale, T), M * 768 * 6.0, "byte"),
    "ln_res": (lambda: ops.ln_modulate(x, shift, scale, T, delta=xn), M * 768 * 12.0, "byte"),
}
names = [k for k in args.kernels.split(",") if k] or list(KERNELS)
for k in names:
    if k not in KERNELS:
        sys.exit(f"unknown kernel {k}; have {sorted(KERNELS)}")

if args.once:
    for k in names:
        KERNELS[k][0]()
    torch.cuda.synchronize()
    for k in names:
        KERNELS[k][0]()
    torch.cuda.synchronize()
    print("ok")
    sys.exit(0)

for k in names:
    for _ in range(3):
        KERNELS[k][0]()
torch.cuda.synchronize()
times = {k: [] for k in names}
for _ in range(args.rounds):
    for k in names:
        fn = KERNELS[k][0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        times[k].append(e0.elapsed_time(e1) / args.iters)
out = {}
for k in names:
    med, mn = statistics.median(times[k]), min(times[k])
    work, kind = KERNELS[k][1], KERNELS[k][2]
    rate = work / (med * 1e-3) / (1e12 if kind == "flop" else 1e9)
    out[k] = {"ms_median": med, "ms_min": mn, "rate": rate, "unit": "TFLOP/s" if kind == "flop" else "GB/s"}
    print(f"{k:12s} median {med * 1e3:8.1f} us   min {mn * 1e3:8.1f} us   {rate:8.1f} {out[k]['unit']}")
if args.json:
    json.dump(out, open(args.json, "w"), indent=1)

#!/usr/bin/env python3
"""GPU diagnostics for the training kernels, each group in its own subprocess (see tools/gpu_check.py)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), (a - b).abs().max().item()


def g_wgrad():
    import torch
    from jpdvt_mt_ntnu_b200 import ops
    torch.manual_seed(0)
    res = {}
    for (m, r, c) in [(64, 256, 256), (128, 256, 256), (1000, 768, 768), (4608, 2304, 768), (4608, 768, 3072), (432, 64, 768),
                      (3, 768, 256), (27, 1536, 768), (300, 128, 128)]:
        p = torch.randn(m, r, device="cuda").bfloat16()
        q = torch.randn(m, c, device="cuda").bfloat16()
        ref = p.float().t() @ q.float()
        got = ops.gemm_wgrad(p, q)
        torch.cuda.synchronize()
        res[f"wgrad m={m} {r}x{c}"] = rel(got, ref)
        if res[f"wgrad m={m} {r}x{c}"][0] > 1e-3:
            res[f"  diag m={m} {r}x{c}"] = {"got": got[:2, :4].tolist(), "ref": ref[:2, :4].tolist(),
                                            "got_T_match": rel(got, (q.float().t() @ p.float()).t() if r == c else ref)[0]}
    return res


def g_elem():
    import torch
    import torch.nn.functional as F
    from jpdvt_mt_ntnu_b200 import ops
    torch.manual_seed(1)
    res = {}
    m, T = 432, 144
    a = torch.randn(m, 768, device="cuda").bfloat16()
    w = (torch.randn(3072, 768, device="cuda") * 0.05).bfloat16()
    pre = torch.randn(m, 3072, device="cuda").bfloat16()
    x = pre.float().requires_grad_(True)
    F.gelu(x, approximate="tanh").sum().backward()
    res["dgelu"] = rel(ops.gemm_dgelu(a, w, x.grad.bfloat16()).float(), (a.float() @ w.float().t()) * x.grad.bfloat16().float())
    dx = torch.randn(m, 768, device="cuda")
    y = torch.randn(m, 768, device="cuda").bfloat16()
    gate = torch.randn(3, 768, device="cuda")
    dy, dgate, dbias = ops.gate_bwd(dx, y, gate, T)
    gfull = gate.repeat_interleave(T, 0)
    res["gate_bwd dy"] = rel(dy.float(), gfull * dx)
    res["gate_bwd dgate"] = rel(dgate, (dx * y.float()).reshape(3, T, 768).sum(1))
    res["gate_bwd dbias"] = rel(dbias, (gfull * dx).sum(0))
    xx = (torch.randn(m, 768, device="cuda") * 2 + 0.3).requires_grad_(True)
    shift = torch.randn(3, 768, device="cuda", requires_grad=True)
    scale = (torch.randn(3, 768, device="cuda") * 0.5).requires_grad_(True)
    idx = torch.arange(m, device="cuda") // T
    out = F.layer_norm(xx, (768,), eps=1e-6) * (1 + scale[idx]) + shift[idx]
    dxn = torch.randn(m, 768, device="cuda")
    out.backward(dxn)
    base = torch.randn(m, 768, device="cuda")
    got_dx, dsh, dsc, dxb = ops.ln_modulate_bwd(xx.detach(), dxn, scale.detach(), T, dx=base.clone())
    res["ln_bwd dx(acc)"] = rel(got_dx, base + xx.grad)
    res["ln_bwd dshift"] = rel(dsh, shift.grad)
    res["ln_bwd dscale"] = rel(dsc, scale.grad)
    res["ln_bwd dx_bf16"] = rel(dxb.float(), base + xx.grad)
    got_dx2, _, _, _ = ops.ln_modulate_bwd(xx.detach(), dxn, scale.detach(), T)
    res["ln_bwd dx(no acc)"] = rel(got_dx2, xx.grad)
    res["colsum bf16"] = rel(ops.colsum(pre), pre.float().sum(0))
    res["colsum f32"] = rel(ops.colsum(dx), dx.sum(0))
    return res


def g_attn():
    import torch
    import torch.nn.functional as F
    from jpdvt_mt_ntnu_b200 import ops
    torch.manual_seed(2)
    res = {}
    for B, T in ((2, 144), (3, 9), (2, 256), (1, 324), (2, 100), (1, 36)):
        qkv = (torch.randn(B * T, 2304, device="cuda") * 1.2).bfloat16()
        d_o = torch.randn(B * T, 768, device="cuda").bfloat16()
        x = qkv.float().requires_grad_(True)
        q, k, v = x.reshape(B, T, 3, 12, 64).permute(2, 0, 3, 1, 4)
        ref_o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, 768)
        ref_o.backward(d_o.float())
        o, lse = ops.attention(qkv, B, T, return_lse=True)
        s = (q @ k.transpose(-1, -2)) * 0.125
        res[f"lse B={B} T={T}"] = rel(lse, torch.logsumexp(s, -1).detach() * 1.4426950408889634)
        dqkv = ops.attention_bwd(qkv, o, d_o, lse, B, T)
        g = x.grad
        res[f"dq B={B} T={T}"] = rel(dqkv[:, :768].float(), g[:, :768])
        res[f"dk B={B} T={T}"] = rel(dqkv[:, 768:1536].float(), g[:, 768:1536])
        res[f"dv B={B} T={T}"] = rel(dqkv[:, 1536:].float(), g[:, 1536:])
    return res


GROUPS = {"wgrad": g_wgrad, "elem": g_elem, "attn": g_attn}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        print("RESULT " + json.dumps(GROUPS[sys.argv[2]]()))
        sys.exit(0)
    for name in (sys.argv[1:] or list(GROUPS)):
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", name], capture_output=True, text=True,
                               timeout=float(os.environ.get("CHECK_TIMEOUT", "180")))
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            out = json.loads(lines[-1][7:]) if lines else {"error": r.stdout[-1500:] + "\n" + r.stderr[-3000:]}
        except subprocess.TimeoutExpired:
            out = {"error": "TIMEOUT"}
        print(f"=== {name} ({time.time() - t0:.1f}s)")
        for k, v in out.items():
            print(f"  {k}: {v}")
        sys.stdout.flush()

#!/usr/bin/env python3
"""First-contact GPU diagnostics: runs every kernel of libjpdvt_sm100.so against a torch fp32 restatement, each group in
its own subprocess with a timeout (a hung or faulting kernel cannot take the others down).

    python tools/gpu_check.py [group ...]      # groups: gemm elementwise attention assign forward sampling
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    import torch
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), (a - b).abs().max().item()


def g_gemm():
    import torch
    from jpdvt_mt_ntnu_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    res = {}
    for (m, n, k) in [(128, 256, 64), (128, 256, 768), (256, 768, 768), (432, 2304, 768), (1000, 768, 3072), (27, 3072, 768),
                      (4608, 2304, 768), (300, 128, 128)]:
        a = torch.randn(m, k, device=dev).bfloat16()
        w = (torch.randn(n, k, device=dev) * 0.05).bfloat16()
        bias = torch.randn(n, device=dev)
        ref = a.float() @ w.float().t() + bias
        out = ops.gemm_bias(a, w, bias)
        torch.cuda.synchronize()
        res[f"bias_bf16 {m}x{n}x{k}"] = rel(out.float(), ref)
        if res[f"bias_bf16 {m}x{n}x{k}"][0] > 1e-2:
            d = (out.float() - ref).abs()
            bad_rows = (d.max(dim=1).values > 0.1).nonzero().flatten()[:8].tolist()
            bad_cols = (d.max(dim=0).values > 0.1).nonzero().flatten()[:8].tolist()
            res[f"  diag {m}x{n}x{k}"] = {"bad_rows": bad_rows, "bad_cols": bad_cols, "out00": out[0, :4].float().tolist(),
                                          "ref00": ref[0, :4].tolist()}
    m, n, k, T = 432, 768, 768, 144
    a = torch.randn(m, k, device=dev).bfloat16(); w = (torch.randn(n, k, device=dev) * 0.05).bfloat16(); bias = torch.randn(n, device=dev)
    res["bias_f32"] = rel(ops.gemm_bias_f32(a, w, bias), a.float() @ w.float().t() + bias)
    o16, o32 = ops.gemm_bias(a, w, bias, want_f32_copy=True)
    res["bias_bf16+f32 copy"] = rel(o32, a.float() @ w.float().t() + bias)
    w4 = (torch.randn(3072, k, device=dev) * 0.05).bfloat16(); b4 = torch.randn(3072, device=dev)
    res["bias_gelu"] = rel(ops.gemm_bias_gelu(a, w4, b4).float(),
                           torch.nn.functional.gelu(a.float() @ w4.float().t() + b4, approximate="tanh"))
    for ncond in (3, 1):
        gate = torch.randn(ncond, n, device=dev)
        g = gate.repeat_interleave(T, 0)[:m] if ncond > 1 else gate
        res[f"gate_bf16 ncond={ncond}"] = rel(ops.gemm_bias_gate(a, w, bias, gate, T).float(), g * (a.float() @ w.float().t() + bias))
    img = torch.rand(3, 3, 192, 192, device=dev) * 2 - 1
    cols = ops.patchify(img)
    ref_cols = img.reshape(3, 3, 12, 16, 12, 16).permute(0, 2, 4, 1, 3, 5).reshape(432, 768)
    res["patchify"] = rel(cols.float(), ref_cols.bfloat16().float())
    wp = (torch.randn(768, 768, device=dev) * 0.05).bfloat16(); xt = torch.randn(432, 8, device=dev)
    w_in_t = torch.randn(8, 768, device=dev) * 0.1; pos = torch.randn(144, 768, device=dev)
    ref = cols.float() @ wp.float().t() + bias + pos.repeat(3, 1) + xt @ w_in_t
    res["patch_embed"] = rel(ops.gemm_patch_embed(cols, wp, bias, xt, w_in_t, pos, 144), ref)
    y = torch.randn(432, 768, device=dev).bfloat16(); w1 = (torch.randn(64, 768, device=dev) * 0.05).bfloat16()
    b1 = torch.randn(64, device=dev) * 0.1; w2 = torch.randn(8, 64, device=dev) * 0.2; b2 = torch.randn(8, device=dev)
    ref = torch.nn.functional.silu(y.float() @ w1.float().t() + b1) @ w2.t() + b2
    res["final_head"] = rel(ops.final_head(y, w1, b1, w2, b2), ref)
    y32 = torch.randn(432, 768, device=dev)
    ref = y32.reshape(3, 12, 12, 16, 16, 3).permute(0, 5, 1, 3, 2, 4).reshape(3, 3, 192, 192)
    res["unpatchify"] = rel(ops.unpatchify(y32, 3, 192), ref)
    return res


def g_elementwise():
    import torch
    from jpdvt_mt_ntnu_b200 import ops
    from jpdvt_mt_ntnu_b200.diffusion import create_diffusion
    torch.manual_seed(1)
    dev = "cuda"
    res = {}
    for rows, T, ncond in ((432, 144, 3), (432, 144, 1), (27, 9, 3), (1001, 143, 7)):
        x = torch.randn(rows, 768, device=dev) * 2 + 0.3
        sh, sc = torch.randn(ncond, 768, device=dev), torch.randn(ncond, 768, device=dev) * 0.5
        ln = torch.nn.functional.layer_norm(x, (768,), eps=1e-6)
        idx = (torch.arange(rows, device=dev) // T).clamp_max(ncond - 1) if ncond > 1 else torch.zeros(rows, dtype=torch.long, device=dev)
        ref = ln * (1 + sc[idx]) + sh[idx]
        res[f"ln_modulate {rows}/{T}/{ncond}"] = rel(ops.ln_modulate(x, sh, sc, T).float(), ref)
        delta = torch.randn(rows, 768, device=dev).bfloat16()
        x2 = x.clone()
        y2 = ops.ln_modulate(x2, sh, sc, T, delta=delta)
        xr = x + delta.float()
        res[f"ln_modulate+delta {rows}/{T}/{ncond}"] = (rel(y2.float(), torch.nn.functional.layer_norm(xr, (768,), eps=1e-6) * (1 + sc[idx]) + sh[idx]),
                                                      bool(torch.equal(x2, xr)))
    t = torch.tensor([0, 1, 5, 250, 999, 37, 512, 4, 8, 991, 995], device=dev)
    w0, b0 = torch.randn(768, 256, device=dev) * 0.02, torch.randn(768, device=dev) * 0.02
    w2, b2 = torch.randn(768, 768, device=dev) * 0.02, torch.randn(768, device=dev) * 0.02
    half = 128
    freqs = torch.exp(-torch.log(torch.tensor(10000.0)) * torch.arange(half, dtype=torch.float32) / half).to(dev)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([args.cos(), args.sin()], -1)
    cref = torch.nn.functional.silu(emb @ w0.t() + b0) @ w2.t() + b2
    c, sc_ = ops.timestep_embed(t, w0, b0, w2, b2)
    res["timestep_embed c"] = rel(c, cref)
    res["timestep_embed silu"] = rel(sc_, torch.nn.functional.silu(cref))
    wall = (torch.randn(12 * 4608 + 1536, 768, device=dev) * 0.05).bfloat16(); ball = torch.randn(12 * 4608 + 1536, device=dev)
    for r in (1, 3, 8, 11):
        s = torch.randn(r, 768, device=dev)
        res[f"adaln_table rows={r}"] = rel(ops.adaln_table(s, wall, ball), s @ wall.float().t() + ball)
    d = create_diffusion("250")
    tabs = d.device_tables(torch.device(dev))
    x0, xt, nz = torch.randn(4, 144, 8, device=dev), torch.randn(4, 144, 8, device=dev), torch.randn(4, 144, 8, device=dev)
    tt = torch.tensor([0, 1, 100, 249], device=dev)
    mean, sample = ops.posterior_step(x0, xt, nz, tabs["coef1"], tabs["coef2"], tabs["logvar"], tt)
    import numpy as np
    g = lambda a: torch.from_numpy(a).to(dev)[tt].float().view(-1, 1, 1)
    mref = g(d.posterior_mean_coef1) * x0 + g(d.posterior_mean_coef2) * xt
    sref = mref + (tt != 0).float().view(-1, 1, 1) * torch.exp(0.5 * g(d.posterior_log_variance_clipped)) * nz
    res["posterior mean (max abs)"] = rel(mean, mref)
    res["posterior sample"] = rel(sample, sref)
    res["posterior mean bit-exact"] = bool(torch.equal(mean, mref))
    qs = ops.q_sample(x0, nz, tabs["sqrt_ac"], tabs["sqrt_1mac"], tt)
    qref = g(d.sqrt_alphas_cumprod) * x0 + g(d.sqrt_one_minus_alphas_cumprod) * nz
    res["q_sample"] = rel(qs, qref)
    res["q_sample bit-exact"] = bool(torch.equal(qs, qref))
    return res


def g_attention():
    import torch
    from jpdvt_mt_ntnu_b200 import ops
    torch.manual_seed(2)
    dev = "cuda"
    res = {}
    for B, T in ((2, 144), (3, 9), (2, 256), (2, 324), (1, 36), (5, 64), (2, 100)):
        qkv = (torch.randn(B * T, 2304, device=dev) * 1.5).bfloat16()
        q, k, v = qkv.float().reshape(B, T, 3, 12, 64).permute(2, 0, 3, 1, 4)
        ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, 768)
        out = ops.attention(qkv, B, T)
        res[f"attention B={B} T={T}"] = rel(out.float(), ref)
    return res


def g_assign():
    import numpy as np
    import torch
    from jpdvt_mt_ntnu_b200 import ops
    from oracle import jpdvt_oracle as orc
    g = np.load(os.path.join(ROOT, "tests", "golden", "assignment.npz"))
    res = {}
    bad = 0
    for tag, sentinel in (("1e9", 1e9), ("2024", 2024.0)):
        for i in range(int(g["n"])):
            sc = torch.from_numpy(g[f"scores_{i}"]).cuda().unsqueeze(0)
            order, pred = ops.assign_from_scores(sc, sentinel)
            ok = np.array_equal(order[0].cpu().numpy(), g[f"order_{tag}_{i}"])
            order_is_perm = sorted(g[f"order_{tag}_{i}"].tolist()) == list(range(sc.shape[1]))
            okp = np.array_equal(pred[0].cpu().numpy(), g[f"pred_{tag}_{i}"]) or not order_is_perm
            if not (ok and okp):
                bad += 1
                res[f"mismatch {tag} {i}"] = {"got": order[0].tolist(), "want": g[f"order_{tag}_{i}"].tolist(),
                                              "gotp": pred[0].tolist(), "wantp": g[f"pred_{tag}_{i}"].tolist()}
    res["from_scores mismatches"] = bad
    torch.manual_seed(3)
    for G, tok in ((3, 4), (4, 4), (3, 6), (4, 3), (5, 2)):
        lat = torch.randn(64, G * G * tok * tok, 8)
        order, pred, scores = ops.assign_greedy_l1(lat.cuda(), torch.from_numpy(orc.sincos_2d(8, G)).float().cuda(), G, 1e9, True)
        nbad, smax = 0, 0.0
        for b in range(64):
            o, p, sc = orc.solve(lat[b], G, tok)
            smax = max(smax, float(np.abs(scores[b].cpu().numpy() - sc).max()))
            nbad += int(list(order[b].cpu().numpy()) != list(o)) + int(list(pred[b].cpu().numpy()) != list(p))
        res[f"greedy_l1 G={G} tok={tok}"] = {"mismatches": nbad, "score_maxdiff": smax}
    return res


def g_forward():
    import numpy as np
    import torch
    from jpdvt_mt_ntnu_b200.models import DiT
    from oracle import cases, jpdvt_oracle as orc
    res = {}
    for name in ("tiny48", "d2_192", "hot192", "d2_288", "d2_256", "full192"):
        case = cases.FORWARD_CASES[name]
        st = cases.state_for(case)
        m = DiT(input_size=case["size"], depth=case["depth"], hidden_size=768, patch_size=16, num_heads=12)
        m.load_state_dict(st); m.cuda()
        img, t, x_t = cases.forward_inputs(case)
        with torch.no_grad():
            o_img, o_te = m(img.cuda(), t.cuda(), x_t.cuda())
        torch.cuda.synchronize()
        gold = np.load(os.path.join(ROOT, "tests", "golden", f"forward_{name}.npz"))
        res[f"{name} te vs golden"] = rel(o_te.cpu(), torch.from_numpy(gold["te"]))
        res[f"{name} img vs golden(sample)"] = rel(o_img.cpu()[:, :, ::7, ::5], torch.from_numpy(gold["img_sample"]))
        r_img, r_te = orc.OracleDenoiser(st, depth=case["depth"])(img, t, x_t)
        res[f"{name} te vs oracle"] = rel(o_te.cpu(), r_te)
        res[f"{name} img vs oracle"] = rel(o_img.cpu(), r_img)
    return res


def g_sampling():
    import numpy as np
    import torch
    from jpdvt_mt_ntnu_b200.models import DiT
    from jpdvt_mt_ntnu_b200.diffusion import create_diffusion
    from jpdvt_mt_ntnu_b200 import assignment
    from oracle import cases
    res = {}
    for name in ("tiny48_s10", "d2_256g4_s25", "d2_192_s250", "full192_s250"):
        case = cases.SAMPLING_CASES[name]
        st = cases.state_for(case)
        m = DiT(input_size=case["size"], depth=case["depth"], hidden_size=768, patch_size=16, num_heads=12)
        m.load_state_dict(st); m.cuda()
        d = create_diffusion(case["respacing"])
        cond, noise = cases.sampling_inputs(case)
        torch.manual_seed(case["loop_seed"])
        step_noise = torch.stack([torch.randn_like(noise) for _ in range(d.num_timesteps)]).cuda()
        gold = np.load(os.path.join(ROOT, "tests", "golden", f"sampling_{name}.npz"))
        t0 = time.time()
        outs = list(d.p_sample_loop_progressive(m.forward, cond.cuda(), noise.shape, noise.cuda(), clip_denoised=False,
                                                step_noise=step_noise))
        torch.cuda.synchronize()
        for n in cases.kept_steps(len(outs)):
            res[f"{name} step{n} sample"] = rel(outs[n]["sample"].cpu(), torch.from_numpy(gold[f"step{n}_sample"]))
            res[f"{name} step{n} x0"] = rel(outs[n]["pred_xstart"].cpu(), torch.from_numpy(gold[f"step{n}_x0"]))
        final = d.p_sample_loop(m.forward, cond.cuda(), noise.shape, noise.cuda(), clip_denoised=False, step_noise=step_noise)
        res[f"{name} final"] = rel(final.cpu(), torch.from_numpy(gold["final"]))
        order, pred, scores = assignment.solve_puzzles(final, case["grid"], return_scores=True)
        res[f"{name} order"] = {"got": order.cpu().tolist(), "want": gold["order"].tolist()}
        res[f"{name} dist maxdiff"] = float(np.abs(scores.cpu().numpy() - gold["dist"]).max())
        res[f"{name} seconds"] = time.time() - t0
    return res


GROUPS = {"gemm": g_gemm, "elementwise": g_elementwise, "attention": g_attention, "assign": g_assign, "forward": g_forward,
          "sampling": g_sampling}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        out = GROUPS[sys.argv[2]]()
        print("RESULT " + json.dumps(out))
        sys.exit(0)
    names = sys.argv[1:] or list(GROUPS)
    summary = {}
    for name in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", name], capture_output=True, text=True,
                               timeout=float(os.environ.get("CHECK_TIMEOUT", "240")))
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if lines:
                summary[name] = json.loads(lines[-1][7:])
            else:
                summary[name] = {"error": (r.stdout[-1500:] + "\n" + r.stderr[-3000:])}
        except subprocess.TimeoutExpired as e:
            summary[name] = {"error": "TIMEOUT", "stderr": (e.stderr or b"")[-2000:].decode("utf8", "replace") if isinstance(e.stderr, bytes) else str(e.stderr)[-2000:]}
        print(f"=== {name} ({time.time() - t0:.1f}s)")
        for k, v in summary[name].items():
            print(f"  {k}: {v}")
        sys.stdout.flush()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(summary, open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w"), indent=1)

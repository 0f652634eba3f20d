#!/usr/bin/env python3
"""Headline benchmark of the JPDVT hot path on B200: puzzles/s for 250-step p_sample_loop sampling + assignment.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4|c5] [--batch B]

One "step" = one pass of the hot path over one batch of synthetic puzzles: all 250 denoiser forwards of
SpacedDiffusion.p_sample_loop (the reference's loop quirk preserved, none of the "dead" steps skipped), the fused
posterior update after each, and the position-to-grid assignment of every puzzle.  Workload (BASELINE.json configs[1]):
JPDVT 3x3 @192 px, batch 256 PER GPU (weak scaling: puzzles are independent, reference shards image_paths[rank::world]).

Prints ONE JSON line (rank 0).  `value` = whole-job puzzles/s with inputs resident in HBM; `e2e` = the same through the
public API from pinned host buffers (H2D of the scrambled images + noise, D2H of the placements inside the timed
region); `roofline` = the dominant kernel (fc1 tcgen05 GEMM) against the measured bf16 peak; `cpu_baseline` = the CPU
oracle port (torch fp32, all host threads) on a bounded sample.  `--impl reference` times that CPU port as its own line.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c2": dict(name="JPDVT 3x3 @192px sampling, 250 steps, batch 256/GPU", size=192, grid=3, batch=256, steps=250),
    "c4": dict(name="JPDVT 4x4 @256px sampling, 250 steps, batch 128/GPU", size=256, grid=4, batch=128, steps=250),
    "c5": dict(name="JPDVT 3x3 @288px masked sampling, 250 steps, batch 128/GPU", size=288, grid=3, batch=128, steps=250),
    # BASELINE.json configs[2]: data-parallel training (fwd + bwd + NCCL gradient all-reduce + AdamW + EMA), metric train img/s
    "c3": dict(name="train_JPDVT 3x3 @192px bf16 data-parallel training, batch 128/GPU", size=192, grid=3, batch=128, steps=0, train=True),
    "c4t": dict(name="train_JPDVT 4x4 @256px bf16 data-parallel training, batch 64/GPU", size=256, grid=4, batch=64, steps=0, train=True),
}
DEPTH = 12


def flops_per_forward(T: int) -> float:
    """algorithmic FLOPs (2*MAC) of one denoiser forward per sample."""
    D = 768
    return (2 * T * D * D + 2 * T * 8 * D + 2 * (256 * D + D * D)
            + 12 * (2 * D * 6 * D + 2 * T * D * 3 * D + 4 * T * T * D + 2 * T * D * D + 2 * (2 * T * D * 4 * D))
            + 2 * D * 2 * D + 2 * T * D * D + 2 * T * D * 64 + 2 * T * 64 * 8)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm_gbs=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


def synthetic_inputs(wl, batch, seed=0):
    """SURVEY.md 8(d) recipe: images rand*2-1 (manual_seed), per-puzzle np permutations, one randn(1,T,8) noise row
    repeated over the batch (inferencetexmet.py:313); C5 additionally zeroes 1-2 random slots of the scrambled image."""
    import numpy as np
    import torch
    G, S = wl["grid"], wl["size"]
    T = (S // 16) ** 2
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(batch, 3, S, S, generator=g) * 2 - 1
    rs = np.random.RandomState(seed)
    perms = np.stack([rs.permutation(G * G) for _ in range(batch)])
    p = S // G
    pieces = img.reshape(batch, 3, G, p, G, p).permute(0, 1, 2, 4, 3, 5).reshape(batch, 3, G * G, p, p)
    idx = torch.from_numpy(perms).long()
    pieces = torch.gather(pieces, 2, idx[:, None, :, None, None].expand(batch, 3, G * G, p, p))
    if wl is WORKLOADS["c5"]:
        for b in range(batch):
            for slot in rs.choice(G * G, size=rs.randint(1, 3), replace=False):
                pieces[b, :, slot] = 0
    cond = pieces.reshape(batch, 3, G, G, p, p).permute(0, 1, 2, 4, 3, 5).reshape(batch, 3, S, S).contiguous()
    noise = torch.randn(1, T, 8, generator=g).repeat(batch, 1, 1)
    return cond, noise, perms


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples taken DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ---------------------------------------------------------------------------------------------------- CPU arm
_CPU_CACHE = {}


def sampling_config(wl, batch):
    """`config` of the sampling line - one dict for both arms, so the driver's same-config check compares like with like."""
    S = wl["size"]
    T = (S // 16) ** 2
    return {"workload": wl["name"], "batch_per_gpu": batch, "diffusion_steps": wl["steps"], "grid": wl["grid"],
            "image_size": S, "tokens": T, "weights": "random N(0,0.02), seed 1234 (fresh init outputs zeros)",
            "l2": f"activations per launch (>= {batch * T * 768 * 2 / 1e6:.0f} MB bf16 / {batch * T * 768 * 4 / 1e6:.0f} MB fp32) exceed "
                  "or rival L2; no flush needed",
            "denoiser_forwards_per_step": wl["steps"], "sharding": "independent puzzle batches per rank, no data-path collective",
            "precision": "ours: bf16 tensor-core operands, fp32 accumulation and residual stream; reference arm: fp32 on the host cores"}


def cpu_port_puzzles_per_s(wl, sample_batch=16, sample_steps=2):
    """The oracle port (torch fp32 CPU restatement of the reference path) on a BOUNDED sample: `sample_batch` puzzles x
    `sample_steps` of the 250 diffusion steps (every step is identical work - the reference feeds the same
    (condition, noise) to all of them) + the assignment, scaled to the full step count."""
    import torch
    from oracle import jpdvt_oracle as orc          # allowed here: bench.py's cpu_baseline / reference legs only
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    S, G = wl["size"], wl["grid"]
    T = (S // 16) ** 2
    if S not in _CPU_CACHE:
        _CPU_CACHE[S] = orc.OracleDenoiser(orc.seeded_state(orc.blank_state(S, DEPTH), seed=1234), depth=DEPTH)
    model = _CPU_CACHE[S]
    sched = orc.Schedule(str(wl["steps"]))
    cond, noise, _ = synthetic_inputs(wl, sample_batch)
    with torch.no_grad():
        t = torch.full((sample_batch,), sched.num_timesteps - 1, dtype=torch.long)
        sched.p_step(model, cond, noise, t, torch.randn_like(noise))          # warm-up (thread pools, allocator)
        t0 = time.perf_counter()
        out = None
        for k in range(sample_steps):
            t = torch.full((sample_batch,), sched.num_timesteps - 1 - k, dtype=torch.long)
            out = sched.p_step(model, cond, noise, t, torch.randn_like(noise))
        t_steps = time.perf_counter() - t0
        t0 = time.perf_counter()
        for b in range(sample_batch):
            orc.solve(out["sample"][b], G, S // (16 * G))
        t_assign = time.perf_counter() - t0
    total = t_steps / sample_steps * wl["steps"] + t_assign
    return sample_batch / total, cores, f"{sample_batch} puzzles x {sample_steps} of {wl['steps']} diffusion steps + assignment, scaled x{wl['steps']}/{sample_steps}", t_steps + t_assign


REF_STAGED = os.path.join(ROOT, "baseline", "_ref", "image_model")
_REF_MODULES = {}


def staged_reference():
    """The UNMODIFIED reference modules (image_model/models.py, diffusion/, inference.py) from the git-ignored staging
    directory baseline/_ref/ (tools/stage_reference.sh copies them there in the build container; the directory travels to the
    GPU box with the snapshot).  The two packages they import that this image lacks (timm, matplotlib) come from the stand-ins
    the golden generator uses (oracle/standins).  None when the staging directory is absent or does not import."""
    if "mods" in _REF_MODULES:
        return _REF_MODULES["mods"]
    mods = None
    if os.path.isfile(os.path.join(REF_STAGED, "models.py")):
        saved = list(sys.path)
        sys.path[:0] = [os.path.join(ROOT, "oracle", "standins"), REF_STAGED]
        try:
            import importlib
            mods = tuple(importlib.import_module(n) for n in ("models", "diffusion", "inference"))
        except Exception as e:  # noqa: BLE001  (a missing third-party import on this box: the port is the fallback)
            print(f"[bench] staged reference not importable ({type(e).__name__}: {e}); using the oracle port", file=sys.stderr)
            mods = None
        finally:
            sys.path[:] = saved
    _REF_MODULES["mods"] = mods
    return mods


def cpu_reference_puzzles_per_s(wl, sample_batch=16, sample_steps=2):
    """The reference's own CPU path on a BOUNDED sample, through its stock calls: `DiT_models["JPDVT"]`, `create_diffusion`,
    `p_sample_loop_progressive` (stopped after `sample_steps` of the 250 steps - every step is identical work) and the
    assignment snippet of inference.py:294-306 (`rearrange`, sklearn `pairwise_distances`, `find_permutation`), scaled to
    the full step count.  Falls back to the oracle port when baseline/_ref is not staged.  -> (..., kind)"""
    mods = staged_reference()
    if mods is None:
        return cpu_port_puzzles_per_s(wl, sample_batch, sample_steps) + ("port",)
    import itertools
    import torch
    from einops import rearrange
    from sklearn.metrics import pairwise_distances
    from oracle import jpdvt_oracle as orc          # seeded weights only (the same state the GPU arm loads)
    ref_models, ref_diffusion, ref_inference = mods
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    S, G = wl["size"], wl["grid"]
    tok = S // (16 * G)
    key = ("ref", S)
    if key not in _CPU_CACHE:
        m = ref_models.DiT_models["JPDVT"](input_size=S)
        m.load_state_dict(orc.seeded_state(m.state_dict(), seed=1234))
        m.train()                                    # inference.py:213-214 (a no-op: no dropout / batch norm)
        _CPU_CACHE[key] = m
    model = _CPU_CACHE[key]
    diffusion = ref_diffusion.create_diffusion(str(wl["steps"]))
    cond, noise, _ = synthetic_inputs(wl, sample_batch)

    def steps(n):
        gen = diffusion.p_sample_loop_progressive(model.forward, cond, noise.shape, noise, clip_denoised=False, model_kwargs=None,
                                                  device="cpu", progress=False)
        return list(itertools.islice(gen, n))[-1]
    with torch.no_grad():
        steps(1)                                     # warm-up (thread pools, allocator)
        t0 = time.perf_counter()
        out = steps(sample_steps)
        t_steps = time.perf_counter() - t0
        canon = torch.tensor(ref_models.get_2d_sincos_pos_embed(8, G)).unsqueeze(0).float()
        t0 = time.perf_counter()
        for b in range(sample_batch):
            lat = rearrange(out["sample"][b], "(p1 h1 p2 w1) d -> (p1 p2) (h1 w1) d", p1=G, p2=G, h1=tok, w1=tok).mean(1)
            dist = pairwise_distances(lat.cpu().numpy(), canon[0].cpu().numpy(), metric="manhattan")
            ref_inference.find_permutation(dist)
        t_assign = time.perf_counter() - t0
    total = t_steps / sample_steps * wl["steps"] + t_assign
    sample = (f"unmodified reference modules (baseline/_ref, timm stand-in): {sample_batch} puzzles x {sample_steps} of {wl['steps']} "
              f"diffusion steps + assignment, scaled x{wl['steps']}/{sample_steps}")
    return sample_batch / total, cores, sample, t_steps + t_assign, "reference"


def stock_torch_gpu_puzzles_per_s(wl, batch, precision="fp32", sample_steps=5):
    """SURVEY.md 8(d) "same box" comparator: the oracle's stock torch ops (what the reference's fp32 nn.Modules execute)
    moved to cuda:0 - cuBLAS / ATen kernels, none of this repo's - on the bench batch, `sample_steps` of the 250 steps
    scaled up + the host assignment loop.  precision: fp32 (inference.py), tf32 (train_JPDVT.py:5-6), bf16 (autocast)."""
    import torch
    from oracle import jpdvt_oracle as orc
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = precision == "tf32"
    torch.backends.cudnn.allow_tf32 = precision == "tf32"
    S, G = wl["size"], wl["grid"]
    model = orc.OracleDenoiser(orc.seeded_state(orc.blank_state(S, DEPTH), seed=1234), depth=DEPTH, device=dev)
    sched = orc.Schedule(str(wl["steps"]))
    cond, noise, _ = synthetic_inputs(wl, batch)
    cond, noise = cond.to(dev), noise.to(dev)
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if precision == "bf16" else contextlib.nullcontext()

    def step(k):
        t = torch.full((batch,), sched.num_timesteps - 1 - k, dtype=torch.long, device=dev)
        return sched.p_step(model, cond, noise, t, torch.randn_like(noise))

    with torch.no_grad(), ctx:
        for k in range(2):
            step(k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(sample_steps):
            out = step(k)
        torch.cuda.synchronize()
        t_steps = time.perf_counter() - t0
    t0 = time.perf_counter()
    lat = out["sample"].float().cpu()
    for b in range(batch):
        orc.solve(lat[b], G, S // (16 * G))
    t_assign = time.perf_counter() - t0
    total = t_steps / sample_steps * wl["steps"] + t_assign
    return batch / total, 1000.0 * t_steps / sample_steps, 1000.0 * t_assign


def run_reference_gpu(args, wl):
    """`--impl reference --ref-device cuda`: not part of the driver contract (its reference arm is the CPU one below);
    prints one line per precision for DESIGN.md's comparison table."""
    batch = args.batch or wl["batch"]
    for prec in args.ref_precision.split(","):
        v, ms, ms_assign = stock_torch_gpu_puzzles_per_s(wl, batch, prec, sample_steps=max(1, args.steps))
        print(json.dumps({"impl": "reference", "device": "cuda:0 (stock torch ops)", "precision": prec,
                          "metric": "puzzles/sec (3x3 @192px sampling)", "value": v, "unit": "puzzles/s",
                          "ms_per_diffusion_step": ms, "assignment_ms": ms_assign,
                          "config": {"workload": wl["name"], "batch": batch, "diffusion_steps": wl["steps"]}}))


def run_reference(args, wl):
    """`--impl reference`: the reference path's CPU implementation on the host cores - the unmodified reference modules
    from the staged baseline/_ref when the snapshot carries them (`kind: "reference"`), else the oracle port
    (`kind: "port"`).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.ref_device == "cuda":
        return run_reference_gpu(args, wl)
    vals, spent = [], 0.0
    for i in range(args.warmup + args.steps):
        v, cores, sample, dt, kind = cpu_reference_puzzles_per_s(wl, sample_batch=16, sample_steps=2)
        if i >= args.warmup:
            vals.append(v); spent += dt
    value = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": "puzzles/sec (3x3 @192px sampling)", "value": value, "unit": "puzzles/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * spent / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": sampling_config(wl, args.batch or wl["batch"]),
        "cpu_baseline": {"value": value, "unit": "puzzles/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "puzzles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------- GPU arm
NCU_SUMMARY = os.path.join(ROOT, "profiles", "r1j_ncu_full_kernels.json")


def ncu_traffic(label):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `label` from the committed `ncu --set full` capture
    (tools/ncu_summary.py output; same shape as the bench: M = 36864).  None when the capture is absent."""
    try:
        for d in json.load(open(NCU_SUMMARY)):
            if d.get("launch") == label:
                mb = float(d["dram__bytes_read.sum [Mbyte]"]) + float(d["dram__bytes_write.sum [Mbyte]"])
                return mb * 1e6, os.path.relpath(NCU_SUMMARY, ROOT)
    except Exception:
        pass
    return None, None


def kernel_roofline(model, wl, batch, peaks):
    """Per-kernel CUDA-event timings of one denoiser forward at the bench shape, through the per-kernel C-ABI entry
    points (ops.*) on the launching stream; inputs are larger than L2 at this batch.  Returns the roofline object of
    the dominant kernel plus the full breakdown."""
    import torch
    from jpdvt_mt_ntnu_b200 import ops
    T = (wl["size"] // 16) ** 2
    M = batch * T
    dev = torch.device("cuda")
    eng = model.engine(dev)
    W = eng.weights.tensors
    xn = torch.randn(M, 768, device=dev).bfloat16()
    x = torch.randn(M, 768, device=dev)
    hid = torch.randn(M, 3072, device=dev).bfloat16() * 0.1
    qkv = torch.randn(M, 2304, device=dev).bfloat16()
    gate = torch.randn(1, 768, device=dev) * 0.1
    shift, scale = torch.randn(1, 768, device=dev), torch.randn(1, 768, device=dev)
    calls = {
        "gemm_qkv (tcgen05, bias)": (lambda: ops.gemm_bias(xn, W["w_qkv"][0], W["b_qkv"][0]), 2.0 * M * 768 * 2304, "flop"),
        "gemm_fc1 (tcgen05, bias+gelu)": (lambda: ops.gemm_bias_gelu(xn, W["w_fc1"][0], W["b_fc1"][0]), 2.0 * M * 768 * 3072, "flop"),
        "gemm_fc2 (tcgen05, gated residual via TMA ring)": (lambda: ops.gemm_bias_gate_residual(x, hid, W["w_fc2"][0], W["b_fc2"][0], gate, T), 2.0 * M * 3072 * 768, "flop"),
        "gemm_proj (tcgen05, gated residual via TMA ring)": (lambda: ops.gemm_bias_gate_residual(x, xn, W["w_proj"][0], W["b_proj"][0], gate, T), 2.0 * M * 768 * 768, "flop"),
        # HBM-bound since the remainder-warp / P-in-TMEM kernels: algorithmic bytes = qkv in (2304 bf16) + output (768 bf16) per row
        "attention (tcgen05 + TMEM, HBM-bound)": (lambda: ops.attention(qkv, batch, T), M * (2304 + 768) * 2.0, "byte"),
        "proj as bytes (fp32 residual RMW + bf16 operand)": (lambda: ops.gemm_bias_gate_residual(x, xn, W["w_proj"][0], W["b_proj"][0], gate, T), M * 768 * (4.0 + 4.0 + 2.0), "byte"),
        "ln_modulate (fp32->bf16)": (lambda: ops.ln_modulate(x, shift, scale, T), M * 768 * 6.0, "byte"),
    }
    out = {}
    for name, (fn, work, kind) in calls.items():
        for _ in range(3):
            fn()
        n = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        if kind == "flop":
            ach = work / (ms * 1e-3) / 1e12
            out[name] = {"ms": ms, "achieved": ach, "unit": "TFLOP/s", "frac": ach / peaks["bf16"]}
        else:
            ach = work / (ms * 1e-3) / 1e9
            out[name] = {"ms": ms, "achieved": ach, "unit": "GB/s", "frac": ach / peaks["hbm_gbs"]}
    top = "gemm_fc1 (tcgen05, bias+gelu)"
    traffic, traffic_src = ncu_traffic("fc1")
    roof = {"bound": "tensor", "kernel": top, "achieved": out[top]["achieved"], "peak": peaks["bf16"],
            "unit": "TFLOP/s", "frac": out[top]["frac"], "traffic": traffic, "traffic_unit": "bytes/launch (dram read + write)",
            "traffic_source": traffic_src, "algorithmic_bytes": M * 768 * 2.0 + 3072 * 768 * 2.0 + M * 3072 * 2.0,
            "peak_source": f"{peaks['source']} bf16 burst (the kernel is timed alone, 10 back-to-back launches; every entry of "
                           "`kernels` uses the same burst / HBM peaks)",
            "frac_of_sustained": out[top]["achieved"] / peaks["bf16_sustained"],
            "how": "CUDA events on the launching stream over 10 back-to-back launches at the bench shape "
                   f"(M={M}, activations > L2); algorithmic FLOPs 2*M*768*3072"}
    return roof, out


def build_sampler(wl, batch, rank, dev):
    """Model (random seeded weights), diffusion and the synthetic puzzle batch of one rank; inputs pinned on the host."""
    from jpdvt_mt_ntnu_b200.diffusion import create_diffusion
    from jpdvt_mt_ntnu_b200.models import DiT_models
    from jpdvt_mt_ntnu_b200.weights import seeded_state
    model = DiT_models["JPDVT"](input_size=wl["size"])
    model.load_state_dict(seeded_state(model.state_dict(), seed=1234))
    model.to(dev)
    diffusion = create_diffusion(str(wl["steps"]))
    cond_h, noise_h, perms = synthetic_inputs(wl, batch, seed=rank)
    return model, diffusion, cond_h.pin_memory(), noise_h.pin_memory(), perms


def measure_strong(wl, total, world, rank, dev, steps, model=None, diffusion=None):
    """The sharded-batch form of a sampling config (reference: image_paths[rank::world_size],
    inference_ddp.py:325): `total` puzzles split over the ranks, each rank solving total / world of them with the loop
    replayed from a CUDA graph (small per-GPU batches are launch-bound), one all_gather of the placements per pass."""
    import torch
    import torch.distributed as dist
    from jpdvt_mt_ntnu_b200 import _lib, assignment
    per = total // world
    G = wl["grid"]
    if model is None:
        model, diffusion, cond_pin, noise_pin, _ = build_sampler(wl, per, rank, dev)
    else:
        cond_h, noise_h, _ = synthetic_inputs(wl, per, seed=100 + rank)
        cond_pin, noise_pin = cond_h.pin_memory(), noise_h.pin_memory()
    cond, noise = cond_pin.to(dev), noise_pin.to(dev)
    gathered = [torch.empty(per, G * G, dtype=torch.int32, device=dev) for _ in range(world)] if world > 1 else None
    graph = per <= 64

    def one(c, z):
        sample = diffusion.p_sample_loop(model.forward, c, z.shape, z, clip_denoised=False, model_kwargs=None, progress=False,
                                         device=dev, graph=graph)
        order, pred = assignment.solve_puzzles(sample, G)
        if world > 1:
            dist.all_gather(gathered, pred)
        return pred

    for _ in range(2):
        one(cond, noise)
    ms, _ = _timed(lambda: one(cond, noise), steps, world, dev)
    ms_e2e, _ = _timed(lambda: one(cond_pin.to(dev, non_blocking=True), noise_pin.to(dev, non_blocking=True)).cpu(), steps, world, dev)
    tflops = flops_per_forward((wl["size"] // 16) ** 2) * wl["steps"] * per * steps / (ms * 1e-3) / 1e12     # per GPU
    return {"workload": wl["name"].split(",")[0], "total_puzzles": per * world, "batch_per_gpu": per, "graph_replay": graph,
            "value": per * world * steps / (ms * 1e-3), "unit": "puzzles/s", "ms_per_pass": ms / steps,
            "model_tflops_per_gpu": tflops, "model_frac_of_bf16_sustained": tflops / measured_peaks()["bf16_sustained"],
            "e2e": {"value": per * world * steps / (ms_e2e * 1e-3), "unit": "puzzles/s", "ms_per_pass": ms_e2e / steps}}


def run_ours(args, wl):
    import numpy as np
    import torch
    import torch.distributed as dist
    from jpdvt_mt_ntnu_b200 import _lib, assignment

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    batch = args.batch or wl["batch"]
    S, G = wl["size"], wl["grid"]
    T = (S // 16) ** 2

    model, diffusion, cond_pin, noise_pin, perms = build_sampler(wl, batch, rank, dev)
    cond, noise = cond_pin.to(dev), noise_pin.to(dev)
    torch.manual_seed(rank)        # seeds the per-step noise p_sample draws inside the loop (gaussian_diffusion.py:424)
    gathered = [torch.empty(batch, G * G, dtype=torch.int32, device=dev) for _ in range(world)] if world > 1 else None

    def one_step(c, z):
        sample = diffusion.p_sample_loop(model.forward, c, z.shape, z, clip_denoised=False, model_kwargs=None,
                                         progress=False, device=dev)
        order, pred = assignment.solve_puzzles(sample, G)
        if world > 1:
            dist.all_gather(gathered, pred)          # the only cross-rank traffic: int32 placements (inference_ddp.py:485-495)
        return pred

    for _ in range(args.warmup):
        one_step(cond, noise)
    clocks = ClockSampler(local)
    clocks.start()
    n0 = _lib.launch_count()
    ms, pred = _timed(lambda: one_step(cond, noise), args.steps, world, dev)
    launches = _lib.launch_count() - n0
    value = batch * world * args.steps / (ms * 1e-3)

    def e2e_step():
        c = cond_pin.to(dev, non_blocking=True)
        z = noise_pin.to(dev, non_blocking=True)
        return one_step(c, z).cpu()

    e2e_step()
    ms_e2e, pred_host = _timed(e2e_step, args.steps, world, dev)
    clock_info = clocks.stop()
    e2e_value = batch * world * args.steps / (ms_e2e * 1e-3)

    # ---- the sharded-batch (strong-scaling) form: the config's puzzle count is the WHOLE job's, split over the ranks
    # (the extra legs must never cost the headline its line: a failure - symmetric across the ranks, e.g. out of memory - is
    # recorded in place of the leg's numbers)
    strong = None
    if not args.no_extras and wl is WORKLOADS["c2"]:
        strong = {}
        try:
            strong["c2"] = measure_strong(wl, 256, world, rank, dev, max(2, args.steps), model, diffusion)
        except Exception as e:  # noqa: BLE001
            strong["c2"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        del model
        torch.cuda.empty_cache()
        model = None
        # the other sampling configs of BASELINE.json (configs[3], [4]): 4x4 @256 px and 3x3 @288 px with missing pieces,
        # 128 puzzles for the whole job each, so that every --gpus N line carries the whole reporting matrix
        for key in ("c4", "c5"):
            try:
                strong[key] = measure_strong(WORKLOADS[key], 128, world, rank, dev, 2)
            except Exception as e:  # noqa: BLE001
                strong[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()
    # ---- BASELINE.json's other metric: train img/s (configs[2]) with the gradient exchange at this N
    train = None
    if not args.no_extras and wl is WORKLOADS["c2"]:
        tclocks = ClockSampler(local)
        tclocks.start()
        try:
            train = measure_training(WORKLOADS["c3"], WORKLOADS["c3"]["batch"], 10, 3, world, rank, dev)
        except Exception as e:  # noqa: BLE001
            train = {"error": f"{type(e).__name__}: {e}"[:300]}
        train["clocks"] = tclocks.stop()

    if rank == 0:
        peaks = measured_peaks()
        if model is None:
            model = build_sampler(wl, batch, rank, dev)[0]
        roof, breakdown = kernel_roofline(model, wl, batch, peaks)
        cpu_obj, stock = None, None
        if world == 1:           # reported at N=1 only (torchrun pins the ranks to one host thread each)
            cpu_v, cores, sample, _, kind = cpu_reference_puzzles_per_s(wl)
            cpu_obj = {"value": cpu_v, "unit": "puzzles/s", "cores": cores, "kind": kind, "sample": sample}
            if not args.no_extras:
                del model
                torch.cuda.empty_cache()
                stock = {"what": "the reference path's stock torch ops (cuBLAS / ATen / fused SDPA - none of this repo's kernels) on "
                                 "the same B200 and batch, a few of the 250 steps scaled up + the host assignment loop",
                         "unit": "puzzles/s"}
                for prec in ("fp32", "tf32", "bf16"):
                    v, ms_step, _ = stock_torch_gpu_puzzles_per_s(wl, batch, prec, sample_steps=3)
                    stock[prec] = {"value": v, "ms_per_diffusion_step": ms_step}
        solved = float((pred_host.numpy() == perms).all(axis=1).mean())
        flops = flops_per_forward(T) * wl["steps"] * batch * args.steps
        line = {
            "metric": "puzzles/sec (3x3 @192px sampling)" if wl is WORKLOADS["c2"] else f"puzzles/sec ({wl['name']})",
            "value": value, "unit": "puzzles/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": sampling_config(wl, batch),
            "e2e": {"value": e2e_value, "unit": "puzzles/s", "h2d_bytes_per_step": int(cond_pin.numel() * 4 + noise_pin.numel() * 4),
                    "d2h_bytes_per_step": int(batch * G * G * 4), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "gpu_launches_how": "jpdvt_launch_count() before / after the timed region (every launch site of the library counts itself)",
            "clocks": clock_info,
            "roofline": roof,
            "kernels": breakdown,
            "model_tflops": flops / (ms * 1e-3) / 1e12,
            "model_frac_of_bf16_sustained": flops / (ms * 1e-3) / 1e12 / peaks["bf16_sustained"] / world,
            "cpu_baseline": cpu_obj,
            "gpu_stock_baseline": stock,
            "strong": strong,
            "train": train,
            "puzzles_solved_frac_random_weights": solved,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()



def cpu_port_train_img_per_s(wl, sample_batch=2):
    """CPU port of one training step (fp32 autograd through the oracle + torch AdamW) on a bounded sample."""
    import numpy as np
    import torch
    from oracle import jpdvt_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    S, G = wl["size"], wl["grid"]
    T = (S // 16) ** 2
    st = {k: v.clone().requires_grad_(k != "pos_embed") for k, v in orc.seeded_state(orc.blank_state(S, DEPTH), seed=1234).items()}
    model = orc.OracleDenoiser.__new__(orc.OracleDenoiser)
    model.w, model.depth, model.heads, model.patch = st, DEPTH, 12, 16
    opt = torch.optim.AdamW([v for k, v in st.items() if k != "pos_embed"], lr=1e-4, weight_decay=0)
    sched = orc.Schedule("")
    piece = torch.from_numpy(orc.sincos_2d(8, G)).float().unsqueeze(0)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(sample_batch, 3, S, S, generator=g) * 2 - 1
    t = torch.randint(0, 1000, (sample_batch,), generator=g)
    times = []
    for it in range(2):
        t0 = time.perf_counter()
        o = orc.training_losses(sched, model, x, t, piece, np.random.RandomState(it).permutation(G * G), torch.randn_like(x),
                                torch.randn(sample_batch, T, 8), block_size=S // G, grid=G, masks=None)
        opt.zero_grad()
        o["loss"].mean().backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    return sample_batch / times[-1], cores, f"{sample_batch} images, one fwd+bwd+AdamW step (second of two)", sum(times)


def _timed(fn, k, world, dev):
    """K calls of fn bracketed by barrier + synchronize on both sides, CUDA events on the launching stream, MAX over ranks."""
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for _ in range(k):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item(), out


def measure_training(wl, batch, steps, warmup, world, rank, dev):
    """BASELINE.json configs[2]: data-parallel training steps (training_losses -> backward -> NCCL gradient all-reduce ->
    fused AdamW + EMA, train_JPDVT.py:357-372) through `Trainer`; returns the numbers of the `train` object."""
    import torch
    from jpdvt_mt_ntnu_b200.diffusion import create_diffusion
    from jpdvt_mt_ntnu_b200.models import DiT_models, get_2d_sincos_pos_embed
    from jpdvt_mt_ntnu_b200.trainer import Trainer
    from jpdvt_mt_ntnu_b200.weights import seeded_state
    S, G = wl["size"], wl["grid"]
    T = (S // 16) ** 2
    model = DiT_models["JPDVT"](input_size=S)
    model.load_state_dict(seeded_state(model.state_dict(), seed=1234))
    model.to(dev)
    diffusion = create_diffusion("")
    trainer = Trainer(model, diffusion, lr=1e-4, weight_decay=0.0, ema_decay=0.9999)
    g = torch.Generator().manual_seed(rank)
    x_pin = (torch.rand(batch, 3, S, S, generator=g) * 2 - 1).pin_memory()
    x = x_pin.to(dev)
    piece = torch.tensor(get_2d_sincos_pos_embed(8, G)).unsqueeze(0).float().to(dev)
    kw = dict(block_size=S // G, patch_size=16, add_mask=False, grid_size=G)
    torch.manual_seed(rank)

    # the whole step replayed from a CUDA graph (no host in the loop, train_JPDVT.py:340-372) where the exchange allows it:
    # one GPU or the peer-memory step; an NCCL all-reduce stays on host-issued launches.  JPDVT_TRAIN_GRAPH=0: never.
    graph = os.environ.get("JPDVT_TRAIN_GRAPH", "1")[:1] != "0" and (world == 1 or trainer.px is not None)

    def one_step(xin):
        t = torch.randint(0, diffusion.num_timesteps, (batch,), device=dev)          # train_JPDVT.py:354
        return trainer.step(xin, t, piece, graph=graph, **kw)

    from jpdvt_mt_ntnu_b200 import _lib
    for _ in range(max(warmup, 3)):
        one_step(x)
    n0 = _lib.launch_count() + trainer.replayed_launches
    ms, loss = _timed(lambda: one_step(x), steps, world, dev)
    launches = _lib.launch_count() + trainer.replayed_launches - n0
    # end to end through the trainer's own host-side pieces: every step's batch comes from pinned host memory (copied one
    # step ahead on a copy stream - BatchPrefetcher), every step's loss goes back to the host (pinned ring - LossLog)
    from jpdvt_mt_ntnu_b200.trainer import BatchPrefetcher, LossLog
    log = LossLog()

    def e2e_steps(n=steps):
        for xin in BatchPrefetcher((x_pin for _ in range(n)), dev):
            log.push(one_step(xin))
        return log.values(dev)[-1]
    e2e_steps(2)                                   # copy stream, device slots and pinned ring exist before the timed region
    ms_e2e, loss_h = _timed(e2e_steps, 1, world, dev)
    flops = 3.0 * flops_per_forward(T) * batch * steps
    mode = "none (1 GPU)"
    if world > 1:
        mode = trainer.allreduce_description()
    out = {"value": batch * world * steps / (ms * 1e-3), "unit": "img/s", "ms_per_step": ms / steps, "steps": steps,
           "batch_per_gpu": batch, "workload": wl["name"], "params": trainer.total, "allreduce": mode,
           "e2e": {"value": batch * world * steps / (ms_e2e * 1e-3), "unit": "img/s", "h2d_bytes_per_step": int(x_pin.numel() * 4),
                   "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / steps},
           "model_tflops_per_gpu": flops / (ms * 1e-3) / 1e12, "final_loss": float(loss_h),
           "gpu_launches": launches, "graph_replay": graph}
    del trainer, model
    torch.cuda.empty_cache()
    return out


def run_train(args, wl):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    batch = args.batch or wl["batch"]
    S, G = wl["size"], wl["grid"]
    T = (S // 16) ** 2
    steps = max(args.steps, 10)
    clocks = ClockSampler(local)
    clocks.start()
    tr = measure_training(wl, batch, steps, args.warmup, world, rank, dev)
    clock_info = clocks.stop()
    if rank == 0:
        peaks = measured_peaks()
        cpu_obj = None
        if world == 1:
            v, cores, sample, _ = cpu_port_train_img_per_s(wl)
            cpu_obj = {"value": v, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": "train img/s", "value": tr["value"], "unit": "img/s", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": tr["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": wl["name"], "batch_per_gpu": batch, "image_size": S, "grid": G, "tokens": T,
                       "optimizer": "AdamW lr 1e-4 wd 0 + EMA 0.9999 (fused, fp32 state)", "params": tr["params"],
                       "allreduce": tr["allreduce"], "l2": "activations per launch exceed L2; no flush needed"},
            "e2e": tr["e2e"], "gpu_launches": tr["gpu_launches"], "clocks": clock_info,
            "model_tflops": tr["model_tflops_per_gpu"] * world,
            "model_frac_of_bf16_sustained": tr["model_tflops_per_gpu"] / peaks["bf16_sustained"],
            "roofline": {"bound": "tensor", "kernel": "whole training step (3 x forward FLOPs: fwd + dgrad + wgrad)",
                         "achieved": tr["model_tflops_per_gpu"], "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": tr["model_tflops_per_gpu"] / peaks["bf16_sustained"], "traffic": None},
            "cpu_baseline": cpu_obj, "final_loss": tr["final_loss"],
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: cuda = stock torch ops on the GPU (same-box comparator, not the contract arm)")
    ap.add_argument("--ref-precision", default="fp32,tf32,bf16")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="puzzles per GPU (default: the workload's)")
    ap.add_argument("--no-extras", action="store_true",
                    help="only the headline sampling measurement (skip the strong-scaling, training and stock-torch legs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3            # timing rule: at least 3 warm-up steps
    wl = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # torchrun pins OMP_NUM_THREADS=1; the CPU arm (rank 0 only) must see every host core.  torch is imported lazily
    # below, so setting the variables here takes effect.
    if args.impl == "reference" or world == 1:
        for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[var] = str(os.cpu_count() or 1)
    # NCCL logs to stdout by default; stdout carries the one JSON line, so its log (whatever NCCL_DEBUG level the caller
    # set - the driver reads the communicator's rank count from it) goes to stderr instead of being switched off
    if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
    if args.impl == "reference":
        run_reference(args, wl)
    elif wl.get("train"):
        run_train(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()

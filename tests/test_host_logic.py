"""Host-side mirror of the reference interface: schedules, module layout, error behaviour, C-ABI exports (no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from jpdvt_mt_ntnu_b200 import _lib, parallel
from jpdvt_mt_ntnu_b200.diffusion import create_diffusion, space_timesteps
from jpdvt_mt_ntnu_b200.diffusion import gaussian_diffusion as gd
from jpdvt_mt_ntnu_b200.models import DiT, DiT_models, get_2d_sincos_pos_embed
from jpdvt_mt_ntnu_b200.weights import seeded_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pos_embed_bit_exact(golden):
    g = golden("static")
    assert np.array_equal(get_2d_sincos_pos_embed(8, 3), g["pe_8_3"])
    assert np.array_equal(get_2d_sincos_pos_embed(8, 4), g["pe_8_4"])
    assert np.array_equal(get_2d_sincos_pos_embed(768, 12), g["pe_768_12"])


@pytest.mark.parametrize("name,spec", [("full", ""), ("s250", "250"), ("s10", "10"), ("ddim50", "ddim50"), ("sec", "10,15,20")])
def test_create_diffusion_tables_bit_exact(golden, name, spec):
    g = golden("static")
    d = create_diffusion(spec)
    assert np.array_equal(np.asarray(d.timestep_map), g[f"{name}_map"])
    for key, arr in (("betas", d.betas), ("sqrt_ac", d.sqrt_alphas_cumprod), ("sqrt_1mac", d.sqrt_one_minus_alphas_cumprod),
                     ("post_var", d.posterior_variance), ("post_logvar", d.posterior_log_variance_clipped),
                     ("coef1", d.posterior_mean_coef1), ("coef2", d.posterior_mean_coef2)):
        assert np.array_equal(arr, g[f"{name}_{key}"]), key
    assert d.num_timesteps == len(g[f"{name}_map"])


def test_create_diffusion_defaults():
    d = create_diffusion("250")
    assert d.model_mean_type == gd.ModelMeanType.START_X
    assert d.model_var_type == gd.ModelVarType.FIXED_SMALL
    assert d.loss_type == gd.LossType.MSE
    assert d.original_num_steps == 1000
    assert create_diffusion(None).num_timesteps == 1000
    assert create_diffusion("", predict_xstart=False).model_mean_type == gd.ModelMeanType.EPSILON


def test_space_timesteps_errors_and_forms():
    assert space_timesteps(1000, "250") == space_timesteps(1000, [250])
    assert sorted(space_timesteps(300, [10, 15, 20]))[:3] == [0, 11, 22]
    with pytest.raises(ValueError):
        space_timesteps(10, "20")
    with pytest.raises(ValueError):
        space_timesteps(1000, "ddim999")


def test_state_dict_layout_and_param_count(golden):
    g = golden("static")
    m = DiT_models["JPDVT"](input_size=192)
    st = m.state_dict()
    assert list(st.keys()) == [str(k) for k in g["state_keys"]]
    assert [v.numel() for v in st.values()] == g["state_numel"].tolist()
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params_192"]) == 130857800
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == int(g["n_trainable_192"]) == 130747208
    assert not m.pos_embed.requires_grad
    assert sum(p.numel() for p in DiT_models["JPDVT"](input_size=288).parameters()) == 130996040


def test_fresh_init_scheme():
    torch.manual_seed(0)
    m = DiT(input_size=96, depth=2, hidden_size=768, patch_size=16, num_heads=12)
    for blk in m.blocks:                                     # adaLN-Zero (models.py:216-225)
        assert blk.adaLN_modulation[-1].weight.abs().max() == 0 and blk.adaLN_modulation[-1].bias.abs().max() == 0
    assert m.final_layer.linear.weight.abs().max() == 0 and m.final_layer.adaLN_modulation[-1].weight.abs().max() == 0
    assert abs(m.time_emb_in.weight.std().item() - 0.02) < 2e-3
    assert abs(m.t_embedder.mlp[0].weight.std().item() - 0.02) < 1e-3
    w = m.blocks[0].attn.qkv.weight                          # xavier uniform: bound sqrt(6/(fan_in+fan_out))
    assert w.abs().max().item() <= (6.0 / (768 + 2304)) ** 0.5 + 1e-6
    assert m.blocks[0].attn.qkv.bias.abs().max() == 0
    assert np.array_equal(m.pos_embed[0].numpy(), get_2d_sincos_pos_embed(768, 6).astype(np.float32))


def test_state_dict_round_trip_and_deepcopy():
    import copy
    m = DiT(input_size=48, depth=1, hidden_size=768, patch_size=16, num_heads=12)
    st = seeded_state(m.state_dict(), seed=3)
    m.load_state_dict(st)
    m2 = copy.deepcopy(m)                                    # EMA copy (train_JPDVT.py:235)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    missing = m2.load_state_dict({k: v for k, v in st.items() if "blocks" not in k}, strict=False)
    assert all("blocks" in k for k in missing.missing_keys)


def test_no_cpu_fallback():
    """The product path must fail loudly without a B200 - never silently compute on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    m = DiT(input_size=48, depth=1, hidden_size=768, patch_size=16, num_heads=12)
    with torch.no_grad(), pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 48, 48), torch.zeros(1, dtype=torch.long), torch.zeros(1, 9, 8))
    d = create_diffusion("10")
    with pytest.raises(RuntimeError):
        d.q_sample(torch.zeros(1, 9, 8), torch.zeros(1, dtype=torch.long), torch.zeros(1, 9, 8))
    with pytest.raises(RuntimeError):
        _lib.require_device()


def test_unsupported_configs_raise():
    m = DiT_models["JPDVT-T"](input_size=256)               # patch 64: the reference forward itself crashes (SURVEY headline 6)
    with pytest.raises(NotImplementedError):
        m._check_supported()
    with pytest.raises(NotImplementedError):
        DiT_models["DiT-S/2"](input_size=32)._check_supported()


def test_cabi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "jpdvt_b200.h")).read()
    declared = set(re.findall(r"\b(jpdvt_[a-z0-9_]+)\s*\(", header))
    declared -= {"jpdvt_status"}
    assert len(declared) >= 20
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/jpdvt_b200.h but not exported"
    assert set(_lib.PROTOTYPES) | set(_lib.OTHER_SYMBOLS) == declared
    assert lib.jpdvt_abi_version() == _lib.ABI_VERSION == 6
    assert isinstance(lib.jpdvt_last_error_string(), bytes)


def test_cabi_struct_layout_matches_header():
    # 4 int32 + 24 pointers; int64 + 2 int32 + 20 pointers; 2 int32 + 6 ptr + int64 + 5 ptr
    assert ctypes.sizeof(_lib.Weights) == 16 + 24 * 8
    assert ctypes.sizeof(_lib.Workspace) == 16 + 20 * 8
    assert ctypes.sizeof(_lib.Sampler) == 8 + 6 * 8 + 8 + 5 * 8
    header = open(os.path.join(ROOT, "include", "jpdvt_b200.h")).read()
    for struct, cls in (("jpdvt_weights", _lib.Weights), ("jpdvt_workspace", _lib.Workspace), ("jpdvt_sampler", _lib.Sampler),
                        ("jpdvt_tape", _lib.Tape), ("jpdvt_bwd_scratch", _lib.BwdScratch), ("jpdvt_weights_t", _lib.WeightsT),
                        ("jpdvt_peer_step", _lib.PeerStep)):
        body = header[header.index(f"typedef struct {struct} {{"):header.index(f"}} {struct};")]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = re.findall(r"\b([a-z_0-9]+)(?:\[[^\]]+\])?;", body)
        assert names == [f[0] for f in cls._fields_], struct


def test_block_and_strided_shards():
    for total, w in ((256, 8), (10, 3), (7, 8), (0, 2)):
        spans = [parallel.block_shard(total, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1
    items = list(range(11))
    got = sorted(sum((parallel.strided_shard(items, r, 4) for r in range(4)), []))
    assert got == items


def test_documented_knobs_exist_in_the_sources():
    """Every JPDVT_* environment knob DESIGN.md's table names is read somewhere in the package (and vice versa for the
    knobs the C sources read), so the A/B documentation cannot drift from the code."""
    design = open(os.path.join(ROOT, "DESIGN.md")).read()
    table = design[design.index("## 9. A/B knobs"):design.index("## 10.")]
    documented = set(re.findall(r"`(JPDVT_[A-Z0-9_]+)", table))
    src = ""
    pkg = os.path.join(ROOT, "jpdvt_mt_ntnu_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".py")):
                src += open(os.path.join(base, f), errors="ignore").read()
    read_in_c = set(re.findall(r'getenv\("(JPDVT_[A-Z0-9_]+)"\)', src))
    read_in_py = set(re.findall(r'environ(?:\.get)?[\(\[]\s*"(JPDVT_[A-Z0-9_]+)"', src))
    missing = {k for k in documented if k not in src}
    assert not missing, f"documented but not read anywhere: {sorted(missing)}"
    internal = {"JPDVT_FORCE_BUILD", "JPDVT_KEEP_NCCL_DEBUG"}          # build / bench plumbing, not A/B knobs
    undocumented = (read_in_c | read_in_py) - documented - internal
    assert not undocumented, f"read by the code but missing from DESIGN.md section 9: {sorted(undocumented)}"


def test_peer_shard_bounds_and_f32_ranges():
    """Host logic of the peer-memory optimizer step (jpdvt_mt_ntnu_b200/peer.py): the flat parameter space is cut into equal
    contiguous slices of whole 2,048-parameter tiles that cover it exactly once; the fp32-replicated index ranges are the
    sorted union of the non-bf16 fields."""
    from jpdvt_mt_ntnu_b200 import peer
    from jpdvt_mt_ntnu_b200._lib import JpdvtError, MAX_F32_RANGES
    for total in (130747208, 2048, 2049, 5, 16 * 2048):
        for world in (2, 3, 4, 8):
            spans = [peer.shard_bounds(total, world, r) for r in range(world)]
            chunk = spans[0][0]
            assert chunk % 2048 == 0 and all(c == chunk for c, _, _ in spans)
            assert spans[0][1] == 0 and all(spans[r][2] == spans[r + 1][1] for r in range(world - 1))
            assert spans[-1][2] == world * chunk >= total > world * chunk - world * 2048
    assert peer.merge_ranges([(10, 20), (0, 5), (20, 30), (28, 40), (50, 50), (60, 61)]) == [(0, 5), (10, 40), (60, 61)]
    with pytest.raises(JpdvtError):
        peer.merge_ranges([(3 * i, 3 * i + 1) for i in range(MAX_F32_RANGES + 1)])
    # the JPDVT layout: everything that is not one of the eight bf16 matrices collapses into eight ranges
    from jpdvt_mt_ntnu_b200.trainer import _BF16_FIELDS
    from jpdvt_mt_ntnu_b200.training import _grad_layout
    import math
    off, ranges, f32_total = 0, [], 0
    for name, shape in _grad_layout(12):
        n = math.prod(shape)
        if name not in _BF16_FIELDS:
            ranges.append((off, off + n))
            f32_total += n
        off += n
    merged = peer.merge_ranges(ranges)
    assert len(merged) == 8 and sum(e - b for b, e in merged) == f32_total and off == 130747208


def test_wgrad_split_plan_fills_whole_waves():
    """The weight-gradient GEMMs' contraction split (csrc/gemm.cu: wgrad_plan, read back through the scratch-size query - a pure
    host function, 148 SMs assumed without a device): the number of work items tiles x splits must land just under a whole
    number of waves on the 74 CTA pairs, not between waves.  At the C3 shapes (M = 128 x 144 rows): fc1 / fc2 have 36 output
    tiles -> 2 splits = 72 items (one wave; the first plan's 5 splits = 180 items = 2.43 waves left a quarter of the SM time
    idle), qkv 27 tiles -> 8 splits = 216 items (2.92 waves), proj 9 tiles -> 8 splits = 72 items."""
    lib = _lib.load()
    m = 128 * 144
    pairs = 74
    for name, (rows, cols), want in (("qkv", (2304, 768), 8), ("proj", (768, 768), 8), ("fc1", (3072, 768), 2), ("fc2", (768, 3072), 2)):
        floats = int(lib.jpdvt_wgrad_scratch_floats(m, rows, cols))
        split = floats // (rows * cols) if floats else 1
        assert split == want, (name, split)
        tiles = ((rows + 255) // 256) * (cols // 256)
        items = tiles * split
        waves = -(-items // pairs)
        assert items / (waves * pairs) > 0.95, (name, items, waves)
    # tiny contractions (the conditioning layers: m = batch rows) never split below one 64-row block per item
    for rows, cols in ((768, 768), (768, 256)):
        floats = int(lib.jpdvt_wgrad_scratch_floats(128, rows, cols))
        split = floats // (rows * cols) if floats else 1
        assert 1 <= split <= 2


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's reference arm) runs without a GPU and prints ONE JSON line with the contract's
    keys; `cpu_baseline.kind` says which CPU implementation was timed: the unmodified reference modules from the staged
    baseline/_ref when the snapshot carries them, the oracle port otherwise."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "puzzles/sec (3x3 @192px sampling)" and line["unit"] == "puzzles/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    staged = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "image_model", "models.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and "scaled x250/2" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "puzzles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("JPDVT 3x3 @192px sampling") and line["config"]["batch_per_gpu"] == 256

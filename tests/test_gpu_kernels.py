"""Per-kernel parity on a B200: every C-ABI entry point against a torch fp32 restatement of the same op.

Tolerances (stated per SURVEY.md 8c): bf16-output kernels rel-L2 <= 5e-3 (bf16 rounding of the result alone is ~1.7e-3);
fp32-output kernels with bf16 operands / fp32 accumulation rel-L2 <= 1e-5 against the same bf16-rounded operands;
pure fp32 elementwise kernels bit-exact; integer kernels bit-exact.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
BF16_TOL, F32_TOL = 5e-3, 1e-5


@pytest.fixture(scope="module")
def ops(cuda):
    from jpdvt_mt_ntnu_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (256, 768, 768), (432, 2304, 768), (1000, 768, 3072), (27, 3072, 768),
                                   (4608, 2304, 768), (300, 128, 128), (1, 256, 64), (129, 384, 192)])
def test_gemm_bias_shapes(ops, m, n, k):
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(n, device="cuda")
    ref = a.float() @ w.float().t() + bias
    assert rel_l2(ops.gemm_bias(a, w, bias).float(), ref) < BF16_TOL
    assert rel_l2(ops.gemm_bias_f32(a, w, bias), ref) < F32_TOL


def test_gemm_empty_and_bad_shapes(ops):
    from jpdvt_mt_ntnu_b200._lib import JpdvtError
    a = torch.zeros(0, 64, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16)
    assert ops.gemm_bias(a, w, torch.zeros(128, device="cuda")).shape == (0, 128)      # empty batch is a no-op
    with pytest.raises(JpdvtError):                                                     # K not a multiple of 64
        ops.gemm_bias(torch.zeros(8, 40, device="cuda", dtype=torch.bfloat16), torch.zeros(128, 40, device="cuda", dtype=torch.bfloat16),
                      torch.zeros(128, device="cuda"))
    with pytest.raises(JpdvtError):                                                     # N not a multiple of 128
        ops.gemm_bias(torch.zeros(8, 64, device="cuda", dtype=torch.bfloat16), torch.zeros(100, 64, device="cuda", dtype=torch.bfloat16),
                      torch.zeros(100, device="cuda"))
    with pytest.raises(JpdvtError):                                                     # wrong dtype is rejected, not converted
        ops.gemm_bias(torch.zeros(8, 64, device="cuda"), w, torch.zeros(128, device="cuda"))


def test_gemm_epilogues(ops):
    torch.manual_seed(0)
    m, n, k, T = 432, 768, 768, 144
    a = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(n, device="cuda")
    lin = a.float() @ w.float().t() + bias
    o16, o32 = ops.gemm_bias(a, w, bias, want_f32_copy=True)
    assert rel_l2(o32, lin) < F32_TOL and rel_l2(o16.float(), lin) < BF16_TOL
    w4 = (torch.randn(3072, k, device="cuda") * 0.05).bfloat16()
    b4 = torch.randn(3072, device="cuda")
    assert rel_l2(ops.gemm_bias_gelu(a, w4, b4).float(), F.gelu(a.float() @ w4.float().t() + b4, approximate="tanh")) < BF16_TOL
    for ncond in (3, 1):
        gate = torch.randn(ncond, n, device="cuda")
        g = gate.repeat_interleave(T, 0) if ncond > 1 else gate
        assert rel_l2(ops.gemm_bias_gate(a, w, bias, gate, T).float(), g * lin) < BF16_TOL
        # the gated residual update as the GEMM epilogue: fp32 read-modify-write of x, in place (models.py:120-121)
        x0 = torch.randn(m, n, device="cuda")
        x = x0.clone()
        assert ops.gemm_bias_gate_residual(x, a, w, bias, gate, T) is x
        assert rel_l2(x, x0 + g * lin) < F32_TOL
        assert rel_l2(x - x0, g * lin) < 1e-4          # the update itself, not hidden behind |x0|


def test_gemm_192_wide_tiles(ops):
    """M = 4,608 (32 puzzles of 144 tokens - the per-GPU shard of C2 on 8 GPUs): the launcher picks 256 x 192 tiles for qkv,
    proj and fc2 there (gemm.cu: small_tile_width - 72 / 216 tiles on 74 CTA pairs instead of 54 / 162) and keeps 256-wide
    ones for fc1.  Same epilogues, same arithmetic per element; checked against fp32 torch like the wide tiles."""
    torch.manual_seed(5)
    m, T = 4608, 144
    a = torch.randn(m, 768, device="cuda").bfloat16()
    for n in (2304, 3072):
        w = (torch.randn(n, 768, device="cuda") * 0.05).bfloat16()
        b = torch.randn(n, device="cuda")
        lin = a.float() @ w.float().t() + b
        assert rel_l2(ops.gemm_bias(a, w, b).float(), lin) < BF16_TOL
        assert rel_l2(ops.gemm_bias_gelu(a, w, b).float(), F.gelu(lin, approximate="tanh")) < BF16_TOL
    for k in (768, 3072):                                       # proj (eight epilogue warps) and fc2 (four, six stages)
        a2 = torch.randn(m, k, device="cuda").bfloat16()
        w = (torch.randn(768, k, device="cuda") * 0.03).bfloat16()
        b = torch.randn(768, device="cuda")
        lin = a2.float() @ w.float().t() + b
        for ncond in (32, 1):
            gate = torch.randn(ncond, 768, device="cuda")
            g = gate.repeat_interleave(T, 0) if ncond > 1 else gate
            x0 = torch.randn(m, 768, device="cuda")
            x = x0.clone()
            ops.gemm_bias_gate_residual(x, a2, w, b, gate, T)
            assert rel_l2(x - x0, g * lin) < 1e-4


def test_patch_embed_head_and_patchify(ops):
    torch.manual_seed(1)
    img = torch.rand(3, 3, 192, 192, device="cuda") * 2 - 1
    cols = ops.patchify(img)
    ref_cols = img.reshape(3, 3, 12, 16, 12, 16).permute(0, 2, 4, 1, 3, 5).reshape(432, 768)
    assert torch.equal(cols, ref_cols.bfloat16())                                       # pure data movement + rounding: exact
    wp = (torch.randn(768, 768, device="cuda") * 0.05).bfloat16()
    bias, xt = torch.randn(768, device="cuda"), torch.randn(432, 8, device="cuda")
    w_in_t, pos = torch.randn(8, 768, device="cuda") * 0.1, torch.randn(144, 768, device="cuda")
    ref = cols.float() @ wp.float().t() + bias + pos.repeat(3, 1) + xt @ w_in_t
    assert rel_l2(ops.gemm_patch_embed(cols, wp, bias, xt, w_in_t, pos, 144), ref) < F32_TOL
    y = torch.randn(432, 768, device="cuda").bfloat16()
    w1 = (torch.randn(64, 768, device="cuda") * 0.05).bfloat16()
    b1, w2, b2 = torch.randn(64, device="cuda") * 0.1, torch.randn(8, 64, device="cuda") * 0.2, torch.randn(8, device="cuda")
    ref = F.silu(y.float() @ w1.float().t() + b1) @ w2.t() + b2
    assert rel_l2(ops.final_head(y, w1, b1, w2, b2), ref) < F32_TOL
    y32 = torch.randn(432, 768, device="cuda")
    assert torch.equal(ops.unpatchify(y32, 3, 192), y32.reshape(3, 12, 12, 16, 16, 3).permute(0, 5, 1, 3, 2, 4).reshape(3, 3, 192, 192))


@pytest.mark.parametrize("rows,T,ncond", [(432, 144, 3), (432, 144, 1), (27, 9, 3), (1001, 143, 7), (1, 1, 1)])
def test_ln_modulate(ops, rows, T, ncond):
    torch.manual_seed(rows)
    x = torch.randn(rows, 768, device="cuda") * 2 + 0.3
    sh, sc = torch.randn(ncond, 768, device="cuda"), torch.randn(ncond, 768, device="cuda") * 0.5
    idx = torch.arange(rows, device="cuda") // T if ncond > 1 else torch.zeros(rows, dtype=torch.long, device="cuda")
    ref = F.layer_norm(x, (768,), eps=1e-6) * (1 + sc[idx]) + sh[idx]
    got = ops.ln_modulate(x, sh, sc, T).float()
    assert rel_l2(got, ref) < BF16_TOL
    assert (got - ref).abs().max() <= 2.0 ** -8 * ref.abs().max() + 1e-6            # each element within bf16 rounding
    # fused residual add: x += delta (in place, exact fp32 add of the bf16 branch), then the same LN + modulate
    delta = torch.randn(rows, 768, device="cuda").bfloat16()
    x2 = x.clone()
    got2 = ops.ln_modulate(x2, sh, sc, T, delta=delta).float()
    assert torch.equal(x2, x + delta.float())
    ref2 = F.layer_norm(x2, (768,), eps=1e-6) * (1 + sc[idx]) + sh[idx]
    assert rel_l2(got2, ref2) < BF16_TOL


@pytest.mark.parametrize("B,T", [(2, 144), (3, 9), (2, 256), (2, 324), (1, 36), (5, 64), (2, 100), (1, 1), (1, 17),
                                 (1, 144), (7, 144), (31, 144), (1, 256), (5, 256), (13, 324), (40, 324)])   # T in {144, 256, 324}: tcgen05
def test_attention(ops, B, T):
    torch.manual_seed(T)
    qkv = (torch.randn(B * T, 2304, device="cuda") * 1.5).bfloat16()
    q, k, v = qkv.float().reshape(B, T, 3, 12, 64).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, 768)
    assert rel_l2(ops.attention(qkv, B, T).float(), ref) < BF16_TOL


def test_timestep_embed_and_adaln(ops):
    torch.manual_seed(2)
    t = torch.tensor([0, 1, 5, 250, 999, 37, 512, 4, 8, 991, 995], device="cuda")
    w0, b0 = torch.randn(768, 256, device="cuda") * 0.02, torch.randn(768, device="cuda") * 0.02
    w2, b2 = torch.randn(768, 768, device="cuda") * 0.02, torch.randn(768, device="cuda") * 0.02
    freqs = torch.exp(-torch.log(torch.tensor(10000.0)) * torch.arange(128, dtype=torch.float32) / 128).cuda()
    args = t[:, None].float() * freqs[None]
    cref = F.silu(torch.cat([args.cos(), args.sin()], -1) @ w0.t() + b0) @ w2.t() + b2
    c, sc = ops.timestep_embed(t, w0, b0, w2, b2)
    # fp32 throughout; the tolerance covers 1-ulp differences of expf on arguments up to 999 rad
    assert rel_l2(c, cref) < 5e-5 and rel_l2(sc, F.silu(cref)) < 5e-5
    wall = (torch.randn(12 * 4608 + 1536, 768, device="cuda") * 0.05).bfloat16()
    ball = torch.randn(12 * 4608 + 1536, device="cuda")
    for r in (1, 3, 8, 11):
        s = torch.randn(r, 768, device="cuda")
        assert rel_l2(ops.adaln_table(s, wall, ball), s @ wall.float().t() + ball) < F32_TOL


def test_diffusion_elementwise_bit_exact(ops):
    from jpdvt_mt_ntnu_b200.diffusion import create_diffusion
    torch.manual_seed(3)
    d = create_diffusion("250")
    tabs = d.device_tables(torch.device("cuda"))
    x0, xt, nz = (torch.randn(4, 144, 8, device="cuda") for _ in range(3))
    tt = torch.tensor([0, 1, 100, 249], device="cuda")
    g = lambda a: torch.from_numpy(a).cuda()[tt].float().view(-1, 1, 1)               # gaussian_diffusion.py:917-929
    mean, sample = ops.posterior_step(x0, xt, nz, tabs["coef1"], tabs["coef2"], tabs["logvar"], tt)
    mref = g(d.posterior_mean_coef1) * x0 + g(d.posterior_mean_coef2) * xt
    assert torch.equal(mean, mref)
    sref = mref + (tt != 0).float().view(-1, 1, 1) * torch.exp(0.5 * g(d.posterior_log_variance_clipped)) * nz
    assert rel_l2(sample, sref) < 1e-6 and torch.equal(sample[0], mref[0])              # t == 0 adds no noise
    qs = ops.q_sample(x0, nz, tabs["sqrt_ac"], tabs["sqrt_1mac"], tt)
    assert torch.equal(qs, g(d.sqrt_alphas_cumprod) * x0 + g(d.sqrt_one_minus_alphas_cumprod) * nz)
    keep = (torch.rand(4, 144, 8, device="cuda") > 0.5).float()
    qk = ops.q_sample(x0, nz, tabs["sqrt_ac"], tabs["sqrt_1mac"], tt, keep)
    assert torch.equal(qk, qs * (1 - keep) + keep * x0)


def test_ddim_step_vs_oracle(ops):
    """One DDIM update against the oracle restatement (itself pinned to the reference's DDIM code: test_oracle_golden.py)."""
    from jpdvt_mt_ntnu_b200.diffusion import create_diffusion
    from oracle import jpdvt_oracle as orc
    torch.manual_seed(4)
    d, s = create_diffusion("50"), orc.Schedule("50")
    x0, xt, nz = (torch.randn(4, 36, 8) for _ in range(3))
    tt = torch.tensor([0, 1, 25, 49])
    for eta in (0.0, 0.7):
        want = s.ddim_step(lambda c, ts, x: (None, x0), None, xt, tt, nz, eta)["sample"]
        got = ops.ddim_step(x0.cuda(), xt.cuda(), nz.cuda(), d._ddim_tables(torch.device("cuda"), eta), tt.cuda())
        assert rel_l2(got.cpu(), want) < 1e-5


def test_assignment_bit_exact_on_golden_scores(ops, golden):
    g = golden("assignment")
    n = int(g["n"])
    for tag, sentinel in (("1e9", 1e9), ("2024", 2024.0)):
        for size in (9, 16):
            ids = [i for i in range(n) if g[f"scores_{i}"].shape[0] == size]
            sc = torch.from_numpy(np.stack([g[f"scores_{i}"] for i in ids])).cuda()
            order, pred = ops.assign_from_scores(sc, sentinel)
            for row, i in enumerate(ids):
                want = g[f"order_{tag}_{i}"]
                assert order[row].cpu().numpy().tolist() == want.tolist(), (tag, i)
                if sorted(want.tolist()) == list(range(size)):
                    assert pred[row].cpu().numpy().tolist() == g[f"pred_{tag}_{i}"].tolist(), (tag, i)


@pytest.mark.parametrize("G,tok", [(3, 4), (4, 4), (3, 6), (4, 3), (5, 2), (2, 1)])
def test_assignment_from_latents_matches_oracle(ops, G, tok):
    from oracle import jpdvt_oracle as orc
    torch.manual_seed(G * 10 + tok)
    lat = torch.randn(64, G * G * tok * tok, 8)
    canon = torch.from_numpy(orc.sincos_2d(8, G)).float()
    order, pred, scores = ops.assign_greedy_l1(lat.cuda(), canon.cuda(), G, 1e9, True)
    for b in range(64):
        o, p, sc = orc.solve(lat[b], G, tok)
        # the per-slot fp32 mean may differ from torch's by an ulp; scores agree to 1e-6 and random latents have wide margins
        assert np.abs(scores[b].cpu().numpy() - sc).max() < 1e-6
        assert order[b].cpu().numpy().tolist() == list(o) and pred[b].cpu().numpy().tolist() == list(p)


def test_assignment_perfect_latents_round_trip(ops):
    """Size-independent property at full batch: canonical embeddings in scrambled order => pred == scramble."""
    from jpdvt_mt_ntnu_b200 import assignment
    rs = np.random.RandomState(0)
    for G, tok, B in ((3, 4, 256), (4, 4, 128), (3, 6, 128)):
        canon = assignment.canonical_embeddings(G, "cuda")
        perms = np.stack([rs.permutation(G * G) for _ in range(B)])
        te = canon[torch.from_numpy(perms).cuda()]                                       # [B, n, 8]
        lat = te.reshape(B, G, 1, G, 1, 8).expand(B, G, tok, G, tok, 8).reshape(B, -1, 8).contiguous()
        order, pred = assignment.solve_puzzles(lat, G)
        assert np.array_equal(pred.cpu().numpy(), perms)
        ok, matches = assignment.accuracy(pred, torch.from_numpy(perms))
        assert bool(ok.all()) and int(matches.sum()) == B * G * G


@pytest.mark.parametrize("m,k,T", [(432, 768, 144), (1000, 3072, 100), (36864, 768, 144), (300, 768, 9), (256, 768, 256), (77, 3072, 77)])
def test_gemm_residual_layernorm_fused(ops, m, k, T):
    """One `x = x + gate * branch(...)` line plus the modulate(LN(x)) that opens the next one, in one kernel
    (models.py:19-20,120-121): x against fp32 torch on the same bf16 operands, xn within bf16 rounding."""
    torch.manual_seed(m + k)
    n = 768
    a = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(n, device="cuda")
    ncond = (m + T - 1) // T
    for nc in (ncond, 1):
        gate, shift, scale = (torch.randn(nc, n, device="cuda") * 0.5 for _ in range(3))
        idx = torch.arange(m, device="cuda") // T if nc > 1 else torch.zeros(m, dtype=torch.long, device="cuda")
        x0 = torch.randn(m, n, device="cuda") * 2 + 0.7            # non-zero mean: exercises the merged-moment statistics
        x = x0.clone()
        x_out, xn = ops.gemm_bias_gate_residual_ln(x, a, w, bias, gate, shift, scale, T)
        ref_x = x0 + gate[idx] * (a.float() @ w.float().t() + bias)
        ref_xn = F.layer_norm(ref_x, (n,), eps=1e-6) * (1 + scale[idx]) + shift[idx]
        assert x_out is x and rel_l2(x, ref_x) < F32_TOL
        assert rel_l2(xn.float(), ref_xn) < BF16_TOL
        assert (xn.float() - ref_xn).abs().max() < 0.05 * ref_xn.abs().max()


@pytest.mark.parametrize("m,k,T", [(432, 768, 144), (1000, 3072, 100), (36864, 768, 144), (300, 3072, 9), (256, 768, 256), (77, 768, 77)])
def test_layernorm_folded_into_consumer(ops, m, k, T):
    """The sampling loop's LayerNorm fold (csrc/fold.cu): residual GEMM leaves bf16(x) + row sums, the weights are folded
    with (shift, scale), the consumer GEMM finishes the LayerNorm per row.  Against fp32 torch of the reference lines
    modulate(norm(x), shift, scale) -> Linear (models.py:19-20,120-121) within bf16 rounding."""
    torch.manual_seed(m + k + 1)
    n, depth = 768, 2
    a = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(n, device="cuda")
    gate = torch.randn(1, n, device="cuda") * 0.5
    x0 = torch.randn(m, n, device="cuda") * 2 + 0.7                 # non-zero row mean: exercises the rank-one mean correction
    x = x0.clone()
    x_out, xb, stats = ops.gemm_bias_gate_residual_copy(x, a, w, bias, gate, T)
    ref_x = x0 + gate * (a.float() @ w.float().t() + bias)
    assert x_out is x and rel_l2(x, ref_x) < F32_TOL
    assert torch.equal(xb, x.bfloat16())                            # the copy is the rounded residual stream, bit for bit
    tot = stats.sum(1)
    assert torch.allclose(tot[:, 0], x.sum(1), rtol=1e-4, atol=1e-2) and torch.allclose(tot[:, 1], (x * x).sum(1), rtol=1e-4)

    w_qkv = (torch.randn(depth, 2304, 768, device="cuda") * 0.03).bfloat16()
    w_fc1 = (torch.randn(depth, 3072, 768, device="cuda") * 0.03).bfloat16()
    b_qkv, b_fc1 = torch.randn(depth, 2304, device="cuda") * 0.1, torch.randn(depth, 3072, device="cuda") * 0.1
    mod = torch.randn(depth * 6 * 768 + 2 * 768, device="cuda") * 0.5
    wf, u, v = ops.fold_ln_weights(w_qkv, w_fc1, b_qkv, b_fc1, mod)
    for blk in range(depth):
        mm = mod[blk * 4608:(blk + 1) * 4608].reshape(6, 768)
        for name, wt, bt, sh, sc, rows, gelu in (("qkv", w_qkv[blk], b_qkv[blk], mm[0], mm[1], slice(0, 2304), False),
                                                 ("fc1", w_fc1[blk], b_fc1[blk], mm[3], mm[4], slice(2304, 5376), True)):
            assert torch.equal(wf[blk, rows], (wt.float() * (1 + sc)).bfloat16()), name
            assert torch.allclose(u[blk, rows], wf[blk, rows].float().sum(1), rtol=1e-4, atol=1e-3), name
            assert torch.allclose(v[blk, rows], bt + wt.float() @ sh, rtol=1e-4, atol=1e-3), name
            out = ops.gemm_ln_folded(xb, stats, wf[blk, rows].contiguous(), u[blk, rows].contiguous(), v[blk, rows].contiguous(), gelu=gelu)
            ref = (F.layer_norm(ref_x, (n,), eps=1e-6) * (1 + sc) + sh) @ wt.float().t() + bt
            if gelu:
                ref = F.gelu(ref, approximate="tanh")
            assert rel_l2(out.float(), ref) < BF16_TOL, (name, rel_l2(out.float(), ref))


def test_timestep_embed_many_row_groups(ops):
    """More row groups than one wave of CTAs (the per-step table of a 250-step loop, training batches > 96): the hidden
    activations live in a scratch of their own, so CTAs that finish early cannot overwrite what later ones still read."""
    torch.manual_seed(9)
    n = 2048
    t = torch.randint(0, 1000, (n,), device="cuda")
    w0, b0 = torch.randn(768, 256, device="cuda") * 0.05, torch.randn(768, device="cuda") * 0.1
    w2, b2 = torch.randn(768, 768, device="cuda") * 0.05, torch.randn(768, device="cuda") * 0.1
    half = 128
    freqs = torch.exp(-np.log(10000.0) * torch.arange(half, dtype=torch.float32, device="cuda") / half)
    args = t[:, None].float() * freqs[None]
    feat = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    cref = F.silu(feat @ w0.t() + b0) @ w2.t() + b2
    for _ in range(3):
        c, sc = ops.timestep_embed(t, w0, b0, w2, b2)
        assert rel_l2(c, cref) < 1e-5 and rel_l2(sc, F.silu(cref)) < 1e-5


def test_mse_loss_kernels_vs_torch(ops):
    """jpdvt_mse_loss_{fwd,bwd} against mean_flat((target - out)**2 [* (1 - masks)]) and its autograd
    (gaussian_diffusion.py:18-22, 835-838)."""
    torch.manual_seed(11)
    for B, S, G in ((3, 96, 3), (2, 128, 4), (5, 192, 3)):
        T = (S // 16) ** 2
        te_out = torch.randn(B, T, 8, device="cuda", requires_grad=True)
        te_tgt = torch.randn(B, T, 8, device="cuda")
        x_out = torch.randn(B, 3, S, S, device="cuda", requires_grad=True)
        x_tgt = torch.randn(B, 3, S, S, device="cuda")
        keep = (torch.rand(B, G * G, device="cuda") > 0.4).float()
        p = S // G
        full = keep.view(B, 1, G, 1, G, 1).expand(B, 3, G, p, G, p).reshape(B, 3, S, S)
        mf = lambda v: v.mean(dim=list(range(1, v.dim())))
        ref = mf((te_tgt - te_out) ** 2) + mf((x_tgt - x_out) ** 2 * (1 - full))
        dl = torch.randn(B, device="cuda")
        ref.backward(dl)
        got = ops.mse_loss_fwd(te_out.detach(), te_tgt, x_out.detach(), x_tgt, keep, G)
        assert rel_l2(got, ref.detach()) < 1e-6
        d_te, d_img = ops.mse_loss_bwd(dl, te_out.detach(), te_tgt, x_out.detach(), x_tgt, keep, G)
        assert rel_l2(d_te, te_out.grad) < 1e-6 and rel_l2(d_img, x_out.grad) < 1e-6
        only = ops.mse_loss_fwd(te_out.detach(), te_tgt)
        assert rel_l2(only, mf((te_tgt - te_out.detach()) ** 2)) < 1e-6
        d_te2, none = ops.mse_loss_bwd(dl, te_out.detach(), te_tgt)
        assert none is None and rel_l2(d_te2, 2 * (te_out.detach() - te_tgt) * dl.view(-1, 1, 1) / (T * 8)) < 1e-6


def test_philox_step_noise(ops):
    """The in-kernel per-step noise: raw words bit-exact against the numpy Philox4x32-10 restatement, normals against its
    fp64 Box-Muller, unit moments, and posterior_step_philox == posterior_step fed the same numbers."""
    from jpdvt_mt_ntnu_b200.diffusion import create_diffusion
    from oracle import jpdvt_oracle as orc
    seed, call, step, n = 0x1234567ABCDEF, 5, 17, 4096
    key = ops.philox_key(seed, call, "cuda")
    z, raw = ops.philox_normal(key, n, step, return_raw=True)
    groups = np.arange(n // 4, dtype=np.uint64)
    ctr = np.stack([groups, np.zeros_like(groups), np.full_like(groups, step), np.full_like(groups, call)], 1)
    want_raw = orc.philox4x32_10(ctr, (seed & 0xFFFFFFFF, seed >> 32)).reshape(-1)
    assert np.array_equal(raw.cpu().numpy().view(np.uint32), want_raw)
    want = orc.philox_normals(n, seed, call, step)
    assert np.abs(z.cpu().numpy().astype(np.float64) - want).max() < 2e-5
    big = ops.philox_normal(key, 1 << 22, 0)
    assert abs(big.mean().item()) < 3e-3 and abs(big.std().item() - 1.0) < 3e-3
    assert abs((big ** 4).mean().item() - 3.0) < 5e-2                                   # Gaussian kurtosis
    assert not torch.equal(ops.philox_normal(key, n, step + 1), z)                       # another step, other numbers
    assert torch.equal(ops.philox_normal(key, n, step), z)                               # counter-based: repeatable
    d = create_diffusion("250")
    tabs = d.device_tables(torch.device("cuda"))
    x0, xt = torch.randn(4, 144, 8, device="cuda"), torch.randn(4, 144, 8, device="cuda")
    tt = torch.tensor([0, 1, 100, 249], device="cuda")
    eps = ops.philox_normal(key, x0.numel(), step).view_as(x0)
    _, want_s = ops.posterior_step(x0, xt, eps, tabs["coef1"], tabs["coef2"], tabs["logvar"], tt)
    got_s = ops.posterior_step_philox(x0, xt, key, step, tabs["coef1"], tabs["coef2"], tabs["logvar"], tt)
    assert torch.equal(got_s, want_s)

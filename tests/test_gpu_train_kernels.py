"""Per-kernel parity of the training kernels on a B200 against torch autograd (fp32) of the same op."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    from jpdvt_mt_ntnu_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("m,r,c", [(64, 256, 256), (1000, 768, 768), (4608, 2304, 768), (4608, 768, 3072), (432, 64, 768),
                                   (3, 768, 256), (27, 1536, 768), (300, 128, 128), (1, 256, 128)])
def test_wgrad_mn_major_gemm(ops, m, r, c):
    torch.manual_seed(m + r)
    p = torch.randn(m, r, device="cuda").bfloat16()
    q = torch.randn(m, c, device="cuda").bfloat16()
    assert rel_l2(ops.gemm_wgrad(p, q), p.float().t() @ q.float()) < 1e-5       # fp32 accumulate, fp32 out, split-K reduce


@pytest.mark.parametrize("m,k_out,n_in", [(4608, 768, 3072), (1000, 3072, 768), (432, 2304, 768), (300, 64, 768), (128, 768, 768),
                                          (7, 1536, 768)])
def test_dgrad_from_the_untransposed_weight(ops, m, k_out, n_in):
    """dX = dY . W with W [out, in] read as an MN-major tcgen05 operand (no transposed weight copy): fp32, bf16 and dGELU
    outputs against fp32 torch, and bit-identical to the K-major GEMM on the transposed copy (same products, same order)."""
    torch.manual_seed(m + k_out)
    dy = torch.randn(m, k_out, device="cuda").bfloat16()
    w = (torch.randn(k_out, n_in, device="cuda") * 0.05).bfloat16()
    ref = dy.float() @ w.float()
    got = ops.gemm_dgrad(dy, w)
    assert rel_l2(got, ref) < 1e-5
    assert torch.equal(got, ops.gemm_bias_f32(dy, w.t().contiguous(), torch.zeros(n_in, device="cuda")))
    assert rel_l2(ops.gemm_dgrad(dy, w, out_dtype=torch.bfloat16).float(), ref) < 5e-3
    gp = torch.rand(m, n_in, device="cuda").bfloat16()
    dh = ops.gemm_dgrad(dy, w, gprime=gp)
    assert rel_l2(dh.float(), ref * gp.float()) < 5e-3
    dh2, cs = ops.gemm_dgrad(dy, w, gprime=gp, with_colsum=True)      # bias gradient of the layer below, from the same epilogue
    assert torch.equal(dh2, dh) and rel_l2(cs, dh.float().sum(0)) < 1e-5


def test_dgelu_gate_ln_colsum(ops):
    torch.manual_seed(1)
    m, T = 432, 144
    a = torch.randn(m, 768, device="cuda").bfloat16()
    w = (torch.randn(3072, 768, device="cuda") * 0.05).bfloat16()
    gp = torch.rand(m, 3072, device="cuda").bfloat16()
    assert rel_l2(ops.gemm_dgelu(a, w, gp).float(), (a.float() @ w.float().t()) * gp.float()) < 5e-3
    dx = torch.randn(m, 768, device="cuda")
    y = torch.randn(m, 768, device="cuda").bfloat16()
    gate = torch.randn(3, 768, device="cuda")
    dy, dgate, dbias = ops.gate_bwd(dx, y, gate, T)
    gfull = gate.repeat_interleave(T, 0)
    assert rel_l2(dy.float(), gfull * dx) < 5e-3
    assert rel_l2(dgate, (dx * y.float()).reshape(3, T, 768).sum(1)) < 1e-5
    assert rel_l2(dbias, (gfull * dx).sum(0)) < 1e-5
    xx = (torch.randn(m, 768, device="cuda") * 2 + 0.3).requires_grad_(True)
    shift = torch.randn(3, 768, device="cuda", requires_grad=True)
    scale = (torch.randn(3, 768, device="cuda") * 0.5).requires_grad_(True)
    idx = torch.arange(m, device="cuda") // T
    dxn = torch.randn(m, 768, device="cuda")
    (F.layer_norm(xx, (768,), eps=1e-6) * (1 + scale[idx]) + shift[idx]).backward(dxn)
    base = torch.randn(m, 768, device="cuda")
    got_dx, dsh, dsc, dxb = ops.ln_modulate_bwd(xx.detach(), dxn, scale.detach(), T, dx=base.clone())
    assert rel_l2(got_dx, base + xx.grad) < 1e-5 and rel_l2(dsh, shift.grad) < 1e-5 and rel_l2(dsc, scale.grad) < 1e-5
    assert rel_l2(dxb.float(), base + xx.grad) < 5e-3
    assert rel_l2(ops.ln_modulate_bwd(xx.detach(), dxn, scale.detach(), T)[0], xx.grad) < 1e-5
    assert rel_l2(ops.colsum(gp), gp.float().sum(0)) < 1e-5 and rel_l2(ops.colsum(dx), dx.sum(0)) < 1e-5


@pytest.mark.parametrize("B,T", [(3, 144), (128, 144), (1, 9), (5, 324), (40, 36)])
def test_ln_gate_bwd_fused(ops, B, T):
    """The fused kernel of the training backward (LayerNorm-modulate backward + the gate backward below it, models.py:19-20,
    120-121) against torch autograd of both ops, with and without the gate half / the accumulate flag."""
    torch.manual_seed(B * 1000 + T)
    m = B * T
    xx = (torch.randn(m, 768, device="cuda") * 2 + 0.3).requires_grad_(True)
    shift = torch.randn(B, 768, device="cuda", requires_grad=True)
    scale = (torch.randn(B, 768, device="cuda") * 0.5).requires_grad_(True)
    idx = torch.arange(m, device="cuda") // T
    dxn = torch.randn(m, 768, device="cuda")
    (F.layer_norm(xx, (768,), eps=1e-6) * (1 + scale[idx]) + shift[idx]).backward(dxn)
    base = torch.randn(m, 768, device="cuda")
    y = torch.randn(m, 768, device="cuda").bfloat16()
    gate = torch.randn(B, 768, device="cuda")
    dx, dsh, dsc, dxb, dy, dgate, dbias = ops.ln_gate_bwd(xx.detach(), dxn, scale.detach(), T, dx=base.clone(), y=y, gate=gate)
    want = base + xx.grad
    assert rel_l2(dx, want) < 1e-5 and rel_l2(dsh, shift.grad) < 1e-5 and rel_l2(dsc, scale.grad) < 1e-5
    assert rel_l2(dxb.float(), want) < 5e-3
    assert rel_l2(dy.float(), gate[idx] * want) < 5e-3
    assert rel_l2(dgate, (want * y.float()).reshape(B, T, 768).sum(1)) < 1e-5
    assert rel_l2(dbias, (gate[idx] * want).sum(0)) < 1e-5
    # the two halves run separately give the same numbers (same arithmetic per element)
    ref_dx, ref_dsh, ref_dsc, _ = ops.ln_modulate_bwd(xx.detach(), dxn, scale.detach(), T, dx=base.clone())
    assert torch.equal(dx, ref_dx)
    assert torch.equal(dy, ops.gate_bwd(ref_dx, y, gate, T)[0])
    only = ops.ln_gate_bwd(xx.detach(), dxn, scale.detach(), T)
    assert rel_l2(only[0], xx.grad) < 1e-5 and only[4] is None


@pytest.mark.parametrize("B,T", [(2, 144), (1, 144), (13, 144), (40, 144), (3, 9), (2, 256), (13, 256), (40, 256), (1, 324), (5, 324),
                                 (13, 324), (2, 100), (1, 36)])
def test_attention_backward(ops, B, T):
    torch.manual_seed(T)
    qkv = (torch.randn(B * T, 2304, device="cuda") * 1.2).bfloat16()
    d_o = torch.randn(B * T, 768, device="cuda").bfloat16()
    x = qkv.float().requires_grad_(True)
    q, k, v = x.reshape(B, T, 3, 12, 64).permute(2, 0, 3, 1, 4)
    F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, 768).backward(d_o.float())
    o, lse = ops.attention(qkv, B, T, return_lse=True)
    s = (q @ k.transpose(-1, -2)).detach() * 0.125
    assert rel_l2(lse, torch.logsumexp(s, -1) * 1.4426950408889634) < 1e-5
    dqkv = ops.attention_bwd(qkv, o, d_o, lse, B, T)
    for lo in (0, 768, 1536):
        assert rel_l2(dqkv[:, lo:lo + 768].float(), x.grad[:, lo:lo + 768]) < 6e-3
    # the qkv bias gradient folded into the epilogue: column sums of exactly the bf16 values written
    dqkv2, dbias = ops.attention_bwd(qkv, o, d_o, lse, B, T, with_bias_grad=True)
    assert torch.equal(dqkv2, dqkv)
    assert rel_l2(dbias, dqkv.float().sum(0)) < 1e-5

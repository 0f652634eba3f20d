import sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo')
from jpdvt_mt_ntnu_b200 import ops
for (B, T) in [(1, 144), (3, 144), (7, 144), (2, 256), (256, 144)]:
    torch.manual_seed(T + B)
    qkv = (torch.randn(B * T, 2304, device="cuda") * 1.5).bfloat16()
    q, k, v = qkv.float().reshape(B, T, 3, 12, 64).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, 768)
    got = ops.attention(qkv, B, T).float()
    torch.cuda.synchronize()
    err = ((got - ref).norm() / ref.norm()).item()
    # per (sample, head, row-block) error map to localise layout bugs
    e = (got - ref).reshape(B, T, 12, 64).pow(2).sum(-1).sqrt() / ref.reshape(B, T, 12, 64).pow(2).sum(-1).sqrt()
    bad = (e > 2e-2).nonzero()
    print(f"B={B} T={T} rel_l2={err:.3e} bad_rows={bad.shape[0]}", bad[:8].tolist(), flush=True)

import os, sys, torch
os.environ["JPDVT_ATTN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jpdvt_mt_ntnu_b200 import ops
B, T = int(os.environ.get('BATCH', '256')), int(os.environ.get('TOKENS', '144'))
qkv = (torch.randn(B * T, 2304, device="cuda") * 1.5).bfloat16()
ops.attention(qkv, B, T)
torch.cuda.synchronize()

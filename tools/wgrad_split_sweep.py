#!/usr/bin/env python3
"""Weight-gradient GEMM time against the number of contraction splits (JPDVT_WGRAD_SPLIT is read once per process, so
the driver loop below starts one process per value): the four shapes of a DiT block at M = BATCH x 144 rows.
    python tools/wgrad_split_sweep.py            # sweep 0 (= the costed plan), 1..16
"""
import os
import subprocess
import sys

if os.environ.get("_SWEEP_CHILD") != "1":
    for s in [0] + list(range(1, 17)):
        env = dict(os.environ, _SWEEP_CHILD="1")
        if s:
            env["JPDVT_WGRAD_SPLIT"] = str(s)
        else:
            env.pop("JPDVT_WGRAD_SPLIT", None)
        out = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
        print(f"split={s or 'plan':>4}: " + (out.stdout.strip() or out.stderr.strip()[-300:]), flush=True)
    sys.exit(0)

import torch                                                       # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jpdvt_mt_ntnu_b200 import ops                                 # noqa: E402

M = int(os.environ.get("BATCH", "128")) * int(os.environ.get("TOKENS", "144"))
dev = torch.device("cuda")
torch.manual_seed(0)
bf = lambda *s: (torch.randn(*s, device=dev) * 0.1).bfloat16()
shapes = {"qkv": (2304, 768), "proj": (768, 768), "fc1": (3072, 768), "fc2": (768, 3072)}
res = []
for name, (o, i) in shapes.items():
    p, q = bf(M, o), bf(M, i)
    for _ in range(3):
        ops.gemm_wgrad(p, q)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm_wgrad(p, q)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    res.append(f"{name} {best * 1e3:6.1f} us")
print("   ".join(res))

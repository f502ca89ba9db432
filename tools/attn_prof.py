#!/usr/bin/env python
"""One fused-attention shape, a few launches: the target of `ncu --set full` captures (tools/gpu_r02_attn4.sh).
Usage: python tools/attn_prof.py [rows=32] [d=40] [T=4096] [kernel=0]   (kernel: see ops.attention / mobi_attention)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobi_b200 import ops  # noqa: E402

rows, d, T, kern = (int(x, 0) for x in (sys.argv[1:5] + ["32", "40", "4096", "0"][len(sys.argv) - 1:]))
H = 8
q, k, v = (torch.randn(rows * H, T, d, device="cuda").to(torch.bfloat16) for _ in range(3))
out = torch.empty(rows, T, H * d, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    ops.attention(q, k, v, rows, H, d, T, T, out=out, kernel=kern, v_rowmajor=True)
torch.cuda.synchronize()
print("ok", rows, d, T, hex(kern))

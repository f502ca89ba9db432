"""Attention-backward micro-benchmark at the L0 self-attention shape of the training step (B = 4 rows, 8 heads, d = 40,
T = 4096): statistics pass + flash kernels vs the materialised-tile path, CUDA events, warm.
Usage: python tools/attn_bwd_bench.py [--iters 5]"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobi_b200.training import UNetTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--shape", type=int, nargs=4, default=[4, 8, 40, 4096], metavar=("B", "H", "D", "T"))
a = ap.parse_args()
B, H, D, T = a.shape
C = H * D
g = torch.Generator(device="cuda").manual_seed(0)
q = (torch.randn(B * H, T, D, device="cuda", generator=g) * D ** -0.5 * math.log2(math.e)).to(torch.bfloat16)
k = torch.randn(B * H, T, D, device="cuda", generator=g).to(torch.bfloat16)
v = torch.randn(B * H, T, D, device="cuda", generator=g).to(torch.bfloat16)
do = torch.randn(B * T, C, device="cuda", generator=g).to(torch.bfloat16)
tr = UNetTrainer.__new__(UNetTrainer)
tr.device, tr._ws, tr.lse_backward = torch.device("cuda"), {}, False
out = torch.empty(B * T, 3 * C, device="cuda", dtype=torch.bfloat16)
for flash in (True, False):
    tr.flash_backward = flash
    for _ in range(2):
        tr._attn_bwd(q, k, v, do, B, H, D, T, out, 0, out, C, 2 * C)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        tr._attn_bwd(q, k, v, do, B, H, D, T, out, 0, out, C, 2 * C)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    fl = 5 * 2.0 * B * H * T * T * D
    print("%s: %.3f ms  (%.0f TFLOP/s on the 5 algorithmic products)" % ("flash" if flash else "tiles+gemm", ms, fl / ms / 1e9))

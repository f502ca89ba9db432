"""LayerNorm-backward micro-benchmark (events, warm, inputs re-used: L2-resident above ~60 MB is not possible, so sizes
at L0 stream from HBM).  Usage: python tools/ln_bwd_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobi_b200 import train_ops as tops  # noqa: E402

dev = "cuda:0"
for rows, C in ((16384, 320), (8192, 320), (4096, 640), (1024, 1280)):
    for train in (False, True):
        for dy_dt in (torch.bfloat16, torch.float32):
            x = torch.randn(rows, C, device=dev)
            dy = torch.randn(rows, C, device=dev).to(dy_dt)
            g = torch.randn(C, device=dev)
            dx = torch.zeros(rows, C, device=dev)
            dg, db = (torch.zeros(C, device=dev), torch.zeros(C, device=dev)) if train else (None, None)
            for _ in range(3):
                tops.layernorm_bwd(x, g, dy, dx, dgamma=dg, dbeta=db)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                tops.layernorm_bwd(x, g, dy, dx, dgamma=dg, dbeta=db)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            nbytes = rows * C * (4 + dy.element_size() + 8)
            print("rows %6d C %4d dgamma %d dy %s: %7.1f us  %6.0f GB/s" % (
                rows, C, int(train), "f32" if dy_dt == torch.float32 else "bf16", us, nbytes / us / 1e3))

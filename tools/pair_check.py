#!/usr/bin/env python
"""Bring-up check of the CTA-pair (cta_group::2) GEMM / conv variant against torch, then a timing A/B."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobi_b200 import ops  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


torch.manual_seed(0)
for (M, N, K, tile) in [(512, 256, 1024, 0), (256, 160, 128, 160), (1000, 320, 1280, 160), (8192, 1280, 1280, 256),
                        (384, 64, 640, 64)]:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    ref = a.float() @ w.float().t() + bias + res
    out = ops.gemm(a, w, bias=bias, residual=res, out_dtype=torch.float32, pair=1, tile_n=tile)
    torch.cuda.synchronize()
    print("gemm pair M=%d N=%d K=%d tile=%d: rel err %.3e" % (M, N, K, tile, rel(out, ref)), flush=True)

for (R, C, Co, S) in [(2, 64, 128, 16), (4, 320, 320, 64), (3, 640, 640, 32)]:
    x = torch.randn(R, S, S, C, device="cuda").bfloat16()
    w4 = torch.randn(Co, C, 3, 3, device="cuda") * (9 * C) ** -0.5
    wp = w4.permute(0, 2, 3, 1).reshape(Co, 9 * C).bfloat16().contiguous()
    bias = torch.randn(Co, device="cuda")
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wp.float().reshape(Co, 3, 3, C).permute(0, 3, 1, 2),
                                     bias, padding=1).permute(0, 2, 3, 1)
    out = ops.conv_implicit(x, wp, 3, 3, 1, 1, bias=bias, pair=1)
    torch.cuda.synchronize()
    print("conv pair R=%d %d->%d @%d: rel err %.3e" % (R, C, Co, S, rel(out, ref)), flush=True)

R = 32
for (C, Co, S) in ((320, 320, 64), (640, 640, 32), (1280, 1280, 16), (2560, 1280, 16), (960, 320, 64)):
    x = torch.randn(R, S, S, C, device="cuda").bfloat16()
    w = (torch.randn(Co, 9 * C, device="cuda") * (9 * C) ** -0.5).bfloat16()
    bias = torch.randn(Co, device="cuda")
    res = torch.randn(R, S, S, Co, device="cuda")
    out = torch.empty(R, S, S, Co, device="cuda")
    fl = 2.0 * R * S * S * Co * 9 * C
    for pair in (-1, 1):
        ms = timeit(lambda: ops.conv_implicit(x, w, 3, 3, 1, 1, bias=bias, residual=res, out=out, pair=pair))
        print("conv3x3 %d->%d @%d pair=%2d  %.3f ms  %.1f TF/s" % (C, Co, S, pair, ms, fl / ms / 1e9), flush=True)
for (M, N, K) in ((32768, 5120, 640), (32768, 640, 2560), (8192, 10240, 1280), (8192, 1280, 5120), (8192, 3840, 1280)):
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for pair in (-1, 1):
        ms = timeit(lambda: ops.gemm(a, w, out=o, pair=pair))
        print("gemm M=%d N=%d K=%d pair=%2d  %.3f ms  %.1f TF/s" % (M, N, K, pair, ms, 2.0 * M * N * K / ms / 1e9), flush=True)

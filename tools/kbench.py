#!/usr/bin/env python
"""Kernel micro-benchmarks at the shapes of one mobi_nusc_512 UNet call (32 rows = 8 joint samples x CFG).

Times each C-ABI kernel in isolation with CUDA events (3 warm-up + `reps` timed launches, inputs cycled through
enough buffers to exceed the 126 MB L2 where the real call would also miss) and prints achieved TFLOP/s or GB/s.
Usage: python tools/kbench.py [attn] [gemm] [conv] [ln] [gn]  -> also appends JSON lines to gpurun_out/kbench.jsonl
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mobi_b200 import _lib as L  # noqa: E402
from mobi_b200 import ops  # noqa: E402

R = int(os.environ.get("KB_ROWS", "32"))
GK = int(os.environ.get("KB_GEMM_KERNEL", "0"))  # 1 = force the one-tile GEMM kernel
GT = int(os.environ.get("KB_GEMM_TILE", "0"))    # 0 = the library's choice, else tile_n (64 / 128 / 160 / 256)
GP = int(os.environ.get("KB_GEMM_PAIR", "0"))    # 0 = the library's choice, 1 = CTA pairs, -1 = never
OUT = os.path.join(ROOT, "gpurun_out", "kbench.jsonl")
os.makedirs(os.path.dirname(OUT), exist_ok=True)


def timeit(fn, reps=int(os.environ.get("KB_REPS", "10"))):
    for _ in range(int(os.environ.get("KB_WARMUP", "3"))):     # KB_WARMUP=0 KB_REPS=1: one launch per shape (ncu captures)
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms, flops=0.0, nbytes=0.0, **kw):
    rec = dict(name=name, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1) if flops else None,
               gbs=round(nbytes / ms / 1e6, 1) if nbytes else None, **kw)
    print("%-46s %8.3f ms  %8s TF/s  %8s GB/s" % (name, ms, rec["tflops"], rec["gbs"]), flush=True)
    with open(OUT, "a") as f:
        f.write(json.dumps(rec) + "\n")


def rnd(*shape, dtype=torch.bfloat16, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(dtype)


def bench_attn():
    for (rows, H, D, T, tag) in [(R, 8, 40, 4096, "self L0"), (R // 2, 8, 40, 4096, "cross L0"), (R, 8, 80, 1024, "self L1"),
                                 (R, 8, 160, 256, "self L2")]:
        q, k = rnd(rows * H, T, D), rnd(rows * H, T, D)
        vt = rnd(rows * H, D, T)
        vrow = rnd(rows * H, T, D)
        out = torch.empty(rows, T, H * D, device="cuda", dtype=torch.bfloat16)
        fl = 4.0 * rows * H * T * T * D
        # 0x100: row-major V, attention4 (P in TMEM); 0x1N0: its exp-on-FMA-pipe shares; 0x103 / 0x104: attention3 (4 / 2 tiles)
        # 0x105: attention4's first TMEM plan (NT = 4, P over S: S(j + 1) after PV(j))
        # 0x106 / 0x1N6: attention4 with the split softmax (two warps per TMEM lane group)
        kerns = (1,) if D > 128 else ((0x100, 0x106, 0x116, 0x126, 0x146, 0x105, 0x103) if D <= 48 else (0x100, 0x106, 0x103, 1))
        for kern in kerns:
            rowv = kern >= 0x100
            ms = timeit(lambda: ops.attention(q, k, vrow if rowv else vt, rows, H, D, T, T, out=out, kernel=kern & 0xff,
                                              v_rowmajor=rowv))
            report("attention %s d=%d T=%d kernel=0x%x" % (tag, D, T, kern), ms, fl, clk_per_tile=round(
                ms * 1e-3 * 1.9e9 * 148 / (rows * H * (T / 128) ** 2), 0))


def bench_gemm():
    shapes = []
    for (C, T) in ((320, 4096), (640, 1024), (1280, 256)):
        M = R * T
        shapes += [("qkv C=%d" % C, M, 3 * C, C, "qkv"), ("to_out+res C=%d" % C, M, C, C, "res"),
                   ("ff1 geglu C=%d" % C, M, 8 * C, C, "geglu"), ("ff2+res C=%d" % C, M, C, 4 * C, "res"),
                   ("proj_in C=%d" % C, M, C, C, "f32")]
    for name, M, N, K, kind in shapes:
        a, w = rnd(M, K), rnd(N, K, scale=K ** -0.5)
        bias = rnd(N, dtype=torch.float32)
        fl = 2.0 * M * N * K
        if kind == "qkv":
            C = K
            H, D, T = 8, C // 8, M // R
            q = torch.empty(R * H, T, D, device="cuda", dtype=torch.bfloat16)
            k = torch.empty_like(q)
            vt = torch.empty(R * H, D, T, device="cuda", dtype=torch.bfloat16)
            vr = torch.empty_like(q)
            ep = L.EPI_QKV_ROW if D <= 128 else L.EPI_QKV
            fn = lambda: ops.gemm(a, w, epilogue=ep, heads=H, head_dim=D, tokens=T, out=q, out2=k,
                                  out3=vr if D <= 128 else vt, kernel=GK, tile_n=GT, pair=GP if GP != 2 else 0)
            nb = M * K * 2 + M * N * 2
        elif kind == "res":
            x = rnd(M, N, dtype=torch.float32)
            fn = lambda: ops.gemm(a, w, bias=bias, residual=x, out=x, kernel=GK, tile_n=GT, pair=GP)
            nb = M * K * 2 + M * N * 8
        elif kind == "geglu":
            from mobi_b200.packing import interleave_geglu_pairs
            w2, b2 = interleave_geglu_pairs(w.float(), bias)
            w2 = w2.to(torch.bfloat16)
            o = torch.empty(M, N // 2, device="cuda", dtype=torch.bfloat16)
            fn = lambda: ops.gemm(a, w2, bias=b2, epilogue=L.EPI_GEGLU2, out=o, kernel=GK, tile_n=GT, pair=GP if GP != 2 else 0)
            nb = M * K * 2 + M * N
        else:
            o = torch.empty(M, N, device="cuda", dtype=torch.float32)
            fn = lambda: ops.gemm(a, w, bias=bias, out=o, kernel=GK, tile_n=GT, pair=GP)
            nb = M * K * 2 + M * N * 4
        report("gemm %s M=%d N=%d K=%d" % (name, M, N, K), timeit(fn), fl, nb)


def bench_conv():
    for (C, Co, S) in ((320, 320, 64), (640, 640, 32), (1280, 1280, 16), (1280, 1280, 8), (2560, 1280, 16), (960, 320, 64)):
        x = rnd(R, S, S, C)
        w = rnd(Co, 9 * C, scale=(9 * C) ** -0.5)
        bias = rnd(Co, dtype=torch.float32)
        res = rnd(R, S, S, Co, dtype=torch.float32)
        out = torch.empty(R, S, S, Co, device="cuda", dtype=torch.float32)
        fl = 2.0 * R * S * S * Co * 9 * C
        ms = timeit(lambda: ops.conv_implicit(x, w, 3, 3, 1, 1, bias=bias, residual=res, out=out, kernel=GK, tile_n=GT, pair=GP))
        report("conv3x3 %d->%d @%d" % (C, Co, S), ms, fl, x.numel() * 2 + out.numel() * 8)


def bench_ln():
    for (C, T) in ((320, 4096), (640, 1024), (1280, 256)):
        x = rnd(R * T, C, dtype=torch.float32)
        g, b = rnd(C, dtype=torch.float32), rnd(C, dtype=torch.float32)
        vec = rnd(R, C, dtype=torch.float32)
        E = R * T * C
        report("layernorm C=%d" % C, timeit(lambda: ops.layernorm(x, g, b)), 0, E * 6)
        report("layernorm+vec C=%d" % C, timeit(lambda: ops.layernorm(x, g, b, add_vec=vec, add_rows_per_vec=T)), 0, E * 10)
        report("ln_dual C=%d" % C, timeit(lambda: ops.ln_dual(x, R, T, [(ops.LN_NORM, g, b), (ops.LN_CAST, None, None)])), 0, E * 6)
        Ug, Z = rnd(R, 16, C, dtype=torch.float32, scale=C ** -0.5), rnd(R, 16, C, dtype=torch.float32, scale=0.01)
        sb, zb = rnd(R, 16, dtype=torch.float32), rnd(C, dtype=torch.float32, scale=0.01)
        report("ln_adapter C=%d" % C, timeit(lambda: ops.ln_adapter(
            x, R, T, g, b, Ug, sb, Z, zb, [(ops.LN_NORM, g, b), (ops.LN_CAST, None, None)], pair=True, add_vec=vec)), 0, E * 10)
        xn = rnd(R * T, C)
        U4, Z4 = Ug.reshape(R, 2, 8, C), Z.reshape(R, 2, 8, C)
        report("ctx_attention(old) C=%d" % C, timeit(lambda: ops.ctx_attention(xn, U4, Z4, zb, x, R, T, 8, 2)), 0, E * 10)


def bench_gn():
    for (C1, C2, S) in ((320, 0, 64), (640, 320, 64), (640, 0, 32), (1280, 640, 32), (1280, 1280, 16), (1280, 0, 8)):
        x1 = rnd(R, S, S, C1, dtype=torch.float32)
        x2 = rnd(R, S, S, C2, dtype=torch.float32) if C2 else None
        C = C1 + C2
        g, b = rnd(C, dtype=torch.float32), rnd(C, dtype=torch.float32)
        E = R * S * S * C
        report("groupnorm+silu C=%d+%d @%d" % (C1, C2, S), timeit(lambda: ops.groupnorm(x1, g, b, 1e-5, x2=x2)), 0, E * 10)
        report("groupnorm+silu C=%d+%d @%d two-pass" % (C1, C2, S),
               timeit(lambda: ops.groupnorm(x1, g, b, 1e-5, x2=x2, two_pass=True)), 0, E * 10)
        # producer-side statistics (timing only: the statistics buffers hold noise)
        x1._colstats = torch.rand(2 * (R * S * S // 32) * C1, device="cuda")
        if x2 is not None:
            x2._colstats = torch.rand(2 * (R * S * S // 32) * C2, device="cuda")
        report("groupnorm+silu C=%d+%d @%d producer statistics" % (C1, C2, S),
               timeit(lambda: ops.groupnorm(x1, g, b, 1e-5, x2=x2)), 0, E * 6)
        del x1._colstats


if __name__ == "__main__":
    which = sys.argv[1:] or ["attn", "gemm", "conv", "ln", "gn"]
    print("device:", torch.cuda.get_device_name(0), "rows", R)
    for w in which:
        globals()["bench_" + w]()

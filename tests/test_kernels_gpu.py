"""Kernel-level parity: every CUDA kernel, called through the C ABI, against plain PyTorch fp32 math on the
same seeded inputs.  Tolerances: bf16 operands with fp32 accumulation -> relative-to-max error <= 1e-2
(typically ~3e-3); fp32 elementwise kernels <= 1e-5."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from mobi_b200 import ops
    return ops


def _L():
    from mobi_b200 import _lib
    return _lib


def relerr(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def describe(a, b):
    """Where does a mismatch live? (row/col pattern helps to find layout bugs)"""
    d = (a.float() - b.float()).abs()
    d2 = d.reshape(-1, d.shape[-1])
    rows = d2.max(dim=1).values
    cols = d2.max(dim=0).values
    br = torch.nonzero(rows > 0.05 * d.max()).flatten()[:16].tolist()
    bc = torch.nonzero(cols > 0.05 * d.max()).flatten()[:16].tolist()
    return "max %.4g ref-max %.4g bad rows %s bad cols %s" % (d.max().item(), b.float().abs().max().item(), br, bc)


def rnd(*shape, seed=0, dtype=torch.bfloat16, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to("cuda").to(dtype)


@pytest.mark.parametrize("M,N,K,tile_n", [
    (128, 64, 64, 64), (128, 128, 64, 128), (256, 256, 128, 256), (128, 160, 64, 160),
    (512, 320, 320, 0), (1000, 320, 320, 64), (4096, 640, 640, 0), (300, 1280, 1280, 256),
    (4, 1280, 320, 0), (260, 4, 2880, 64), (128, 128, 88, 128), (1024, 2560, 320, 0),
    (128, 23680, 1280, 0), (384, 12800, 1280, 160),      # CTA pairs on an odd number of 128-row tiles (the embedding GEMM)
])
@pytest.mark.parametrize("kernel", [0, 1])
def test_gemm_plain(M, N, K, tile_n, kernel):
    """kernel 0 = automatic (persistent kernel when its vector epilogue applies), 1 = one-tile kernel."""
    ops = _ops()
    a = rnd(M, K, seed=1)
    w = rnd(N, K, seed=2, scale=K ** -0.5)
    bias = rnd(N, seed=3, dtype=torch.float32)
    ref = a.float() @ w.float().t() + bias
    out = ops.gemm(a, w, bias=bias, out_dtype=torch.float32, tile_n=tile_n, kernel=kernel)
    torch.cuda.synchronize()
    assert relerr(out, ref) < 2e-3, describe(out, ref)
    out16 = ops.gemm(a, w, bias=bias, out_dtype=torch.bfloat16, tile_n=tile_n, kernel=kernel)
    assert relerr(out16, ref) < 1e-2, describe(out16, ref)


@pytest.mark.parametrize("Rh,T,C,K", [(2, 256, 320, 320), (3, 1024, 640, 640), (2, 64, 1280, 1280), (4, 128, 320, 1280)])
@pytest.mark.parametrize("off", [0, 1])
def test_gemm_row_segments_in_place(Rh, T, C, K, off):
    """The cross-modal output projection (attention.py:255-262): rows of one modality, T at a time, land in every other
    T-row segment of the interleaved residual stream, which is also the residual (in place).  T % 128 == 0 takes the
    straight-line epilogue (+ the L2 prefetch of the residual boxes), T = 64 the general one."""
    ops = _ops()
    a = rnd(Rh * T, K, seed=1)
    w = rnd(C, K, seed=2, scale=K ** -0.5)
    bias = rnd(C, seed=3, dtype=torch.float32)
    x = rnd(2 * Rh * T, C, seed=4, dtype=torch.float32)
    ref = x.clone().reshape(Rh, 2, T, C)
    ref[:, off] += (a.float() @ w.float().t() + bias).reshape(Rh, T, C)
    ops.gemm(a, w, bias=bias, residual=x, out=x, ldo=C, out_seg=T, out_seg_stride=2 * T, out_seg_offset=off * T)
    torch.cuda.synchronize()
    assert relerr(x, ref.reshape(2 * Rh * T, C)) < 2e-3, describe(x, ref.reshape(2 * Rh * T, C))
    assert torch.equal(x.reshape(Rh, 2, T, C)[:, 1 - off], ref[:, 1 - off])       # the other modality's rows are untouched


def test_gemm_residual_rowbias_act():
    ops = _ops()
    M, N, K = 2 * 384, 320, 640
    a = rnd(M, K, seed=1)
    w = rnd(N, K, seed=2, scale=K ** -0.5)
    bias = rnd(N, seed=3, dtype=torch.float32)
    rb = rnd(2, 3 * N, seed=4, dtype=torch.float32)  # wider row stride than N: slice [N:2N]
    res = rnd(M, N, seed=5, dtype=torch.float32)
    ref = a.float() @ w.float().t() + bias + rb[:, N:2 * N].repeat_interleave(384, 0) + res
    out = ops.gemm(a, w, bias=bias, row_bias=rb[:, N:2 * N], rows_per_group=384, ld_row_bias=3 * N, residual=res,
                   out_dtype=torch.float32)
    assert relerr(out, ref) < 2e-3, describe(out, ref)
    # in-place residual (out aliases residual), bf16 residual
    res16 = res.to(torch.bfloat16)
    ref2 = a.float() @ w.float().t() + res16.float()
    out2 = ops.gemm(a, w, residual=res16, out=res16.clone(), out_dtype=torch.bfloat16)
    assert relerr(out2, ref2) < 1e-2, describe(out2, ref2)
    x = res.clone()
    ops.gemm(a, w, residual=x, out=x)
    assert relerr(x, a.float() @ w.float().t() + res) < 2e-3
    # SiLU activation
    ref3 = torch.nn.functional.silu(a.float() @ w.float().t() + bias)
    out3 = ops.gemm(a, w, bias=bias, act=1, out_dtype=torch.float32)
    assert relerr(out3, ref3) < 2e-3, describe(out3, ref3)


def test_gemm_geglu():
    ops = _ops()
    L = _L()
    M, C = 512, 320
    a = rnd(M, C, seed=1)
    w = rnd(8 * C, C, seed=2, scale=C ** -0.5)      # reference layout: [value rows | gate rows]
    b = rnd(8 * C, seed=3, dtype=torch.float32)
    h = a.float() @ w.float().t() + b
    val, gate = h.chunk(2, dim=-1)
    ref = val * torch.nn.functional.gelu(gate)
    from mobi_b200.packing import interleave_geglu, interleave_geglu_pairs
    wp, bp = interleave_geglu(w, b)
    out = ops.gemm(a, wp, bias=bp, epilogue=L.EPI_GEGLU)
    assert out.shape == (M, 4 * C)
    assert relerr(out, ref) < 1e-2, describe(out, ref)
    # (value, gate) pair layout: what the transformer blocks use; both kernels
    w2, b2 = interleave_geglu_pairs(w, b)
    for kernel in (0, 1):
        out2 = ops.gemm(a, w2, bias=b2, epilogue=L.EPI_GEGLU2, kernel=kernel)
        assert out2.shape == (M, 4 * C)
        assert relerr(out2, ref) < 1e-2, describe(out2, ref)
    out1 = ops.gemm(a, wp, bias=bp, epilogue=L.EPI_GEGLU, kernel=1)
    assert relerr(out1, ref) < 1e-2, describe(out1, ref)


@pytest.mark.parametrize("B,T,H,D", [(2, 256, 8, 40), (1, 128, 2, 16), (2, 64, 8, 160), (1, 1024, 8, 80)])
def test_gemm_head_layouts(B, T, H, D):
    ops = _ops()
    L = _L()
    C = H * D
    a = rnd(B * T, C, seed=1)
    w = rnd(3 * C, C, seed=2, scale=C ** -0.5)
    ref = (a.float() @ w.float().t()).reshape(B, T, 3, H, D)
    q = torch.empty(B * H, T, D, device="cuda", dtype=torch.bfloat16)
    k = torch.empty_like(q)
    vt = torch.empty(B * H, D, T, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, epilogue=L.EPI_QKV, heads=H, head_dim=D, tokens=T, out=q, out2=k, out3=vt)
    rq = ref[:, :, 0].permute(0, 2, 1, 3).reshape(B * H, T, D)
    rk = ref[:, :, 1].permute(0, 2, 1, 3).reshape(B * H, T, D)
    rv = ref[:, :, 2].permute(0, 2, 3, 1).reshape(B * H, D, T)
    assert relerr(q, rq) < 1e-2, describe(q, rq)
    assert relerr(k, rk) < 1e-2, describe(k, rk)
    assert relerr(vt, rv) < 1e-2, describe(vt, rv)
    # row-major V variants (v written like k)
    v_row = torch.empty_like(q)
    q3, k3 = torch.empty_like(q), torch.empty_like(q)
    rvr = ref[:, :, 2].permute(0, 2, 1, 3).reshape(B * H, T, D)
    for kernel in (0, 1):
        ops.gemm(a, w, epilogue=L.EPI_QKV_ROW, heads=H, head_dim=D, tokens=T, out=q3, out2=k3, out3=v_row, kernel=kernel)
        assert relerr(q3, rq) < 1e-2 and relerr(k3, rk) < 1e-2 and relerr(v_row, rvr) < 1e-2, describe(v_row, rvr)
        ops.gemm(a, w[C:], epilogue=L.EPI_KV_ROW, heads=H, head_dim=D, tokens=T, out=k3, out2=v_row, kernel=kernel)
        assert relerr(k3, rk) < 1e-2 and relerr(v_row, rvr) < 1e-2
    q2 = torch.empty_like(q)
    ops.gemm(a, w[:C], epilogue=L.EPI_HEADS, heads=H, head_dim=D, tokens=T, out=q2)
    assert relerr(q2, rq) < 1e-2
    v2 = torch.empty_like(vt)
    ops.gemm(a, w[2 * C:], epilogue=L.EPI_HEADS_T, heads=H, head_dim=D, tokens=T, out=v2)
    assert relerr(v2, rv) < 1e-2


@pytest.mark.parametrize("N,H,W,C,Cout,kh,kw", [
    (2, 64, 64, 320, 320, 3, 3), (4, 32, 32, 640, 640, 3, 3), (4, 16, 16, 1280, 1280, 3, 3),
    (4, 8, 8, 2560, 1280, 3, 3), (3, 8, 8, 1280, 1280, 3, 3), (8, 4, 4, 1280, 1280, 3, 3),
    (2, 64, 64, 320, 4, 3, 3), (1, 128, 128, 128, 128, 3, 3), (1, 256, 256, 128, 128, 1, 5),
    (2, 32, 32, 960, 640, 3, 3), (1, 64, 64, 64, 64, 3, 3),
])
@pytest.mark.parametrize("kernel", [0, 1])
def test_conv_implicit(N, H, W, C, Cout, kh, kw, kernel):
    ops = _ops()
    x = rnd(N, C, H, W, seed=1, dtype=torch.float32)
    w = rnd(Cout, C, kh, kw, seed=2, dtype=torch.float32, scale=(C * kh * kw) ** -0.5)
    b = rnd(Cout, seed=3, dtype=torch.float32)
    emb = rnd(N, Cout, seed=4, dtype=torch.float32)
    res = rnd(N, H, W, Cout, seed=5, dtype=torch.float32)
    xb = x.to(torch.bfloat16)
    wb = w.to(torch.bfloat16)
    ref = torch.nn.functional.conv2d(xb.float(), wb.float(), b, padding=(kh // 2, kw // 2))
    ref = ref + emb[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1) + res
    x_nhwc = xb.permute(0, 2, 3, 1).contiguous()
    wp = wb.permute(0, 2, 3, 1).reshape(Cout, kh * kw * C).contiguous()
    assert ops.conv_implicit_ok(H, W, C)
    out = ops.conv_implicit(x_nhwc, wp, kh, kw, kh // 2, kw // 2, bias=b, row_bias=emb, residual=res, kernel=kernel)
    assert relerr(out, ref) < 2e-3, describe(out.reshape(-1, Cout), ref.reshape(-1, Cout))


@pytest.mark.parametrize("N,H,W,C,Cout,tile_n", [
    (2, 64, 64, 320, 320, 160), (4, 32, 32, 640, 640, 160), (4, 16, 16, 1280, 1280, 160), (4, 8, 8, 2560, 1280, 160),
    (3, 8, 8, 1280, 1280, 128), (1, 128, 128, 128, 256, 128), (2, 64, 64, 960, 320, 160), (8, 4, 4, 1280, 512, 256),
    (5, 64, 64, 64, 320, 160),
])
def test_conv_implicit_quad_clusters(N, H, W, C, Cout, tile_n):
    """pair = 2: clusters of two CTA pairs on neighbouring n-tiles, each CTA fetching half of its A tile and multicasting it
    to its counterpart.  Same result as the pair kernel bit for bit (same MMAs, same order), and the PyTorch reference."""
    ops = _ops()
    x = rnd(N, H, W, C, seed=1)
    w = rnd(Cout, 9 * C, seed=2, scale=(9 * C) ** -0.5)
    b = rnd(Cout, seed=3, dtype=torch.float32)
    emb = rnd(N, Cout, seed=4, dtype=torch.float32)
    res = rnd(N, H, W, Cout, seed=5, dtype=torch.float32)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().reshape(Cout, 3, 3, C).permute(0, 3, 1, 2), b,
                                     padding=1).permute(0, 2, 3, 1) + emb[:, None, None, :] + res
    pair = ops.conv_implicit(x, w, 3, 3, 1, 1, bias=b, row_bias=emb, residual=res, tile_n=tile_n, pair=1)
    quad = ops.conv_implicit(x, w, 3, 3, 1, 1, bias=b, row_bias=emb, residual=res, tile_n=tile_n, pair=2, colstats=True)
    torch.cuda.synchronize()
    assert relerr(quad, ref) < 2e-3, describe(quad.reshape(-1, Cout), ref.reshape(-1, Cout))
    assert torch.equal(quad, pair)
    if hasattr(quad, "_colstats") and (N * H * W) % 32 == 0:
        g = N * H * W // 32
        want = quad.reshape(g, 32, Cout).sum(1)
        assert relerr(quad._colstats[:g * Cout].reshape(g, Cout), want) < 1e-5


@pytest.mark.parametrize("N,H,W,C,Cout", [
    (2, 64, 64, 320, 320), (4, 32, 32, 640, 640), (4, 16, 16, 1280, 1280), (4, 8, 8, 2560, 1280), (3, 8, 8, 1280, 640),
    (2, 64, 64, 960, 320), (6, 64, 64, 64, 320), (2, 32, 32, 320, 480), (1, 16, 16, 128, 4),
])
def test_conv_implicit_wide_pairs(N, H, W, C, Cout):
    """pair = 3: a CTA pair owns a 256 x 320 tile as two N = 160 MMAs per k-step on the same A tile (three rotating TMEM
    slots).  Bit-identical to the pair kernel with 160-wide tiles (same MMAs on the same operands)."""
    ops = _ops()
    x = rnd(N, H, W, C, seed=1)
    w = rnd(Cout, 9 * C, seed=2, scale=(9 * C) ** -0.5)
    b = rnd(Cout, seed=3, dtype=torch.float32)
    emb = rnd(N, Cout, seed=4, dtype=torch.float32)
    res = rnd(N, H, W, Cout, seed=5, dtype=torch.float32)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().reshape(Cout, 3, 3, C).permute(0, 3, 1, 2), b,
                                     padding=1).permute(0, 2, 3, 1) + emb[:, None, None, :] + res
    pair = ops.conv_implicit(x, w, 3, 3, 1, 1, bias=b, row_bias=emb, residual=res, tile_n=160, pair=1)
    wide = ops.conv_implicit(x, w, 3, 3, 1, 1, bias=b, row_bias=emb, residual=res, tile_n=160, pair=3, colstats=True)
    torch.cuda.synchronize()
    assert relerr(wide, ref) < 2e-3, describe(wide.reshape(-1, Cout), ref.reshape(-1, Cout))
    assert torch.equal(wide, pair)
    if hasattr(wide, "_colstats") and (N * H * W) % 32 == 0:
        g = N * H * W // 32
        want = wide.reshape(g, 32, Cout).sum(1)
        assert relerr(wide._colstats[:g * Cout].reshape(g, Cout), want) < 1e-5


@pytest.mark.parametrize("M,N,K", [(4096, 320, 320), (1000, 640, 1280), (8192, 1280, 5120), (300, 960, 64), (128, 23680, 1280),
                                   (256, 200, 640)])
def test_gemm_wide_pairs(M, N, K):
    ops = _ops()
    a = rnd(M, K, seed=1)
    w = rnd(N, K, seed=2, scale=K ** -0.5)
    bias = rnd(N, seed=3, dtype=torch.float32)
    res = rnd(M, N, seed=4, dtype=torch.float32)
    ref = a.float() @ w.float().t() + bias + res
    pair = ops.gemm(a, w, bias=bias, residual=res, out_dtype=torch.float32, tile_n=160, pair=1)
    wide = ops.gemm(a, w, bias=bias, residual=res, out_dtype=torch.float32, tile_n=160, pair=3)
    torch.cuda.synchronize()
    assert relerr(wide, ref) < 2e-3, describe(wide, ref)
    assert torch.equal(wide, pair)
    # the bf16 head-split and GEGLU epilogues on wide tiles
    if N == 960:
        H, D, T = 8, 40, 100
        outs = []
        for pr in (1, 3):
            q = torch.zeros(3 * H, T, D, device="cuda", dtype=torch.bfloat16)
            k, v = torch.zeros_like(q), torch.zeros_like(q)
            ops.gemm(a, w, epilogue=_L().EPI_QKV_ROW, heads=H, head_dim=D, tokens=T, out=q, out2=k, out3=v, tile_n=160, pair=pr)
            outs.append((q, k, v))
        torch.cuda.synchronize()
        assert all(torch.equal(x, y) for x, y in zip(*outs))


@pytest.mark.parametrize("M,N,K,tile_n", [(4096, 320, 320, 160), (1000, 640, 1280, 160), (8192, 1280, 5120, 160),
                                          (300, 512, 64, 128), (2048, 1024, 2560, 256)])
def test_gemm_quad_clusters(M, N, K, tile_n):
    ops = _ops()
    a = rnd(M, K, seed=1)
    w = rnd(N, K, seed=2, scale=K ** -0.5)
    bias = rnd(N, seed=3, dtype=torch.float32)
    res = rnd(M, N, seed=4, dtype=torch.float32)
    ref = a.float() @ w.float().t() + bias + res
    pair = ops.gemm(a, w, bias=bias, residual=res, out_dtype=torch.float32, tile_n=tile_n, pair=1)
    quad = ops.gemm(a, w, bias=bias, residual=res, out_dtype=torch.float32, tile_n=tile_n, pair=2)
    torch.cuda.synchronize()
    assert relerr(quad, ref) < 2e-3, describe(quad, ref)
    assert torch.equal(quad, pair)


def _fuzz_cases(n, seed):
    import random
    rng = random.Random(seed)
    cases = []
    for i in range(n):
        M = rng.choice([4, 64, 128, 256, 384, 1000, 2048, 4096, 8192, 20000, 128 * rng.randint(1, 80)])
        N = rng.choice([4, 32, 64, 160, 320, 640, 960, 1280, 2560, 4 * rng.randint(1, 500), 32 * rng.randint(1, 60)])
        K = rng.choice([64, 320, 640, 1280, 2880, 8 * rng.randint(1, 400)])
        cases.append((i, M, N, K, rng.random() < 0.5, rng.random() < 0.6, rng.random() < 0.4, rng.choice([0, 0, 1, 2]),
                      rng.random() < 0.3))
    return cases


@pytest.mark.parametrize("case", _fuzz_cases(36, 2024), ids=lambda c: "M%d-N%d-K%d-%d" % (c[1], c[2], c[3], c[0]))
def test_gemm_policy_fuzz(case):
    """Whatever tile shape / CTA grouping / epilogue class the library picks for a problem (160- and 256-wide tiles, pairs,
    the straight-line f32 epilogue, the residual prefetch, ...) gives the result of the plainest configuration (64-wide
    tiles, one CTA per tile) on the same operands -- the accumulation order per output element is the same -- and both
    match PyTorch.  Seeded random shapes incl. tails in M, N and K, odd tile counts and every PLAIN epilogue option."""
    i, M, N, K, out_f32, use_res, use_rb, act, stats = case
    ops = _ops()
    a = rnd(M, K, seed=10 + i)
    w = rnd(N, K, seed=50 + i, scale=K ** -0.5)
    bias = rnd(N, seed=90 + i, dtype=torch.float32)
    res = rnd(M, N, seed=130 + i, dtype=torch.float32) if use_res else None
    rpg = 32 if M % 32 == 0 else 1
    rb = rnd((M + rpg - 1) // rpg, N, seed=170 + i, dtype=torch.float32) if use_rb else None
    kw = dict(bias=bias, residual=res, row_bias=rb, rows_per_group=rpg, act=act,
              out_dtype=torch.float32 if out_f32 else torch.bfloat16)
    ref = a.float() @ w.float().t() + bias
    if rb is not None:
        ref = ref + rb.repeat_interleave(rpg, 0)[:M]
    if act == 1:
        ref = torch.nn.functional.silu(ref)
    elif act == 2:
        ref = torch.nn.functional.gelu(ref)
    if res is not None:
        ref = ref + res
    plain = ops.gemm(a, w, tile_n=64, pair=-1, **kw)
    auto = ops.gemm(a, w, colstats=stats and out_f32 and act == 0, **kw)
    torch.cuda.synchronize()
    tol = 2e-3 if out_f32 else 1e-2
    assert relerr(auto, ref) < tol, describe(auto, ref)
    assert relerr(auto, plain) < 1e-6 if out_f32 else torch.equal(auto, plain), describe(auto, plain)
    st = getattr(auto, "_colstats", None)
    if st is not None and M % 32 == 0:
        g = M // 32
        assert relerr(st[:g * N].reshape(g, N), auto.reshape(g, 32, N).sum(1)) < 1e-5


@pytest.mark.parametrize("case", [(2, 64, 64, 320, 320), (1, 32, 32, 640, 1280), (7, 16, 16, 1280, 640), (3, 8, 8, 1280, 1280),
                                  (5, 64, 64, 64, 320), (1, 64, 64, 320, 4), (9, 32, 32, 128, 960), (16, 16, 16, 320, 320)],
                         ids=lambda c: "n%d-%dx%d-%d-%d" % c)
def test_conv_policy_vs_plainest_tiles(case):
    """The same for the implicit-GEMM convolution: the automatic choice (wide pair tiles where N % 320 == 0 and the tiles
    fill the machine) against 64-wide tiles on single CTAs."""
    N, H, W, C, Cout = case
    ops = _ops()
    x = rnd(N, H, W, C, seed=1)
    w = rnd(Cout, 9 * C, seed=2, scale=(9 * C) ** -0.5)
    b = rnd(Cout, seed=3, dtype=torch.float32)
    emb = rnd(N, Cout, seed=4, dtype=torch.float32)
    res = rnd(N, H, W, Cout, seed=5, dtype=torch.float32)
    plain = ops.conv_implicit(x, w, 3, 3, 1, 1, bias=b, row_bias=emb, residual=res, tile_n=64, pair=-1)
    auto = ops.conv_implicit(x, w, 3, 3, 1, 1, bias=b, row_bias=emb, residual=res, colstats=True)
    torch.cuda.synchronize()
    assert relerr(auto, plain) < 1e-6, describe(auto.reshape(-1, Cout), plain.reshape(-1, Cout))


@pytest.mark.parametrize("N,H,W,C,Cout,k,stride,pads", [
    (2, 64, 64, 320, 320, 3, 2, (1, 1)), (2, 32, 32, 9, 320, 3, 1, (1, 1)), (1, 64, 64, 128, 128, 3, 2, (0, 0)),
])
def test_conv_im2col(N, H, W, C, Cout, k, stride, pads):
    ops = _ops()
    x = rnd(N, C, H, W, seed=1, dtype=torch.float32).to(torch.bfloat16)
    w = rnd(Cout, C, k, k, seed=2, dtype=torch.float32, scale=(C * k * k) ** -0.5).to(torch.bfloat16)
    b = rnd(Cout, seed=3, dtype=torch.float32)
    if pads == (0, 0):  # VAE downsample: pad right/bottom by one, then valid conv (model.py:72-76)
        xp = torch.nn.functional.pad(x.float(), (0, 1, 0, 1))
        ref = torch.nn.functional.conv2d(xp, w.float(), b, stride=stride)
    else:
        ref = torch.nn.functional.conv2d(x.float(), w.float(), b, stride=stride, padding=pads)
    ho, wo = ref.shape[2], ref.shape[3]
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    cols = ops.im2col(x_nhwc, k, k, stride, pads[0], pads[1], ho, wo)
    from mobi_b200.packing import pack_conv_weight
    wp = pack_conv_weight(w.float(), kpad=cols.shape[1])
    out = ops.gemm(cols, wp, bias=b, out_dtype=torch.float32).reshape(N, ho, wo, Cout)
    assert relerr(out, ref.permute(0, 2, 3, 1)) < 2e-3, describe(out, ref.permute(0, 2, 3, 1))


@pytest.mark.parametrize("B,H,D,Tq,Tk", [
    (1, 1, 64, 128, 128), (2, 8, 40, 256, 256), (1, 8, 40, 1024, 1024), (2, 8, 80, 256, 256),
    (2, 8, 160, 256, 256), (2, 8, 160, 64, 64), (1, 2, 16, 64, 64), (1, 2, 32, 16, 16),
    (1, 8, 40, 4096, 4096), (1, 4, 80, 384, 200), (2, 3, 40, 300, 136), (1, 2, 128, 512, 320), (1, 2, 64, 256, 1024),
    (40, 8, 40, 1024, 1024), (1, 2, 48, 700, 100),
])
@pytest.mark.parametrize("kernel", [1, 3, 4, 5, 6, 7, 8])
def test_attention(B, H, D, Tq, Tk, kernel):
    """kernel 1 = the one-tile kernel (takes V^T; what head_dim 160 runs); 3.. take V row-major (MN-major tcgen05
    operand): 3 = attention4 (P in TMEM, row sums on the tensor core: what the UNet runs for head_dim <= 112),
    4 = attention3 (P staged in shared memory; head_dim 113..128 and the A/B arm), 5 = its two-tile variant,
    6 = attention4 with a larger share of the exponentials on the FMA pipe, 7 = attention4's first TMEM plan (P over S), 8 = attention4 with the split softmax (two warps per lane group)."""
    rowv = kernel >= 3
    if rowv and D > 128:
        pytest.skip("row-major V needs head_dim <= 128")
    if not rowv and Tk % 8 != 0:
        pytest.skip("transposed V needs Tk % 8 == 0")
    ops = _ops()
    q = rnd(B * H, Tq, D, seed=1, dtype=torch.float32)
    k = rnd(B * H, Tk, D, seed=2, dtype=torch.float32)
    v = rnd(B * H, Tk, D, seed=3, dtype=torch.float32)
    # make the softmax peaky in places so the running-max / rescale path is exercised
    k[:, Tk // 2:] *= 3.0
    scale = D ** -0.5
    qs = (q * (scale * math.log2(math.e))).to(torch.bfloat16)
    kb, vb = k.to(torch.bfloat16), v.to(torch.bfloat16)
    s = torch.einsum("bid,bjd->bij", qs.float(), kb.float()) * math.log(2.0)
    ref = torch.einsum("bij,bjd->bid", s.softmax(-1), vb.float())
    ref = ref.reshape(B, H, Tq, D).permute(0, 2, 1, 3).reshape(B, Tq, H * D)
    vt = vb.contiguous() if rowv else vb.transpose(1, 2).contiguous()
    out = ops.attention(qs, kb, vt, B, H, D, Tq, Tk, kernel={3: 0, 4: 3, 5: 4, 6: 0x40, 7: 5, 8: 6}.get(kernel, kernel), v_rowmajor=rowv)
    torch.cuda.synchronize()
    assert relerr(out, ref) < 1e-2, describe(out, ref)


@pytest.mark.parametrize("N,HW,C1,C2,dtype,eps", [
    (2, 4096, 320, 0, torch.float32, 1e-5), (2, 1024, 1280, 640, torch.float32, 1e-5),
    (3, 64, 1280, 1280, torch.float32, 1e-5), (1, 65536, 128, 0, torch.bfloat16, 1e-6),
    (2, 256, 64, 0, torch.float32, 1e-6), (2, 4096, 640, 320, torch.float32, 1e-5),
])
def test_groupnorm(N, HW, C1, C2, dtype, eps):
    ops = _ops()
    side = int(HW ** 0.5)
    x1 = (rnd(N, side, side, C1, seed=1, dtype=torch.float32) * 2 + 0.5).to(dtype)
    x2 = (rnd(N, side, side, C2, seed=2, dtype=torch.float32) - 1.0).to(dtype) if C2 else None
    C = C1 + C2
    g = rnd(C, seed=3, dtype=torch.float32)
    b = rnd(C, seed=4, dtype=torch.float32)
    xc = torch.cat([x1, x2], -1) if C2 else x1
    ref = torch.nn.functional.group_norm(xc.float().permute(0, 3, 1, 2), 32, g, b, eps)
    ref = torch.nn.functional.silu(ref).permute(0, 2, 3, 1)
    out, cat = ops.groupnorm(x1, g, b, eps, x2=x2, silu=True, want_concat=True)
    assert relerr(out, ref) < 8e-3, describe(out, ref)
    assert torch.equal(cat, xc.to(torch.bfloat16))
    out2 = ops.groupnorm(x1, g, b, eps, x2=x2, silu=False)
    ref2 = torch.nn.functional.group_norm(xc.float().permute(0, 3, 1, 2), 32, g, b, eps).permute(0, 2, 3, 1)
    assert relerr(out2, ref2) < 8e-3, describe(out2, ref2)


def test_layernorm_variants():
    ops = _ops()
    R, T, C = 4, 256, 320
    x = rnd(R, T, C, seed=1, dtype=torch.float32) * 1.5 + 0.3
    g = rnd(C, seed=2, dtype=torch.float32)
    b = rnd(C, seed=3, dtype=torch.float32)
    ref = torch.nn.functional.layer_norm(x, (C,), g, b, 1e-5)
    out = ops.layernorm(x, g, b)
    assert relerr(out, ref.reshape(-1, C)) < 8e-3
    # camera rows (even) / lidar rows (odd) gathered from the interleaved batch (attention.py:246-247)
    cam = ops.layernorm(x, g, b, rows=R // 2 * T, seg=T, seg_stride=2 * T, seg_offset=0)
    lid = ops.layernorm(x, None, None, rows=R // 2 * T, seg=T, seg_stride=2 * T, seg_offset=T)
    assert relerr(cam, ref[0::2].reshape(-1, C)) < 8e-3
    assert torch.equal(lid, x[1::2].reshape(-1, C).to(torch.bfloat16))
    # broadcast add + write back (attn2 with one key, attention.py:235)
    vec = rnd(R, C, seed=4, dtype=torch.float32)
    x2 = x.clone()
    out2 = ops.layernorm(x2, g, b, add_vec=vec, add_rows_per_vec=T)
    xr = x + vec[:, None, :]
    assert torch.allclose(x2, xr, atol=1e-6)
    assert relerr(out2, torch.nn.functional.layer_norm(xr, (C,), g, b, 1e-5).reshape(-1, C)) < 8e-3
    for C2 in (640, 1280, 32):
        xx = rnd(64, C2, seed=5, dtype=torch.float32)
        gg = rnd(C2, seed=6, dtype=torch.float32)
        bb = rnd(C2, seed=7, dtype=torch.float32)
        assert relerr(ops.layernorm(xx, gg, bb), torch.nn.functional.layer_norm(xx, (C2,), gg, bb, 1e-5)) < 8e-3


def test_layernorm_dual():
    """mobi_ln_dual: camera rows normalised + lidar rows cast (and the reverse) in one pass (attention.py:246-261)."""
    ops = _ops()
    for R, T, C in ((4, 256, 320), (2, 64, 1280), (6, 16, 640), (2, 128, 64)):
        x = rnd(R, T, C, seed=1, dtype=torch.float32) * 1.5 + 0.3
        g = rnd(C, seed=2, dtype=torch.float32)
        b = rnd(C, seed=3, dtype=torch.float32)
        ref = torch.nn.functional.layer_norm(x, (C,), g, b, 1e-5)
        qn, ctx = ops.ln_dual(x.reshape(R * T, C), R, T, [(ops.LN_NORM, g, b), (ops.LN_CAST, None, None)])
        assert relerr(qn, ref[0::2].reshape(-1, C)) < 8e-3
        assert torch.equal(ctx, x[1::2].reshape(-1, C).to(torch.bfloat16))
        ctx, qn = ops.ln_dual(x.reshape(R * T, C), R, T, [(ops.LN_CAST, None, None), (ops.LN_NORM, g, b)])
        assert relerr(qn, ref[1::2].reshape(-1, C)) < 8e-3
        assert torch.equal(ctx, x[0::2].reshape(-1, C).to(torch.bfloat16))
        full, none = ops.ln_dual(x.reshape(R * T, C), R, T, [(ops.LN_NORM, g, b)], pair=False)
        assert none is None and relerr(full, ref.reshape(-1, C)) < 8e-3


@pytest.mark.parametrize("R,T,C,H,multimodal", [(4, 256, 320, 8, True), (2, 64, 1280, 8, True), (4, 30, 640, 8, False),
                                                (2, 16, 64, 4, True), (2, 4096, 320, 8, True)])
def test_ln_adapter(R, T, C, H, multimodal):
    """mobi_ln_adapter vs the unfused math of attention.py:235-247 in fp32: attn2 vector add, cond_adapter_norm,
    2-key cross attention in folded (U, Z) form, connector, then the following LayerNorm / cast."""
    ops = _ops()
    x = rnd(R, T, C, seed=1, dtype=torch.float32) * 1.3 + 0.2
    vec = rnd(R, C, seed=2, dtype=torch.float32)
    g, b = 1 + 0.1 * rnd(C, seed=3, dtype=torch.float32), 0.1 * rnd(C, seed=4, dtype=torch.float32)
    g2, b2 = 1 + 0.1 * rnd(C, seed=5, dtype=torch.float32), 0.1 * rnd(C, seed=6, dtype=torch.float32)
    U = rnd(R, 2, H, C, seed=7, dtype=torch.float32, scale=2.0 * C ** -0.5)
    Z = rnd(R, 2, H, C, seed=8, dtype=torch.float32)
    zb = rnd(C, seed=9, dtype=torch.float32)
    # reference
    x1 = x + vec[:, None]
    xn = torch.nn.functional.layer_norm(x1, (C,), g, b, 1e-5)
    s = torch.einsum("btc,bkhc->btkh", xn, U)
    pr = s.softmax(dim=2)
    x2 = x1 + torch.einsum("btkh,bkhc->btc", pr, Z) + zb
    # tables as BasicTransformerBlock.context_tables builds them
    Ug = torch.zeros(R, 2, 8, C, device="cuda")
    Zp = torch.zeros_like(Ug)
    sb = torch.zeros(R, 2, 8, device="cuda")
    Ug[:, :, :H], Zp[:, :, :H], sb[:, :, :H] = U * g, Z, (U * b).sum(-1)
    xx = x.clone().reshape(R * T, C)
    slots = [(ops.LN_NORM, g2, b2), (ops.LN_CAST, None, None)] if multimodal else [(ops.LN_NORM, g2, b2)]
    outs = ops.ln_adapter(xx, R, T, g, b, Ug.reshape(R, 16, C), sb.reshape(R, 16), Zp.reshape(R, 16, C), zb, slots,
                          pair=multimodal, add_vec=vec)
    torch.cuda.synchronize()
    assert relerr(xx, x2.reshape(-1, C)) < 1e-5, describe(xx, x2.reshape(-1, C))
    ln2 = torch.nn.functional.layer_norm(x2, (C,), g2, b2, 1e-5)
    if multimodal:
        assert relerr(outs[0], ln2[0::2].reshape(-1, C)) < 8e-3
        assert relerr(outs[1], x2[1::2].reshape(-1, C)) < 8e-3
    else:
        assert relerr(outs[0], ln2.reshape(-1, C)) < 8e-3


def test_small_kernels():
    ops = _ops()
    t = torch.tensor([981, 1, 500, 21], device="cuda", dtype=torch.int64)
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half).cuda()
    args = t[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1)
    out = ops.timestep_embedding(t, 320)
    assert (out.float() - ref).abs().max() < 1e-2
    x = rnd(3, 9, 32, 32, seed=1, dtype=torch.float32)
    nhwc = ops.nchw_to_nhwc(x)
    assert torch.equal(nhwc, x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(ops.nhwc_to_nchw(nhwc), x)
    assert torch.equal(ops.nchw_to_nhwc(x, torch.bfloat16), x.permute(0, 2, 3, 1).to(torch.bfloat16))
    y = rnd(2, 8, 8, 64, seed=2, dtype=torch.float32)
    up = ops.upsample_nearest2x(y, torch.bfloat16)
    ref = torch.nn.functional.interpolate(y.permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up, ref.to(torch.bfloat16))
    assert relerr(ops.silu(y), torch.nn.functional.silu(y)) < 8e-3
    assert torch.equal(ops.add_f32(y, y), y + y)
    assert torch.equal(ops.cast_bf16(y), y.to(torch.bfloat16))
    assert torch.allclose(ops.scale_f32(y, 1 / 0.18215), y * (1 / 0.18215), rtol=1e-6)


def test_ctx_attention():
    ops = _ops()
    B, T, C, H, Kk = 4, 256, 320, 8, 2
    xn = rnd(B * T, C, seed=1)
    U = rnd(B, Kk * H, C, seed=2, dtype=torch.float32, scale=C ** -0.5)
    Z = rnd(B, Kk * H, C, seed=3, dtype=torch.float32)
    zb = rnd(C, seed=4, dtype=torch.float32)
    x = rnd(B * T, C, seed=5, dtype=torch.float32)
    s = torch.einsum("btc,bjc->btj", xn.float().reshape(B, T, C), U).reshape(B, T, Kk, H)
    p = s.softmax(dim=2).reshape(B, T, Kk * H)
    ref = x.reshape(B, T, C) + torch.einsum("btj,bjc->btc", p, Z) + zb
    out = ops.ctx_attention(xn, U, Z, zb, x.clone(), B, T, H, Kk)
    assert relerr(out, ref.reshape(-1, C)) < 1e-4, describe(out, ref.reshape(-1, C))


def test_sampler_kernels():
    ops = _ops()
    B, H = 4, 32
    x = rnd(B, 4, H, H, seed=1, dtype=torch.float32)
    eps = rnd(2 * B, 4, H, H, seed=2, dtype=torch.float32)
    old = [rnd(B, 4, H, H, seed=3 + i, dtype=torch.float32) for i in range(3)]
    noise = rnd(B, 4, H, H, seed=9, dtype=torch.float32)
    a_t, a_prev, sigma, s = 0.3, 0.4, 0.1, 5.0
    e = eps[:B] + s * (eps[B:] - eps[:B])
    ep = (55 * e - 59 * old[0] + 37 * old[1] - 9 * old[2]) / 24
    pred = (x - math.sqrt(1 - a_t) * ep) / math.sqrt(a_t)
    xp = math.sqrt(a_prev) * pred + math.sqrt(1 - a_prev - sigma ** 2) * ep + sigma * noise
    e_out = torch.empty_like(x)
    x_prev, pred_x0 = ops.sampler_update(
        eps, x, cfg=True, scale=s, coefs=(55 / 24, -59 / 24, 37 / 24, -9 / 24), sqrt_one_minus_at=math.sqrt(1 - a_t),
        sqrt_at=math.sqrt(a_t), sqrt_a_prev=math.sqrt(a_prev), dir_coef=math.sqrt(1 - a_prev - sigma ** 2),
        sigma_temp=sigma, noise=noise, old=old, e_out=e_out)
    assert torch.allclose(e_out, e, atol=1e-5)
    assert torch.allclose(pred_x0, pred, atol=1e-4, rtol=1e-5)
    assert torch.allclose(x_prev, xp, atol=1e-4, rtol=1e-5)
    img = rnd(B, 4, H, H, seed=20, dtype=torch.float32)
    mask = (rnd(B, 1, H, H, seed=21, dtype=torch.float32) > 0).float()
    x_in = torch.empty(2 * B, 9, H, H, device="cuda")
    ops.assemble_input(x, img, mask, x_in, cfg=True)
    ref = torch.cat([x, img, mask], 1)
    assert torch.equal(x_in, torch.cat([ref, ref]))
    bm = (rnd(B, 1, H, H, seed=22, dtype=torch.float32) > 0).float()
    x0 = rnd(B, 4, H, H, seed=23, dtype=torch.float32)
    nz = rnd(B, 4, H, H, seed=24, dtype=torch.float32)
    xb = x.clone()
    ops.assemble_input(xb, img, mask, x_in, cfg=False, blend=(bm, x0, nz, 0.8, 0.6))
    xr = (0.8 * x0 + 0.6 * nz) * bm + (1 - bm) * x
    assert torch.allclose(xb, xr, atol=1e-6)
    assert torch.allclose(x_in[:B], torch.cat([xr, img, mask], 1), atol=1e-6)


# ----------------------------------------------------------------------------------------------- producer-side GroupNorm statistics
@pytest.mark.parametrize("M,N,K,res", [(4096, 320, 320, True), (2100, 640, 128, False), (8192, 1280, 64, True)])
def test_gemm_colstats(M, N, K, res):
    """mobi_gemm_args.colstats: per-column sums / sums of squares of the FINAL output over every 32-row group."""
    ops = _ops()
    a = rnd(M, K, seed=1, dtype=torch.float32).to(torch.bfloat16)
    w = rnd(N, K, seed=2, dtype=torch.float32, scale=K ** -0.5).to(torch.bfloat16)
    b = rnd(N, seed=3, dtype=torch.float32)
    r = rnd(M, N, seed=4, dtype=torch.float32) if res else None
    out = ops.gemm(a, w, bias=b, residual=r, out_dtype=torch.float32, colstats=True)
    st = out._colstats.view(2, (M + 31) // 32, N)
    torch.cuda.synchronize()
    pad = (-M) % 32
    o = torch.nn.functional.pad(out, (0, 0, 0, pad)).view(-1, 32, N)
    assert relerr(st[0], o.sum(1)) < 1e-5 and relerr(st[1], (o * o).sum(1)) < 1e-5


@pytest.mark.parametrize("N,HW,C1,C2,silu", [(2, 4096, 320, 0, True), (3, 1024, 1280, 640, True), (2, 64, 1280, 1280, False),
                                             (2, 256, 640, 320, True)])
def test_groupnorm_with_producer_statistics(N, HW, C1, C2, silu):
    """GroupNorm fed by the column statistics its producers left (ops.carry_colstats): same result as the statistics pass,
    incl. a concatenation whose groups straddle the two inputs (1280 + 640 channels -> groups of 60)."""
    ops = _ops()
    side = int(HW ** 0.5)

    def produced(C, seed):   # a conv-like producer: 1x1 GEMM over [N*HW, C] with colstats
        a = rnd(N * HW, 64, seed=seed, dtype=torch.float32).to(torch.bfloat16)
        w = rnd(C, 64, seed=seed + 1, dtype=torch.float32, scale=0.125).to(torch.bfloat16)
        bias = rnd(C, seed=seed + 2, dtype=torch.float32) * 2      # a mean well away from zero
        out = ops.gemm(a, w, bias=bias, out_dtype=torch.float32, colstats=True)
        return ops.carry_colstats(out.reshape(N, side, side, C), out)

    x1 = produced(C1, 10)
    x2 = produced(C2, 20) if C2 else None
    C = C1 + C2
    g, b = 1 + 0.1 * rnd(C, seed=5, dtype=torch.float32), 0.1 * rnd(C, seed=6, dtype=torch.float32)
    got = ops.groupnorm(x1, g, b, 1e-5, x2=x2, silu=silu)
    plain1 = x1.clone()
    plain2 = x2.clone() if x2 is not None else None                  # clones carry no statistics: the two-read path
    want = ops.groupnorm(plain1, g, b, 1e-5, x2=plain2, silu=silu)
    xx = x1 if x2 is None else torch.cat([x1, x2], -1)
    ref = torch.nn.functional.group_norm(xx.permute(0, 3, 1, 2), 32, g, b, 1e-5)
    ref = (torch.nn.functional.silu(ref) if silu else ref).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert relerr(got, ref) < 1e-2 and relerr(got.float(), want.float()) < 1e-2
    assert (got.float() - want.float()).abs().max().item() <= 2 * 2 ** -7 * ref.abs().max().item()   # <= a bf16 ulp apart


def test_conv_colstats_and_cfg_duplicate():
    ops = _ops()
    from mobi_b200.attention import repeat_rows2
    x = rnd(2, 32, 32, 64, seed=1, dtype=torch.float32).to(torch.bfloat16)
    w = rnd(128, 9 * 64, seed=2, dtype=torch.float32, scale=(9 * 64) ** -0.5).to(torch.bfloat16)
    out = ops.conv_implicit(x, w, 3, 3, 1, 1, bias=rnd(128, seed=3, dtype=torch.float32), colstats=True)
    st = out._colstats.view(2, -1, 128)
    o = out.reshape(-1, 32, 128)
    assert relerr(st[0], o.sum(1)) < 1e-5 and relerr(st[1], (o * o).sum(1)) < 1e-5
    dup = repeat_rows2(out)
    assert dup._colstats.numel() == 2 * st.numel() and torch.equal(dup._colstats.view(2, -1, 128)[:, st.shape[1]:], st)

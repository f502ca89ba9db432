"""Tensor-level wrappers over the training-step entries of the C ABI (include/mobi_b200.h, "Training step").
Same rules as mobi_b200.ops: CUDA tensors only, outputs from torch.empty, launches on the current stream.
"""
import ctypes as C
import math

import torch

from . import _lib as L
from . import ops
from .ops import _cuda, _timed


def transpose(x, *, rows, cols, ld_in=None, batch=1, in_batch_stride=0, out=None):
    """out[b][c, r] = bf16(x[b][r, c]).  x is addressed as base + b * in_batch_stride + r * ld_in + c (elements)."""
    assert x.is_cuda and x.dtype in (torch.float32, torch.bfloat16)
    ld_in = cols if ld_in is None else ld_in
    if out is None:
        out = torch.empty((batch, cols, rows), device=x.device, dtype=torch.bfloat16)
    with _timed("transpose", nbytes=batch * rows * cols * (x.element_size() + 2)):
        L.check(L.load().mobi_transpose_bf16(x.data_ptr(), L.dt(x), out.data_ptr(), batch, rows, cols, ld_in,
                                             in_batch_stride, rows, rows * cols, L.stream()), "transpose_bf16")
    return out


def layernorm_bwd(x, gamma, dy, dx, *, rows=None, seg=0, seg_stride=0, seg_offset=0, dgamma=None, dbeta=None,
                  accumulate=True, eps=1e-5):
    _cuda(x, gamma, dy, dx, dgamma, dbeta)
    assert x.dtype == torch.float32 and dx.dtype == torch.float32
    Cc = x.shape[-1]
    a = L.LayerNormBwdArgs()
    a.x, a.gamma, a.dy, a.dx = x.data_ptr(), L.ptr(gamma), dy.data_ptr(), dx.data_ptr()
    a.dgamma, a.dbeta = L.ptr(dgamma), L.ptr(dbeta)
    a.rows = dy.numel() // Cc if rows is None else rows
    a.C = Cc
    a.seg, a.seg_stride, a.seg_offset = seg, seg_stride, seg_offset
    a.eps, a.dy_dtype, a.accumulate = eps, L.dt(dy), 1 if accumulate else 0
    with _timed("ln_bwd", nbytes=a.rows * Cc * (4 + dy.element_size() + 8)):
        L.check(L.load().mobi_layernorm_bwd(C.byref(a), L.stream()), "layernorm_bwd")
    return dx


def groupnorm_bwd(x1, gamma, beta, dy, eps, *, x2=None, silu=True, dres=None, groups=32):
    """Returns (dx1, dx2) f32 NHWC; dx2 is None without a second source."""
    _cuda(x1, x2, gamma, beta, dy, dres)
    assert x1.dtype == torch.float32 and (x2 is None or x2.dtype == torch.float32)
    n, h, w, c1 = x1.shape
    c2 = 0 if x2 is None else x2.shape[-1]
    assert dy.numel() == n * h * w * (c1 + c2)
    dx1 = torch.empty_like(x1)
    dx2 = None if x2 is None else torch.empty_like(x2)
    a = L.GroupNormBwdArgs()
    a.x1, a.x2, a.gamma, a.beta = x1.data_ptr(), L.ptr(x2), gamma.data_ptr(), beta.data_ptr()
    a.dy, a.dres, a.dx1, a.dx2 = dy.data_ptr(), L.ptr(dres), dx1.data_ptr(), L.ptr(dx2)
    a.n_img, a.hw, a.c1, a.c2, a.groups = n, h * w, c1, c2, groups
    a.silu, a.dy_dtype, a.eps = 1 if silu else 0, L.dt(dy), eps
    with _timed("gn_bwd", nbytes=n * h * w * (c1 + c2) * 16):
        L.check(L.load().mobi_groupnorm_bwd(C.byref(a), L.stream()), "groupnorm_bwd")
    return dx1, dx2


def geglu(g):
    _cuda(g)
    rows, two_f = g.shape
    out = torch.empty((rows, two_f // 2), device=g.device, dtype=torch.bfloat16)
    with _timed("geglu", nbytes=rows * two_f * 3):
        L.check(L.load().mobi_geglu(g.data_ptr(), out.data_ptr(), rows, two_f // 2, L.stream()), "geglu")
    return out


def geglu_bwd(g, dh):
    _cuda(g, dh)
    rows, two_f = g.shape
    assert dh.shape == (rows, two_f // 2) and dh.dtype == torch.bfloat16
    dg = torch.empty_like(g)
    with _timed("geglu_bwd", nbytes=rows * two_f * 5):
        L.check(L.load().mobi_geglu_bwd(g.data_ptr(), dh.data_ptr(), dg.data_ptr(), rows, two_f // 2, L.stream()),
                "geglu_bwd")
    return dg


def attn_softmax_bwd(S, dP, dS, dSt, Pt, stats, tq, tk, dscale, batch=1):
    a = L.AttnSoftmaxBwdArgs()
    a.S, a.dP, a.dS, a.dSt, a.Pt, a.stats = (S.data_ptr(), dP.data_ptr(), dS.data_ptr(), dSt.data_ptr(), Pt.data_ptr(),
                                             stats.data_ptr())
    a.batch, a.tq, a.tk, a.dscale = batch, tq, tk, dscale
    with _timed("attn_softmax_bwd", nbytes=batch * tq * tk * 22, kernels=2):
        L.check(L.load().mobi_attn_softmax_bwd(C.byref(a), L.stream()), "attn_softmax_bwd")


def attn_softmax_bwd_lse(S, dP, dS, dSt, Pt, stats, tq, tk, dscale, lse, o, d_o, ld, head_dim, batch=1):
    """Like attn_softmax_bwd, with the row statistics from the forward's log-sum-exp and Delta = rowsum(dO * O)."""
    a = L.AttnSoftmaxBwdArgs()
    a.S, a.dP, a.dS, a.dSt, a.Pt, a.stats = (S.data_ptr(), dP.data_ptr(), dS.data_ptr(), dSt.data_ptr(), Pt.data_ptr(),
                                             stats.data_ptr())
    a.batch, a.tq, a.tk, a.dscale = batch, tq, tk, dscale
    with _timed("attn_softmax_bwd", nbytes=batch * tq * tk * 14, kernels=2):
        L.check(L.load().mobi_attn_softmax_bwd_lse(C.byref(a), lse.data_ptr(), o.data_ptr(), d_o.data_ptr(), ld, head_dim,
                                                   L.stream()), "attn_softmax_bwd_lse")


def attn_bwd_flash(q, k, v, d_o, stats, dq, dk, dv, *, heads, tokens, head_dim, dscale, batch_rows=1):
    """Statistics pass + flash-style dQ' / dK / dV (see mobi_attn_bwd_flash): q, k, v bf16 [batch_rows * heads, T, D]; d_o,
    dq, dk, dv token-major bf16 matrix VIEWS [batch_rows * T, heads * D] (row strides taken from the views)."""
    a = L.AttnBwdTilesArgs()
    a.q, a.k, a.v, a.d_o = q.data_ptr(), k.data_ptr(), v.data_ptr(), d_o.data_ptr()
    a.stats = stats.data_ptr()
    a.heads, a.tokens, a.head_dim, a.stats_only, a.ld_do, a.dscale = heads, tokens, head_dim, 1, d_o.stride(0), dscale
    a.batch_rows = batch_rows
    z = batch_rows * heads
    with _timed("attn_bwd_stats", 2 * 2.0 * z * tokens * tokens * head_dim):
        L.check(L.load().mobi_attn_bwd_tiles(C.byref(a), L.stream()), "attn_bwd_tiles(stats)")
    f = L.AttnBwdFlashArgs()
    f.q, f.k, f.v, f.d_o, f.stats = q.data_ptr(), k.data_ptr(), v.data_ptr(), d_o.data_ptr(), stats.data_ptr()
    f.dq, f.dk, f.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    f.ld_do, f.ld_dq, f.ld_dk, f.ld_dv = d_o.stride(0), dq.stride(0), dk.stride(0), dv.stride(0)
    f.heads, f.tokens, f.head_dim, f.batch_rows, f.dscale = heads, tokens, head_dim, batch_rows, dscale
    with _timed("attn_bwd_flash", 7 * 2.0 * z * tokens * tokens * head_dim, kernels=2):
        L.check(L.load().mobi_attn_bwd_flash(C.byref(f), L.stream()), "attn_bwd_flash")


def attn_bwd_tiles(q, k, v, d_o, stats, dS, dSt, Pt, *, heads, tokens, head_dim, ld_do, dscale, batch_rows=1):
    """Fused S / dP recompute + softmax backward for all heads of `batch_rows` batch rows (see mobi_attn_bwd_tiles).
    dSt None: `Pt` receives P row-major and no transposed tiles are written."""
    a = L.AttnBwdTilesArgs()
    a.q, a.k, a.v, a.d_o = q.data_ptr(), k.data_ptr(), v.data_ptr(), d_o.data_ptr()
    a.stats, a.dS, a.dSt, a.Pt = stats.data_ptr(), dS.data_ptr(), L.ptr(dSt), Pt.data_ptr()
    a.heads, a.tokens, a.head_dim, a.stats_only, a.ld_do, a.dscale = heads, tokens, head_dim, 0, ld_do, dscale
    a.batch_rows = batch_rows
    fl = 2 * 2 * 2.0 * batch_rows * heads * tokens * tokens * head_dim      # S and dP, each computed in both passes
    with _timed("attn_bwd_tiles", fl, nbytes=batch_rows * heads * tokens * tokens * (4.0 if dSt is None else 6.0), kernels=2):
        L.check(L.load().mobi_attn_bwd_tiles(C.byref(a), L.stream()), "attn_bwd_tiles")


def ctx_attn_qspace(q, k, v, batch, tokens, heads, *, d_o=None):
    """Forward (d_o None): returns o.  Backward: returns (dq, dk, dv)."""
    _cuda(q, k, v, d_o)
    Cc = q.shape[-1]
    keys = k.shape[1]
    a = L.CtxAttnQspaceArgs()
    a.q, a.k, a.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    a.batch, a.tokens, a.C, a.heads, a.keys = batch, tokens, Cc, heads, keys
    if d_o is None:
        o = torch.empty_like(q)
        a.o, a.backward = o.data_ptr(), 0
        with _timed("ctx_attn"):
            L.check(L.load().mobi_ctx_attn_qspace(C.byref(a), L.stream()), "ctx_attn_qspace")
        return o
    dq = torch.empty_like(q)
    dk = torch.zeros_like(k)
    dv = torch.zeros_like(v)
    a.d_o, a.dq, a.dk, a.dv, a.backward = d_o.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), 1
    with _timed("ctx_attn_bwd"):
        L.check(L.load().mobi_ctx_attn_qspace(C.byref(a), L.stream()), "ctx_attn_qspace")
    return dq, dk, dv


def colsum(x, out, *, rows=None, cols=None, ld=None, rows_per_group=0):
    """out[g, :] += column sums of the rows of group g (out f32, accumulated)."""
    assert x.is_cuda and out.is_cuda and out.dtype == torch.float32
    cols = x.shape[-1] if cols is None else cols
    rows = x.numel() // x.shape[-1] if rows is None else rows
    ld = x.stride(-2) if ld is None else ld
    with _timed("colsum", nbytes=rows * cols * x.element_size()):
        L.check(L.load().mobi_colsum(x.data_ptr(), L.dt(x), rows, cols, ld, rows_per_group, out.data_ptr(), L.stream()),
                "colsum")
    return out


def wgrad_small(A, B, out):
    """out[n, k] += sum_m A[m, n] * B[m, k] (f32, tiny m)."""
    _cuda(A, B)
    assert A.dtype == B.dtype == out.dtype == torch.float32 and out.stride(-1) == 1
    m, n = A.shape
    k = B.shape[1]
    with _timed("wgrad_small"):
        L.check(L.load().mobi_wgrad_small(A.data_ptr(), B.data_ptr(), out.data_ptr(), m, n, k, A.stride(0), B.stride(0),
                                          out.stride(0), L.stream()), "wgrad_small")
    return out


def scatter_add_rows(src, dst, *, rows, Cc, seg=0, seg_stride=0, seg_offset=0):
    _cuda(src, dst)
    with _timed("scatter_add", nbytes=rows * Cc * (8 + src.element_size())):
        L.check(L.load().mobi_scatter_add_rows(src.data_ptr(), L.dt(src), dst.data_ptr(), rows, Cc, seg, seg_stride,
                                               seg_offset, L.stream()), "scatter_add_rows")
    return dst


def zero_insert2x(dy):
    _cuda(dy)
    n, h, w, c = dy.shape
    z = torch.empty((n, 2 * h, 2 * w, c), device=dy.device, dtype=torch.bfloat16)
    with _timed("zero_insert"):
        L.check(L.load().mobi_zero_insert2x(dy.data_ptr(), L.dt(dy), z.data_ptr(), n, h, w, c, L.stream()), "zero_insert2x")
    return z


def sum2x2(d):
    _cuda(d)
    n, h2, w2, c = d.shape
    out = torch.empty((n, h2 // 2, w2 // 2, c), device=d.device, dtype=torch.float32)
    with _timed("sum2x2"):
        L.check(L.load().mobi_sum2x2(d.data_ptr(), out.data_ptr(), n, h2 // 2, w2 // 2, c, L.stream()), "sum2x2")
    return out


def q_sample(x0, noise, sqrt_ac, sqrt_1mac, t, c_noised):
    _cuda(x0, noise, sqrt_ac, sqrt_1mac, t)
    assert t.dtype == torch.int64 and x0.dtype == torch.float32 and noise.dtype == torch.float32
    b, c, h, w = x0.shape
    out = torch.empty_like(x0)
    with _timed("q_sample"):
        L.check(L.load().mobi_q_sample(x0.data_ptr(), noise.data_ptr(), sqrt_ac.data_ptr(), sqrt_1mac.data_ptr(),
                                       t.data_ptr(), out.data_ptr(), b, c, c_noised, h * w, L.stream()), "q_sample")
    return out


def mse_grad(pred, target, loss_sum, grad_scale):
    _cuda(pred, target, loss_sum)
    grad = torch.empty_like(pred)
    with _timed("mse_grad"):
        L.check(L.load().mobi_mse_grad(pred.data_ptr(), target.data_ptr(), grad.data_ptr(), loss_sum.data_ptr(),
                                       pred.numel(), grad_scale, L.stream()), "mse_grad")
    return grad


def silu_bwd(pre, dy):
    """dx (f32) = dy * silu'(pre)."""
    _cuda(pre, dy)
    dx = torch.empty(pre.shape, device=pre.device, dtype=torch.float32)
    with _timed("silu_bwd"):
        L.check(L.load().mobi_silu_bwd(pre.data_ptr(), L.dt(pre), dy.data_ptr(), L.dt(dy), dx.data_ptr(), pre.numel(),
                                       L.stream()), "silu_bwd")
    return dx


def adamw(p, g, m, v, *, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2, step=1, grad_scale=1.0):
    _cuda(p, g, m, v)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    with _timed("adamw", nbytes=p.numel() * 28):
        L.check(L.load().mobi_adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2, eps,
                                    weight_decay, bc1, bc2, grad_scale, L.stream()), "adamw")


def adamw_segments(p, g, m, v, *, bounds, flags, steps, state, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2,
                   grad_scale=1.0):
    """AdamW over a flat buffer cut at `bounds` = (bound1, bound2) into <= 3 segments; flags f32 [>= 3] (device): a
    segment with flag <= 0 got no gradient this step and is skipped like torch.optim.AdamW skips grad-None parameters;
    steps int32 [3] (device, in/out): per-segment step counts; state f32 [9] scratch.  See mobi_adamw_segments."""
    _cuda(p, g, m, v, flags, steps, state)
    assert flags.dtype == torch.float32 and steps.dtype == torch.int32 and state.dtype == torch.float32
    assert flags.numel() >= 3 and steps.numel() >= 3 and state.numel() >= 9
    with _timed("adamw", nbytes=p.numel() * 28, kernels=2):
        L.check(L.load().mobi_adamw_segments(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(),
                                             int(bounds[0]), int(bounds[1]), flags.data_ptr(), steps.data_ptr(),
                                             state.data_ptr(), lr, beta1, beta2, eps, weight_decay, grad_scale, L.stream()),
                "adamw_segments")


# ------------------------------------------------------------------------------------------------ composites
def wgrad_splits(M, n_out, k_in, sms=148):
    """How many K-slices a weight gradient over M token rows is cut into: enough that tiles x slices fill the SMs, slices
    of at least 512 rows, a power of two that divides M."""
    tiles = ((n_out + 127) // 128) * ((k_in + 159) // 160)
    s = 1
    while s < 32 and tiles * s < sms and M % (2 * s) == 0 and M // (2 * s) >= 512:
        s *= 2
    return s


def wgrad(dy_bf, x_bf, out, *, M, n_out, k_in):
    """out[n_out, k_in] += dy^T x over M token rows (dy_bf [M, n_out], x_bf [M, k_in], bf16; row strides taken from the
    views): ONE tcgen05 GEMM that reads both operands MN-major (no transposed copies).  Few output tiles and a long token
    dimension (320 x 320 over 16,384 rows = 6 tiles) would leave most SMs idle, so the token rows are cut into K-slices
    that run as the batch entries of one launch and meet in the f32 gradient through atomic adds."""
    lda, ldb = dy_bf.stride(-2), x_bf.stride(-2)
    s = wgrad_splits(M, n_out, k_in)
    if s == 1:
        return ops.gemm(dy_bf, x_bf, out=out, residual=out, out_dtype=torch.float32, M=n_out, N=k_in, K=M, lda=lda, ldb=ldb,
                        a_mn=True, b_mn=True)
    ks = M // s
    return ops.gemm(dy_bf, x_bf, out=out, out_dtype=torch.float32, M=n_out, N=k_in, K=ks, lda=lda, ldb=ldb, a_mn=True,
                    b_mn=True, batch=s, a_batch_stride=ks * lda, b_batch_stride=ks * ldb, out_batch_stride=0,
                    atomic_out=True)


LN2 = math.log(2.0)

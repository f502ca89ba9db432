"""Tensor-level wrappers over the C ABI: they check devices/dtypes/contiguity, allocate outputs with
torch.empty (the library never allocates) and launch on PyTorch's current stream.
"""
import ctypes as C

import os

import torch

from . import _lib as L


_SHAPE_KINDS = os.environ.get("MOBI_GEMM_SHAPES", "0") == "1"


class Stats:
    """Launch accounting (every C-ABI call = the number of CUDA kernels it launches) and optional per-op CUDA-event
    timing on the launching stream (bench.py's roofline leg)."""
    launches = 0
    records = None  # list of (kind, flops, bytes, ev_start, ev_end) when profiling

    @classmethod
    def begin_profile(cls):
        cls.records = []

    @classmethod
    def end_profile(cls):
        torch.cuda.synchronize()
        out = {}
        for kind, flops, nbytes, e0, e1 in cls.records:
            d = out.setdefault(kind, dict(launches=0, flops=0.0, bytes=0.0, ms=0.0))
            d["launches"] += 1
            d["flops"] += flops
            d["bytes"] += nbytes
            d["ms"] += e0.elapsed_time(e1)
        cls.records = None
        return out


class _timed:
    def __init__(self, kind, flops=0.0, nbytes=0.0, kernels=1):
        self.kind, self.flops, self.nbytes, self.kernels = kind, flops, nbytes, kernels

    def __enter__(self):
        Stats.launches += self.kernels
        if Stats.records is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if Stats.records is not None and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            Stats.records.append((self.kind, self.flops, self.nbytes, self.e0, e1))
        return False


def _cuda(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("mobi_b200 ops need CUDA tensors (no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError("mobi_b200 ops need contiguous tensors, got strides %s" % (t.stride(),))


# Producer-side GroupNorm statistics (mobi_gemm_args.colstats).  A GEMM / conv launched with colstats=True leaves the column
# sums of its output in a side buffer, kept as the attribute `_colstats` of the tensor it returns; groupnorm() picks it up
# from its inputs and then streams them once.  Views lose Python attributes: carry_colstats() hands the buffer on.
COLSTATS = os.environ.get("MOBI_GN_COLSTATS", "1") == "1"


def _colstats_buffer(M, N, device):
    return torch.empty(2 * ((M + 31) // 32) * N, device=device, dtype=torch.float32)


def carry_colstats(dst, src):
    st = getattr(src, "_colstats", None)
    if st is not None:
        dst._colstats = st
    return dst


def _rows_view(t, what):
    """Checks that `t` is a CUDA matrix view (unit inner stride, uniform row stride) and returns its row stride."""
    if not t.is_cuda:
        raise RuntimeError("mobi_b200 ops need CUDA tensors (no CPU fallback)")
    if t.stride(-1) != 1:
        raise RuntimeError("mobi_b200.gemm: %s must have a unit inner stride, got %s" % (what, t.stride(),))
    if t.dim() == 1:
        return t.shape[0]
    ld = t.stride(-2)
    for d in range(t.dim() - 2):  # leading dims must collapse onto the row dim
        if t.shape[d] != 1 and t.stride(d) != t.stride(d + 1) * t.shape[d + 1]:
            raise RuntimeError("mobi_b200.gemm: %s rows are not uniformly strided: %s" % (what, t.stride(),))
    return ld


def gemm(a, w, *, bias=None, row_bias=None, rows_per_group=1, ld_row_bias=0, residual=None, out=None,
         out_dtype=torch.bfloat16, epilogue=L.EPI_PLAIN, act=0, heads=0, head_dim=0, tokens=0, out2=None, out3=None,
         tile_n=0, M=None, K=None, lda=None, ldb=None, ldo=None, out_seg=0, out_seg_stride=0, out_seg_offset=0,
         kernel=0, N=None, batch=1, a_batch_stride=0, b_batch_stride=0, out_batch_stride=0, pair=0, a_mn=False, b_mn=False,
         atomic_out=False, batch_inner=0, a_batch2_stride=0, b_batch2_stride=0, out_batch2_stride=0, colstats=False):
    """out[M,N] = a[M,K] @ w[N,K]^T (+bias +row_bias +residual), bf16 operands, fp32 accumulate.
    colstats: also leave the per-column statistics of the output for the GroupNorm that consumes it (out._colstats).

    Mirrors torch.nn.functional.linear(a, w, bias); see include/mobi_b200.h for the epilogues.  a, w and out
    may be column slices of wider matrices (row strides are taken from the views or given explicitly).
    a_mn / b_mn: the operand is given MN-major, i.e. a is the row-major [K, M] matrix / w the row-major [K, N] matrix
    (out = a^T @ w): pass M, N, K explicitly.
    """
    _cuda(bias, out2, out3)
    if row_bias is not None:  # may be a column slice of a wider matrix (ld_row_bias = its row stride)
        assert row_bias.is_cuda and row_bias.stride(-1) == 1
        if ld_row_bias == 0 and row_bias.dim() == 2:
            ld_row_bias = row_bias.stride(0)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16, "gemm operands must be bf16"
    assert not (a_mn or b_mn) or None not in (M, N, K), "MN-major operands: pass M, N and K explicitly"
    lda_v = _rows_view(a, "a")
    ldb_v = _rows_view(w, "w")
    lda = lda_v if lda is None else lda
    ldb = ldb_v if ldb is None else ldb
    K = w.shape[1] if K is None else K
    N = w.shape[0] if N is None else N
    if M is None:
        M = a.numel() // a.shape[-1]
    if epilogue in (L.EPI_PLAIN, L.EPI_GEGLU, L.EPI_GEGLU2):
        n_out = N if epilogue == L.EPI_PLAIN else N // 2
        if out is None:
            out = torch.empty((M, n_out), device=a.device, dtype=out_dtype if epilogue == L.EPI_PLAIN else torch.bfloat16)
        ldo_v = _rows_view(out, "out") if ldo is None else ldo
        if residual is not None and residual is not out:
            assert _rows_view(residual, "residual") == ldo_v, "residual must share the output row stride"
    else:
        assert out is not None, "head layouts need preallocated outputs"
        _cuda(out)
        ldo_v = 0
    args = L.GemmArgs()
    args.A, args.B, args.out = a.data_ptr(), w.data_ptr(), out.data_ptr()
    args.out2, args.out3 = L.ptr(out2), L.ptr(out3)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= N
    if row_bias is not None:
        assert row_bias.dtype == torch.float32
    args.bias, args.row_bias, args.residual = L.ptr(bias), L.ptr(row_bias), L.ptr(residual)
    args.M, args.N, args.K = M, N, K
    args.lda, args.ldb, args.ldo = lda, ldb, ldo_v
    args.rows_per_group, args.ld_row_bias = rows_per_group, ld_row_bias
    args.out_dtype = L.dt(out)
    args.res_dtype = L.dt(residual) if residual is not None else L.DT_F32
    args.epilogue, args.act = epilogue, act
    args.heads, args.head_dim, args.tokens = heads, head_dim, tokens
    args.out_seg, args.out_seg_stride, args.out_seg_offset = out_seg, out_seg_stride, out_seg_offset
    args.conv = 0
    args.tile_n = tile_n
    args.kernel = kernel
    args.batch = batch
    args.a_batch_stride, args.b_batch_stride, args.out_batch_stride = a_batch_stride, b_batch_stride, out_batch_stride
    args.pair = pair
    args.a_mn_major, args.b_mn_major = (1 if a_mn else 0), (1 if b_mn else 0)
    args.atomic_out = 1 if atomic_out else 0
    args.batch_inner = batch_inner
    args.a_batch2_stride, args.b_batch2_stride, args.out_batch2_stride = a_batch2_stride, b_batch2_stride, out_batch2_stride
    st = None
    if colstats and COLSTATS and epilogue == L.EPI_PLAIN and out.dtype == torch.float32 and N % 4 == 0 and kernel == 0:
        st = _colstats_buffer(M, N, a.device)
        args.colstats = st.data_ptr()
    kind = "gemm"
    if _SHAPE_KINDS:   # MOBI_GEMM_SHAPES=1: per-shape accounting for tools/train_bench.py
        kind = "gemm M%d N%d K%d b%d%s%s%s" % (M, N, K, batch, " amn" if a_mn else "", " bmn" if b_mn else "",
                                                " atomic" if atomic_out else "")
    with _timed(kind, 2.0 * M * N * K * max(1, batch)):
        L.check(L.load().mobi_gemm(C.byref(args), L.stream()), "gemm")
    if st is not None:
        out._colstats = st
    return out


def conv_implicit(x, w, kh, kw, pad_h, pad_w, *, bias=None, row_bias=None, ld_row_bias=0, residual=None, out=None,
                  out_dtype=torch.float32, tile_n=0, kernel=0, pair=0, colstats=False):
    """Stride-1 'same' convolution of an NHWC bf16 image by implicit GEMM.

    x: [N, H, W, C] bf16; w: [Cout, kh*kw*C] bf16 with K ordered (kh, kw, c). Returns [N, H, W, Cout].
    row_bias: f32 [N, Cout]-like (one row per image): the timestep-embedding add (openaimodel.py:264-272).
    """
    _cuda(x, w, bias, residual, out)
    if row_bias is not None:
        assert row_bias.is_cuda and row_bias.stride(-1) == 1
        if ld_row_bias == 0 and row_bias.dim() == 2:
            ld_row_bias = row_bias.stride(0)
    n, h, wd, c = x.shape
    cout = w.shape[0]
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert w.shape[1] == kh * kw * c, (w.shape, kh, kw, c)
    if out is None:
        out = torch.empty((n, h, wd, cout), device=x.device, dtype=out_dtype)
    args = L.GemmArgs()
    args.A, args.B, args.out = x.data_ptr(), w.data_ptr(), out.data_ptr()
    args.bias, args.row_bias, args.residual = L.ptr(bias), L.ptr(row_bias), L.ptr(residual)
    args.M, args.N, args.K = n * h * wd, cout, kh * kw * c
    args.lda, args.ldb, args.ldo = c, w.stride(0), cout
    args.rows_per_group, args.ld_row_bias = h * wd, ld_row_bias
    args.out_dtype = L.dt(out)
    args.res_dtype = L.dt(residual) if residual is not None else L.DT_F32
    args.epilogue, args.act = L.EPI_PLAIN, 0
    args.conv, args.n_img, args.H, args.W, args.C = 1, n, h, wd, c
    args.KH, args.KW, args.pad_h, args.pad_w = kh, kw, pad_h, pad_w
    args.tile_n = tile_n
    args.kernel = kernel
    args.pair = pair
    st = None
    if colstats and COLSTATS and out.dtype == torch.float32 and cout % 4 == 0 and kernel == 0:
        st = _colstats_buffer(n * h * wd, cout, x.device)
        args.colstats = st.data_ptr()
    with _timed("conv", 2.0 * n * h * wd * cout * kh * kw * c):
        L.check(L.load().mobi_gemm(C.byref(args), L.stream()), "gemm(conv)")
    if st is not None:
        out._colstats = st
    return out


def conv_implicit_ok(h, w, c):
    """Shapes the implicit path tiles (see mobi_gemm): everything else goes through im2col."""
    if c % 8 != 0 or c < 64:
        return False
    if w >= 128:
        return w % 128 == 0
    if 128 % w != 0:
        return False
    bh = min(128 // w, h)
    return h % bh == 0 and 128 % (w * bh) == 0


def im2col(x, kh, kw, stride, pad_top, pad_left, ho, wo):
    _cuda(x)
    n, h, w, c = x.shape
    k = kh * kw * c
    kpad = (k + 7) // 8 * 8
    out = torch.empty((n * ho * wo, kpad), device=x.device, dtype=torch.bfloat16)
    a = L.Im2colArgs()
    a.x, a.out, a.in_dtype = x.data_ptr(), out.data_ptr(), L.dt(x)
    a.n, a.h, a.w, a.c, a.kh, a.kw = n, h, w, c, kh, kw
    a.stride, a.pad_top, a.pad_left, a.ho, a.wo, a.kpad = stride, pad_top, pad_left, ho, wo, kpad
    with _timed("im2col", 0.0, out.numel() * 2.0 + x.numel() * x.element_size()):
        L.check(L.load().mobi_im2col(C.byref(a), L.stream()), "im2col")
    return out


def attention(q, k, vt, batch, heads, head_dim, tq, tk, out=None, kernel=0, v_rowmajor=False, lse=None):
    """q,k: bf16 [batch*heads, T, d]; vt: bf16 [batch*heads, d, Tk] (or V itself, [batch*heads, Tk, d], with
    v_rowmajor=True, head_dim <= 128); returns bf16 [batch, tq, heads*d]."""
    _cuda(q, k, vt, out)
    assert q.dtype == k.dtype == vt.dtype == torch.bfloat16
    if out is None:
        out = torch.empty((batch, tq, heads * head_dim), device=q.device, dtype=torch.bfloat16)
    a = L.AttnArgs()
    a.q, a.k, a.vt, a.out = q.data_ptr(), k.data_ptr(), vt.data_ptr(), out.data_ptr()
    a.batch, a.heads, a.head_dim, a.tq, a.tk = batch, heads, head_dim, tq, tk
    a.ld_out = heads * head_dim
    a.kernel = kernel
    a.v_rowmajor = int(v_rowmajor)
    if lse is not None:
        assert lse.is_cuda and lse.dtype == torch.float32 and lse.numel() == batch * heads * tq and lse.is_contiguous()
    a.lse = L.ptr(lse)
    with _timed("attention", 4.0 * batch * heads * tq * tk * head_dim):
        L.check(L.load().mobi_attention(C.byref(a), L.stream()), "attention")
    return out


def groupnorm(x1, gamma, beta, eps, *, x2=None, silu=True, groups=32, want_concat=False, out_dtype=torch.bfloat16,
              two_pass=False):
    """x1 (and optionally x2, concatenated on channels): NHWC [N, H, W, C] f32|bf16 -> bf16 NHWC."""
    _cuda(x1, x2, gamma, beta)
    n = x1.shape[0]
    hw = x1.numel() // (n * x1.shape[-1])
    c1 = x1.shape[-1]
    c2 = x2.shape[-1] if x2 is not None else 0
    if x2 is not None:
        assert x2.dtype == x1.dtype and x2.shape[:-1] == x1.shape[:-1]
    c = c1 + c2
    out = torch.empty(x1.shape[:-1] + (c,), device=x1.device, dtype=out_dtype)
    cat = torch.empty(out.shape, device=x1.device, dtype=torch.bfloat16) if want_concat else None
    lib = L.load()
    nbytes = lib.mobi_groupnorm_scratch_bytes(n, hw, c, groups)
    partials = torch.empty((nbytes // 4,), device=x1.device, dtype=torch.float32)
    a = L.GroupNormArgs()
    a.x1, a.x2, a.gamma, a.beta = x1.data_ptr(), L.ptr(x2), gamma.data_ptr(), beta.data_ptr()
    a.out, a.out_concat, a.partials = out.data_ptr(), L.ptr(cat), partials.data_ptr()
    a.n_img, a.hw, a.c1, a.c2, a.groups = n, hw, c1, c2, groups
    a.in_dtype, a.silu, a.eps, a.out_dtype = L.dt(x1), int(silu), eps, L.dt(out)
    a.force_two_pass = int(two_pass)
    # statistics left by the producers of x1 / x2 (see carry_colstats): one streaming pass instead of two reads
    st1 = getattr(x1, "_colstats", None)
    st2 = getattr(x2, "_colstats", None) if x2 is not None else None
    fused_stats = (st1 is not None and (x2 is None or st2 is not None) and hw % 32 == 0 and hw >= 256 and not two_pass
                   and st1.numel() == 2 * (n * hw // 32) * c1 and (st2 is None or st2.numel() == 2 * (n * hw // 32) * c2))
    if fused_stats:
        a.colstats1, a.colstats2 = st1.data_ptr(), L.ptr(st2)
        kernels, nbytes = 2, x1.numel() * x1.element_size() + (x2.numel() * x2.element_size() if x2 is not None else 0)
    else:
        kernels = lib.mobi_groupnorm_launches(hw, c, groups, a.in_dtype, a.force_two_pass)
        nbytes = x1.numel() * x1.element_size() * 2 + (x2.numel() * x2.element_size() * 2 if x2 is not None else 0)
    with _timed("groupnorm", 0.0, nbytes + out.numel() * 2 * (2 if want_concat else 1), kernels=kernels):
        L.check(lib.mobi_groupnorm(C.byref(a), L.stream()), "groupnorm")
    return (out, cat) if want_concat else out


def layernorm(x, gamma, beta, *, rows=None, seg=0, seg_stride=0, seg_offset=0, add_vec=None, add_rows_per_vec=0,
              eps=1e-5):
    """x: f32 [*, C] -> bf16 [rows, C]. gamma=None casts only. See mobi_layernorm for the row gather."""
    _cuda(x, gamma, beta, add_vec)
    assert x.dtype == torch.float32
    c = x.shape[-1]
    if rows is None:
        rows = x.numel() // c
    out = torch.empty((rows, c), device=x.device, dtype=torch.bfloat16)
    a = L.LayerNormArgs()
    a.x, a.gamma, a.beta, a.out, a.add_vec = x.data_ptr(), L.ptr(gamma), L.ptr(beta), out.data_ptr(), L.ptr(add_vec)
    a.rows, a.C = rows, c
    a.seg, a.seg_stride, a.seg_offset, a.add_rows_per_vec = seg, seg_stride, seg_offset, add_rows_per_vec
    a.eps = eps
    with _timed("layernorm", 0.0, rows * c * (4 + 2 + (8 if add_vec is not None else 0))):
        L.check(L.load().mobi_layernorm(C.byref(a), L.stream()), "layernorm")
    return out


LN_SKIP, LN_NORM, LN_CAST = 0, 1, 2


def _dual_spec(spec, slots, pair, batch, tokens, c, device):
    """slots: [(mode, gamma, beta)] for slot 0 (even batch rows, or all rows when pair=False) and slot 1."""
    outs = []
    spec.pair = int(pair)
    for i in range(2):
        mode, gamma, beta = slots[i] if i < len(slots) else (LN_SKIP, None, None)
        out = None
        if mode != LN_SKIP:
            _cuda(gamma, beta)
            out = torch.empty(((batch // 2 if pair else batch) * tokens, c), device=device, dtype=torch.bfloat16)
        spec.mode[i] = mode
        spec.gamma[i], spec.beta[i], spec.out[i] = L.ptr(gamma), L.ptr(beta), L.ptr(out)
        outs.append(out)
    return outs


def ln_dual(x, batch, tokens, slots, *, pair=True, eps=1e-5):
    """One pass over x f32 [batch*tokens, C]; see mobi_ln_dual.  Returns the per-slot bf16 outputs (None if skipped)."""
    _cuda(x)
    assert x.dtype == torch.float32
    c = x.shape[-1]
    spec = L.LnDualSpec()
    outs = _dual_spec(spec, slots, pair, batch, tokens, c, x.device)
    n_out = sum(o.numel() for o in outs if o is not None)
    with _timed("layernorm", 0.0, n_out * 6.0):
        L.check(L.load().mobi_ln_dual(x.data_ptr(), C.byref(spec), batch, tokens, c, eps, L.stream()), "ln_dual")
    return outs


def ln_adapter(x, batch, tokens, gamma, beta, Ug, sb, Z, zb, slots, *, pair, add_vec=None, eps=1e-5):
    """Fused attn2-vector add + folded bbox/reference adapter + following LayerNorm(s); see mobi_ln_adapter."""
    _cuda(x, gamma, beta, Ug, sb, Z, zb, add_vec)
    assert x.dtype == torch.float32 and Ug.dtype == Z.dtype == sb.dtype == zb.dtype == torch.float32
    c = x.shape[-1]
    assert Ug.shape == (batch, 16, c) and Z.shape == (batch, 16, c) and sb.shape == (batch, 16)
    a = L.LnAdapterArgs()
    a.x, a.add_vec, a.gamma, a.beta = x.data_ptr(), L.ptr(add_vec), gamma.data_ptr(), beta.data_ptr()
    a.Ug, a.sb, a.Z, a.zb = Ug.data_ptr(), sb.data_ptr(), Z.data_ptr(), zb.data_ptr()
    a.batch, a.tokens, a.C, a.eps = batch, tokens, c, eps
    outs = _dual_spec(a.next, slots, pair, batch, tokens, c, x.device)
    n_out = sum(o.numel() for o in outs if o is not None)
    with _timed("ln_adapter", 0.0, x.numel() * 8.0 + n_out * 2.0):
        L.check(L.load().mobi_ln_adapter(C.byref(a), L.stream()), "ln_adapter")
    return outs


def timestep_embedding(t, dim, max_period=10000.0):
    _cuda(t)
    assert t.dtype == torch.int64
    out = torch.empty((t.shape[0], dim), device=t.device, dtype=torch.bfloat16)
    Stats.launches += 1
    L.check(L.load().mobi_timestep_embedding(t.data_ptr(), out.data_ptr(), t.shape[0], dim, max_period, L.stream()),
            "timestep_embedding")
    return out


def fourier_embed(x, num_freqs, ld_out=None):
    """x f32 [..., dims] -> bf16 [rows, ld_out]: [x, sin(x 2^i), cos(x 2^i) ...] per row of `dims` values (see
    mobi_fourier_embed); ld_out pads the row (zeros) to a multiple of 8 for the GEMM that follows."""
    _cuda(x)
    assert x.dtype == torch.float32
    dims = x.shape[-1]
    rows = x.numel() // dims
    per = dims * (1 + 2 * num_freqs)
    ld_out = per if ld_out is None else ld_out
    out = torch.empty((rows, ld_out), device=x.device, dtype=torch.bfloat16)
    Stats.launches += 1
    L.check(L.load().mobi_fourier_embed(x.data_ptr(), out.data_ptr(), rows, dims, num_freqs, ld_out, L.stream()),
            "fourier_embed")
    return out


def silu(x):
    _cuda(x)
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    Stats.launches += 1
    L.check(L.load().mobi_silu(x.data_ptr(), L.dt(x), out.data_ptr(), x.numel(), L.stream()), "silu")
    return out


def nchw_to_nhwc(x, out_dtype=torch.float32):
    _cuda(x)
    assert x.dtype == torch.float32
    n, c, h, w = x.shape
    out = torch.empty((n, h, w, c), device=x.device, dtype=out_dtype)
    Stats.launches += 1
    L.check(L.load().mobi_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), L.dt(out), n, c, h * w, L.stream()),
            "nchw_to_nhwc")
    return out


def nhwc_to_nchw(x):
    _cuda(x)
    n, h, w, c = x.shape
    out = torch.empty((n, c, h, w), device=x.device, dtype=torch.float32)
    Stats.launches += 1
    L.check(L.load().mobi_nhwc_to_nchw(x.data_ptr(), L.dt(x), out.data_ptr(), n, c, h * w, L.stream()),
            "nhwc_to_nchw")
    return out


def upsample_nearest2x(x, out_dtype=None):
    _cuda(x)
    n, h, w, c = x.shape
    out = torch.empty((n, 2 * h, 2 * w, c), device=x.device, dtype=out_dtype or x.dtype)
    Stats.launches += 1
    L.check(L.load().mobi_upsample_nearest2x(x.data_ptr(), L.dt(x), out.data_ptr(), L.dt(out), n, h, w, c,
                                             L.stream()), "upsample_nearest2x")
    return out


def ctx_attention(xn, U, Z, zb, x, batch, tokens, heads, keys):
    _cuda(xn, U, Z, zb, x)
    assert xn.dtype == torch.bfloat16 and x.dtype == torch.float32
    assert U.dtype == Z.dtype == zb.dtype == torch.float32
    a = L.CtxAttnArgs()
    a.xn, a.U, a.Z, a.zb, a.x = xn.data_ptr(), U.data_ptr(), Z.data_ptr(), zb.data_ptr(), x.data_ptr()
    a.batch, a.tokens, a.C, a.heads, a.keys = batch, tokens, x.shape[-1], heads, keys
    with _timed("ctx_attention", 0.0, xn.numel() * 2.0 + x.numel() * 8.0):
        L.check(L.load().mobi_ctx_attention(C.byref(a), L.stream()), "ctx_attention")
    return x


def softmax_rows(s, out=None):
    """s: f32 [rows, cols] scores in log2 units -> bf16 probabilities (see mobi_softmax_rows)."""
    _cuda(s, out)
    assert s.dtype == torch.float32 and s.dim() == 2
    if out is None:
        out = torch.empty(s.shape, device=s.device, dtype=torch.bfloat16)
    with _timed("softmax", 0.0, s.numel() * 14.0):
        L.check(L.load().mobi_softmax_rows(s.data_ptr(), out.data_ptr(), s.shape[0], s.shape[1], s.stride(0),
                                           out.stride(0), L.stream()), "softmax_rows")
    return out


def add_f32(a, b, out=None):
    _cuda(a, b, out)
    if out is None:
        out = torch.empty_like(a)
    Stats.launches += 1
    L.check(L.load().mobi_add_f32(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), L.stream()), "add_f32")
    return out


def scale_f32(x, s, out=None):
    _cuda(x, out)
    if out is None:
        out = torch.empty_like(x)
    Stats.launches += 1
    L.check(L.load().mobi_scale_f32(x.data_ptr(), float(s), out.data_ptr(), x.numel(), L.stream()), "scale_f32")
    return out


def cast_bf16(x):
    _cuda(x)
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    Stats.launches += 1
    L.check(L.load().mobi_cast_bf16(x.data_ptr(), out.data_ptr(), x.numel(), L.stream()), "cast_bf16")
    return out


def sampler_update(eps, x, *, cfg, scale, coefs, sqrt_one_minus_at, sqrt_at, sqrt_a_prev, dir_coef, sigma_temp=0.0,
                   noise=None, old=(), e_out=None, x_prev=None, pred_x0=None):
    _cuda(eps, x, noise, e_out, x_prev, pred_x0, *old)
    if x_prev is None:
        x_prev = torch.empty_like(x)
    if pred_x0 is None:
        pred_x0 = torch.empty_like(x)
    a = L.SamplerArgs()
    a.eps, a.x, a.noise = eps.data_ptr(), x.data_ptr(), L.ptr(noise)
    olds = list(old) + [None] * (3 - len(old))
    a.old1, a.old2, a.old3 = L.ptr(olds[0]), L.ptr(olds[1]), L.ptr(olds[2])
    a.e_out, a.x_prev, a.pred_x0 = L.ptr(e_out), x_prev.data_ptr(), pred_x0.data_ptr()
    a.n, a.cfg, a.scale = x.numel(), int(cfg), scale
    cs = list(coefs) + [0.0] * (4 - len(coefs))
    a.c0, a.c1, a.c2, a.c3 = cs
    a.sqrt_one_minus_at, a.sqrt_at, a.sqrt_a_prev, a.dir_coef, a.sigma_temp = (
        sqrt_one_minus_at, sqrt_at, sqrt_a_prev, dir_coef, sigma_temp)
    Stats.launches += 1
    L.check(L.load().mobi_sampler_update(C.byref(a), L.stream()), "sampler_update")
    return x_prev, pred_x0


def assemble_input(x, rest_image, rest_mask, x_in, *, cfg, blend=None):
    """x_in = cat([x, rest...], 1) (twice under CFG). blend = (mask, x0, noise, sqrt_ac, sqrt_1mac) or None."""
    _cuda(x, rest_image, rest_mask, x_in)
    b, _, h, w = x.shape
    a = L.AssembleArgs()
    a.x, a.inpaint_image, a.inpaint_mask, a.x_in = x.data_ptr(), rest_image.data_ptr(), L.ptr(rest_mask), x_in.data_ptr()
    a.B, a.hw, a.cfg = b, h * w, int(cfg)
    a.rest_c = 5 if rest_mask is not None else rest_image.shape[1]
    if blend is not None:
        mask, x0, noise, sa, s1 = blend
        _cuda(mask, x0, noise)
        a.blend_mask, a.blend_x0, a.blend_noise = mask.data_ptr(), x0.data_ptr(), noise.data_ptr()
        a.blend_c, a.sqrt_ac, a.sqrt_1mac = mask.shape[1], sa, s1
    Stats.launches += 1
    L.check(L.load().mobi_assemble_input(C.byref(a), L.stream()), "assemble_input")
    return x_in

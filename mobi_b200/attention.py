"""B200-native drop-ins for ldm/modules/attention.py: same class names, constructor arguments and
state-dict keys as the reference; forward runs on the C-ABI CUDA kernels (mobi_b200.ops).

Internally everything is token-major / channels-last: x is one fp32 residual stream [rows*T, C]; GEMM
operands are bf16; every contraction runs on tcgen05 with fp32 accumulation.  Algebra that removes work
without changing the function (all exact in real arithmetic):
  * attn2 has ONE key, softmax over one key is 1, so attn2(x) = to_out(to_v(c0)) for every token: a
    per-row vector computed once per context and added by the following LayerNorm kernel;
  * cond_adapter_attn has TWO keys: scores are LN(x) . (W_q^T k) and the output is a 2-way blend of
    (W_connector W_out) v, so W_q, W_out and the connector fold into per-row [2*heads, C] tables (U, Z);
  * connector(to_out(o)) is one matrix: (W_connector W_out) o + (W_connector b_out + b_connector).
"""
import math

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .packing import interleave_geglu_pairs

LOG2E = math.log2(math.e)


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


def Normalize(in_channels):
    return nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


def _bf16(t):
    return t.detach().to(torch.bfloat16).contiguous()


def _f32(t):
    return t.detach().float().contiguous()


def repeat_rows2(t):
    """cat([t, t]) along the batch dimension as two device-to-device copies (data movement only); the producer-side
    GroupNorm statistics of `t` (ops.carry_colstats: [2][row groups][C]) are repeated the same way."""
    out = torch.empty((2 * t.shape[0],) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
    out[:t.shape[0]].copy_(t)
    out[t.shape[0]:].copy_(t)
    st = getattr(t, "_colstats", None)
    if st is not None:
        half = st.view(2, -1)
        st2 = torch.empty((2, 2 * half.shape[1]), device=st.device, dtype=st.dtype)
        st2[:, :half.shape[1]].copy_(half)
        st2[:, half.shape[1]:].copy_(half)
        out._colstats = st2.view(-1)
    return out


class GEGLU(nn.Module):
    """attention.py:38-45 (parameter container; the math is the GEGLU epilogue of mobi_gemm)."""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    """attention.py:48-65 with glu=True (the only configuration the UNet uses, attention.py:204)."""

    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.):
        super().__init__()
        if not glu:
            raise NotImplementedError("mobi_b200.FeedForward: only glu=True (gated_ff) is on the hot path")
        inner_dim = int(dim * mult)
        dim_out = dim if dim_out is None else dim_out
        self.net = nn.Sequential(GEGLU(dim, inner_dim), nn.Dropout(dropout), nn.Linear(inner_dim, dim_out))


class CrossAttention(nn.Module):
    """attention.py:153-194 (parameter container; see BasicTransformerBlock for the fused execution)."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner_dim = dim_head * heads
        context_dim = query_dim if context_dim is None else context_dim
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, query_dim), nn.Dropout(dropout))


class BasicTransformerBlock(nn.Module):
    """attention.py:197-266."""

    def __init__(self, dim, n_heads, d_head, dropout=0., context_dim=None, gated_ff=True, checkpoint=False,
                 bbox_cond=False, multimodal=False):
        super().__init__()
        self.bbox_cond = bbox_cond
        self.multimodal = multimodal
        self.dim, self.n_heads, self.d_head = dim, n_heads, d_head
        assert dim == n_heads * d_head
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head,
                                    dropout=dropout)
        if self.bbox_cond:
            self.cond_adapter_attn = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads,
                                                    dim_head=d_head, dropout=dropout)
        if self.multimodal:
            self.cross_modal_attn_camera = CrossAttention(query_dim=dim, context_dim=dim, heads=n_heads,
                                                          dim_head=d_head, dropout=dropout)
            self.cross_modal_attn_lidar = CrossAttention(query_dim=dim, context_dim=dim, heads=n_heads,
                                                         dim_head=d_head, dropout=dropout)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)
        if self.bbox_cond:
            self.cond_adapter_norm = nn.LayerNorm(dim)
            self.cond_adapter_connector = zero_module(nn.Linear(dim, dim))
        if self.multimodal:
            self.cross_modal_norm_camera = nn.LayerNorm(dim)
            self.cross_modal_connector_camera = zero_module(nn.Linear(dim, dim))
            self.cross_modal_norm_lidar = nn.LayerNorm(dim)
            self.cross_modal_connector_lidar = zero_module(nn.Linear(dim, dim))
        self.checkpoint = checkpoint
        self._p = None

    # ------------------------------------------------------------------ packing
    def pack(self):
        C = self.dim
        sc = self.d_head ** -0.5
        p = {}
        a1 = self.attn1
        p["w_qkv"] = _bf16(torch.cat([a1.to_q.weight.float() * (sc * LOG2E), a1.to_k.weight.float(),
                                      a1.to_v.weight.float()], 0))
        p["w_o"], p["b_o"] = _bf16(a1.to_out[0].weight), _f32(a1.to_out[0].bias)
        a2 = self.attn2
        p["w_v2"], p["w_o2"], p["b_o2"] = _bf16(a2.to_v.weight), _bf16(a2.to_out[0].weight), _f32(a2.to_out[0].bias)

        def fold(attn, conn):
            wf = conn.weight.float() @ attn.to_out[0].weight.float()
            bf = conn.weight.float() @ attn.to_out[0].bias.float() + conn.bias.float()
            return _bf16(wf), _f32(bf)

        if self.bbox_cond:
            ca = self.cond_adapter_attn
            p["c_wkv"] = _bf16(torch.cat([ca.to_k.weight.float(), ca.to_v.weight.float()], 0))
            p["c_wqT"] = _bf16((ca.to_q.weight.float() * sc).t())  # [C(c), C(h*d+dd)]
            p["c_wf"], p["c_bf"] = fold(ca, self.cond_adapter_connector)
        if self.multimodal:
            for m in ("camera", "lidar"):
                at = getattr(self, "cross_modal_attn_" + m)
                p[m + "_wq"] = _bf16(at.to_q.weight.float() * (sc * LOG2E))
                p[m + "_wkv"] = _bf16(torch.cat([at.to_k.weight.float(), at.to_v.weight.float()], 0))
                p[m + "_wf"], p[m + "_bf"] = fold(at, getattr(self, "cross_modal_connector_" + m))
        w1, b1 = interleave_geglu_pairs(self.ff.net[0].proj.weight.detach().float(),
                                        self.ff.net[0].proj.bias.detach().float())
        p["w_ff1"], p["b_ff1"] = _bf16(w1), _f32(b1)
        p["w_ff2"], p["b_ff2"] = _bf16(self.ff.net[2].weight), _f32(self.ff.net[2].bias)
        for n in ("norm1", "norm2", "norm3", "cond_adapter_norm", "cross_modal_norm_camera", "cross_modal_norm_lidar"):
            if hasattr(self, n):
                ln = getattr(self, n)
                p[n] = (_f32(ln.weight), _f32(ln.bias))
        old = self._p
        if old is not None and old.keys() == p.keys():
            # refresh IN PLACE: captured CUDA graphs (samplers, UNetTrainer) hold the addresses of these tensors
            for k, v in p.items():
                for dst, src in zip(old[k] if isinstance(v, tuple) else (old[k],), v if isinstance(v, tuple) else (v,)):
                    dst.copy_(src)
            return
        self._p = p

    # ------------------------------------------------------------------ context tables (once per context)
    def context_tables(self, context):
        """context: f32 [R, n_ctx, ctx_dim] on the GPU.  Returns what the block needs from it."""
        p = self._p
        R = context.shape[0]
        C, H, D = self.dim, self.n_heads, self.d_head
        if context.shape[1] > 1 and not self.bbox_cond:
            context = context[:, [0]]  # attention.py:231-232
        c0 = context[:, 0].to(torch.bfloat16).contiguous()
        v0 = ops.gemm(c0, p["w_v2"])                                            # to_v(c0)   bf16 [R, C]
        tab = {"vec2": ops.gemm(v0, p["w_o2"], bias=p["b_o2"], out_dtype=torch.float32)}  # to_out(.) f32 [R, C]
        if self.bbox_cond:
            nk = context.shape[1]
            cc = context.to(torch.bfloat16).reshape(R * nk, -1).contiguous()
            kv = ops.gemm(cc, p["c_wkv"])                                       # bf16 [R*nk, 2C]
            U = torch.empty((R, nk, H, C), device=context.device, dtype=torch.float32)
            Z = torch.empty_like(U)
            for h in range(H):
                # U[r,j,h,:] = sum_dd k[r,j,h*D+dd] * (scale*W_q)[h*D+dd, :]
                ops.gemm(kv[:, h * D:], p["c_wqT"][:, h * D:], out=U[:, :, h], out_dtype=torch.float32, M=R * nk, K=D,
                         lda=2 * C, ldb=C, ldo=H * C)
                ops.gemm(kv[:, C + h * D:], p["c_wf"][:, h * D:], out=Z[:, :, h], out_dtype=torch.float32, M=R * nk,
                         K=D, lda=2 * C, ldb=C, ldo=H * C)
            if nk == 1:
                # one key: softmax == 1, the adapter adds a constant vector per batch row -> merge with attn2's
                tab["vec2"] = (tab["vec2"] + Z.sum(dim=(1, 2)) + p["c_bf"]).contiguous()
                tab["adapter"] = "folded"
            elif nk == 2 and H <= 8:
                # tables of mobi_ln_adapter: slot j = key*8 + head, LayerNorm affine folded in (once per run)
                g, b = p["cond_adapter_norm"]
                Ug = torch.zeros((R, 2, 8, C), device=context.device, dtype=torch.float32)
                Zp = torch.zeros_like(Ug)
                sb = torch.zeros((R, 2, 8), device=context.device, dtype=torch.float32)
                Ug[:, :, :H] = U * g
                Zp[:, :, :H] = Z
                sb[:, :, :H] = (U * b).sum(-1)
                tab["Ug"], tab["Zp"], tab["sb"] = Ug.reshape(R, 16, C), Zp.reshape(R, 16, C), sb.reshape(R, 16)
                tab["adapter"] = "fused"
            else:
                tab["U"], tab["Z"], tab["nk"] = U, Z, nk
                tab["adapter"] = "generic"
        return tab

    # ------------------------------------------------------------------ execution
    def run_attn1(self, x, R, T):
        """Step 1 only (x = attn1(norm1(x)) + x, attention.py:234), in place: the part of the block that does not see
        the context — under classifier-free guidance it is identical for the uncond and cond halves of the batch."""
        p = self._p
        C, H, D = self.dim, self.n_heads, self.d_head
        dev = x.device
        xn = ops.layernorm(x, *p["norm1"])
        q = torch.empty((R * H, T, D), device=dev, dtype=torch.bfloat16)
        k = torch.empty_like(q)
        rowv = D <= 128  # V in its natural layout (MN-major tcgen05 operand); wider heads use the transposed-V kernel
        v = torch.empty_like(q) if rowv else torch.empty((R * H, D, T), device=dev, dtype=torch.bfloat16)
        ops.gemm(xn, p["w_qkv"], epilogue=L.EPI_QKV_ROW if rowv else L.EPI_QKV, heads=H, head_dim=D, tokens=T, out=q,
                 out2=k, out3=v)
        o = ops.attention(q, k, v, R, H, D, T, T, v_rowmajor=rowv)
        ops.gemm(o.reshape(R * T, C), p["w_o"], bias=p["b_o"], residual=x, out=x)
        return x

    def run(self, x, R, T, tab, out_bf16=False, attn1_done=False):
        """x: f32 [R*T, C] residual stream, updated in place.  Returns x (f32) or, with out_bf16, a bf16 copy
        of the block output (what SpatialTransformer.proj_out consumes)."""
        if not attn1_done:
            self.run_attn1(x, R, T)
        return self._run_rest(x, R, T, tab, out_bf16)

    def _run_rest(self, x, R, T, tab, out_bf16):
        p = self._p
        C, H, D = self.dim, self.n_heads, self.d_head
        dev = x.device
        # 2. attn2 == broadcast add of tab["vec2"][row] (attention.py:235), fused into the next normalisation pass
        pending = tab["vec2"]
        cam = (ops.LN_NORM,) + p["cross_modal_norm_camera"] if self.multimodal else None
        lid = (ops.LN_NORM,) + p["cross_modal_norm_lidar"] if self.multimodal else None
        cast = (ops.LN_CAST, None, None)
        n3 = (ops.LN_NORM,) + p["norm3"]
        qn = ctx = xn = None
        # 3. bbox/ref adapter with two keys (attention.py:237-243), fused with the LayerNorm(s) that follow it
        mode = tab.get("adapter") if self.bbox_cond else None
        if mode == "fused":
            g, b = p["cond_adapter_norm"]
            outs = ops.ln_adapter(x, R, T, g, b, tab["Ug"], tab["sb"], tab["Zp"], p["c_bf"],
                                  [cam, cast] if self.multimodal else [n3], pair=self.multimodal, add_vec=pending)
            pending = None
            if self.multimodal:
                qn, ctx = outs
            else:
                xn = outs[0]
        elif mode == "generic":
            xa = ops.layernorm(x, *p["cond_adapter_norm"], add_vec=pending, add_rows_per_vec=T)
            pending = None
            ops.ctx_attention(xa, tab["U"], tab["Z"], p["c_bf"], x, R, T, H, tab["nk"])
        # 4. cross-modal attention (attention.py:245-263): camera rows are even, lidar rows odd
        if self.multimodal:
            Rh = R // 2
            if qn is None:
                if pending is not None:
                    seg = dict(rows=Rh * T, seg=T, seg_stride=2 * T, add_vec=pending, add_rows_per_vec=T)
                    qn = ops.layernorm(x, *p["cross_modal_norm_camera"], seg_offset=0, **seg)
                    ctx = ops.layernorm(x, None, None, seg_offset=T, **seg)
                    pending = None
                else:
                    qn, ctx = ops.ln_dual(x, R, T, [cam, cast])
            self._cross(p, "camera", qn, ctx, x, Rh, T, 0)
            # lidar queries attend to the UPDATED camera tokens (attention.py:259)
            ctx, qn = ops.ln_dual(x, R, T, [cast, lid])
            self._cross(p, "lidar", qn, ctx, x, Rh, T, T)
        # 5. GEGLU feed-forward (attention.py:265)
        if xn is None:
            xn = ops.layernorm(x, *p["norm3"], add_vec=pending, add_rows_per_vec=T)
        hmid = ops.gemm(xn, p["w_ff1"], bias=p["b_ff1"], epilogue=L.EPI_GEGLU2)
        if out_bf16:
            return ops.gemm(hmid, p["w_ff2"], bias=p["b_ff2"], residual=x, out_dtype=torch.bfloat16)
        ops.gemm(hmid, p["w_ff2"], bias=p["b_ff2"], residual=x, out=x)
        return x

    def _cross(self, p, m, qn, ctx, x, Rh, T, row_off):
        C, H, D = self.dim, self.n_heads, self.d_head
        dev = x.device
        q = torch.empty((Rh * H, T, D), device=dev, dtype=torch.bfloat16)
        k = torch.empty_like(q)
        rowv = D <= 128
        v = torch.empty_like(q) if rowv else torch.empty((Rh * H, D, T), device=dev, dtype=torch.bfloat16)
        ops.gemm(qn, p[m + "_wq"], epilogue=L.EPI_HEADS, heads=H, head_dim=D, tokens=T, out=q)
        ops.gemm(ctx, p[m + "_wkv"], epilogue=L.EPI_KV_ROW if rowv else L.EPI_KV, heads=H, head_dim=D, tokens=T, out=k,
                 out2=v)
        o = ops.attention(q, k, v, Rh, H, D, T, T, v_rowmajor=rowv)
        ops.gemm(o.reshape(Rh * T, C), p[m + "_wf"], bias=p[m + "_bf"], residual=x, out=x, ldo=C, out_seg=T,
                 out_seg_stride=2 * T, out_seg_offset=row_off)

    def forward(self, x, context=None):
        """Reference signature: x [B, T, C] f32, context [B, n, ctx] -> [B, T, C]."""
        if self._p is None:
            self.pack()
        B, T, C = x.shape
        xs = x.detach().float().reshape(B * T, C).clone()
        tab = self.context_tables(context.detach().float())
        return self.run(xs, B, T, tab).reshape(B, T, C)


class SpatialTransformer(nn.Module):
    """attention.py:269-313."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0., context_dim=None, bbox_cond=False,
                 multimodal=False):
        super().__init__()
        self.in_channels = in_channels
        inner_dim = n_heads * d_head
        self.inner_dim = inner_dim
        self.norm = Normalize(in_channels)
        self.proj_in = nn.Conv2d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=context_dim,
                                   bbox_cond=bbox_cond, multimodal=multimodal) for _ in range(depth)])
        self.proj_out = zero_module(nn.Conv2d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0))
        self._p = None

    def pack(self):
        self._p = dict(
            gn=(_f32(self.norm.weight), _f32(self.norm.bias)),
            w_in=_bf16(self.proj_in.weight.reshape(self.inner_dim, self.in_channels)), b_in=_f32(self.proj_in.bias),
            w_out=_bf16(self.proj_out.weight.reshape(self.in_channels, self.inner_dim)), b_out=_f32(self.proj_out.bias))
        for b in self.transformer_blocks:
            b.pack()

    def context_tables(self, context):
        return [b.context_tables(context) for b in self.transformer_blocks]

    def run(self, h, tabs, duplicate=False):
        """h: f32 NHWC [R, H, W, C] -> f32 NHWC.
        duplicate: h holds ONE half of a classifier-free-guidance batch whose two halves are identical up to here
        (DDIMSampler.p_sample_ddim feeds cat([x] * 2), ddim.py:180-183).  GroupNorm, proj_in and the first block's
        self-attention do not see the context, so they run once on the half; the residual stream and h are then
        repeated to both halves and the context-dependent rest runs on all rows.  Returns 2R rows."""
        p = self._p
        R, Hh, Ww, C = h.shape
        T = Hh * Ww
        hn = ops.groupnorm(h, p["gn"][0], p["gn"][1], 1e-6, silu=False)
        x = ops.gemm(hn.reshape(R * T, C), p["w_in"], bias=p["b_in"], out_dtype=torch.float32)
        n = len(self.transformer_blocks)
        attn1_done = False
        if duplicate:
            self.transformer_blocks[0].run_attn1(x, R, T)
            x, h = repeat_rows2(x), repeat_rows2(h)
            R, attn1_done = 2 * R, True
        for i, blk in enumerate(self.transformer_blocks):
            x = blk.run(x, R, T, tabs[i], out_bf16=(i == n - 1), attn1_done=(attn1_done and i == 0))
        out = ops.gemm(x, p["w_out"], bias=p["b_out"], residual=h.reshape(R * T, C), out_dtype=torch.float32, colstats=True)
        return ops.carry_colstats(out.reshape(R, Hh, Ww, C), out)

    def forward(self, x, context=None):
        """Reference signature: x NCHW f32 -> NCHW f32."""
        if self._p is None:
            self.pack()
        h = ops.nchw_to_nhwc(x.detach().float().contiguous())
        out = self.run(h, self.context_tables(context.detach().float()))
        return ops.nhwc_to_nchw(out)

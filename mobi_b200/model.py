"""B200-native drop-in for ldm/modules/diffusionmodules/model.py: the first-stage / range autoencoder networks
(`Encoder` 368-489, `Decoder` 492-630, `ResnetBlock` 82-141, `AttnBlock` 150-202, `Upsample` 42-57, `Downsample`
60-79) with the reference's constructor arguments and state-dict keys, including the lidar adapter ((1,5)-kernel
ResnetBlocks, model.py:384-401, 559-578, 615-622).  forward() takes/returns NCHW fp32 like the reference; inside, feature
maps are NHWC (fp32 residual stream, bf16 GEMM/conv operands) and every op is a call into the C-ABI CUDA library: the same
tcgen05 implicit-GEMM conv, GEMM and GroupNorm kernels as the UNet.  The single-head d=512 attention of the mid block
runs as GEMM (QK^T) -> row softmax -> GEMM (PV) per image: its scores (64 MB fp32 at 4096 tokens) stay L2-resident.
"""
import math

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .attention import _bf16, _f32
from .openaimodel import Conv3x3

LOG2E = math.log2(math.e)


def Normalize(in_channels, num_groups=32):
    return nn.GroupNorm(num_groups=num_groups, num_channels=in_channels, eps=1e-6, affine=True)


def _gn(norm):
    return _f32(norm.weight), _f32(norm.bias)


class Upsample(nn.Module):
    """model.py:42-57."""

    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if self.with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)
        self._p = None

    def pack(self):
        self._p = Conv3x3.pack(self.conv) if self.with_conv else {}

    def run(self, h):
        if not self.with_conv:
            return ops.upsample_nearest2x(h)
        return Conv3x3.run(self._p, ops.upsample_nearest2x(h, torch.bfloat16), colstats=True)


class Downsample(nn.Module):
    """model.py:60-79: zero pad right/bottom by one, then a VALID stride-2 conv."""

    def __init__(self, in_channels, with_conv):
        super().__init__()
        if not with_conv:
            raise NotImplementedError("mobi_b200.model.Downsample: avg-pool variant is not used by MObI's configs")
        self.with_conv = with_conv
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)
        self._p = None

    def pack(self):
        self._p = Conv3x3.pack(self.conv)

    def run(self, h):
        return Conv3x3.run(self._p, h, pad_override=(0, 0, 1, 1), colstats=True)


class ResnetBlock(nn.Module):
    """model.py:82-141 with temb_channels = 0 (the autoencoders have no timestep embedding)."""

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout, temb_channels=512,
                 kernel_size=3, padding=1):
        super().__init__()
        if conv_shortcut:
            raise NotImplementedError("mobi_b200.model.ResnetBlock: conv_shortcut is not used by MObI's configs")
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.use_conv_shortcut = conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=1, padding=padding)
        if temb_channels > 0:
            self.temb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=kernel_size, stride=1, padding=padding)
        if self.in_channels != self.out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)
        self._p = None

    def pack(self):
        p = dict(gn1=_gn(self.norm1), gn2=_gn(self.norm2), conv1=Conv3x3.pack(self.conv1), conv2=Conv3x3.pack(self.conv2))
        if self.in_channels != self.out_channels:
            p["w_skip"] = _bf16(self.nin_shortcut.weight.reshape(self.out_channels, self.in_channels))
            p["b_skip"] = _f32(self.nin_shortcut.bias)
        self._p = p

    def run(self, h):
        """h: f32 (or bf16) NHWC residual stream -> f32 NHWC."""
        p = self._p
        n, hh, ww, _ = h.shape
        need_raw = "w_skip" in p
        g = ops.groupnorm(h, p["gn1"][0], p["gn1"][1], 1e-6, silu=True, want_concat=need_raw)
        hn, raw = g if need_raw else (g, None)
        h1 = Conv3x3.run(p["conv1"], hn, colstats=True)
        hn2 = ops.groupnorm(h1, p["gn2"][0], p["gn2"][1], 1e-6, silu=True)
        if need_raw:
            res = ops.gemm(raw.reshape(n * hh * ww, -1), p["w_skip"], bias=p["b_skip"], out_dtype=torch.float32)
            res = res.reshape(n, hh, ww, self.out_channels)
        else:
            res = h
        return Conv3x3.run(p["conv2"], hn2, residual=res, colstats=True)


class AttnBlock(nn.Module):
    """model.py:150-202: single-head self-attention over all pixels, scale C^-0.5."""

    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.k = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.v = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.proj_out = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self._p = None

    def pack(self):
        c = self.in_channels
        sc = float(int(c) ** (-0.5)) * LOG2E  # softmax scale and the exp -> exp2 change of base, folded into q
        w = lambda m: m.weight.detach().float().reshape(c, c)
        self._p = dict(
            gn=_gn(self.norm),
            w_qk=_bf16(torch.cat([w(self.q) * sc, w(self.k)], 0)),
            b_qk=_f32(torch.cat([self.q.bias.detach().float() * sc, self.k.bias.detach().float()], 0)),
            w_v=_bf16(w(self.v)), b_v=_f32(self.v.bias), w_o=_bf16(w(self.proj_out)), b_o=_f32(self.proj_out.bias))

    def run(self, h):
        p = self._p
        n, hh, ww, c = h.shape
        t = hh * ww
        hn = ops.groupnorm(h, p["gn"][0], p["gn"][1], 1e-6, silu=False).reshape(n * t, c)
        qk = ops.gemm(hn, p["w_qk"], bias=p["b_qk"])                                   # bf16 [n*t, 2c]
        vt = torch.empty((n, c, t), device=h.device, dtype=torch.bfloat16)             # V^T per image
        ops.gemm(hn, p["w_v"], bias=p["b_v"], epilogue=L.EPI_HEADS_T, heads=1, head_dim=c, tokens=t, out=vt)
        o = torch.empty((n * t, c), device=h.device, dtype=torch.bfloat16)
        # all images in three launches (batched tcgen05 GEMMs around one row softmax); images are processed in groups whose
        # fp32 score matrices (t x t each) stay below ~1 GB
        grp = max(1, min(n, (1 << 30) // (4 * t * t)))
        s = torch.empty((grp, t, t), device=h.device, dtype=torch.float32)
        pr = torch.empty((grp, t, t), device=h.device, dtype=torch.bfloat16)
        for i0 in range(0, n, grp):
            g = min(grp, n - i0)
            qi = qk[i0 * t:(i0 + g) * t]
            ops.gemm(qi[:, :c], qi[:, c:], out=s[:g], out_dtype=torch.float32, M=t, N=t, K=c, lda=2 * c, ldb=2 * c, ldo=t,
                     batch=g, a_batch_stride=t * 2 * c, b_batch_stride=t * 2 * c, out_batch_stride=t * t)   # S = Q K^T (log2 units)
            ops.softmax_rows(s[:g].reshape(g * t, t), out=pr[:g].reshape(g * t, t))
            ops.gemm(pr[:g], vt[i0:i0 + g], out=o[i0 * t:(i0 + g) * t], out_dtype=torch.bfloat16, M=t, N=c, K=t, lda=t,
                     ldb=t, ldo=c, batch=g, a_batch_stride=t * t, b_batch_stride=c * t, out_batch_stride=t * c)   # O = P V
        out = ops.gemm(o, p["w_o"], bias=p["b_o"], residual=h.reshape(n * t, c), out_dtype=torch.float32, colstats=True)
        return ops.carry_colstats(out.reshape(n, hh, ww, c), out)


def make_attn(in_channels, attn_type="vanilla"):
    if attn_type != "vanilla":
        raise NotImplementedError("mobi_b200.model: only attn_type='vanilla' is used by MObI's configs")
    return AttnBlock(in_channels)


def _pack_all(module):
    for m in module.modules():
        if m is not module and hasattr(m, "pack") and hasattr(m, "_p"):
            m.pack()


class Encoder(nn.Module):
    """model.py:368-489."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, lidar_adapter=False,
                 dropout=0.0, resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True,
                 use_linear_attn=False, attn_type="vanilla", **ignore_kwargs):
        super().__init__()
        self.ch = ch
        self.temb_ch = 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.lidar_adapter = lidar_adapter
        if self.lidar_adapter:
            self.conv_in_lidar = nn.Conv2d(in_channels, self.ch, kernel_size=(1, 5), stride=1, padding=(0, 2))
            self.res_block_lidar1 = ResnetBlock(in_channels=self.ch, out_channels=self.ch, temb_channels=0,
                                                kernel_size=(1, 5), padding=(0, 2), dropout=dropout)
            self.res_block_lidar2 = ResnetBlock(in_channels=self.ch, out_channels=self.ch, temb_channels=0,
                                                kernel_size=(1, 5), padding=(0, 2), dropout=dropout)
        else:
            self.conv_in = nn.Conv2d(in_channels, self.ch, kernel_size=3, stride=1, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.in_ch_mult = in_ch_mult
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block = nn.ModuleList()
            attn = nn.ModuleList()
            block_in = ch * in_ch_mult[i_level]
            block_out = ch * ch_mult[i_level]
            for _ in range(self.num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=0, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            down = nn.Module()
            down.block = block
            down.attn = attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res = curr_res // 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, kernel_size=3, stride=1, padding=1)
        self._p = None

    def pack(self):
        _pack_all(self)
        self._p = dict(conv_in=Conv3x3.pack(self.conv_in_lidar if self.lidar_adapter else self.conv_in),
                       gn_out=_gn(self.norm_out), conv_out=Conv3x3.pack(self.conv_out))

    def run(self, h):
        """h: NHWC f32 image -> NHWC f32 [N, h/8, w/8, 2*z]."""
        p = self._p
        h = Conv3x3.run(p["conv_in"], h, colstats=True)
        if self.lidar_adapter:
            h = self.res_block_lidar1.run(h)
            h = self.res_block_lidar2.run(h)
        for i_level in range(self.num_resolutions):
            for i_block in range(self.num_res_blocks):
                h = self.down[i_level].block[i_block].run(h)
                if len(self.down[i_level].attn) > 0:
                    h = self.down[i_level].attn[i_block].run(h)
            if i_level != self.num_resolutions - 1:
                h = self.down[i_level].downsample.run(h)
        h = self.mid.block_1.run(h)
        h = self.mid.attn_1.run(h)
        h = self.mid.block_2.run(h)
        hn = ops.groupnorm(h, p["gn_out"][0], p["gn_out"][1], 1e-6, silu=True)
        return Conv3x3.run(p["conv_out"], hn)

    @torch.no_grad()
    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("mobi_b200.model.Encoder runs on CUDA only (no CPU fallback)")
        if self._p is None:
            self.pack()
        return ops.nhwc_to_nchw(self.run(ops.nchw_to_nhwc(x.detach().float().contiguous())))


class Decoder(nn.Module):
    """model.py:492-630."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, lidar_adapter=False,
                 dropout=0.0, resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False,
                 tanh_out=False, use_linear_attn=False, attn_type="vanilla", **ignorekwargs):
        super().__init__()
        if give_pre_end or tanh_out:
            raise NotImplementedError("mobi_b200.model.Decoder: give_pre_end / tanh_out are not used by MObI's configs")
        self.ch = ch
        self.temb_ch = 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.give_pre_end = give_pre_end
        self.tanh_out = tanh_out
        self.lidar_adapter = lidar_adapter
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = nn.Conv2d(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block = nn.ModuleList()
            attn = nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(self.num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=0, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            up = nn.Module()
            up.block = block
            up.attn = attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res = curr_res * 2
            self.up.insert(0, up)
        if self.lidar_adapter:
            self.res_block_lidar1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0,
                                                kernel_size=(1, 5), padding=(0, 2), dropout=dropout)
            self.norm_out_lidar1 = Normalize(block_in)
            self.res_block_lidar2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0,
                                                kernel_size=(1, 5), padding=(0, 2), dropout=dropout)
            self.norm_out_lidar2 = Normalize(block_in)
            self.conv_out_lidar = nn.Conv2d(block_in, out_ch, kernel_size=(1, 5), stride=1, padding=(0, 2))
        else:
            self.norm_out = Normalize(block_in)
            self.conv_out = nn.Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)
        self._p = None

    def pack(self):
        _pack_all(self)
        p = dict(conv_in=Conv3x3.pack(self.conv_in))
        if self.lidar_adapter:
            p.update(gn_l1=_gn(self.norm_out_lidar1), gn_l2=_gn(self.norm_out_lidar2),
                     conv_out=Conv3x3.pack(self.conv_out_lidar))
        else:
            p.update(gn_out=_gn(self.norm_out), conv_out=Conv3x3.pack(self.conv_out))
        self._p = p

    def run(self, z):
        """z: NHWC f32 latent [N, h, w, z_channels] -> NHWC f32 image [N, 8h, 8w, out_ch]."""
        p = self._p
        self.last_z_shape = z.shape
        h = Conv3x3.run(p["conv_in"], z, colstats=True)
        h = self.mid.block_1.run(h)
        h = self.mid.attn_1.run(h)
        h = self.mid.block_2.run(h)
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks + 1):
                h = self.up[i_level].block[i_block].run(h)
                if len(self.up[i_level].attn) > 0:
                    h = self.up[i_level].attn[i_block].run(h)
            if i_level != 0:
                h = self.up[i_level].upsample.run(h)
        if self.lidar_adapter:
            h = self.res_block_lidar1.run(h)
            # norm + swish between the two lidar blocks ("a small mistake" the checkpoints were trained with,
            # model.py:617-618): its output is the INPUT of the next block, i.e. the residual stream itself
            h = ops.groupnorm(h, p["gn_l1"][0], p["gn_l1"][1], 1e-6, silu=True, out_dtype=torch.float32)
            h = self.res_block_lidar2.run(h)
            hn = ops.groupnorm(h, p["gn_l2"][0], p["gn_l2"][1], 1e-6, silu=True)
        else:
            hn = ops.groupnorm(h, p["gn_out"][0], p["gn_out"][1], 1e-6, silu=True)
        return Conv3x3.run(p["conv_out"], hn)

    @torch.no_grad()
    def forward(self, z):
        if not z.is_cuda:
            raise RuntimeError("mobi_b200.model.Decoder runs on CUDA only (no CPU fallback)")
        if self._p is None:
            self.pack()
        return ops.nhwc_to_nchw(self.run(ops.nchw_to_nhwc(z.detach().float().contiguous())))

"""B200-native drop-ins for the conditioning encoders that run right before the denoising loop:
ldm/modules/encoders/modules.py (FrozenCLIPImageEmbedder minus the CLIP tower, BBoxEmbedder) and
ldm/modules/encoders/xf.py (Transformer / ResidualAttentionBlock / MultiheadAttention / MLP / LayerNorm), with the
reference's class names, constructor arguments and state-dict keys; forward runs on the C-ABI CUDA kernels.

The frozen CLIP vision tower itself (a Hugging Face `CLIPVisionModel`, 304 M parameters, one 224 x 224 reference crop per
sample) is NOT re-implemented: `FrozenCLIPImageEmbedder` takes it as an injected module (or loads it with
`from_pretrained` when the weights are available) and starts from its `pooler_output`.

Algebra (exact in real arithmetic): the mapper is a Transformer over ONE token, so every attention is a softmax over a
single key: attn(x) = c_proj(W_v LN(x) + b_v); q, k and the softmax are dead, and c_proj o W_v folds into one matrix.
"""
import torch
import torch.nn as nn

from . import ops
from .attention import _bf16, _f32


class LayerNorm(nn.LayerNorm):
    """xf.py:22-28 (parameter container: fp32 statistics are what the CUDA LayerNorm kernel does anyway)."""


class MultiheadAttention(nn.Module):
    """xf.py:31-45 (parameter container)."""

    def __init__(self, n_ctx, width, heads):
        super().__init__()
        self.n_ctx, self.width, self.heads = n_ctx, width, heads
        self.c_qkv = nn.Linear(width, width * 3)
        self.c_proj = nn.Linear(width, width)


class MLP(nn.Module):
    """xf.py:48-59 (parameter container)."""

    def __init__(self, width):
        super().__init__()
        self.width = width
        self.c_fc = nn.Linear(width, width * 4)
        self.c_proj = nn.Linear(width * 4, width)


class ResidualAttentionBlock(nn.Module):
    """xf.py:79-98."""

    def __init__(self, n_ctx, width, heads):
        super().__init__()
        self.attn = MultiheadAttention(n_ctx, width, heads)
        self.ln_1 = LayerNorm(width)
        self.mlp = MLP(width)
        self.ln_2 = LayerNorm(width)


class Transformer(nn.Module):
    """xf.py:101-130 for n_ctx = 1 (the only use in MObI: `Transformer(1, 1024, 5, 1)`, modules.py:155)."""

    def __init__(self, n_ctx, width, layers, heads):
        super().__init__()
        if n_ctx != 1:
            raise NotImplementedError("mobi_b200.encoders.Transformer: MObI maps ONE CLIP token (n_ctx = 1)")
        self.n_ctx, self.width, self.layers = n_ctx, width, layers
        self.resblocks = nn.ModuleList([ResidualAttentionBlock(n_ctx, width, heads) for _ in range(layers)])
        self._p = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    def invalidate(self):
        self._p = None

    @torch.no_grad()
    def pack(self):
        p = []
        for blk in self.resblocks:
            a, w = blk.attn, self.width
            heads = a.heads
            ch = w // heads
            # c_qkv output is viewed as [heads, 3*ch] and split into q | k | v per head (xf.py:69-71): the v rows
            idx = torch.cat([torch.arange(h * 3 * ch + 2 * ch, (h + 1) * 3 * ch) for h in range(heads)]).to(a.c_qkv.weight.device)
            wv, bv = a.c_qkv.weight.detach().float()[idx], a.c_qkv.bias.detach().float()[idx]
            wp, bp = a.c_proj.weight.detach().float(), a.c_proj.bias.detach().float()
            p.append(dict(ln1=(_f32(blk.ln_1.weight), _f32(blk.ln_1.bias)), ln2=(_f32(blk.ln_2.weight), _f32(blk.ln_2.bias)),
                          w_attn=_bf16(wp @ wv), b_attn=_f32(wp @ bv + bp),
                          w_fc=_bf16(blk.mlp.c_fc.weight), b_fc=_f32(blk.mlp.c_fc.bias),
                          w_proj=_bf16(blk.mlp.c_proj.weight), b_proj=_f32(blk.mlp.c_proj.bias)))
        self._p = p

    def run(self, x):
        """x: f32 [B, width], updated in place."""
        if self._p is None:
            self.pack()
        for p in self._p:
            h = ops.layernorm(x, *p["ln1"])
            ops.gemm(h, p["w_attn"], bias=p["b_attn"], residual=x, out=x)                 # x + c_proj(v)
            h = ops.layernorm(x, *p["ln2"])
            m = ops.gemm(h, p["w_fc"], bias=p["b_fc"], act=2)                              # GELU(c_fc(.))
            ops.gemm(m, p["w_proj"], bias=p["b_proj"], residual=x, out=x)
        return x

    def forward(self, x):
        """Reference signature: [B, 1, width] -> [B, 1, width]."""
        if not x.is_cuda:
            raise RuntimeError("mobi_b200.encoders run on CUDA only (no CPU fallback)")
        b, n, w = x.shape
        assert n == 1 and w == self.width
        return self.run(x.detach().float().reshape(b, w).clone()).reshape(b, 1, w)


class BBoxEmbedder(nn.Module):
    """modules.py:181-213: Fourier features of the 8 box corners -> 4 Linears (SiLU after the 2nd and 3rd)."""

    def __init__(self, embedder_num_freqs=4, proj_dims=(768, 512, 512, 768)):
        super().__init__()
        self.num_freqs = embedder_num_freqs
        self.out_dim = 3 * (1 + 2 * embedder_num_freqs)
        proj_dims = list(proj_dims)
        self.bbox_proj = nn.Linear(self.out_dim * 8, proj_dims[0])
        self.second_linear = nn.Sequential(nn.Linear(proj_dims[0], proj_dims[1]), nn.SiLU(),
                                           nn.Linear(proj_dims[1], proj_dims[2]), nn.SiLU(),
                                           nn.Linear(proj_dims[2], proj_dims[3]))
        self._p = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    def invalidate(self):
        self._p = None

    @torch.no_grad()
    def pack(self):
        lin = [self.bbox_proj, self.second_linear[0], self.second_linear[2], self.second_linear[4]]
        self._p = [(_bf16(m.weight), _f32(m.bias)) for m in lin]

    def forward(self, bbox, out=None):
        """bbox [B, 8, 3] -> [B, 1, proj_dims[-1]] f32 (written into the row view `out` [B, dim] when given)."""
        if not bbox.is_cuda:
            raise RuntimeError("mobi_b200.encoders run on CUDA only (no CPU fallback)")
        if self._p is None:
            self.pack()
        b = bbox.shape[0]
        assert tuple(bbox.shape[1:]) == (8, 3)
        e = ops.fourier_embed(bbox.detach().float().contiguous(), self.num_freqs).reshape(b, 8 * self.out_dim)
        p = self._p
        h = ops.gemm(e, p[0][0], bias=p[0][1])
        h = ops.gemm(h, p[1][0], bias=p[1][1], act=1)
        h = ops.gemm(h, p[2][0], bias=p[2][1], act=1)
        return ops.gemm(h, p[3][0], bias=p[3][1], out=out, out_dtype=torch.float32).unsqueeze(1)

    def encode(self, cond):
        return {"ref_bbox_token": self(cond["ref_bbox"])}


class FrozenCLIPImageEmbedder(nn.Module):
    """modules.py:141-179.  `transformer` (the CLIP vision tower) is injected or loaded from `version`; everything after
    its pooler_output runs on the CUDA kernels."""

    def __init__(self, conditions, version="openai/clip-vit-large-patch14", transformer=None):
        super().__init__()
        if "ref_image" in conditions:
            if transformer is None:
                from transformers import CLIPVisionModel  # needs the checkpoint on disk: there is no network here
                transformer = CLIPVisionModel.from_pretrained(version)
            self.transformer = transformer
            self.final_ln = LayerNorm(1024)
            self.mapper = Transformer(1, 1024, 5, 1)
        if "ref_bbox" in conditions:
            self.bbox_embedder = BBoxEmbedder()
        self.freeze()

    def freeze(self):
        if hasattr(self, "transformer"):
            self.transformer = self.transformer.eval()
        for param in self.parameters():
            param.requires_grad = False

    def map_pooled(self, pooled):
        """pooler_output [B, 1024] -> final_ln(mapper(.)) [B, 1, 1024] (modules.py:166-169)."""
        x = self.mapper.run(pooled.detach().float().contiguous().clone())
        ln = ops.layernorm(x, _f32(self.final_ln.weight), _f32(self.final_ln.bias))      # bf16 [B, 1024]
        return ln.unsqueeze(1)

    @torch.no_grad()
    def forward(self, image):
        return self.map_pooled(self.transformer(pixel_values=image).pooler_output)

    def encode(self, cond):
        ret = {}
        if "ref_image" in cond:
            ret["ref_image_token"] = self(cond["ref_image"])
        if "ref_bbox" in cond:
            ret["ref_bbox_token"] = self.bbox_embedder(cond["ref_bbox"])
        return ret


@torch.no_grad()
def learned_conditioning(cond_stage_model, proj_out, cond, cond_stage_key=("ref_image", "ref_bbox")):
    """LatentDiffusion.get_learned_conditioning (ddpm.py:610-630): encode, project the image token to the context
    width with `proj_out` (an nn.Linear(1024, 768) of the LatentDiffusion module) and concatenate -> [B, n, 768]."""
    keys = [k for k in ("ref_image", "ref_bbox") if k in cond_stage_key]
    b = cond[keys[0]].shape[0]
    dim = proj_out.weight.shape[0]
    out = torch.empty((b, len(keys), dim), device=cond[keys[0]].device, dtype=torch.float32)   # the cat of ddpm.py:623-630
    for i, k in enumerate(keys):
        if k == "ref_image":
            tok = cond_stage_model(cond["ref_image"])                         # bf16 [B, 1, 1024]
            ops.gemm(tok.reshape(b, -1), _bf16(proj_out.weight), bias=_f32(proj_out.bias), out=out[:, i])
        else:
            cond_stage_model.bbox_embedder(cond["ref_bbox"], out=out[:, i])
    return out

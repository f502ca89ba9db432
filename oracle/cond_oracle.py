"""TEST INFRASTRUCTURE ONLY (oracle) — fp32 restatement of the conditioning encoders that run right before the denoising
loop: everything of `LatentDiffusion.get_learned_conditioning` after the frozen CLIP vision tower.

  ref image : CLIP pooler_output [B, 1024] -> Transformer(n_ctx=1, width=1024, layers=5, heads=1) -> LayerNorm(1024)
              (FrozenCLIPImageEmbedder.forward, ldm/modules/encoders/modules.py:163-169; xf.py:22-130)
              -> proj_out Linear(1024, 768)                                   (ldm/models/diffusion/ddpm.py:622)
  bbox      : 8 corners x 3 coords -> Fourier features (include_input, 4 log-spaced frequencies, [sin, cos]) -> 216
              -> Linear(216, 768) -> Linear(768, 512) -> SiLU -> Linear(512, 512) -> SiLU -> Linear(512, 768)
              (BBoxEmbedder, modules.py:181-213; Embedder / get_embedder, modules.py:215-262)
  cond      : cat([ref_image_token, ref_bbox_token], dim=1)  -> [B, 2, 768]   (ddpm.py:623-630)

Pure functions over reference-format state dicts (keys as under `cond_stage_model.` / `proj_out.` of the reference's
LatentDiffusion).  Pinned by tests/golden/cond_tiny.npz, produced by the unmodified reference modules.
"""
import math

import torch
import torch.nn.functional as F


def shapes(width=1024, layers=5, out_dim=768, bbox_dims=(768, 512, 512, 768), num_freqs=4):
    s = {}
    for i in range(layers):
        p = "cond_stage_model.mapper.resblocks.%d." % i
        s[p + "attn.c_qkv.weight"] = (3 * width, width)
        s[p + "attn.c_qkv.bias"] = (3 * width,)
        s[p + "attn.c_proj.weight"] = (width, width)
        s[p + "attn.c_proj.bias"] = (width,)
        s[p + "ln_1.weight"] = (width,)
        s[p + "ln_1.bias"] = (width,)
        s[p + "mlp.c_fc.weight"] = (4 * width, width)
        s[p + "mlp.c_fc.bias"] = (4 * width,)
        s[p + "mlp.c_proj.weight"] = (width, 4 * width)
        s[p + "mlp.c_proj.bias"] = (width,)
        s[p + "ln_2.weight"] = (width,)
        s[p + "ln_2.bias"] = (width,)
    s["cond_stage_model.final_ln.weight"] = (width,)
    s["cond_stage_model.final_ln.bias"] = (width,)
    s["proj_out.weight"] = (out_dim, width)
    s["proj_out.bias"] = (out_dim,)
    fdim = 3 * (1 + 2 * num_freqs) * 8
    b = "cond_stage_model.bbox_embedder."
    s[b + "bbox_proj.weight"] = (bbox_dims[0], fdim)
    s[b + "bbox_proj.bias"] = (bbox_dims[0],)
    dims = list(bbox_dims)
    for j, idx in enumerate((0, 2, 4)):
        s[b + "second_linear.%d.weight" % idx] = (dims[j + 1], dims[j])
        s[b + "second_linear.%d.bias" % idx] = (dims[j + 1],)
    return s


def fourier_embed(x, num_freqs=4):
    """Embedder.__call__ (modules.py:215-252): [x, sin(x f0), cos(x f0), sin(x f1), ...], f = 2^linspace(0, num_freqs-1)."""
    outs = [x]
    for f in 2.0 ** torch.linspace(0.0, num_freqs - 1, steps=num_freqs):
        outs.append(torch.sin(x * f))
        outs.append(torch.cos(x * f))
    return torch.cat(outs, -1)


def bbox_token(sd, bbox, num_freqs=4):
    """BBoxEmbedder.forward (modules.py:203-209): bbox [B, 8, 3] -> [B, 1, 768]."""
    p = "cond_stage_model.bbox_embedder."
    e = fourier_embed(bbox.float(), num_freqs).reshape(bbox.shape[0], -1)
    h = F.linear(e, sd[p + "bbox_proj.weight"], sd[p + "bbox_proj.bias"])
    h = F.linear(h, sd[p + "second_linear.0.weight"], sd[p + "second_linear.0.bias"])
    h = F.linear(F.silu(h), sd[p + "second_linear.2.weight"], sd[p + "second_linear.2.bias"])
    h = F.linear(F.silu(h), sd[p + "second_linear.4.weight"], sd[p + "second_linear.4.bias"])
    return h.unsqueeze(1)


def _attention_one_head(qkv, heads):
    """QKVMultiheadAttention.forward (xf.py:63-76) for any n_ctx."""
    bs, n_ctx, width = qkv.shape
    ch = width // heads // 3
    scale = 1 / math.sqrt(math.sqrt(ch))
    qkv = qkv.view(bs, n_ctx, heads, -1)
    q, k, v = torch.split(qkv, ch, dim=-1)
    w = torch.softmax(torch.einsum("bthc,bshc->bhts", q * scale, k * scale).float(), dim=-1)
    return torch.einsum("bhts,bshc->bthc", w, v).reshape(bs, n_ctx, -1)


def image_token(sd, pooled, layers=5, heads=1):
    """FrozenCLIPImageEmbedder.forward after the CLIP tower (modules.py:165-169) + proj_out (ddpm.py:622)."""
    x = pooled.float().unsqueeze(1)
    for i in range(layers):
        p = "cond_stage_model.mapper.resblocks.%d." % i
        h = F.layer_norm(x, (x.shape[-1],), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        h = F.linear(h, sd[p + "attn.c_qkv.weight"], sd[p + "attn.c_qkv.bias"])
        h = _attention_one_head(h, heads)
        x = x + F.linear(h, sd[p + "attn.c_proj.weight"], sd[p + "attn.c_proj.bias"])
        h = F.layer_norm(x, (x.shape[-1],), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        h = F.gelu(F.linear(h, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]))
        x = x + F.linear(h, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
    x = F.layer_norm(x, (x.shape[-1],), sd["cond_stage_model.final_ln.weight"], sd["cond_stage_model.final_ln.bias"], 1e-5)
    return F.linear(x, sd["proj_out.weight"], sd["proj_out.bias"])


def learned_conditioning(sd, pooled, bbox):
    """get_learned_conditioning (ddpm.py:610-630) with cond_stage_key = [ref_image, ref_bbox]: [B, 2, 768]."""
    return torch.cat([image_token(sd, pooled), bbox_token(sd, bbox)], dim=1)

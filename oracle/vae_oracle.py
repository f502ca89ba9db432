"""TEST INFRASTRUCTURE ONLY (oracle) — fp32 restatement of the first-stage / range autoencoder.

Restates ldm/modules/diffusionmodules/model.py (Encoder 368-489, Decoder 492-630, ResnetBlock 82-141, AttnBlock
150-202, Upsample 42-57, Downsample 60-79) and ldm/models/autoencoder.py:63-72 (AutoencoderKL.encode/decode) as
pure functions over a reference-format state_dict (`AutoencoderKL.state_dict()` keys).
"""
import torch
import torch.nn.functional as F


def default_ddconfig(lidar=False, **kw):
    """configs/mobi_nusc_512.yaml:84-127."""
    cfg = dict(double_z=True, z_channels=4, resolution=512, in_channels=2 if lidar else 3, out_ch=2 if lidar else 3,
               ch=128, ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0, lidar_adapter=lidar)
    cfg.update(kw)
    return cfg


def tiny_ddconfig(lidar=False, **kw):
    cfg = default_ddconfig(lidar, resolution=64, ch=64, ch_mult=[1, 2], num_res_blocks=1)
    cfg.update(kw)
    return cfg


def _norm(x, sd, p):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)  # model.py:38-39


def _conv(x, sd, p, padding):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=padding)


def resnet_block(sd, p, x, padding=1):
    """ResnetBlock.forward, model.py:121-141 (temb is None)."""
    h = _conv(F.silu(_norm(x, sd, p + ".norm1")), sd, p + ".conv1", padding)
    h = _conv(F.silu(_norm(h, sd, p + ".norm2")), sd, p + ".conv2", padding)
    if (p + ".nin_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + ".nin_shortcut.weight"], sd[p + ".nin_shortcut.bias"])
    return x + h


def attn_block(sd, p, x):
    """AttnBlock.forward, model.py:178-202: single head, scale C^-0.5."""
    h_ = _norm(x, sd, p + ".norm")
    q = F.conv2d(h_, sd[p + ".q.weight"], sd[p + ".q.bias"])
    k = F.conv2d(h_, sd[p + ".k.weight"], sd[p + ".k.bias"])
    v = F.conv2d(h_, sd[p + ".v.weight"], sd[p + ".v.bias"])
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    h_ = F.conv2d(h_, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return x + h_


def decoder_forward(sd, cfg, z, prefix="decoder"):
    """Decoder.forward, model.py:587-630."""
    nres = len(cfg["ch_mult"])
    p = prefix
    h = _conv(z, sd, p + ".conv_in", 1)
    h = resnet_block(sd, p + ".mid.block_1", h)
    h = attn_block(sd, p + ".mid.attn_1", h)
    h = resnet_block(sd, p + ".mid.block_2", h)
    for i_level in reversed(range(nres)):
        for i_block in range(cfg["num_res_blocks"] + 1):
            h = resnet_block(sd, "%s.up.%d.block.%d" % (p, i_level, i_block), h)
            # attn_resolutions is [] in every shipped config (configs/mobi_nusc_512.yaml:99)
        if i_level != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(h, sd, "%s.up.%d.upsample.conv" % (p, i_level), 1)
    if cfg.get("lidar_adapter"):
        h = resnet_block(sd, p + ".res_block_lidar1", h, padding=(0, 2))
        h = F.silu(_norm(h, sd, p + ".norm_out_lidar1"))  # the acknowledged extra norm+swish, model.py:617-618
        h = resnet_block(sd, p + ".res_block_lidar2", h, padding=(0, 2))
        h = F.silu(_norm(h, sd, p + ".norm_out_lidar2"))
        h = _conv(h, sd, p + ".conv_out_lidar", (0, 2))
    else:
        h = F.silu(_norm(h, sd, p + ".norm_out"))
        h = _conv(h, sd, p + ".conv_out", 1)
    return h


def encoder_forward(sd, cfg, x, prefix="encoder"):
    """Encoder.forward, model.py:454-489."""
    p = prefix
    nres = len(cfg["ch_mult"])
    if cfg.get("lidar_adapter"):
        h = _conv(x, sd, p + ".conv_in_lidar", (0, 2))
        h = resnet_block(sd, p + ".res_block_lidar1", h, padding=(0, 2))
        h = resnet_block(sd, p + ".res_block_lidar2", h, padding=(0, 2))
    else:
        h = _conv(x, sd, p + ".conv_in", 1)
    for i_level in range(nres):
        for i_block in range(cfg["num_res_blocks"]):
            h = resnet_block(sd, "%s.down.%d.block.%d" % (p, i_level, i_block), h)
        if i_level != nres - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)  # model.py:72-76
            q = "%s.down.%d.downsample.conv" % (p, i_level)
            h = F.conv2d(h, sd[q + ".weight"], sd[q + ".bias"], stride=2, padding=0)
    h = resnet_block(sd, p + ".mid.block_1", h)
    h = attn_block(sd, p + ".mid.attn_1", h)
    h = resnet_block(sd, p + ".mid.block_2", h)
    h = F.silu(_norm(h, sd, p + ".norm_out"))
    return _conv(h, sd, p + ".conv_out", 1)


def vae_decode(sd, cfg, z):
    """AutoencoderKL.decode, autoencoder.py:69-72."""
    z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    return decoder_forward(sd, cfg, z)


def vae_encode_moments(sd, cfg, x):
    """AutoencoderKL.encode, autoencoder.py:63-67: returns the moments (mean | logvar) tensor; the posterior
    mode is moments[:, :z] (distributions.py:24-43)."""
    h = encoder_forward(sd, cfg, x)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


def decode_first_stage(sd, cfg, z, scale_factor=0.18215):
    """LatentDiffusion.decode_first_stage, ddpm.py:837-849, 893-899 (first_stage_key == 'inpaint' keeps z[:, :4])."""
    return vae_decode(sd, cfg, (1.0 / scale_factor * z)[:, :4])


def decode_sample(sample):
    """LatentDiffusion.decode_sample, ddpm.py:1420-1447 for use_camera and use_lidar with equal latent sizes."""
    return sample[::2], sample[1::2]


def state_dict_shapes(cfg, embed_dim=4, with_encoder=True):
    """name -> shape for AutoencoderKL.state_dict() (encoder.*, decoder.*, quant_conv, post_quant_conv)."""
    shapes = {}
    ch, mult, nrb, z = cfg["ch"], cfg["ch_mult"], cfg["num_res_blocks"], cfg["z_channels"]
    lidar = cfg.get("lidar_adapter", False)

    def conv(p, o, i, kh, kw):
        shapes[p + ".weight"] = (o, i, kh, kw)
        shapes[p + ".bias"] = (o,)

    def norm(p, c):
        shapes[p + ".weight"] = (c,)
        shapes[p + ".bias"] = (c,)

    def res(p, i, o, k=(3, 3)):
        norm(p + ".norm1", i)
        conv(p + ".conv1", o, i, *k)
        norm(p + ".norm2", o)
        conv(p + ".conv2", o, o, *k)
        if i != o:
            conv(p + ".nin_shortcut", o, i, 1, 1)

    def attn(p, c):
        norm(p + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv(p + "." + n, c, c, 1, 1)

    if with_encoder:
        p = "encoder"
        if lidar:
            conv(p + ".conv_in_lidar", ch, cfg["in_channels"], 1, 5)
            res(p + ".res_block_lidar1", ch, ch, (1, 5))
            res(p + ".res_block_lidar2", ch, ch, (1, 5))
        else:
            conv(p + ".conv_in", ch, cfg["in_channels"], 3, 3)
        in_mult = (1,) + tuple(mult)
        bi = ch
        for lvl in range(len(mult)):
            bi = ch * in_mult[lvl]
            bo = ch * mult[lvl]
            for b in range(nrb):
                res("%s.down.%d.block.%d" % (p, lvl, b), bi, bo)
                bi = bo
            if lvl != len(mult) - 1:
                conv("%s.down.%d.downsample.conv" % (p, lvl), bi, bi, 3, 3)
        res(p + ".mid.block_1", bi, bi)
        attn(p + ".mid.attn_1", bi)
        res(p + ".mid.block_2", bi, bi)
        norm(p + ".norm_out", bi)
        conv(p + ".conv_out", 2 * z if cfg["double_z"] else z, bi, 3, 3)
        conv("quant_conv", 2 * embed_dim, 2 * z, 1, 1)
    p = "decoder"
    bi = ch * mult[-1]
    conv(p + ".conv_in", bi, z, 3, 3)
    res(p + ".mid.block_1", bi, bi)
    attn(p + ".mid.attn_1", bi)
    res(p + ".mid.block_2", bi, bi)
    for lvl in reversed(range(len(mult))):
        bo = ch * mult[lvl]
        for b in range(nrb + 1):
            res("%s.up.%d.block.%d" % (p, lvl, b), bi, bo)
            bi = bo
        if lvl != 0:
            conv("%s.up.%d.upsample.conv" % (p, lvl), bi, bi, 3, 3)
    if lidar:
        res(p + ".res_block_lidar1", bi, bi, (1, 5))
        norm(p + ".norm_out_lidar1", bi)
        res(p + ".res_block_lidar2", bi, bi, (1, 5))
        norm(p + ".norm_out_lidar2", bi)
        conv(p + ".conv_out_lidar", cfg["out_ch"], bi, 1, 5)
    else:
        norm(p + ".norm_out", bi)
        conv(p + ".conv_out", cfg["out_ch"], bi, 3, 3)
    conv("post_quant_conv", z, embed_dim, 1, 1)
    return shapes

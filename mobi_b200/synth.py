"""Synthetic weights and inputs for benchmarking (there is no network for checkpoints or nuScenes).

Random init of the reference architecture is not usable as-is: `zero_module` (openaimodel.py:229-231, 836;
attention.py:218-223, 296-300) makes a fresh UNet output exactly 0.  Every matrix/filter is therefore drawn
N(0, 1/fan_in), biases N(0, 0.1^2), norm scales 1 + N(0, 0.1^2), on the device, from a seeded generator.
"""
import math

import torch


@torch.no_grad()
def init_synthetic_(module, seed=0):
    dev = next(module.parameters()).device
    g = torch.Generator(device=dev).manual_seed(seed)
    for name, p in module.named_parameters():
        if p.dim() >= 2:
            fan_in = p[0].numel()
            p.copy_(torch.randn(p.shape, generator=g, device=dev) * fan_in ** -0.5)
        elif name.endswith(".weight"):
            p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g, device=dev))
        else:
            p.copy_(0.1 * torch.randn(p.shape, generator=g, device=dev))
    if hasattr(module, "invalidate"):
        module.invalidate()
    for m in module.modules():
        if hasattr(m, "invalidate"):
            m.invalidate()
    return module


def mobi_unet_config(latent=64, use_lidar=True):
    """configs/mobi_nusc_512.yaml:62-81 (latent 64) / mobi_nusc_256.yaml (latent 32) / pbe.yaml (use_lidar False)."""
    return dict(image_size=latent, in_channels=9, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                transformer_depth=1, context_dim=768, use_checkpoint=False, legacy=False,
                add_conv_in_front_of_unet=False, bbox_cond=True, use_camera=True, use_lidar=use_lidar)


def vae_ddconfig(latent=64, lidar=False):
    """configs/mobi_nusc_512.yaml:84-127 (first_stage_config / lidar_stage_config ddconfig)."""
    return dict(double_z=True, z_channels=4, resolution=8 * latent, in_channels=2 if lidar else 3,
                out_ch=2 if lidar else 3, ch=128, ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[],
                dropout=0.0, lidar_adapter=lidar)


def build_synthetic_ldm(latent=64, use_lidar=True, device="cuda", seed=0, unet_cfg=None, with_vae=False, with_cond=False):
    """with_cond: also the conditioning stage after the CLIP tower (mapper Transformer, final_ln, BBoxEmbedder, proj_out),
    entered at the tower's pooler_output (PooledFeatureTower)."""
    from .ddpm import LatentDiffusion
    cfg = unet_cfg or mobi_unet_config(latent, use_lidar)
    conditions = ["ref_image", "ref_bbox"]
    vae = lambda lidar: dict(target="mobi_b200.autoencoder.AutoencoderKL",
                             params=dict(ddconfig=vae_ddconfig(latent, lidar), embed_dim=4,
                                         lossconfig=dict(target="torch.nn.Identity")))
    with torch.device("meta"):
        ldm = LatentDiffusion(unet_config=dict(target="mobi_b200.openaimodel.UNetModel", params=cfg),
                              first_stage_config=vae(False) if with_vae else None,
                              lidar_stage_config=vae(True) if (with_vae and use_lidar) else None,
                              cond_stage_config=dict(target="mobi_b200.encoders.FrozenCLIPImageEmbedder",
                                                     params=dict(conditions=conditions, transformer=PooledFeatureTower()))
                              if with_cond else None,
                              cond_stage_key=conditions if with_cond else "image",
                              linear_start=0.00085, linear_end=0.0120, timesteps=1000, first_stage_key="inpaint",
                              image_size=cfg["image_size"], channels=4, conditioning_key="crossattn",
                              scale_factor=0.18215, lidar_scale_factor=0.18215, use_camera=True, use_lidar=use_lidar)
    ldm = ldm.to_empty(device=device)
    ldm.register_schedule(linear_start=0.00085, linear_end=0.0120, timesteps=1000)  # buffers were meta: rebuild
    ldm = ldm.to(device)
    init_synthetic_(ldm.model.diffusion_model, seed)
    if with_vae:
        init_synthetic_(ldm.first_stage_model, seed + 1)
        if ldm.lidar_stage_model is not None:
            init_synthetic_(ldm.lidar_stage_model, seed + 2)
    g = torch.Generator(device=next(ldm.parameters()).device).manual_seed(seed + 3)
    with torch.no_grad():   # to_empty() left these uninitialised: [learnable_vector, bbox_uncond_vector] ~ N(0, 1)
        for p in (ldm.learnable_vector, ldm.bbox_uncond_vector):
            p.copy_(torch.randn(p.shape, generator=g, device=p.device))
    if with_cond:
        init_synthetic_(ldm.cond_stage_model, seed + 4)
        init_synthetic_(ldm.proj_out, seed + 5)
    return ldm.eval()


def synthetic_inputs(n_joint, latent, context_dim=768, seed=1, device="cpu", rows_per_sample=2, n_ctx=2, pin=False):
    """x_T, inpaint_image ~ N(0,1); mask = 1 outside a centred ~20 % box; cond, uc ~ N(0,1)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = rows_per_sample * n_joint
    x_T = torch.randn(r, 4, latent, latent, generator=g)
    inpaint_image = torch.randn(r, 4, latent, latent, generator=g)
    mask = torch.ones(r, 1, latent, latent)
    side = max(1, int(round(latent * (0.2 ** 0.5))))
    lo = (latent - side) // 2
    mask[:, :, lo:lo + side, lo:lo + side] = 0.0
    cond = torch.randn(r, n_ctx, context_dim, generator=g)
    uc = torch.randn(1, n_ctx, context_dim, generator=g).repeat(r, 1, 1)
    out = dict(x_T=x_T, inpaint_image=inpaint_image, inpaint_mask=mask, cond=cond, uc=uc)
    if pin and torch.cuda.is_available():
        out = {k: v.pin_memory() for k, v in out.items()}
    if device != "cpu":
        out = {k: v.to(device) for k, v in out.items()}
    return out


def synthetic_lidar_batch(n, px=512, H=32, W=1096, seed=3, device="cuda", obj_range_m=10.0):
    """Synthetic stand-in for `batch["lidar"]` of the reference dataset (ldm/data/nuscenes.py:470-489), a decoded lidar
    image [n, 3, px, px] and one edited box per sample [n, 8, 3] (corner order of box_np_ops.center_to_corner_box3d,
    box_np_ops.py:48-78, 212-237), for the range-view post-processing bench: a box ~obj_range_m ahead whose direction sits
    in the middle of the crop window, object-normalised depths in the middle of the crop, 15 % holes in the sweep."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    u = lambda *s: torch.rand(*s, generator=g)                                              # noqa: E731
    beams = torch.tensor([0.0232 * x for x in range(-23, 9)])
    col = torch.randint(0, W, (n,), generator=g)
    wc = torch.tensor([64, 128, 256, 512])[torch.arange(n) % 4].clamp(max=px)
    yaw0 = math.pi * (torch.arange(W, dtype=torch.float32) / W * 2 - 1)
    pitch = beams.flip(0)[torch.linspace(0, 31, H).round().long()][None, :, None].expand(n, H, W) + (u(n, H, W) - 0.5) * 0.004
    yaw = yaw0[None, None].expand(n, H, W) + (u(n, H, W) - 0.5) * 0.002
    d_orig = u(n, H, W) * 1.85 - 0.95
    d_orig[u(n, H, W) < 0.15] = -1
    gt = torch.zeros(n, H, W)
    for b in range(n):
        gt[b, H // 2:H // 2 + 3, [(int(col[b]) + k) % W for k in range(-2, 3)]] = 1
    d_obj = 2 * obj_range_m / 54 - 1
    dec = u(n, 3, px, px) * 2.2 - 1.1                                                       # the clamp has work to do
    dec[:, 0, :, int(px * 0.3):int(px * 0.7)] = u(n, px, int(px * 0.7) - int(px * 0.3)) * 1.5 - 0.75
    yc = yaw0[col]
    center = torch.stack([obj_range_m * torch.cos(yc), -obj_range_m * torch.sin(yc), torch.full((n,), -0.4)], 1)
    norm = torch.tensor([[0, 0, 0], [0, 0, 1], [0, 1, 1], [0, 1, 0], [1, 0, 0], [1, 0, 1], [1, 1, 1], [1, 1, 0]],
                        dtype=torch.float32) - 0.5
    c = norm[None] * torch.tensor([4.6, 2.2, 1.8])
    s, co = torch.sin(-yc), torch.cos(-yc)
    z, o = torch.zeros(n), torch.ones(n)
    rot_t = torch.stack([torch.stack([co, -s, z], 1), torch.stack([s, co, z], 1), torch.stack([z, z, o], 1)], 1)
    bbox = torch.bmm(c.expand(n, 8, 3).contiguous(), rot_t) + center[:, None]
    batch = dict(range_depth_orig=d_orig, range_int_orig=u(n, H, W) * 0.6, range_pitch=pitch.contiguous(),
                 range_yaw=yaw.contiguous(), range_instance_mask_orig=gt, range_shift_left=(col - wc // 2) % W + W,
                 width_crop=wc, min_depth_obj=torch.full((n,), d_obj - 0.03), max_depth_obj=torch.full((n,), d_obj + 0.03))
    return ({k: v.to(device) for k, v in batch.items()}, dec.to(device), bbox.to(device))


class PooledFeatureTower(torch.nn.Module):
    """Stand-in for the CLIP vision tower (out of scope: SURVEY.md §2; no checkpoint without a network): the "image" it
    is given already IS the tower's pooler_output [B, 1024], so FrozenCLIPImageEmbedder.forward runs everything that
    follows the tower (mapper Transformer, final_ln; then proj_out in get_learned_conditioning) unchanged."""

    def forward(self, pixel_values=None):
        from types import SimpleNamespace
        return SimpleNamespace(pooler_output=pixel_values)


def synthetic_dataset_batch(n, px=512, seed=7, pin=True):
    """A batch with the layout of the reference dataset (ldm/data/nuscenes.py:452-489) filled with synthetic data, on the
    HOST (pinned): camera / range images in [-1, 1], hole masks (1 = keep) with a centred ~20 % box, box corners in [0, 1]
    image coordinates, CLIP pooler features [n, 1024] in place of the reference crop (see PooledFeatureTower), and the
    sweep-side tensors of synthetic_lidar_batch for the post-processing."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    u = lambda *s: torch.rand(*s, generator=g) * 2 - 1                                      # noqa: E731
    mask = torch.ones(n, 1, px, px)
    side = int(round(px * 0.2 ** 0.5))
    lo = (px - side) // 2
    mask[:, :, lo:lo + side, lo:lo + side] = 0.0
    sweep, _, bbox_3d = synthetic_lidar_batch(n, px=px, seed=seed + 1, device="cpu")
    img_gt, rng_gt = u(n, 3, px, px), u(n, 2, px, px)
    batch = dict(
        image=dict(GT=img_gt, inpaint_image=img_gt * mask, inpaint_mask=mask.clone(),
                   cond=dict(ref_image=torch.randn(n, 1024, generator=g), ref_bbox=torch.rand(n, 8, 3, generator=g))),
        lidar=dict(range_data=rng_gt, range_data_inpaint=rng_gt * mask, range_mask=mask.clone(),
                   cond=dict(ref_image=torch.randn(n, 1024, generator=g), ref_bbox=torch.rand(n, 8, 3, generator=g)), **sweep),
        bbox_3d=bbox_3d)
    if pin and torch.cuda.is_available():
        pinned = lambda d: {k: (pinned(v) if isinstance(v, dict) else v.contiguous().pin_memory()) for k, v in d.items()}  # noqa: E731
        batch = pinned(batch)
    return batch

"""TEST INFRASTRUCTURE ONLY (oracle) — fp32 restatement of the input assembly that feeds the training step and the
inpainting loop: LatentDiffusion.encode_all_stages (ddpm.py:1010-1033) and the tail of
LatentDiffusion.get_input (ddpm.py:758-834), with the posterior noise passed in (the reference draws it from the global
RNG inside DiagonalGaussianDistribution.sample, distributions.py:35-37).  Pinned by tests/golden/get_input_tiny.npz.
"""
import torch
import torch.nn.functional as F

from . import vae_oracle as vo


def posterior_sample(moments, noise):
    """DiagonalGaussianDistribution (distributions.py:24-37): mean + exp(0.5 * clamp(logvar, -30, 20)) * noise."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    return mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise


def encode_stage(sd, cfg, gt, inpaint, mask, noise_gt, noise_inpaint, scale_factor):
    """One modality of encode_all_stages (ddpm.py:1013-1031): [N, 9, h, w] = [z | z_inpaint | mask]."""
    z = scale_factor * posterior_sample(vo.vae_encode_moments(sd, cfg, gt), noise_gt)
    zi = scale_factor * posterior_sample(vo.vae_encode_moments(sd, cfg, inpaint), noise_inpaint)
    m = F.interpolate(mask, size=z.shape[-1], mode="nearest")
    return torch.cat((z, zi, m), dim=1)


def align_lidar(z_lidar, bbox, image_size):
    """get_input, ddpm.py:797-816: centre-crop the lidar latent's width to the image latent, pad (or crop, when negative)
    its rows, and move the box corners' x / y into the cropped frame.  Returns (z, bbox)."""
    W = z_lidar.shape[-1]
    left, right = W // 2 - image_size // 2, W // 2 + image_size // 2
    pad = (image_size - z_lidar.shape[-2]) // 2
    z = F.pad(z_lidar[..., left:right], (0, 0, pad, pad), mode="constant", value=0)
    bbox = bbox.clone()
    bbox[..., 0] = (bbox[..., 0] * W - left) / image_size
    bbox[..., 1] += pad / image_size
    return z, bbox


def cat_interleave(tensors):
    """ldm/util.py:213-221."""
    t = torch.cat([u.unsqueeze(1) for u in tensors], dim=1)
    return t.reshape(-1, *t.shape[2:])


def get_input(cam_sd, cam_cfg, lid_sd, lid_cfg, g, image_size, scale_factor=0.18215, lidar_scale_factor=0.18215):
    """g: dict of tensors named like the golden file.  Returns dict(z, z_lidar, bbox)."""
    z_image = encode_stage(cam_sd, cam_cfg, g["image_gt"], g["image_inpaint"], g["image_mask"], g["noise_cam_gt"],
                           g["noise_cam_inpaint"], scale_factor)
    z_lidar = encode_stage(lid_sd, lid_cfg, g["range_gt"], g["range_inpaint"], g["range_mask"], g["noise_lid_gt"],
                           g["noise_lid_inpaint"], lidar_scale_factor)
    z_l, bbox_l = align_lidar(z_lidar, g["bbox_lidar_in"], image_size)
    return dict(z=cat_interleave([z_image, z_l]), z_lidar=z_lidar[:, :4], bbox=cat_interleave([g["bbox_camera"], bbox_l]))

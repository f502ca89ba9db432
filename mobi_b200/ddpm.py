"""Drop-in for the parts of ldm/models/diffusion/ddpm.py that sit on the sampling hot path:
`LatentDiffusion.apply_model` (ddpm.py:1060-1157), `DiffusionWrapper` (1681-1722), `q_sample` (284-287),
`register_schedule` (127-179), `decode_first_stage` (837-901) and `decode_sample` (1420-1447).

Not rebuilt (out of scope): training_step/optimizers, log_images, get_input's dataset plumbing,
the CLIP/bbox conditioning encoders — `apply_model` starts from the already-encoded [R, n_ctx, 768] context.
State-dict keys match the reference (`model.diffusion_model.*`, `first_stage_model.*`, `lidar_stage_model.*`,
schedule buffers), so reference checkpoints load with strict=False exactly as the reference does.
"""
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from .util import instantiate_from_config


def extract_into_tensor(a, t, x_shape):
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """ldm/modules/diffusionmodules/util.py:21-43 ("linear" only: the schedule every config uses)."""
    if schedule != "linear":
        raise NotImplementedError("mobi_b200: only the 'linear' beta schedule is on the hot path")
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64, device="cpu") ** 2).numpy()


class DiffusionWrapper(nn.Module):
    """ddpm.py:1681-1722."""

    def __init__(self, diff_model_config, conditioning_key):
        super().__init__()
        self.diffusion_model = instantiate_from_config(diff_model_config).eval()
        self.conditioning_key = conditioning_key
        assert self.conditioning_key in [None, "concat", "crossattn", "hybrid", "adm"]

    def forward(self, x, t, c_concat=None, c_crossattn=None):
        if self.conditioning_key == "crossattn":
            cc = c_crossattn[0] if len(c_crossattn) == 1 else torch.cat(c_crossattn, 1)
            return self.diffusion_model(x, t, context=cc)
        raise NotImplementedError("mobi_b200: conditioning_key %r is not used by MObI" % (self.conditioning_key,))


class LatentDiffusion(nn.Module):
    """Inference-side LatentDiffusion with the reference's constructor names (extra kwargs are accepted and ignored
    so the YAML `params:` block can be passed through unchanged)."""

    def __init__(self, unet_config, first_stage_config=None, lidar_stage_config=None, cond_stage_config=None,
                 timesteps=1000, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3,
                 conditioning_key="crossattn", parameterization="eps", first_stage_key="image", image_size=256,
                 channels=3, scale_factor=1.0, lidar_scale_factor=1.0, use_camera=True, use_lidar=False,
                 v_posterior=0.0, **ignored):
        super().__init__()
        assert parameterization == "eps"
        self.parameterization = parameterization
        self.first_stage_key = first_stage_key
        self.image_size = image_size
        self.channels = channels
        self.scale_factor = scale_factor
        self.lidar_scale_factor = lidar_scale_factor
        self.use_camera = use_camera
        self.use_lidar = use_lidar
        self.v_posterior = v_posterior
        self.model = DiffusionWrapper(unet_config, conditioning_key)
        self.first_stage_model = instantiate_from_config(first_stage_config) if (first_stage_config and use_camera) else None
        self.lidar_stage_model = instantiate_from_config(lidar_stage_config) if (lidar_stage_config and use_lidar) else None
        self.learnable_vector = nn.Parameter(torch.randn((1, 1, 768)), requires_grad=False)     # ddpm.py:476
        self.bbox_uncond_vector = nn.Parameter(torch.randn((1, 1, 768)), requires_grad=False)   # ddpm.py:477
        self.register_schedule(beta_schedule=beta_schedule, timesteps=timesteps, linear_start=linear_start,
                               linear_end=linear_end, cosine_s=cosine_s)

    @property
    def device(self):
        return self.betas.device

    def register_schedule(self, given_betas=None, beta_schedule="linear", timesteps=1000, linear_start=1e-4,
                          linear_end=2e-2, cosine_s=8e-3):
        """ddpm.py:127-179 (float64 numpy -> float32 buffers)."""
        betas = given_betas if given_betas is not None else make_beta_schedule(
            beta_schedule, timesteps, linear_start=linear_start, linear_end=linear_end, cosine_s=cosine_s)
        alphas = 1.0 - betas
        alphas_cumprod = np.cumprod(alphas, axis=0)
        alphas_cumprod_prev = np.append(1.0, alphas_cumprod[:-1])
        self.num_timesteps = int(betas.shape[0])
        self.linear_start, self.linear_end = linear_start, linear_end
        to_torch = partial(torch.tensor, dtype=torch.float32, device="cpu")
        self.register_buffer("betas", to_torch(betas))
        self.register_buffer("alphas_cumprod", to_torch(alphas_cumprod))
        self.register_buffer("alphas_cumprod_prev", to_torch(alphas_cumprod_prev))
        self.register_buffer("sqrt_alphas_cumprod", to_torch(np.sqrt(alphas_cumprod)))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", to_torch(np.sqrt(1.0 - alphas_cumprod)))
        self.register_buffer("log_one_minus_alphas_cumprod", to_torch(np.log(1.0 - alphas_cumprod)))
        self.register_buffer("sqrt_recip_alphas_cumprod", to_torch(np.sqrt(1.0 / alphas_cumprod)))
        self.register_buffer("sqrt_recipm1_alphas_cumprod", to_torch(np.sqrt(1.0 / alphas_cumprod - 1)))
        posterior_variance = (1 - self.v_posterior) * betas * (1.0 - alphas_cumprod_prev) / (
            1.0 - alphas_cumprod) + self.v_posterior * betas
        self.register_buffer("posterior_variance", to_torch(posterior_variance))
        self.register_buffer("posterior_log_variance_clipped", to_torch(np.log(np.maximum(posterior_variance, 1e-20))))
        self.register_buffer("posterior_mean_coef1", to_torch(betas * np.sqrt(alphas_cumprod_prev) / (1.0 - alphas_cumprod)))
        self.register_buffer("posterior_mean_coef2",
                             to_torch((1.0 - alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - alphas_cumprod)))

    def q_sample(self, x_start, t, noise=None):
        """ddpm.py:284-287."""
        noise = torch.randn_like(x_start) if noise is None else noise
        return (extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start +
                extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    @torch.no_grad()
    def apply_model(self, x_noisy, t, cond, return_ids=False):
        """ddpm.py:1060-1157 (the split_input_params branch raises NotImplementedError in the reference too)."""
        if not isinstance(cond, dict):
            if not isinstance(cond, list):
                cond = [cond]
            key = "c_concat" if self.model.conditioning_key == "concat" else "c_crossattn"
            cond = {key: cond}
        x_recon = self.model(x_noisy, t, **cond)
        if isinstance(x_recon, tuple) and not return_ids:
            return x_recon[0]
        return x_recon

    @torch.no_grad()
    def decode_first_stage(self, z, predict_cids=False, force_not_quantize=False, module_name="first_stage_model"):
        """ddpm.py:837-901 (AutoencoderKL branch)."""
        assert module_name in ["first_stage_model", "lidar_stage_model"]
        assert not predict_cids
        module = getattr(self, module_name)
        sf = self.scale_factor if module_name == "first_stage_model" else self.lidar_scale_factor
        if self.first_stage_key == "inpaint":
            z = z[:, :4, :, :]
        z = z.detach().float().contiguous()
        if z.is_cuda:
            from . import ops
            z = ops.scale_f32(z, 1.0 / sf)       # z / scale_factor (ddpm.py:846-849)
        else:
            z = 1.0 / sf * z
        return module.decode(z)

    def decode_sample(self, sample, z_lidar=None):
        """ddpm.py:1420-1447."""
        h_camera, h_lidar = None, None
        if self.use_camera and self.use_lidar:
            h_camera = sample[::2]
            bottom = (sample[1::2].shape[-2] - z_lidar.shape[-2]) // 2
            top = bottom + z_lidar.shape[-2]
            h_lidar = sample[1::2][:, :, bottom:top, :]
            if self.image_size != z_lidar.shape[-1]:
                z_lidar[..., z_lidar.shape[-1] // 2 - self.image_size // 2: z_lidar.shape[-1] // 2 + self.image_size // 2] = h_lidar
                h_lidar = z_lidar
        elif self.use_camera:
            h_camera = sample
        else:
            bottom = (sample[1::2].shape[-2] - z_lidar.shape[-2]) // 2
            top = bottom + z_lidar.shape[-2]
            h_lidar = sample[:, :, bottom:top, :]
            if self.image_size != z_lidar.shape[-1]:
                z_lidar[..., z_lidar.shape[-1] // 2 - self.image_size // 2: z_lidar.shape[-1] // 2 + self.image_size // 2] = h_lidar
                h_lidar = z_lidar
        return h_camera, h_lidar

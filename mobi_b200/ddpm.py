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
                 v_posterior=0.0, cond_stage_key="image", **ignored):
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
        # conditioning stage (ddpm.py:564-586): "__is_unconditional__" / None -> no encoder (the path starts from tokens)
        self.cond_stage_key = list(cond_stage_key) if isinstance(cond_stage_key, (list, tuple)) else cond_stage_key
        self.cond_stage_model = instantiate_from_config(cond_stage_config) if isinstance(cond_stage_config, dict) else None
        self.proj_out = nn.Linear(1024, 768).requires_grad_(False)                               # ddpm.py:479
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

    @torch.no_grad()
    def get_learned_conditioning(self, c):
        """ddpm.py:610-630: encode the raw conditioning dict, project the image token, concatenate -> [B, n, 768]."""
        from . import encoders
        if self.cond_stage_model is None:
            raise RuntimeError("get_learned_conditioning: this LatentDiffusion was built without a cond_stage_config")
        keys = self.cond_stage_key if isinstance(self.cond_stage_key, list) else ["ref_image"]
        return encoders.learned_conditioning(self.cond_stage_model, self.proj_out, c, cond_stage_key=keys)

    # ------------------------------------------------------------------ input assembly
    @torch.no_grad()
    def encode_first_stage(self, x, module_name="first_stage_model"):
        """ddpm.py:970-1008 (the patch-wise split_input_params branch is not used by MObI's configs)."""
        assert module_name in ["first_stage_model", "lidar_stage_model"]
        return getattr(self, module_name).encode(x)

    def get_first_stage_encoding(self, encoder_posterior, scale_factor=1, noise=None):
        """ddpm.py:600-608; `noise` overrides the posterior's RNG draw (parity tests)."""
        from .autoencoder import DiagonalGaussianDistribution
        if isinstance(encoder_posterior, DiagonalGaussianDistribution):
            z = encoder_posterior.sample(noise)
        elif isinstance(encoder_posterior, torch.Tensor):
            z = encoder_posterior
        else:
            raise NotImplementedError("encoder_posterior of type '%s' not yet implemented" % type(encoder_posterior))
        return scale_factor * z

    def _sample_into(self, m_gt, m_in, noise, mask, scale, out, row_stride=1, row_offset=0, left=0, pad=0):
        """ONE kernel: sample both posteriors, nearest-resize the mask, write the 9-channel rows (cropped / padded /
        interleaved) in place (mobi_assemble_latent_input)."""
        import ctypes as C

        from . import _lib as L
        n, _, hs, ws = m_gt.shape
        assert hs == ws, "the reference resizes the mask to a square of the latent width (ddpm.py:1020, 1030)"
        mask = mask.detach().float().contiguous()
        a = L.LatentInputArgs()
        a.moments_gt, a.moments_inpaint = m_gt.data_ptr(), m_in.data_ptr()
        a.noise_gt, a.noise_inpaint = noise[0].data_ptr(), noise[1].data_ptr()
        a.mask, a.out = mask.data_ptr(), out.data_ptr()
        a.n, a.hs, a.ws, a.hm, a.wm = n, hs, ws, mask.shape[-2], mask.shape[-1]
        a.S, a.left, a.pad, a.row_stride, a.row_offset, a.scale = out.shape[-1], left, pad, row_stride, row_offset, scale
        L.check(L.load().mobi_assemble_latent_input(C.byref(a), L.stream()), "assemble_latent_input")
        return out

    def _encode_pair(self, module_name, gt, inpaint, noise):
        m_gt = self.encode_first_stage(gt, module_name).parameters.contiguous()
        m_in = self.encode_first_stage(inpaint, module_name).parameters.contiguous()
        if noise is None:   # DiagonalGaussianDistribution.sample draws (distributions.py:35-37)
            shape = (m_gt.shape[0], 4) + tuple(m_gt.shape[2:])
            noise = (torch.randn(shape, device=m_gt.device), torch.randn(shape, device=m_gt.device))
        return m_gt, m_in, tuple(t.detach().float().contiguous() for t in noise)

    @torch.no_grad()
    def encode_all_stages(self, image_gt, image_inpaint, image_mask, range_gt, range_inpaint, range_mask, noise=None):
        """ddpm.py:1010-1033.  Returns (z_image, z_lidar), each [N, 9, h, w] = [z | z_inpaint | mask].  `noise` =
        optional dict(camera=(n_gt, n_inpaint), lidar=(n_gt, n_inpaint)) replacing the posterior RNG draws."""
        z_image = z_lidar = None
        noise = noise or {}
        if self.use_camera:
            m_gt, m_in, nz = self._encode_pair("first_stage_model", image_gt, image_inpaint, noise.get("camera"))
            z_image = torch.empty((m_gt.shape[0], 9) + tuple(m_gt.shape[2:]), device=m_gt.device, dtype=torch.float32)
            self._sample_into(m_gt, m_in, nz, image_mask, self.scale_factor, z_image)
        if self.use_lidar:
            m_gt, m_in, nz = self._encode_pair("lidar_stage_model", range_gt, range_inpaint, noise.get("lidar"))
            z_lidar = torch.empty((m_gt.shape[0], 9) + tuple(m_gt.shape[2:]), device=m_gt.device, dtype=torch.float32)
            self._sample_into(m_gt, m_in, nz, range_mask, self.lidar_scale_factor, z_lidar)
        return z_image, z_lidar

    @torch.no_grad()
    def get_input(self, batch, k="inpaint", force_c_encode=False, bs=None, return_vae_rec=False, noise=None):
        """ddpm.py:758-834 for the joint camera + lidar model: batch = {"image": {GT, inpaint_image, inpaint_mask, cond},
        "lidar": {range_data, range_data_inpaint, range_mask, cond}}.  Returns {"z": [2N, 9, S, S] rows interleaved
        cam0, lid0, cam1, ..., "cond": {key: interleaved raw conditioning}, "z_lidar": [N, 4, hl, wl]}.  The lidar box
        corners are re-normalised IN PLACE like the reference does (ddpm.py:815-816).  The conditioning encoders are a
        separate step (mobi_b200.encoders.learned_conditioning)."""
        from . import _lib as L
        assert k == "inpaint" and self.use_camera and self.use_lidar, "get_input: joint camera + lidar model only"
        assert not force_c_encode and not return_vae_rec, "get_input: encode conditioning / decode reconstructions separately"
        image_data, lidar_data = batch["image"], batch["lidar"]
        if bs is not None:
            sel = lambda d: {kk: (sel(v) if isinstance(v, dict) else v[:bs]) for kk, v in d.items()}
            image_data, lidar_data = sel(image_data), sel(lidar_data)
        noise = noise or {}
        n, S = image_data["GT"].shape[0], self.image_size
        z = torch.empty((2 * n, 9, S, S), device=image_data["GT"].device, dtype=torch.float32)
        m_gt, m_in, nz = self._encode_pair("first_stage_model", image_data["GT"], image_data["inpaint_image"],
                                           noise.get("camera"))
        assert m_gt.shape[-1] == S and m_gt.shape[-2] == S, "camera latent must have the UNet's image_size"
        self._sample_into(m_gt, m_in, nz, image_data["inpaint_mask"], self.scale_factor, z, 2, 0)
        # the lidar latent is centre-cropped in width and padded (cropped when negative) in height (ddpm.py:797-812)
        m_gt, m_in, nz = self._encode_pair("lidar_stage_model", lidar_data["range_data"], lidar_data["range_data_inpaint"],
                                           noise.get("lidar"))
        hl, wl = m_gt.shape[-2], m_gt.shape[-1]
        left, pad = wl // 2 - S // 2, (S - hl) // 2
        self._sample_into(m_gt, m_in, nz, lidar_data["range_mask"], self.lidar_scale_factor, z, 2, 1, left, pad)
        # un-cropped lidar latent for decode_sample (ddpm.py:819): same kernel, identity window
        z_lidar = torch.empty((n, 9, hl, wl), device=z.device, dtype=torch.float32)
        self._sample_into(m_gt, m_in, nz, lidar_data["range_mask"], self.lidar_scale_factor, z_lidar)
        bbox = lidar_data["cond"]["ref_bbox"]
        assert bbox.is_cuda and bbox.dtype == torch.float32 and bbox.is_contiguous()
        L.check(L.load().mobi_bbox_renorm(bbox.data_ptr(), bbox.numel() // 3, wl, left, S, pad, L.stream()), "bbox_renorm")
        cond = {}
        for key in image_data["cond"]:
            c, l = image_data["cond"][key], lidar_data["cond"][key]
            cond[key] = torch.stack([c, l], dim=1).reshape((-1,) + tuple(c.shape[1:]))    # cat_interleave (ldm/util.py:213-221)
        return {"z": z, "cond": cond, "z_lidar": z_lidar[:, :4]}

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

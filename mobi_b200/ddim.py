"""Drop-in for ldm/models/diffusion/ddim.py (DDIMSampler): same constructor, `make_schedule`, `sample`,
`ddim_sampling` and `p_sample_ddim` signatures, defaults and return values.  The per-step tensor math of
p_sample_ddim (ddim.py:172-212) runs in two fused CUDA kernels; with the native UNet the model call is a
replayed CUDA graph."""
import numpy as np
import torch

from . import ops
from .sampling import SamplerBase, StepCoefficients


class DDIMSampler(SamplerBase):
    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, **kwargs):
        """ddim.py:56-112."""
        if conditioning is not None:
            cbs = conditioning[list(conditioning.keys())[0]].shape[0] if isinstance(conditioning, dict) \
                else conditioning.shape[0]
            if cbs != batch_size:
                print(f"Warning: Got {cbs} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        size = (batch_size, C, H, W)
        if verbose:
            print(f"Data shape for DDIM sampling is {size}, eta {eta}")
        return self.ddim_sampling(conditioning, size, callback=callback, img_callback=img_callback,
                                  quantize_denoised=quantize_x0, mask=mask, x0=x0, ddim_use_original_steps=False,
                                  noise_dropout=noise_dropout, temperature=temperature,
                                  score_corrector=score_corrector, corrector_kwargs=corrector_kwargs, x_T=x_T,
                                  log_every_t=log_every_t, unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning, **kwargs)

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100, temperature=1.,
                      noise_dropout=0., score_corrector=None, corrector_kwargs=None, unconditional_guidance_scale=1.,
                      unconditional_conditioning=None, **kwargs):
        """ddim.py:114-163."""
        self._check_unsupported(quantize_denoised, score_corrector, noise_dropout)
        if ddim_use_original_steps:
            raise NotImplementedError("mobi_b200.DDIMSampler: ddim_use_original_steps is not used by MObI")
        device = self.model.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device).float().clone()
        if timesteps is None:
            timesteps = self.ddim_timesteps
        else:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        # with mask / x0 the step kernel blends the known latent into img IN PLACE (the reference builds a new tensor,
        # ddim.py:148): logged intermediates are snapshots then, so a later step cannot rewrite them
        keep = (lambda t: t.clone()) if mask is not None else (lambda t: t)
        intermediates = {"x_inter": [keep(img)], "pred_x0": [keep(img)]}
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]
        if kwargs.get("verbose_steps", False):
            print(f"Running DDIM Sampling with {total_steps} timesteps")
        rest_image, rest_mask = self._rest_from_kwargs(kwargs)
        rest_c = 5 if rest_mask is not None else rest_image.shape[1]
        self._setup_eval(b, (4 + rest_c,) + tuple(shape[2:]), cond, unconditional_conditioning,
                         unconditional_guidance_scale)
        if mask is not None:
            assert x0 is not None
            mask = mask.to(device).float().contiguous()
            x0 = x0.to(device).float().contiguous()
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            blend = None
            if mask is not None:  # img = q_sample(x0, ts) * mask + (1 - mask) * img   (ddim.py:145-148)
                noise = torch.randn_like(x0)
                blend = (mask, x0, noise, float(self._sqrt_ac_host[int(step)]), float(self._sqrt_1mac_host[int(step)]))
            img, pred_x0 = self._step(img, rest_image, rest_mask, int(step), index, unconditional_guidance_scale,
                                      temperature, blend)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(keep(img))
                intermediates["pred_x0"].append(pred_x0)
        return img, intermediates

    def _step(self, x, rest_image, rest_mask, step, index, scale, temperature, blend):
        ops.assemble_input(x, rest_image, rest_mask, self._x_in, cfg=self._cfg, blend=blend)
        eps = self._eval_model(step)
        c = StepCoefficients(self.ddim_alphas, self.ddim_alphas_prev, self.ddim_sqrt_one_minus_alphas, self.ddim_sigmas,
                             index)
        noise = torch.randn_like(x) if c.sigma != 0.0 else None  # sigma == 0: the reference's noise term is exactly 0
        return ops.sampler_update(eps, x, cfg=self._cfg, scale=float(scale), coefs=(1.0,),
                                  sqrt_one_minus_at=c.sqrt_one_minus_at, sqrt_at=c.sqrt_at, sqrt_a_prev=c.sqrt_a_prev,
                                  dir_coef=c.dir_coef, sigma_temp=c.sigma * float(temperature), noise=noise)

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, **kwargs):
        """ddim.py:165-213, single step (stateless API kept for callers that drive the loop themselves)."""
        self._check_unsupported(quantize_denoised, score_corrector, noise_dropout)
        if use_original_steps or repeat_noise:
            raise NotImplementedError("mobi_b200.DDIMSampler.p_sample_ddim: use_original_steps/repeat_noise")
        rest_image, rest_mask = self._rest_from_kwargs(kwargs)
        x = x.float().contiguous()
        cfg = not (unconditional_conditioning is None or unconditional_guidance_scale == 1.)
        b = x.shape[0]
        rest_c = 5 if rest_mask is not None else rest_image.shape[1]
        x_in = torch.empty(((2 if cfg else 1) * b, 4 + rest_c) + tuple(x.shape[2:]), device=x.device)
        ops.assemble_input(x, rest_image, rest_mask, x_in, cfg=cfg)
        t_in = torch.cat([t] * 2) if cfg else t
        c_in = torch.cat([unconditional_conditioning, c]) if cfg else c
        self._cfg = cfg
        eps = self._apply_model(x_in, t_in, c_in).float().contiguous()
        k = StepCoefficients(self.ddim_alphas, self.ddim_alphas_prev, self.ddim_sqrt_one_minus_alphas, self.ddim_sigmas,
                             index)
        noise = torch.randn_like(x) if k.sigma != 0.0 else None
        return ops.sampler_update(eps, x, cfg=cfg, scale=float(unconditional_guidance_scale), coefs=(1.0,),
                                  sqrt_one_minus_at=k.sqrt_one_minus_at, sqrt_at=k.sqrt_at, sqrt_a_prev=k.sqrt_a_prev,
                                  dir_coef=k.dir_coef, sigma_temp=k.sigma * float(temperature), noise=noise)

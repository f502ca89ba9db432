"""Drop-in for ldm/models/diffusion/plms.py (PLMSSampler): the sampler every shipped MObI script uses
(`--plms`, scripts/realism_test_bench.sh:74).  Same signatures as the reference; the Adams-Bashforth eps
combination (plms.py:221-235) is fused into the sampler-update kernel."""
import numpy as np
import torch

from . import ops
from .sampling import SamplerBase, StepCoefficients

# plms.py:227-235: weights on (e_t, old[-1], old[-2], old[-3]) by number of stored eps
_AB = {1: (3 / 2, -1 / 2), 2: (23 / 12, -16 / 12, 5 / 12), 3: (55 / 24, -59 / 24, 37 / 24, -9 / 24)}


class PLMSSampler(SamplerBase):
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        if ddim_eta != 0:
            raise ValueError("ddim_eta must be 0 for PLMS")  # plms.py:25-26
        super().make_schedule(ddim_num_steps, ddim_discretize, ddim_eta, verbose)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, **kwargs):
        """plms.py:57-113."""
        if conditioning is not None:
            cbs = conditioning[list(conditioning.keys())[0]].shape[0] if isinstance(conditioning, dict) \
                else conditioning.shape[0]
            if cbs != batch_size:
                print(f"Warning: Got {cbs} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        size = (batch_size, C, H, W)
        if verbose:
            print(f"Data shape for PLMS sampling is {size}")
        return self.plms_sampling(conditioning, size, callback=callback, img_callback=img_callback,
                                  quantize_denoised=quantize_x0, mask=mask, x0=x0, ddim_use_original_steps=False,
                                  noise_dropout=noise_dropout, temperature=temperature,
                                  score_corrector=score_corrector, corrector_kwargs=corrector_kwargs, x_T=x_T,
                                  log_every_t=log_every_t, unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning, **kwargs)

    @torch.no_grad()
    def plms_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100, temperature=1.,
                      noise_dropout=0., score_corrector=None, corrector_kwargs=None, unconditional_guidance_scale=1.,
                      unconditional_conditioning=None, **kwargs):
        """plms.py:115-171 with p_sample_plms (173-239) inlined into fused kernels."""
        self._check_unsupported(quantize_denoised, score_corrector, noise_dropout)
        if ddim_use_original_steps:
            raise NotImplementedError("mobi_b200.PLMSSampler: ddim_use_original_steps is not used by MObI")
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
        rest_image, rest_mask = self._rest_from_kwargs(kwargs)
        if rest_mask is None:
            raise Exception("PLMS needs inpaint_image and inpaint_mask (plms.py:218)")
        self._setup_eval(b, (9,) + tuple(shape[2:]), cond, unconditional_conditioning, unconditional_guidance_scale)
        scale = float(unconditional_guidance_scale)
        if mask is not None:
            assert x0 is not None
            mask = mask.to(device).float().contiguous()
            x0 = x0.to(device).float().contiguous()
        old_eps = []
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            step = int(step)
            step_next = int(time_range[min(i + 1, len(time_range) - 1)])
            blend = None
            if mask is not None:  # plms.py:147-150
                noise = torch.randn_like(x0)
                blend = (mask, x0, noise, float(self._sqrt_ac_host[step]), float(self._sqrt_1mac_host[step]))
            c = StepCoefficients(self.ddim_alphas, self.ddim_alphas_prev, self.ddim_sqrt_one_minus_alphas,
                                 self.ddim_sigmas, index)
            kw = dict(cfg=self._cfg, scale=scale, sqrt_one_minus_at=c.sqrt_one_minus_at, sqrt_at=c.sqrt_at,
                      sqrt_a_prev=c.sqrt_a_prev, dir_coef=c.dir_coef)
            ops.assemble_input(img, rest_image, rest_mask, self._x_in, cfg=self._cfg, blend=blend)
            eps = self._eval_model(step)
            e_t = torch.empty_like(img)
            if len(old_eps) == 0:
                # pseudo improved Euler (plms.py:221-226): a provisional step, a second UNet call at t_next
                x_tmp, _ = ops.sampler_update(eps, img, coefs=(1.0,), e_out=e_t, **kw)
                ops.assemble_input(x_tmp, rest_image, rest_mask, self._x_in, cfg=self._cfg)
                eps_next = self._eval_model(step_next)
                img_new, pred_x0 = ops.sampler_update(eps_next, img, coefs=(0.5, 0.5), old=[e_t], **kw)
            else:
                n = min(len(old_eps), 3)
                img_new, pred_x0 = ops.sampler_update(eps, img, coefs=_AB[n], old=list(reversed(old_eps[-n:])),
                                                      e_out=e_t, **kw)
            img = img_new
            old_eps.append(e_t)
            if len(old_eps) >= 4:
                old_eps.pop(0)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(keep(img))
                intermediates["pred_x0"].append(pred_x0)
        return img, intermediates

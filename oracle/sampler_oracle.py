"""TEST INFRASTRUCTURE ONLY (oracle) — fp32 restatement of the reference samplers and schedules.

Restates ldm/models/diffusion/ddim.py (DDIMSampler), plms.py (PLMSSampler), the schedule builders in
ldm/modules/diffusionmodules/util.py:21-74 and ldm/models/diffusion/ddpm.py:127-179, 284-287, with the UNet
passed in as a callable eps = apply_model(x[R,9,h,w], t[R], cond[R,n,ctx]).  Deterministic: eta must be 0 and
x_T must be given (the reference draws device RNG otherwise, SURVEY.md §8c "RNG discipline").
"""
import numpy as np
import torch


def make_beta_schedule(n_timestep=1000, linear_start=0.00085, linear_end=0.0120):
    """util.py:21-26 ("linear" = linear in sqrt(beta)), float64."""
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()


def register_schedule(timesteps=1000, linear_start=0.00085, linear_end=0.0120):
    """ddpm.py:127-150: float64 numpy -> float32 buffers."""
    betas = make_beta_schedule(timesteps, linear_start, linear_end)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return dict(betas=f32(betas), alphas_cumprod=f32(ac), alphas_cumprod_prev=f32(ac_prev),
                sqrt_alphas_cumprod=f32(np.sqrt(ac)), sqrt_one_minus_alphas_cumprod=f32(np.sqrt(1.0 - ac)))


def make_ddim_timesteps(num_ddim, num_ddpm=1000):
    """util.py:46-60, 'uniform'."""
    c = num_ddpm // num_ddim
    return np.asarray(list(range(0, num_ddpm, c))) + 1


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta):
    """util.py:63-74. alphacums: float32 torch tensor (CPU)."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


def q_sample(sched, x_start, t, noise):
    """ddpm.py:284-287."""
    a = sched["sqrt_alphas_cumprod"].to(x_start.device)[t].reshape(-1, 1, 1, 1)
    b = sched["sqrt_one_minus_alphas_cumprod"].to(x_start.device)[t].reshape(-1, 1, 1, 1)
    return a * x_start + b * noise


def _model_out(apply_model, x9, t, c, uc, scale):
    """ddim.py:177-184 / plms.py:177-185: CFG with the doubled batch ordered [uncond ; cond]."""
    if uc is None or scale == 1.0:
        return apply_model(x9, t, c)
    e_u, e_c = apply_model(torch.cat([x9] * 2), torch.cat([t] * 2), torch.cat([uc, c])).chunk(2)
    return e_u + scale * (e_c - e_u)


def _coefs(sched, S, eta):
    ts = make_ddim_timesteps(S, sched["alphas_cumprod"].shape[0])
    sigmas, alphas, alphas_prev = make_ddim_sampling_parameters(sched["alphas_cumprod"].cpu(), ts, eta)
    return ts, sigmas, alphas, alphas_prev, np.sqrt(1.0 - alphas)


def _x_prev(x4, e, idx, alphas, alphas_prev, sqrt_1m, sigmas):
    """ddim.py:195-212 with sigma = 0 noise dropped (eta = 0 => sigma_t * noise == 0)."""
    dev = x4.device
    full = lambda v: torch.full((x4.shape[0], 1, 1, 1), float(v), device=dev)
    a_t, a_prev, sigma_t, s1m = full(alphas[idx]), full(alphas_prev[idx]), full(sigmas[idx]), full(sqrt_1m[idx])
    pred_x0 = (x4 - s1m * e) / a_t.sqrt()
    dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e
    return a_prev.sqrt() * pred_x0 + dir_xt, pred_x0


def ddim_sample(apply_model, sched, S, x_T, cond, uc, scale, inpaint_image, inpaint_mask, mask=None, x0=None,
                blend_noise=None, steps_to_run=None):
    """DDIMSampler.sample / ddim_sampling / p_sample_ddim (ddim.py:56-213) with test_model_kwargs, eta=0.
    mask/x0: optional known-latent blend (ddim.py:145-148); blend_noise[i] replaces q_sample's randn at loop
    iteration i.  Returns (samples, list of per-step (x_prev, pred_x0, e_t))."""
    ts, sigmas, alphas, alphas_prev, sqrt_1m = _coefs(sched, S, 0.0)
    img = x_T
    trace = []
    time_range = np.flip(ts)
    total = ts.shape[0]
    for i, step in enumerate(time_range):
        if steps_to_run is not None and i >= steps_to_run:
            break
        index = total - i - 1
        t = torch.full((img.shape[0],), int(step), device=img.device, dtype=torch.long)
        if mask is not None:
            img = q_sample(sched, x0, t, blend_noise[i]) * mask + (1.0 - mask) * img
        x9 = torch.cat([img, inpaint_image, inpaint_mask], dim=1)
        e = _model_out(apply_model, x9, t, cond, uc, scale)
        img, pred = _x_prev(img, e, index, alphas, alphas_prev, sqrt_1m, sigmas)
        trace.append((img, pred, e))
    return img, trace


def plms_sample(apply_model, sched, S, x_T, cond, uc, scale, inpaint_image, inpaint_mask, steps_to_run=None):
    """PLMSSampler.plms_sampling / p_sample_plms (plms.py:115-239), eta=0."""
    ts, sigmas, alphas, alphas_prev, sqrt_1m = _coefs(sched, S, 0.0)
    img = x_T
    old_eps = []
    trace = []
    time_range = np.flip(ts)
    total = ts.shape[0]
    for i, step in enumerate(time_range):
        if steps_to_run is not None and i >= steps_to_run:
            break
        index = total - i - 1
        b = img.shape[0]
        t = torch.full((b,), int(step), device=img.device, dtype=torch.long)
        t_next = torch.full((b,), int(time_range[min(i + 1, len(time_range) - 1)]), device=img.device, dtype=torch.long)
        e_t = _model_out(apply_model, torch.cat([img, inpaint_image, inpaint_mask], 1), t, cond, uc, scale)
        if len(old_eps) == 0:  # pseudo improved Euler, plms.py:221-226
            x_prev, _ = _x_prev(img, e_t, index, alphas, alphas_prev, sqrt_1m, sigmas)
            e_next = _model_out(apply_model, torch.cat([x_prev, inpaint_image, inpaint_mask], 1), t_next, cond, uc,
                                scale)
            e_prime = (e_t + e_next) / 2
        elif len(old_eps) == 1:
            e_prime = (3 * e_t - old_eps[-1]) / 2
        elif len(old_eps) == 2:
            e_prime = (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
        else:
            e_prime = (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24
        img, pred = _x_prev(img, e_prime, index, alphas, alphas_prev, sqrt_1m, sigmas)
        old_eps.append(e_t)
        if len(old_eps) >= 4:
            old_eps.pop(0)
        trace.append((img, pred, e_t))
    return img, trace

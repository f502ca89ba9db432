"""TEST INFRASTRUCTURE ONLY (oracle) — fp32 restatement of the reference's training step for the UNet.

Restates LatentDiffusion.p_losses (ldm/models/diffusion/ddpm.py:1177-1217: q_sample on the first 4 channels, UNet,
mean squared error against the noise; logvar = 0, l_simple_weight = 1, original_elbo_weight = 0 as in every shipped
config) and the parameter selection of DiffusionWrapper.__init__ (ddpm.py:1686-1698), with torch.autograd over the
functional UNet of oracle/unet_oracle.py standing in for `loss.backward()`, and torch.optim.AdamW semantics
(ddpm.py:1655) restated for one step.  Pinned by tests/golden/train_tiny.npz, generated from the unmodified reference.
"""
import torch

from . import sampler_oracle as so
from . import unet_oracle as uo

TRAINABLE_KEYS = ("cond_adapter", "lidar", "cross_modal")  # ddpm.py:1686-1698


def is_trainable(name):
    return any(k in name for k in TRAINABLE_KEYS)


def p_losses(sd, cfg, sched, x_start, t, noise, cond):
    """ddpm.py:1177-1217 with first_stage_key == 'inpaint', parameterization 'eps', loss_type 'l2'."""
    x_noisy = torch.cat([so.q_sample(sched, x_start[:, :4], t, noise), x_start[:, 4:]], dim=1)
    eps = uo.unet_forward(sd, cfg, x_noisy, t, cond)
    loss_simple = ((eps - noise) ** 2).mean(dim=[1, 2, 3])
    return loss_simple.mean(), eps


def loss_and_grads(sd, cfg, sched, x_start, t, noise, cond, want_d_cond=False):
    """Returns (loss, {name: grad}) for the trainable tensors of the UNet state dict (and, with want_d_cond, the gradient
    w.r.t. the conditioning tokens under the key "__d_cond__": it is what reaches the trainable bbox_embedder,
    ddpm.py:580-586)."""
    leaves = {}
    work = {}
    for k, v in sd.items():
        if is_trainable(k):
            leaves[k] = v.detach().clone().requires_grad_(True)
            work[k] = leaves[k]
        else:
            work[k] = v.detach()
    cond_leaf = cond.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        loss, _ = p_losses(work, cfg, sched, x_start, t, noise, cond_leaf)
        grads = torch.autograd.grad(loss, list(leaves.values()) + [cond_leaf], allow_unused=True)
    out = {}
    for (k, v), g in zip(leaves.items(), grads[:-1]):
        out[k] = torch.zeros_like(v) if g is None else g
    if want_d_cond:
        out["__d_cond__"] = grads[-1]
    return loss.detach(), out


def adamw_step(p, g, m, v, step, lr=8e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
    """torch.optim.AdamW (decoupled weight decay, bias correction), one tensor."""
    p = p * (1.0 - lr * weight_decay)
    m = betas[0] * m + (1.0 - betas[0]) * g
    v = betas[1] * v + (1.0 - betas[1]) * g * g
    bc1, bc2 = 1.0 - betas[0] ** step, 1.0 - betas[1] ** step
    denom = v.sqrt() / bc2 ** 0.5 + eps
    return p - (lr / bc1) * m / denom, m, v


def synth_train_inputs(n_joint, h, context_dim, seed=5, device="cpu"):
    """x_start = [latent | inpaint_image | mask] (9 channels), t, noise, cond for n_joint joint samples."""
    import numpy as np
    inp = uo.synth_inputs(n_joint, h, context_dim=context_dim, seed=seed, device=device)
    rng = np.random.default_rng(seed + 100)
    r = 2 * n_joint
    x_start = torch.cat([inp["x_T"], inp["inpaint_image"], inp["inpaint_mask"]], 1)
    noise = torch.from_numpy(rng.standard_normal((r, 4, h, h), dtype=np.float32)).to(device)
    t = torch.from_numpy(rng.integers(0, 1000, size=(r,))).long().to(device)
    return dict(x_start=x_start, t=t, noise=noise, cond=inp["cond"])

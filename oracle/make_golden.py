"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shims.py) on deterministic synthetic weights and inputs.

Run here (the build container) with:  python -m oracle.make_golden
The GPU box has no /root/reference; tests there read only the committed fixtures.

What is pinned (all with the tiny topology-complete configs from oracle/*_oracle.py so fixtures stay small):
  unet_tiny.npz        UNetModel.forward(x, t, context) through LatentDiffusion.apply_model
  ddim_tiny.npz        DDIMSampler.sample, 4 steps of S=10, CFG 3.0, test_model_kwargs path (+ mask/x0 blend run)
  plms_tiny.npz        PLMSSampler.sample, 4 steps of S=10, CFG 3.0
  vae_tiny.npz         AutoencoderKL.decode / encode moments for the camera and the lidar-adapter autoencoder
  train_tiny.npz       LatentDiffusion.p_losses loss + loss.backward() gradients of the trainable (adapter) parameters,
                       and the parameters after one torch.optim.AdamW step
  get_input_tiny.npz   LatentDiffusion.get_input (encode_all_stages: 4 VAE encodes + posterior samples + nearest mask resize +
                       9-channel concat; lidar crop/pad to the image latent; bbox re-normalisation; cam/lidar interleave)
                       with the tiny VAEs, a 32x32 camera image and a 48x48 range image (so crop AND negative pad run)
  cond_full.npz        conditioning tokens [B, 2, 768] from the reference's mapper Transformer + final_ln + proj_out and
                       BBoxEmbedder at FULL size (weights are regenerated from the seed by tensor name, only inputs and
                       outputs are stored)
  schedule.npz         register_schedule + DDIM parameters for S=50 on the real (1000-step) schedule
  shapes_512.json      state_dict key -> shape of the real mobi_nusc_512 UNet (1,118 tensors) and both VAEs
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import cond_oracle, ref_shims, sampler_oracle, train_oracle, unet_oracle, vae_oracle  # noqa: E402


def build_reference_ldm(unet_cfg, cam_dd=None, lid_dd=None):
    from ldm.models.diffusion.ddpm import LatentDiffusion

    def ae(dd):
        return None if dd is None else dict(target="ldm.models.autoencoder.AutoencoderKL",
                                            params=dict(embed_dim=4, ddconfig=dd,
                                                        lossconfig=dict(target="torch.nn.Identity")))

    model = LatentDiffusion(
        cond_stage_config="__is_unconditional__", first_stage_config=ae(cam_dd), lidar_stage_config=ae(lid_dd),
        unet_config=dict(target="ldm.modules.diffusionmodules.openaimodel.UNetModel", params=unet_cfg),
        linear_start=0.00085, linear_end=0.0120, num_timesteps_cond=1, log_every_t=200, timesteps=1000,
        first_stage_key="inpaint", cond_stage_key=["ref_image", "ref_bbox"], image_size=unet_cfg["image_size"],
        channels=4, cond_stage_trainable=False, conditioning_key="crossattn", monitor=None, u_cond_percent=0.2,
        scale_factor=0.18215, lidar_scale_factor=0.18215, use_ema=False, use_camera=True, use_lidar=True)
    # "__is_unconditional__" resets conditioning_key to None in __init__ (ddpm.py:469-470); restore crossattn,
    # the value every shipped config uses (configs/mobi_nusc_512.yaml:42)
    model.model.conditioning_key = "crossattn"
    return model.eval()


def load_synth(module, shapes_fn_result, seed):
    ref_sd = module.state_dict()
    mine = shapes_fn_result
    assert set(ref_sd.keys()) == set(mine.keys()), (
        sorted(set(ref_sd) - set(mine))[:5], sorted(set(mine) - set(ref_sd))[:5])
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(mine[k]), (k, tuple(v.shape), mine[k])
    sd = unet_oracle.synth_state_dict(mine, seed=seed)
    module.load_state_dict(sd, strict=True)
    return sd


def main():
    ref_shims.install()
    torch.set_grad_enabled(False)
    os.makedirs(GOLDEN, exist_ok=True)
    from ldm.models.diffusion.ddim import DDIMSampler
    from ldm.models.diffusion.plms import PLMSSampler

    # ---- tiny joint model
    ucfg = unet_oracle.tiny_unet_config()
    cam_dd, lid_dd = vae_oracle.tiny_ddconfig(False), vae_oracle.tiny_ddconfig(True)
    ldm = build_reference_ldm(ucfg, cam_dd, lid_dd)
    load_synth(ldm.model.diffusion_model, unet_oracle.state_dict_shapes(ucfg), seed=0)
    load_synth(ldm.first_stage_model, vae_oracle.state_dict_shapes(cam_dd), seed=10)
    load_synth(ldm.lidar_stage_model, vae_oracle.state_dict_shapes(lid_dd), seed=20)

    inp = unet_oracle.synth_inputs(2, 16, context_dim=ucfg["context_dim"], seed=1)
    x9 = torch.cat([inp["x_T"], inp["inpaint_image"], inp["inpaint_mask"]], 1)
    t = torch.tensor([981, 981, 21, 21], dtype=torch.long)
    eps = ldm.apply_model(x9, t, inp["cond"])
    np.savez(os.path.join(GOLDEN, "unet_tiny.npz"), x=x9.numpy(), t=t.numpy(), context=inp["cond"].numpy(),
             eps=eps.numpy())
    print("unet_tiny eps std %.4f absmax %.4f" % (eps.std(), eps.abs().max()))

    # ---- samplers (S=10 schedule, first 4 steps via timesteps subset is awkward: run full S=4 instead)
    S, scale = 4, 3.0
    samples, inter = DDIMSampler(ldm).sample(
        S=S, conditioning=inp["cond"], batch_size=4, shape=[4, 16, 16], verbose=False,
        unconditional_guidance_scale=scale, unconditional_conditioning=inp["uc"], eta=0.0, x_T=inp["x_T"],
        log_every_t=1,
        test_model_kwargs=dict(inpaint_image=inp["inpaint_image"], inpaint_mask=inp["inpaint_mask"]))
    out = dict(samples=samples.numpy(), pred_x0_last=inter["pred_x0"][-1].numpy(),
               x_inter=np.stack([a.numpy() for a in inter["x_inter"]]))
    # known-latent blend path (ddim.py:145-148): q_sample draws randn_like -> seed torch and record the draws
    rng = np.random.default_rng(7)
    bmask = torch.from_numpy((rng.random((4, 1, 16, 16)) > 0.5).astype(np.float32))
    bx0 = torch.from_numpy(rng.standard_normal((4, 4, 16, 16), dtype=np.float32))
    torch.manual_seed(123)
    noises = [torch.randn(4, 4, 16, 16) for _ in range(2 * S)]
    it = iter(noises)
    orig_q = ldm.q_sample
    ldm.q_sample = lambda x_start, tt, noise=None: orig_q(x_start, tt, noise=next(it))
    # eta=0 still draws noise_like every step (ddim.py:209) but multiplies it by sigma=0
    samples_b, _ = DDIMSampler(ldm).sample(
        S=S, conditioning=inp["cond"], batch_size=4, shape=[4, 16, 16], verbose=False,
        unconditional_guidance_scale=scale, unconditional_conditioning=inp["uc"], eta=0.0, x_T=inp["x_T"],
        mask=bmask, x0=bx0,
        test_model_kwargs=dict(inpaint_image=inp["inpaint_image"], inpaint_mask=inp["inpaint_mask"]))
    ldm.q_sample = orig_q
    out.update(blend_mask=bmask.numpy(), blend_x0=bx0.numpy(), blend_noise=np.stack([n.numpy() for n in noises[:S]]),
               samples_blend=samples_b.numpy())
    np.savez(os.path.join(GOLDEN, "ddim_tiny.npz"), S=S, scale=scale, **{k: v.numpy() for k, v in inp.items()}, **out)
    print("ddim_tiny samples std %.4f" % samples.std())

    samples_p, _ = PLMSSampler(ldm).sample(
        S=S, conditioning=inp["cond"], batch_size=4, shape=[4, 16, 16], verbose=False,
        unconditional_guidance_scale=scale, unconditional_conditioning=inp["uc"], eta=0.0, x_T=inp["x_T"],
        inpaint_image=inp["inpaint_image"], inpaint_mask=inp["inpaint_mask"])
    np.savez(os.path.join(GOLDEN, "plms_tiny.npz"), S=S, scale=scale, samples=samples_p.numpy())
    print("plms_tiny samples std %.4f" % samples_p.std())

    # ---- VAEs: decode_sample + decode_first_stage for both modalities, encode moments
    h_cam, h_lid = ldm.decode_sample(samples, samples[1::2].clone())
    img = ldm.decode_first_stage(h_cam)
    rng_img = ldm.decode_first_stage(h_lid, module_name="lidar_stage_model")
    gx = np.random.default_rng(3)
    cam_in = torch.from_numpy(gx.standard_normal((1, 3, 64, 64), dtype=np.float32))
    lid_in = torch.from_numpy(gx.standard_normal((1, 2, 64, 64), dtype=np.float32))
    cam_m = ldm.first_stage_model.encode(cam_in).parameters
    lid_m = ldm.lidar_stage_model.encode(lid_in).parameters
    np.savez(os.path.join(GOLDEN, "vae_tiny.npz"), z=samples.numpy(), image=img.numpy(), range=rng_img.numpy(),
             cam_in=cam_in.numpy(), lid_in=lid_in.numpy(), cam_moments=cam_m.numpy(), lid_moments=lid_m.numpy())
    print("vae_tiny image std %.4f range std %.4f" % (img.std(), rng_img.std()))


    # ---- training step: p_losses + backward + one AdamW step through the reference modules
    tr = train_oracle.synth_train_inputs(2, 16, ucfg["context_dim"], seed=5)
    with torch.enable_grad():
        ldm.train()
        for n_, p_ in ldm.model.diffusion_model.named_parameters():
            assert p_.requires_grad == train_oracle.is_trainable(n_), n_   # DiffusionWrapper.__init__, ddpm.py:1686-1698
        params = [p_ for p_ in ldm.model.diffusion_model.parameters() if p_.requires_grad]
        opt = torch.optim.AdamW(params, lr=8e-5)                           # ddpm.py:1655
        cond_leaf = tr["cond"].clone().requires_grad_(True)        # d loss / d conditioning tokens: what trains bbox_embedder
        loss, _ = ldm.p_losses(tr["x_start"], cond_leaf, tr["t"], noise=tr["noise"])
        loss.backward()
        grads = {n_: p_.grad.detach().clone() for n_, p_ in ldm.model.diffusion_model.named_parameters()
                 if p_.requires_grad}
        opt.step()
        after = {n_: p_.detach().clone() for n_, p_ in ldm.model.diffusion_model.named_parameters() if p_.requires_grad}
        ldm.eval()
    keep = ("input_blocks.1.1.", "middle_block.1.", "output_blocks.3.1.")
    out = dict(x_start=tr["x_start"].numpy(), t=tr["t"].numpy(), noise=tr["noise"].numpy(), cond=tr["cond"].numpy(),
               loss=np.float32(loss.item()), d_cond=cond_leaf.grad.numpy(), names=np.array(sorted(grads)),
               grad_l2=np.array([grads[k].norm().item() for k in sorted(grads)], dtype=np.float64),
               grad_sum=np.array([grads[k].double().sum().item() for k in sorted(grads)], dtype=np.float64))
    for k in sorted(grads):
        if k.startswith(keep):
            out["g:" + k] = grads[k].numpy()
            if "cross_modal_attn_camera.to_q" in k or "cond_adapter_norm" in k:
                out["p1:" + k] = after[k].numpy()
    np.savez_compressed(os.path.join(GOLDEN, "train_tiny.npz"), **out)
    print("train_tiny loss %.6f, %d trainable tensors, %d stored in full" % (
        loss.item(), len(grads), sum(1 for k in out if k.startswith("g:"))))
    for p_ in ldm.model.diffusion_model.parameters():
        p_.grad = None
    load_synth(ldm.model.diffusion_model, unet_oracle.state_dict_shapes(ucfg), seed=0)   # undo the optimizer step
    torch.set_grad_enabled(False)



    # ---- get_input assembly through the unmodified LatentDiffusion.get_input
    gi = np.random.default_rng(21)
    f32 = lambda *s_: torch.from_numpy(gi.standard_normal(s_, dtype=np.float32))
    nb = 2
    mk = lambda h_: torch.from_numpy((gi.random((nb, 1, h_, h_)) > 0.3).astype(np.float32))
    batch = dict(
        image=dict(GT=f32(nb, 3, 32, 32), inpaint_image=f32(nb, 3, 32, 32), inpaint_mask=mk(32),
                   cond=dict(ref_image=f32(nb, 3, 8, 8), ref_bbox=torch.from_numpy(gi.uniform(0, 1, (nb, 8, 3)).astype(np.float32)))),
        lidar=dict(range_data=f32(nb, 2, 48, 48), range_data_inpaint=f32(nb, 2, 48, 48), range_mask=mk(48),
                   cond=dict(ref_image=f32(nb, 3, 8, 8), ref_bbox=torch.from_numpy(gi.uniform(0, 1, (nb, 8, 3)).astype(np.float32)))))
    bbox_lidar_in = batch["lidar"]["cond"]["ref_bbox"].clone()      # get_input re-normalises it IN PLACE (ddpm.py:815-816)
    ldm.process_conditioning = lambda cond, force_c_encode=False: (cond, None)   # conditioning encoders: separate row
    torch.manual_seed(77)
    gout = ldm.get_input(batch, "inpaint")
    torch.manual_seed(77)                                             # DiagonalGaussianDistribution.sample draws, in call order
    noises = [torch.randn(nb, 4, 16, 16), torch.randn(nb, 4, 16, 16), torch.randn(nb, 4, 24, 24), torch.randn(nb, 4, 24, 24)]
    np.savez(os.path.join(GOLDEN, "get_input_tiny.npz"),
             image_gt=batch["image"]["GT"].numpy(), image_inpaint=batch["image"]["inpaint_image"].numpy(),
             image_mask=batch["image"]["inpaint_mask"].numpy(), range_gt=batch["lidar"]["range_data"].numpy(),
             range_inpaint=batch["lidar"]["range_data_inpaint"].numpy(), range_mask=batch["lidar"]["range_mask"].numpy(),
             bbox_camera=batch["image"]["cond"]["ref_bbox"].numpy(), bbox_lidar_in=bbox_lidar_in.numpy(),
             noise_cam_gt=noises[0].numpy(), noise_cam_inpaint=noises[1].numpy(), noise_lid_gt=noises[2].numpy(),
             noise_lid_inpaint=noises[3].numpy(), z=gout["z"].numpy(), z_lidar=gout["z_lidar"].numpy(),
             bbox_out=gout["cond"]["ref_bbox"].numpy())
    print("get_input_tiny z", tuple(gout["z"].shape), "z_lidar", tuple(gout["z_lidar"].shape), "std %.4f" % gout["z"].std())

    # ---- conditioning encoders after the CLIP tower, full size, reference modules
    from ldm.modules.encoders.modules import BBoxEmbedder
    from ldm.modules.encoders.xf import LayerNorm as XfLayerNorm
    from ldm.modules.encoders.xf import Transformer as XfTransformer
    csd = unet_oracle.synth_state_dict(cond_oracle.shapes(), seed=30)
    mapper, final_ln, bbox_emb = XfTransformer(1, 1024, 5, 1), XfLayerNorm(1024), BBoxEmbedder()   # modules.py:150-157
    proj_out = torch.nn.Linear(1024, 768)                                                          # ddpm.py:492
    pre = "cond_stage_model."
    mapper.load_state_dict({k[len(pre + "mapper."):]: v for k, v in csd.items() if k.startswith(pre + "mapper.")})
    final_ln.load_state_dict({k[len(pre + "final_ln."):]: v for k, v in csd.items() if k.startswith(pre + "final_ln.")})
    bbox_emb.load_state_dict({k[len(pre + "bbox_embedder."):]: v for k, v in csd.items()
                              if k.startswith(pre + "bbox_embedder.")})
    proj_out.load_state_dict({"weight": csd["proj_out.weight"], "bias": csd["proj_out.bias"]})
    gc = np.random.default_rng(11)
    pooled = torch.from_numpy(gc.standard_normal((6, 1024), dtype=np.float32))     # CLIP pooler_output stand-in
    bbox = torch.from_numpy(gc.uniform(-1.0, 1.0, (6, 8, 3)).astype(np.float32))   # normalised box corners
    z = final_ln(mapper(pooled.unsqueeze(1)))                                      # modules.py:165-169
    tok_img = proj_out(z)                                                          # ddpm.py:622
    tok_box = bbox_emb(bbox)                                                       # modules.py:203-209
    cond_ref = torch.cat([tok_img, tok_box], dim=1)                                # ddpm.py:623-630
    np.savez(os.path.join(GOLDEN, "cond_full.npz"), pooled=pooled.numpy(), bbox=bbox.numpy(), cond=cond_ref.numpy())
    print("cond_full cond std %.4f" % cond_ref.std())

    # ---- schedule buffers on the real schedule, S=50
    smp = DDIMSampler(ldm)
    smp.make_schedule(ddim_num_steps=50, ddim_eta=0.0, verbose=False)
    np.savez(os.path.join(GOLDEN, "schedule.npz"), betas=ldm.betas.numpy(), alphas_cumprod=ldm.alphas_cumprod.numpy(),
             alphas_cumprod_prev=ldm.alphas_cumprod_prev.numpy(),
             sqrt_alphas_cumprod=ldm.sqrt_alphas_cumprod.numpy(),
             sqrt_one_minus_alphas_cumprod=ldm.sqrt_one_minus_alphas_cumprod.numpy(),
             ddim_timesteps=np.asarray(smp.ddim_timesteps), ddim_alphas=np.asarray(smp.ddim_alphas),
             ddim_alphas_prev=np.asarray(smp.ddim_alphas_prev), ddim_sigmas=np.asarray(smp.ddim_sigmas),
             ddim_sqrt_one_minus_alphas=np.asarray(smp.ddim_sqrt_one_minus_alphas))

    # ---- full-size key/shape inventory (meta device: no memory)
    from ldm.models.autoencoder import AutoencoderKL
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    shapes = {}
    with torch.device("meta"):
        full = UNetModel(**unet_oracle.default_unet_config())
        pbe = UNetModel(**unet_oracle.default_unet_config(use_lidar=False))
        cam = AutoencoderKL(ddconfig=vae_oracle.default_ddconfig(False), lossconfig=dict(target="torch.nn.Identity"),
                            embed_dim=4)
        lid = AutoencoderKL(ddconfig=vae_oracle.default_ddconfig(True), lossconfig=dict(target="torch.nn.Identity"),
                            embed_dim=4)
    for name, m in (("unet_512", full), ("unet_pbe", pbe), ("vae_camera", cam), ("vae_lidar", lid)):
        shapes[name] = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(GOLDEN, "shapes_512.json"), "w") as fh:
        json.dump(shapes, fh, indent=0, sort_keys=True)
    print("unet_512 tensors:", len(shapes["unet_512"]), "params:",
          sum(int(np.prod(s)) for s in shapes["unet_512"].values()))


if __name__ == "__main__":
    main()

"""The whole caller-side path of the reference's test bench (scripts/inference_test_bench.py:403-464, 567-629) on the
drop-in package, at full size: the model is built from the reference's YAML schema (configs/mobi_nusc_512.yaml:
`target: ldm....` strings, `${...}` interpolation) by mobi_b200.config.build_model, filled from a reference-FORMAT
checkpoint file ({"state_dict": {...}} with the reference's key names, EMA / CLIP-tower keys it must ignore), and run
through mobi_b200.pipeline.inpaint_batch on a batch with the dataset's layout.  Every stage is compared with the oracle
on the same weights: get_input (4 VAE encodes at 512 px + assembly), conditioning tokens, the sampler run, both decodes,
and the range-view post-processing (bit-exact masks)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


def cosine(a, b):
    return torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()


def _reference_format_checkpoint(path):
    """Synthetic weights under the key names of the reference's LatentDiffusion checkpoint."""
    from oracle import cond_oracle as co
    from oracle import unet_oracle as uo
    from oracle import vae_oracle as vo
    ucfg = uo.default_unet_config(image_size=64)
    cam, lid = vo.default_ddconfig(False, resolution=512), vo.default_ddconfig(True, resolution=512)
    parts = dict(unet=uo.synth_state_dict(uo.state_dict_shapes(ucfg), seed=0),
                 cam=uo.synth_state_dict(vo.state_dict_shapes(cam), seed=31),
                 lid=uo.synth_state_dict(vo.state_dict_shapes(lid), seed=32),
                 cond=uo.synth_state_dict(co.shapes(), seed=33))
    g = torch.Generator().manual_seed(34)
    sd = {"model.diffusion_model." + k: v for k, v in parts["unet"].items()}
    sd.update({"first_stage_model." + k: v for k, v in parts["cam"].items()})
    sd.update({"lidar_stage_model." + k: v for k, v in parts["lid"].items()})
    sd.update(parts["cond"])                                       # cond_stage_model.* and proj_out.*
    sd["learnable_vector"] = torch.randn(1, 1, 768, generator=g)
    sd["bbox_uncond_vector"] = torch.randn(1, 1, 768, generator=g)
    # things a real checkpoint carries that the drop-in must skip: the EMA copy, the CLIP tower, lightning bookkeeping
    sd["model_ema.decay"] = torch.tensor(0.9999)
    sd["model_ema.num_updates"] = torch.tensor(1)
    sd["cond_stage_model.transformer.vision_model.embeddings.class_embedding"] = torch.zeros(1024)
    torch.save({"state_dict": sd, "epoch": 28, "global_step": 1}, path)
    parts["uc"] = torch.cat([sd["learnable_vector"], sd["bbox_uncond_vector"]], 1)
    parts["cfgs"] = (ucfg, cam, lid)
    return parts


def test_test_bench_pipeline_from_reference_yaml_and_checkpoint(tmp_path):
    from mobi_b200 import config, pipeline, synth
    from oracle import cond_oracle as co
    from oracle import input_oracle as io
    from oracle import range_oracle as ro
    from oracle import sampler_oracle as so
    from oracle import unet_oracle as uo
    from oracle import vae_oracle as vo
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", 0)
    ckpt = str(tmp_path / "mobi_synthetic.ckpt")
    parts = _reference_format_checkpoint(ckpt)
    ucfg, cam_cfg, lid_cfg = parts["cfgs"]
    cfg = config.load_config(os.path.join(ROOT, "configs", "mobi_nusc_512.yaml"))
    assert cfg["model"]["target"] == "ldm.models.diffusion.ddpm.LatentDiffusion"      # the reference's own target strings
    model = config.build_model(cfg, ckpt=ckpt, device=dev, clip_tower=synth.PooledFeatureTower(), meta_init=True)
    os.remove(ckpt)
    assert type(model.model.diffusion_model).__module__ == "mobi_b200.openaimodel" and model.image_size == 64
    N, S, scale = 2, 4, 5.0
    host = synth.synthetic_dataset_batch(N, px=512, seed=11, pin=True)
    batch = pipeline.batch_to_device(host, dev)
    gen = torch.Generator(device="cpu").manual_seed(12)
    nz = lambda: torch.randn(N, 4, 64, 64, generator=gen).to(dev)                                # noqa: E731
    noise = dict(camera=(nz(), nz()), lidar=(nz(), nz()))
    x_T = torch.randn(2 * N, 4, 64, 64, generator=gen).to(dev)
    bbox_cam_in, bbox_lid_in = batch["image"]["cond"]["ref_bbox"].clone(), batch["lidar"]["cond"]["ref_bbox"].clone()
    sampler = pipeline.make_sampler(model, plms=False)
    out = pipeline.inpaint_batch(model, sampler, batch, ddim_steps=S, scale=scale, start_code=x_T, noise=noise)
    torch.cuda.synchronize()

    cu = lambda d: {k: v.to(dev) for k, v in d.items()}                                          # noqa: E731
    usd, csd, lsd, cosd = cu(parts["unet"]), cu(parts["cam"]), cu(parts["lid"]), cu(parts["cond"])
    with torch.no_grad():
        # ---- get_input: 4 encodes at 512 px, posterior samples with the same noise, mask resize, interleave
        g = dict(image_gt=batch["image"]["GT"], image_inpaint=batch["image"]["inpaint_image"],
                 image_mask=batch["image"]["inpaint_mask"], range_gt=batch["lidar"]["range_data"],
                 range_inpaint=batch["lidar"]["range_data_inpaint"], range_mask=batch["lidar"]["range_mask"],
                 noise_cam_gt=noise["camera"][0], noise_cam_inpaint=noise["camera"][1], noise_lid_gt=noise["lidar"][0],
                 noise_lid_inpaint=noise["lidar"][1], bbox_camera=bbox_cam_in, bbox_lidar_in=bbox_lid_in)
        want_in = io.get_input(csd, cam_cfg, lsd, lid_cfg, g, 64)
        e_z = rel(out["z"][:, :8], want_in["z"][:, :8])
        assert torch.equal(out["z"][:, 8], want_in["z"][:, 8])
        # ---- conditioning tokens from the interleaved raw conditioning
        pooled = io.cat_interleave([batch["image"]["cond"]["ref_image"], batch["lidar"]["cond"]["ref_image"]])
        want_c = co.learned_conditioning(cosd, pooled, want_in["bbox"])
        e_c = rel(out["cond"], want_c)
        # ---- the sampler on the oracle's own inputs (end to end) and on the pipeline's inputs (sampler alone)
        apply_ref = lambda x, t, c: uo.unet_forward(usd, ucfg, x, t, c)                          # noqa: E731
        uc = parts["uc"].to(dev).repeat(2 * N, 1, 1)
        sched = so.register_schedule()
        ref_e2e, _ = so.ddim_sample(apply_ref, sched, S, x_T, want_c, uc, scale, want_in["z"][:, 4:8], want_in["z"][:, 8:9])
        ref_own, _ = so.ddim_sample(apply_ref, sched, S, x_T, out["cond"], uc, scale, out["z"][:, 4:8], out["z"][:, 8:9])
        e_s, e_s_e2e = rel(out["samples"], ref_own), rel(out["samples"], ref_e2e)
        # ---- decodes of the pipeline's own latents
        h_cam, h_lid = out["samples"][::2], out["samples"][1::2]
        from oracle.precision import vae_bf16_operand_floor
        want_img = vo.vae_decode(csd, cam_cfg, h_cam / 0.18215)
        want_rng = vo.vae_decode(lsd, lid_cfg, h_lid / 0.18215)
        e_img, e_rng = rel(out["image_decoded"], want_img), rel(out["range_sample"], want_rng)
        f_img = vae_bf16_operand_floor(csd, cam_cfg, h_cam / 0.18215, want_img)[0]
        f_rng = vae_bf16_operand_floor(lsd, lid_cfg, h_lid / 0.18215, want_rng)[0]
        assert torch.equal(out["image_sample"], out["image_decoded"].clamp(-1, 1))         # log_data's clamp, ddpm.py:1494
    print("pipeline vs oracle: get_input z %.3e | cond %.3e | sampler (own inputs) %.3e, end to end %.3e cosine %.6f | "
          "decode camera %.3e (bf16-operand floor %.3e) lidar %.3e (floor %.3e)"
          % (e_z, e_c, e_s, e_s_e2e, cosine(out["samples"], ref_e2e), e_img, f_img, e_rng, f_rng))
    assert e_z < 1e-2 and e_c < 1e-2
    assert e_s < 2e-2 and e_s_e2e < 3e-2 and cosine(out["samples"], ref_e2e) > 0.999
    assert e_img < 2e-2 and e_img < 1.5 * f_img and e_rng < 2e-2 and e_rng < 1.5 * f_rng
    assert out["image_sample"].shape == (N, 3, 512, 512) and out["range_sample"].shape == (N, 2, 512, 512)
    # ---- range-view post-processing of the pipeline's own decoded range image: bit-exact against the NumPy oracle
    lid = {k: v.cpu().numpy() for k, v in batch["lidar"].items() if torch.is_tensor(v)}
    dec = out["range_sample"].cpu().numpy()
    inp = dict(range_depth=dec[:, [0]], range_int=dec[:, [1]], range_depth_orig=lid["range_depth_orig"],
               range_int_orig=lid["range_int_orig"], range_pitch=lid["range_pitch"], range_yaw=lid["range_yaw"],
               range_instance_mask_orig=lid["range_instance_mask_orig"], crop_left=lid["range_shift_left"],
               width_crop=lid["width_crop"], min_depth_obj=lid["min_depth_obj"], max_depth_obj=lid["max_depth_obj"])
    want = ro.run_pipeline(inp, batch["bbox_3d"].cpu().numpy())
    assert np.array_equal(out["pred_instance_mask"].cpu().numpy(), want["pred_instance_mask"].astype(np.uint8))
    assert np.array_equal(out["range_pred"].cpu().numpy(), want["range_pred"].astype(np.float32))
    assert out["n_points"].cpu().tolist() == [len(p) for p in want["pred_points"]]

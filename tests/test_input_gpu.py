"""Parity of the input assembly (LatentDiffusion.get_input / encode_all_stages) against
tests/golden/get_input_tiny.npz produced by the unmodified reference: 4 VAE encodes (bf16 tensor-core path, tolerance
1e-2 like tests/test_vae_gpu.py), posterior samples with the recorded noise, nearest mask resize, lidar centre-crop +
negative pad, box re-normalisation and camera / lidar interleave (exact)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ldm():
    from mobi_b200.ddpm import LatentDiffusion
    from oracle import unet_oracle as uo
    from oracle import vae_oracle as vo
    ucfg = uo.tiny_unet_config()
    vae = lambda lidar: dict(target="mobi_b200.autoencoder.AutoencoderKL",
                             params=dict(ddconfig=vo.tiny_ddconfig(lidar), embed_dim=4, lossconfig=dict(target="torch.nn.Identity")))
    ldm = LatentDiffusion(unet_config=dict(target="mobi_b200.openaimodel.UNetModel", params=ucfg),
                          first_stage_config=vae(False), lidar_stage_config=vae(True), linear_start=0.00085,
                          linear_end=0.0120, timesteps=1000, first_stage_key="inpaint", image_size=16, channels=4,
                          conditioning_key="crossattn", scale_factor=0.18215, lidar_scale_factor=0.18215, use_camera=True,
                          use_lidar=True)
    ldm.first_stage_model.load_state_dict(uo.synth_state_dict(vo.state_dict_shapes(vo.tiny_ddconfig(False)), seed=10))
    ldm.lidar_stage_model.load_state_dict(uo.synth_state_dict(vo.state_dict_shapes(vo.tiny_ddconfig(True)), seed=20))
    return ldm.cuda().eval()


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


def test_get_input_vs_reference_golden():
    g = {k: torch.from_numpy(v).cuda() for k, v in np.load(os.path.join(GOLDEN, "get_input_tiny.npz")).items()}
    ldm = _ldm()
    rnd = lambda *s: torch.zeros(*s, device="cuda")
    batch = dict(image=dict(GT=g["image_gt"], inpaint_image=g["image_inpaint"], inpaint_mask=g["image_mask"],
                            cond=dict(ref_image=rnd(2, 3, 8, 8), ref_bbox=g["bbox_camera"].clone())),
                 lidar=dict(range_data=g["range_gt"], range_data_inpaint=g["range_inpaint"], range_mask=g["range_mask"],
                            cond=dict(ref_image=rnd(2, 3, 8, 8) + 1, ref_bbox=g["bbox_lidar_in"].clone())))
    noise = dict(camera=(g["noise_cam_gt"], g["noise_cam_inpaint"]), lidar=(g["noise_lid_gt"], g["noise_lid_inpaint"]))
    out = ldm.get_input(batch, "inpaint", noise=noise)
    torch.cuda.synchronize()
    assert out["z"].shape == g["z"].shape == (4, 9, 16, 16) and out["z_lidar"].shape == (2, 4, 24, 24)
    e_lat = rel(out["z"][:, :8], g["z"][:, :8])
    e_lid = rel(out["z_lidar"], g["z_lidar"])
    print("get_input latents max-abs-rel vs reference golden: interleaved z %.3e, un-cropped lidar z %.3e" % (e_lat, e_lid))
    assert e_lat < 1e-2 and e_lid < 1e-2
    assert torch.equal(out["z"][:, 8], g["z"][:, 8])                      # mask channel: exact (nearest resize, crop, zero pad)
    assert rel(out["cond"]["ref_bbox"], g["bbox_out"]) < 1e-6             # interleaved, lidar corners re-normalised
    assert torch.equal(out["cond"]["ref_image"][0::2], batch["image"]["cond"]["ref_image"])
    assert torch.equal(out["cond"]["ref_image"][1::2], batch["lidar"]["cond"]["ref_image"])
    # encode_all_stages alone (reference signature), posterior mode path = noise of zeros
    zi, zl = ldm.encode_all_stages(g["image_gt"], g["image_inpaint"], g["image_mask"], g["range_gt"], g["range_inpaint"],
                                   g["range_mask"], noise=noise)
    assert zi.shape == (2, 9, 16, 16) and zl.shape == (2, 9, 24, 24)
    assert rel(zi[:, :8], g["z"][0::2, :8]) < 1e-2 and rel(zl[:, :4], g["z_lidar"]) < 1e-2

"""Config / checkpoint compatibility on CPU: `${...}` interpolation like OmegaConf's, the `target:`
swap, and that a state dict with the REFERENCE's key names (tests/golden/shapes_512.json was dumped from the unmodified
reference modules) loads into the drop-in LatentDiffusion without missing or unexpected keys."""
import json
import os

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

YAML = """
use_camera: True
use_lidar: True
latent_size: 64
conditions: [ref_image, ref_bbox]
sizes: {image: 512, name: "run_${latent_size}"}
model:
  base_learning_rate: 8.0e-05
  target: ldm.models.diffusion.ddpm.LatentDiffusion
  params:
    linear_start: 0.00085
    linear_end: 0.0120
    timesteps: 1000
    first_stage_key: "inpaint"
    cond_stage_key: ${conditions}
    image_size: ${latent_size}
    channels: 4
    conditioning_key: crossattn
    scale_factor: 0.18215
    lidar_scale_factor: 0.18215
    use_camera: ${use_camera}
    use_lidar: ${use_lidar}
    unet_config:
      target: ldm.modules.diffusionmodules.openaimodel.UNetModel
      params:
        image_size: ${latent_size}
        in_channels: 9
        out_channels: 4
        model_channels: 320
        attention_resolutions: [4, 2, 1]
        num_res_blocks: 2
        channel_mult: [1, 2, 4, 4]
        num_heads: 8
        use_spatial_transformer: True
        transformer_depth: 1
        context_dim: 768
        use_checkpoint: False
        legacy: False
        bbox_cond: True
        use_camera: ${use_camera}
        use_lidar: ${use_lidar}
    first_stage_config:
      target: ldm.models.autoencoder.AutoencoderKL
      params:
        embed_dim: 4
        ddconfig: &dd {double_z: true, z_channels: 4, resolution: 512, in_channels: 3, out_ch: 3, ch: 128,
                       ch_mult: [1, 2, 4, 4], num_res_blocks: 2, attn_resolutions: [], dropout: 0.0}
        lossconfig: {target: torch.nn.Identity}
    lidar_stage_config:
      target: ldm.models.autoencoder.AutoencoderKL
      params:
        embed_dim: 4
        ddconfig: {double_z: true, z_channels: 4, resolution: 512, in_channels: 2, out_ch: 2, ch: 128,
                   ch_mult: [1, 2, 4, 4], num_res_blocks: 2, attn_resolutions: [], dropout: 0.0, lidar_adapter: true}
        lossconfig: {target: torch.nn.Identity}
    cond_stage_config: __is_unconditional__
"""


def test_interpolation_and_retarget():
    from mobi_b200 import config as C
    from mobi_b200.util import retarget
    cfg = C.load_config(YAML)
    p = cfg["model"]["params"]
    assert p["cond_stage_key"] == ["ref_image", "ref_bbox"] and p["image_size"] == 64 and p["use_lidar"] is True
    assert p["unet_config"]["params"]["image_size"] == 64 and cfg["sizes"]["name"] == "run_64"
    r = retarget(cfg["model"])
    assert r["target"] == "mobi_b200.ddpm.LatentDiffusion"
    assert r["params"]["unet_config"]["target"] == "mobi_b200.openaimodel.UNetModel"
    assert r["params"]["lidar_stage_config"]["target"] == "mobi_b200.autoencoder.AutoencoderKL"
    with pytest.raises(KeyError):
        C.load_config("a: ${missing.key}\n")
    with pytest.raises(ValueError):
        C.load_config("a: ${b}\nb: ${a}\n")


def test_reference_checkpoint_keys_load_into_the_drop_in():
    from mobi_b200 import config as C
    shapes = json.load(open(os.path.join(GOLDEN, "shapes_512.json")))
    with torch.device("meta"):
        model = C.build_model(C.load_config(YAML), device=None)
    have = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    want = {}
    for prefix, name in (("model.diffusion_model.", "unet_512"), ("first_stage_model.", "vae_camera"),
                         ("lidar_stage_model.", "vae_lidar")):
        for k, shp in shapes[name].items():
            if k.startswith("loss."):
                continue
            want[prefix + k] = tuple(shp)
    missing = sorted(set(want) - set(have))
    assert not missing, missing[:5]
    for k, shp in want.items():
        assert have[k] == shp, (k, have[k], shp)
    extra = sorted(k for k in set(have) - set(want)
                   if not k.startswith(("proj_out.", "learnable_vector", "bbox_uncond_vector")) and "alphas" not in k
                   and "betas" not in k and "posterior" not in k and "sqrt_" not in k and "log_one_minus" not in k)
    assert not extra, extra[:5]
    # a Lightning-style checkpoint dict with an EMA copy loads by key
    fake = {"state_dict": {**{k: torch.zeros(s) for k, s in list(want.items())[:3]}, "model_ema.decay": torch.zeros(())}}
    cpu_model = torch.nn.Module()
    missing_keys, unexpected = C.load_checkpoint(model.to_empty(device="cpu"), fake, strict=False)
    assert not unexpected

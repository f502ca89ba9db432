"""Plug-in glue mirroring ldm/util.py: `instantiate_from_config` (ldm/util.py:76-91) is how every reference
component is created from YAML, and therefore how these drop-ins are selected:

    unet_config:
      target: mobi_b200.openaimodel.UNetModel      # was ldm.modules.diffusionmodules.openaimodel.UNetModel
"""
import importlib

import torch


def get_obj_from_str(string, reload=False):
    module, cls = string.rsplit(".", 1)
    if reload:
        importlib.reload(importlib.import_module(module))
    return getattr(importlib.import_module(module, package=None), cls)


def instantiate_from_config(config):
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**config.get("params", dict()))


def cat_interleave(tensors):
    """ldm/util.py:213-221: [cam0, lid0, cam1, lid1, ...] along the batch dimension."""
    if len(tensors) == 0:
        return tensors
    t = torch.cat([u.unsqueeze(1) for u in tensors], dim=1)
    return t.reshape(-1, *t.shape[2:])


# Reference class paths -> drop-in class paths (what INTEGRATION.md asks a maintainer to change in the YAML).
TARGET_MAP = {
    "ldm.modules.diffusionmodules.openaimodel.UNetModel": "mobi_b200.openaimodel.UNetModel",
    "ldm.models.autoencoder.AutoencoderKL": "mobi_b200.autoencoder.AutoencoderKL",
    "ldm.models.diffusion.ddpm.LatentDiffusion": "mobi_b200.ddpm.LatentDiffusion",
    "ldm.modules.encoders.modules.FrozenCLIPImageEmbedder": "mobi_b200.encoders.FrozenCLIPImageEmbedder",
    "ldm.modules.encoders.modules.BBoxEmbedder": "mobi_b200.encoders.BBoxEmbedder",
}


def retarget(config):
    """Returns a copy of a (nested dict) reference config with hot-path targets swapped for the drop-ins."""
    if isinstance(config, dict):
        out = {k: retarget(v) for k, v in config.items()}
        if out.get("target") in TARGET_MAP:
            out["target"] = TARGET_MAP[out["target"]]
        return out
    if isinstance(config, list):
        return [retarget(v) for v in config]
    return config

"""Plug-in glue mirroring ldm/util.py: `instantiate_from_config` (ldm/util.py:76-91) is how every reference
component is created from YAML, and therefore how these drop-ins are selected:

    unet_config:
      target: mobi_b200.openaimodel.UNetModel      # was ldm.modules.diffusionmodules.openaimodel.UNetModel
"""
import importlib

import torch


_SENTINELS = ("__is_first_stage__", "__is_unconditional__")   # config values the reference treats as "no module"


def get_obj_from_str(string, reload=False):
    """Resolves a dotted "package.module.Name" path to the object it names (same contract as ldm/util.py:86-91)."""
    module_name, _, attr = string.rpartition(".")
    module = importlib.import_module(module_name)
    if reload:
        module = importlib.reload(module)
    return getattr(module, attr)


def instantiate_from_config(config):
    """{target: dotted path, params: kwargs} -> instance (the reference's plug-in mechanism, ldm/util.py:76-83)."""
    target = config.get("target") if hasattr(config, "get") else None
    if target is None:
        if config in _SENTINELS:
            return None
        raise KeyError("Expected key `target` to instantiate.")
    kwargs = config.get("params") or {}
    return get_obj_from_str(target)(**kwargs)


def cat_interleave(tensors):
    """ldm/util.py:213-221: [cam0, lid0, cam1, lid1, ...] along the batch dimension (stack on a new axis 1, fold it
    into the batch)."""
    if not tensors:
        return tensors
    stacked = torch.stack(list(tensors), dim=1)
    return stacked.flatten(0, 1)


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

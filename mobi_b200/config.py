"""Config / checkpoint compatibility: read the reference's YAML files without OmegaConf,
honouring its `${key}` / `${a.b.c}` interpolations (configs/mobi_nusc_512.yaml:38-52, 66-82), swap the hot-path `target:`
strings for the drop-ins and load reference checkpoints by their own keys
(`model.diffusion_model.*`, `first_stage_model.*`, `lidar_stage_model.*`, `cond_stage_model.*`, `proj_out.*`,
`learnable_vector`, `bbox_uncond_vector`; scripts/inference_test_bench.py:150-167, main.py:503-533).
"""
import re

import torch
import yaml

from .util import instantiate_from_config, retarget

_INTERP = re.compile(r"\$\{([^}]+)\}")


def _lookup(root, dotted):
    node = root
    for part in dotted.strip().split("."):
        if isinstance(node, list):
            node = node[int(part)]
        elif isinstance(node, dict) and part in node:
            node = node[part]
        else:
            raise KeyError("config interpolation ${%s}: key %r not found" % (dotted, part))
    return node


def _resolve(node, root, depth=0):
    if depth > 32:
        raise ValueError("config interpolation is cyclic")
    if isinstance(node, dict):
        return {k: _resolve(v, root, depth) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root, depth) for v in node]
    if isinstance(node, str):
        m = _INTERP.fullmatch(node.strip())
        if m:   # the whole value is one reference: keep the referenced node's type (list, int, bool, ...)
            return _resolve(_lookup(root, m.group(1)), root, depth + 1)
        if _INTERP.search(node):   # embedded in a longer string: substitute the text
            return _INTERP.sub(lambda mm: str(_resolve(_lookup(root, mm.group(1)), root, depth + 1)), node)
    return node


def load_config(source):
    """`source`: path of a YAML file or the YAML text itself.  Returns plain nested dicts / lists with every `${...}`
    resolved against the document root (OmegaConf's absolute interpolation, the only form the reference configs use)."""
    text = source
    if "\n" not in source and not source.lstrip().startswith("{"):
        with open(source) as fh:
            text = fh.read()
    raw = yaml.safe_load(text)
    return _resolve(raw, raw)


def build_model(config, ckpt=None, device="cuda", strict=False, verbose=False, clip_tower=None, meta_init=False):
    """config: a loaded reference config (or its `model` sub-dict).  Instantiates the drop-in LatentDiffusion (UNet, both
    autoencoders, conditioning stage when configured) and optionally loads a reference checkpoint.
    clip_tower: the CLIP vision tower module to inject into the conditioning stage (out of scope of this package; without
    it FrozenCLIPImageEmbedder loads `openai/clip-vit-large-patch14` from disk like the reference does).
    meta_init: build on the meta device and materialise on `device` (skips the random init of 1.1 B parameters that a
    checkpoint overwrites anyway); needs `ckpt`."""
    model_cfg = retarget(config["model"] if "model" in config else config)
    if clip_tower is not None:
        model_cfg["params"]["cond_stage_config"]["params"]["transformer"] = clip_tower
    if meta_init:
        assert ckpt is not None and device is not None, "meta_init materialises empty tensors: a checkpoint must fill them"
        with torch.device("meta"):
            model = instantiate_from_config(model_cfg)
        model = model.to_empty(device=device)
        p = model_cfg["params"]
        model.register_schedule(linear_start=p.get("linear_start", 1e-4), linear_end=p.get("linear_end", 2e-2),
                                timesteps=p.get("timesteps", 1000))
        model = model.to(device)
        missing, unexpected = load_checkpoint(model, ckpt, strict=strict, verbose=verbose)
        params = {n for n, _ in model.named_parameters()}   # (schedule buffers were rebuilt by register_schedule above)
        still = [k for k in missing if k in params and not k.startswith("cond_stage_model.transformer.")]
        if still:
            raise RuntimeError("build_model(meta_init=True): the checkpoint leaves %d tensors uninitialised, e.g. %s"
                               % (len(still), still[:3]))
        return model.eval()
    model = instantiate_from_config(model_cfg)
    if ckpt is not None:
        load_checkpoint(model, ckpt, strict=strict, verbose=verbose)
    return model.to(device).eval() if device is not None else model.eval()


def load_checkpoint(model, path_or_state, strict=False, verbose=False):
    """Loads `{"state_dict": ...}` (Lightning checkpoint) or a bare state dict with the reference's key names.
    Returns (missing, unexpected) like nn.Module.load_state_dict."""
    sd = path_or_state
    if isinstance(sd, str):
        sd = torch.load(sd, map_location="cpu")
    if isinstance(sd, dict) and "state_dict" in sd:
        sd = sd["state_dict"]
    # buffers the reference registers on LatentDiffusion / the EMA copy are not parameters of the drop-in
    sd = {k: v for k, v in sd.items() if not k.startswith("model_ema.")}
    res = model.load_state_dict(sd, strict=strict)
    if verbose:
        print("load_checkpoint: %d missing, %d unexpected keys" % (len(res.missing_keys), len(res.unexpected_keys)))
    return res.missing_keys, res.unexpected_keys

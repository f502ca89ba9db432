"""TEST INFRASTRUCTURE ONLY — lets the UNMODIFIED reference under /root/reference be imported in the build
container (it is a Python reference; it cannot travel to the GPU box, so everything it produces is committed
as fixtures under tests/golden/ by oracle/make_golden.py).

Three third-party imports of the reference are absent here: `omegaconf`, `pytorch_lightning`
and `taming`.  None is used on the hot path, so they are shimmed with empty stand-ins; and the samplers'
`register_buffer` is patched not to force "cuda" (ddim.py:19-23, plms.py:18-22) so they run on CPU.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MOBI_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ldm"))


def install():
    import torch
    import torch.nn as nn

    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")
        lc = types.ModuleType("omegaconf.listconfig")

        class ListConfig(list):
            pass

        lc.ListConfig = ListConfig
        oc.listconfig = lc
        oc.ListConfig = ListConfig
        oc.OmegaConf = type("OmegaConf", (), {})
        sys.modules["omegaconf"] = oc
        sys.modules["omegaconf.listconfig"] = lc
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            @property
            def device(self):
                try:
                    return next(self.parameters()).device
                except StopIteration:
                    return torch.device("cpu")

            def log(self, *a, **k):
                pass

            def log_dict(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        util = types.ModuleType("pytorch_lightning.utilities")
        dist = types.ModuleType("pytorch_lightning.utilities.distributed")
        dist.rank_zero_only = lambda f: f
        util.distributed = dist
        pl.utilities = util
        sys.modules["pytorch_lightning"] = pl
        sys.modules["pytorch_lightning.utilities"] = util
        sys.modules["pytorch_lightning.utilities.distributed"] = dist
    if "taming" not in sys.modules:
        names = ["taming", "taming.modules", "taming.modules.vqvae", "taming.modules.vqvae.quantize"]
        mods = {n: types.ModuleType(n) for n in names}
        mods["taming.modules.vqvae.quantize"].VectorQuantizer2 = object
        for n in names:
            sys.modules[n] = mods[n]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    from ldm.models.diffusion import ddim, plms

    def _register_buffer(self, name, attr):
        setattr(self, name, attr)

    ddim.DDIMSampler.register_buffer = _register_buffer
    plms.PLMSSampler.register_buffer = _register_buffer

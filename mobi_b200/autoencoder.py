"""B200-native drop-in for ldm/models/autoencoder.py:AutoencoderKL (encode / decode only: autoencoder.py:63-72) and
ldm/modules/distributions/distributions.py:DiagonalGaussianDistribution (24-62).  Same constructor signature and
state-dict keys (`encoder.*`, `decoder.*`, `quant_conv.*`, `post_quant_conv.*`), so the camera VAE and the
range-view VAE checkpoints load unchanged.  The GAN/LPIPS training step of the autoencoder is a separate workflow of
the reference and is out of scope.
"""
import torch
import torch.nn as nn

from . import ops
from .model import Decoder, Encoder
from .openaimodel import Conv3x3


class DiagonalGaussianDistribution(object):
    """distributions.py:24-62 (tiny elementwise math on the [N, 8, h, w] moments; plain torch ops on the device)."""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if self.deterministic:
            self.var = self.std = torch.zeros_like(self.mean).to(device=self.parameters.device)

    def sample(self, noise=None):
        noise = torch.randn(self.mean.shape, device=self.parameters.device) if noise is None else noise
        return self.mean + self.std * noise

    def kl(self, other=None):
        if self.deterministic:
            return torch.Tensor([0.])
        if other is None:
            return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])
        return 0.5 * torch.sum(torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var - 1.0
                               - self.logvar + other.logvar, dim=[1, 2, 3])

    def mode(self):
        return self.mean


class AutoencoderKL(nn.Module):
    def __init__(self, ddconfig, lossconfig=None, embed_dim=4, ckpt_path=None, ignore_keys=(), image_key="image",
                 colorize_nlabels=None, monitor=None, range_object_norm=False, range_object_norm_scale=0.75,
                 range_int_norm=False, **kwargs):
        super().__init__()
        self.image_key = image_key
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.range_object_norm = range_object_norm
        self.range_object_norm_scale = range_object_norm_scale
        self.range_int_norm = range_int_norm
        assert ddconfig["double_z"]
        self.quant_conv = nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.embed_dim = embed_dim
        self._p = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)

    def init_from_ckpt(self, path, ignore_keys=()):
        sd = torch.load(path, map_location="cpu")["state_dict"]
        for k in list(sd.keys()):
            if any(k.startswith(ik) for ik in ignore_keys):
                del sd[k]
        self.load_state_dict(sd, strict=False)

    def invalidate(self):
        self._p = None
        self.encoder._p = None
        self.decoder._p = None

    @torch.no_grad()
    def pack(self):
        self.encoder.pack()
        self.decoder.pack()
        self._p = dict(quant=Conv3x3.pack(self.quant_conv), post_quant=Conv3x3.pack(self.post_quant_conv))

    @torch.no_grad()
    def encode(self, x):
        """autoencoder.py:63-67: x NCHW f32 image / range image -> posterior over the [N, 4, h/8, w/8] latent."""
        if not x.is_cuda:
            raise RuntimeError("mobi_b200.AutoencoderKL runs on CUDA only (no CPU fallback)")
        if self._p is None:
            self.pack()
        h = self.encoder.run(ops.nchw_to_nhwc(x.detach().float().contiguous()))
        moments = Conv3x3.run(self._p["quant"], h)
        return DiagonalGaussianDistribution(ops.nhwc_to_nchw(moments))

    @torch.no_grad()
    def decode(self, z):
        """autoencoder.py:69-72: z NCHW f32 [N, 4, h, w] -> NCHW f32 [N, out_ch, 8h, 8w]."""
        if not z.is_cuda:
            raise RuntimeError("mobi_b200.AutoencoderKL runs on CUDA only (no CPU fallback)")
        if self._p is None:
            self.pack()
        zz = Conv3x3.run(self._p["post_quant"], ops.nchw_to_nhwc(z.detach().float().contiguous()))
        return ops.nhwc_to_nchw(self.decoder.run(zz))

    def forward(self, input, sample_posterior=True):
        posterior = self.encode(input)
        z = posterior.sample() if sample_posterior else posterior.mode()
        return self.decode(z), posterior

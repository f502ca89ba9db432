import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as RealF
from oracle import vae_oracle as vo, unet_oracle as uo
torch.set_num_threads(8)
class QF:
    def __init__(self, conv=None): self.conv = conv
    def __getattr__(self, n): return getattr(RealF, n)
    def conv2d(self, x, w, b=None, **kw):
        if self.conv is not None and w.shape[1] >= 64:
            x, w = x.to(self.conv).float(), w.to(self.conv).float()
        return RealF.conv2d(x, w, b, **kw)
for lidar in (False, True):
    cfg = vo.default_ddconfig(lidar, resolution=256)
    sd = uo.synth_state_dict(vo.state_dict_shapes(cfg), seed=30 + int(lidar))
    z = torch.randn(2, 4, 32, 32, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref = vo.vae_decode(sd, cfg, z)
        vo.F = QF(torch.bfloat16)
        out = vo.vae_decode(sd, cfg, z)
        vo.F = RealF
    e = ((out - ref).abs().max() / ref.abs().max()).item()
    print("lidar", lidar, "bf16-operand emulation max-abs-rel %.3e" % e, "ref absmax %.3f" % ref.abs().max().item(), "rms rel %.3e" % ((out-ref).pow(2).mean().sqrt()/ref.pow(2).mean().sqrt()).item())

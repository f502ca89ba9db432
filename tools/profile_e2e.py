#!/usr/bin/env python
"""Where does the time go outside the per-kernel sums?  (1) one UNet call: sum of per-op CUDA-event times vs the
whole eager call vs a CUDA-graph replay; (2) camera / range-view VAE decode broken down by kernel class."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mobi_b200 import ops, synth  # noqa: E402

n = int(os.environ.get("PE_SAMPLES", "8"))
latent = 64
dev = torch.device("cuda", 0)
ldm = synth.build_synthetic_ldm(latent=latent, device=dev, seed=0, with_vae=True)
unet = ldm.model.diffusion_model
inp = synth.synthetic_inputs(n, latent, seed=1, device=dev)
x_in = torch.randn(4 * n, 9, latent, latent, device=dev)
t_in = torch.full((4 * n,), 481, device=dev, dtype=torch.long)
c_in = torch.cat([inp["uc"], inp["cond"]]).contiguous()


def ev():
    return torch.cuda.Event(enable_timing=True)


for _ in range(2):
    unet(x_in, t_in, context=c_in)
torch.cuda.synchronize()
ops.Stats.begin_profile()
unet(x_in, t_in, context=c_in)
prof = ops.Stats.end_profile()
print("per-op sum: %.2f ms  %s" % (sum(v["ms"] for v in prof.values()),
                                  {k: (v["launches"], round(v["ms"], 2)) for k, v in sorted(prof.items())}))
e0, e1 = ev(), ev()
e0.record()
for _ in range(3):
    unet(x_in, t_in, context=c_in)
e1.record()
torch.cuda.synchronize()
print("eager whole call: %.2f ms" % (e0.elapsed_time(e1) / 3))
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    unet(x_in, t_in, context=c_in)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
before = ops.Stats.launches
with torch.cuda.graph(g):
    out = unet(x_in, t_in, context=c_in)
print("kernels in graph (C-ABI launches):", ops.Stats.launches - before)
g.replay()
torch.cuda.synchronize()
e0, e1 = ev(), ev()
e0.record()
for _ in range(5):
    g.replay()
e1.record()
torch.cuda.synchronize()
print("graph replay: %.2f ms" % (e0.elapsed_time(e1) / 5))

z = torch.randn(2 * n, 4, latent, latent, device=dev)
for name, mod, zz in (("camera", "first_stage_model", z[::2]), ("lidar", "lidar_stage_model", z[1::2])):
    for _ in range(2):
        ldm.decode_first_stage(zz, module_name=mod)
    torch.cuda.synchronize()
    ops.Stats.begin_profile()
    ldm.decode_first_stage(zz, module_name=mod)
    prof = ops.Stats.end_profile()
    e0, e1 = ev(), ev()
    e0.record()
    ldm.decode_first_stage(zz, module_name=mod)
    e1.record()
    torch.cuda.synchronize()
    print("%s decode of %d images: whole %.2f ms, per-op sum %.2f ms  %s" % (
        name, n, e0.elapsed_time(e1), sum(v["ms"] for v in prof.values()),
        {k: (v["launches"], round(v["ms"], 2), round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 0)) for k, v in sorted(prof.items())}))

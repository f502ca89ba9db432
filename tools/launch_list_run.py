#!/usr/bin/env python
"""The workload of the ncu launch list: two eager DDIM + CFG steps of 8 joint samples (32 UNet rows, mobi_nusc_512 shapes,
synthetic weights) between cudaProfilerStart / Stop, so that `ncu --profile-from-start off` lists exactly the kernels of the
hot path (model construction, weight synthesis and the warm-up call stay outside).
Usage: ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/launch_list_run.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mobi_b200 import synth  # noqa: E402
from mobi_b200.ddim import DDIMSampler  # noqa: E402

N, LATENT, STEPS = int(os.environ.get("LL_SAMPLES", "8")), 64, int(os.environ.get("LL_STEPS", "2"))
dev = torch.device("cuda", 0)
ldm = synth.build_synthetic_ldm(latent=LATENT, use_lidar=True, device=dev, seed=0, with_vae=False, with_cond=False)
sampler = DDIMSampler(ldm, use_cuda_graph=False)
inp = {k: v.to(dev) for k, v in synth.synthetic_inputs(N, LATENT, seed=1, rows_per_sample=2, n_ctx=2).items()}
x_T = torch.randn(2 * N, 4, LATENT, LATENT, device=dev)


def run():
    return sampler.sample(S=STEPS, conditioning=inp["cond"], batch_size=2 * N, shape=[4, LATENT, LATENT], verbose=False,
                          unconditional_guidance_scale=5.0, unconditional_conditioning=inp["uc"], eta=0.0, x_T=x_T,
                          test_model_kwargs=dict(inpaint_image=inp["inpaint_image"], inpaint_mask=inp["inpaint_mask"]))[0]


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launch_list_run: %d joint samples, %d DDIM steps, output %s finite=%s" % (N, STEPS, tuple(out.shape), bool(torch.isfinite(out).all())))

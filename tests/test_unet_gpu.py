"""Parity of the CUDA UNet / samplers (through the drop-in classes and the C ABI) against
 (a) the committed golden vectors produced by the unmodified reference (tiny topology-complete config), and
 (b) the oracle restatement run in fp32 on the same device (TF32 off) at the real model width.
Tolerance: per-call eps max-abs-rel under a fixed cap of 2e-2 AND within 1.5x of the bf16-operand
floor of the same call (the fp32 oracle with only its contraction operands rounded to bf16, oracle/precision.py): the
measured floor of these calls is 0.9-1.1e-2, and the kernel path sits at 0.8-1.0x of it."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_EPS = 2e-2


def relerr(a, b):
    a, b = a.double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


def build_unet(cfg, seed=0):
    from mobi_b200.openaimodel import UNetModel
    from oracle import unet_oracle as uo
    sd = uo.synth_state_dict(uo.state_dict_shapes(cfg), seed=seed)
    with torch.device("meta"):
        net = UNetModel(**cfg)
    net = net.to_empty(device="cuda")
    missing = net.load_state_dict(sd, strict=True)
    return net.eval(), {k: v.cuda() for k, v in sd.items()}


def build_ldm(cfg, seed=0):
    from mobi_b200.ddpm import LatentDiffusion
    from oracle import unet_oracle as uo
    sd = uo.synth_state_dict(uo.state_dict_shapes(cfg), seed=seed)
    ldm = LatentDiffusion(unet_config=dict(target="mobi_b200.openaimodel.UNetModel", params=cfg),
                          linear_start=0.00085, linear_end=0.0120, timesteps=1000, first_stage_key="inpaint",
                          image_size=cfg["image_size"], channels=4, conditioning_key="crossattn", scale_factor=0.18215,
                          lidar_scale_factor=0.18215, use_camera=True, use_lidar=True).cuda().eval()
    ldm.model.diffusion_model.load_state_dict(sd, strict=True)
    return ldm, {k: v.cuda() for k, v in sd.items()}


def test_unet_tiny_vs_reference_golden():
    from oracle import unet_oracle as uo
    g = np.load(os.path.join(GOLDEN, "unet_tiny.npz"))
    cfg = uo.tiny_unet_config()
    net, sd = build_unet(cfg)
    x, t, ctx = (torch.from_numpy(g[k]).cuda() for k in ("x", "t", "context"))
    eps = net(x, t, context=ctx)
    torch.cuda.synchronize()
    e = relerr(eps, g["eps"])
    print("tiny unet eps max-abs-rel vs reference golden: %.3e" % e)
    # The 64-channel golden model averages bf16 operand rounding over 5x fewer terms than the real 320-channel
    # model: an fp32 oracle with every GEMM/conv operand rounded to bf16 (tools/precision_probe.py) is itself
    # 1.4e-2 away from the fp32 reference here (9.4e-3 at full width).  Structural errors show up as O(1).
    assert e < 2.5e-2
    # calling again with the same context reuses the context tables and is deterministic
    assert torch.equal(net(x, t, context=ctx), eps)


@pytest.mark.parametrize("use_lidar,n_ctx", [(True, 2), (False, 1)])
def test_unet_fullwidth_vs_oracle(use_lidar, n_ctx):
    """Config 1 (mobi_nusc-mini_256: N=1 joint sample x CFG = 4 rows, latent 32x32, t=981) and the camera-only
    pbe.yaml variant, at the real model width (1.04 B / 0.92 B parameters)."""
    from oracle import unet_oracle as uo
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = uo.default_unet_config(image_size=32, use_lidar=use_lidar)
    net, sd = build_unet(cfg)
    inp = uo.synth_inputs(1, 32, seed=1, device="cuda")
    x = torch.cat([inp["x_T"], inp["inpaint_image"], inp["inpaint_mask"]], 1)
    x_in = torch.cat([x, x])
    t_in = torch.full((4,), 981, device="cuda", dtype=torch.long)
    c_in = torch.cat([inp["uc"], inp["cond"]])[:, :n_ctx].contiguous()
    eps = net(x_in, t_in, context=c_in)
    with torch.no_grad():
        ref = uo.unet_forward(sd, cfg, x_in, t_in, c_in)
    from oracle.precision import bf16_operand_floor
    floor = bf16_operand_floor(sd, cfg, x_in, t_in, c_in, ref=ref)
    e = relerr(eps, ref)
    cos = torch.nn.functional.cosine_similarity(eps.flatten().double(), ref.flatten().double(), dim=0).item()
    print("full-width unet (lidar=%s) eps max-abs-rel %.3e cosine %.6f ref-absmax %.3f | bf16-operand floor %.3e"
          % (use_lidar, e, cos, ref.abs().max(), floor[0]))
    assert e < TOL_EPS and e < 1.5 * floor[0] and cos > 0.9999


def _sampler_inputs():
    g = np.load(os.path.join(GOLDEN, "ddim_tiny.npz"))
    return g, {k: torch.from_numpy(g[k]).cuda() for k in ("x_T", "cond", "uc", "inpaint_image", "inpaint_mask")}


@pytest.mark.parametrize("graph", [False, True])
def test_ddim_tiny_vs_reference_golden(graph):
    from mobi_b200.ddim import DDIMSampler
    from oracle import unet_oracle as uo
    g, T = _sampler_inputs()
    ldm, _ = build_ldm(uo.tiny_unet_config())
    smp = DDIMSampler(ldm, use_cuda_graph=graph)
    samples, inter = smp.sample(S=int(g["S"]), conditioning=T["cond"], batch_size=4, shape=[4, 16, 16], verbose=False,
                                unconditional_guidance_scale=float(g["scale"]), unconditional_conditioning=T["uc"],
                                eta=0.0, x_T=T["x_T"], log_every_t=1,
                                test_model_kwargs=dict(inpaint_image=T["inpaint_image"], inpaint_mask=T["inpaint_mask"]))
    e = relerr(samples, g["samples"])
    print("ddim tiny (graph=%s) final-latent max-abs-rel vs reference: %.3e" % (graph, e))
    assert e < 2.5e-2  # 4 CFG-3 steps compound the per-call bf16 error (see the tiny-model note above)
    assert len(inter["x_inter"]) == 5 and relerr(inter["x_inter"][1], g["x_inter"][1]) < 2.5e-2
    assert smp.launches == 4


def test_plms_tiny_vs_reference_golden():
    from mobi_b200.plms import PLMSSampler
    from oracle import unet_oracle as uo
    g, T = _sampler_inputs()
    gp = np.load(os.path.join(GOLDEN, "plms_tiny.npz"))
    ldm, _ = build_ldm(uo.tiny_unet_config())
    smp = PLMSSampler(ldm)
    samples, _ = smp.sample(S=int(gp["S"]), conditioning=T["cond"], batch_size=4, shape=[4, 16, 16], verbose=False,
                            unconditional_guidance_scale=float(gp["scale"]), unconditional_conditioning=T["uc"],
                            eta=0.0, x_T=T["x_T"], inpaint_image=T["inpaint_image"], inpaint_mask=T["inpaint_mask"])
    e = relerr(samples, gp["samples"])
    print("plms tiny final-latent max-abs-rel vs reference: %.3e" % e)
    assert e < 2.5e-2
    assert smp.launches == 5  # S + 1 UNet evaluations (plms.py:223-226)


def test_ddim_blend_path_matches_oracle():
    """mask/x0 known-latent blend (ddim.py:145-148): noise is drawn on the device, so compare against the oracle
    fed with the same draws."""
    from mobi_b200.ddim import DDIMSampler
    from oracle import sampler_oracle as so
    from oracle import unet_oracle as uo
    g, T = _sampler_inputs()
    cfg = uo.tiny_unet_config()
    ldm, sd = build_ldm(cfg)
    bmask = torch.from_numpy(g["blend_mask"]).cuda()
    bx0 = torch.from_numpy(g["blend_x0"]).cuda()
    torch.manual_seed(5)
    gen_state = torch.cuda.get_rng_state()
    smp = DDIMSampler(ldm, use_cuda_graph=False)
    samples, _ = smp.sample(S=4, conditioning=T["cond"], batch_size=4, shape=[4, 16, 16], verbose=False,
                            unconditional_guidance_scale=3.0, unconditional_conditioning=T["uc"], eta=0.0,
                            x_T=T["x_T"], mask=bmask, x0=bx0,
                            test_model_kwargs=dict(inpaint_image=T["inpaint_image"], inpaint_mask=T["inpaint_mask"]))
    torch.cuda.set_rng_state(gen_state)
    noises = [torch.randn_like(bx0) for _ in range(4)]
    with torch.no_grad():
        ref, _ = so.ddim_sample(lambda x, t, c: uo.unet_forward(sd, cfg, x, t, c), so.register_schedule(), 4, T["x_T"],
                                T["cond"], T["uc"], 3.0, T["inpaint_image"], T["inpaint_mask"], mask=bmask, x0=bx0,
                                blend_noise=noises)
    e = relerr(samples, ref)
    print("ddim blend path max-abs-rel vs oracle: %.3e" % e)
    assert e < 2.5e-2

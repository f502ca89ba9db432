"""Parity at the shapes BASELINE.json's metric is quoted on (VERDICT r1 "pin the headline config"): everything here runs
the real 1.04 B-parameter joint UNet (or the 0.92 B camera-only one) at latent 64 x 64 = mobi_nusc_512 / pbe.yaml,
through the drop-in classes and the C ABI, against the fp32 oracle on the same GPU (TF32 off) on the shared synthetic
state dict.

Tolerances ("max-abs-rel" = max-abs error relative to the max-abs of the oracle tensor, the north star's measure;
"rel-rms" = ||a - b|| / ||b||):
  * one UNet evaluation (eps, before the CFG combine), also teacher-forced at every one of the 50 DDIM / 51 PLMS
    steps: max-abs-rel <= 2e-2, rel-rms <= 1.5e-2, AND both <= 1.5 x the bf16-operand floor of the same call.
    The north star's example figure (1e-2) is where this model's bf16-OPERAND FLOOR itself sits at latent 64: the fp32
    oracle with nothing changed but the contraction operands rounded to bf16 (oracle/precision.py) is 0.9-1.4e-2 away
    from the fp32 oracle, so no bf16-operand implementation can promise 1e-2 at every step; the floor is measured in
    the test and printed beside the CUDA path's error (round-2 measurement: 32-row call 9.9e-3; per-step curve
    7.1e-3 .. 1.4e-2, mean 9.5e-3; the numbers are recorded in DESIGN.md §2).
  * final latents after the free-running 50-step CFG-5 run: max-abs-rel <= 2e-2, cosine >= 0.9995 (measured 6.8e-3 /
    0.99998: per-step errors are largely independent between steps and DDIM averages them; SURVEY.md §8c had proposed
    5e-2).
  * VAE decode at 512 px: max-abs-rel <= 2e-2 and <= 1.5 x the bf16-operand floor of the same decode (the decoder is a
    chain of ~30 convolutions at up to 512 x 512: its floor is measured in the test like the UNet's; round-2
    measurement 1.07e-2 camera / 1.13e-2 lidar).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_EPS = 2e-2          # max-abs-rel of one UNet evaluation (cap; the binding bar is 1.5 x the bf16-operand floor)
TOL_EPS_RMS = 1.5e-2    # rel-rms of one UNet evaluation (cap; the bf16-operand floor's own rel-rms is 0.7-1.06e-2)
TOL_FINAL = 2e-2        # final latents of a 50-step run
TOL_VAE = 2e-2          # cap; the binding bar is 1.5 x the bf16-operand floor of the same decode


def relerr(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max()).item()


def relrms(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm()).item()


def check_eps(eps, ref, floor, what):
    """The per-call bar: see the module docstring.  floor = (max-abs-rel, rel-rms) of the bf16-operand oracle."""
    e, r = relerr(eps, ref), relrms(eps, ref)
    print("%s: eps max-abs-rel %.3e rel-rms %.3e cosine %.6f | bf16-operand floor of the same call: %.3e / %.3e"
          % (what, e, r, cosine(eps, ref), floor[0], floor[1]))
    assert e < TOL_EPS and r < TOL_EPS_RMS and e < 1.5 * floor[0] and r < 1.5 * floor[1], (what, e, r, floor)
    return e


def cosine(a, b):
    return torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()


def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def build_ldm(latent=64, use_lidar=True, with_vae=False):
    """The drop-in LatentDiffusion at full width with the oracle's synthetic state dict loaded by the reference's keys."""
    from mobi_b200 import synth
    from mobi_b200.ddpm import LatentDiffusion
    from oracle import unet_oracle as uo
    cfg = uo.default_unet_config(image_size=latent, use_lidar=use_lidar)
    sd = uo.synth_state_dict(uo.state_dict_shapes(cfg), seed=0)
    vae = lambda lidar: dict(target="mobi_b200.autoencoder.AutoencoderKL",                       # noqa: E731
                             params=dict(ddconfig=synth.vae_ddconfig(latent, lidar), embed_dim=4,
                                         lossconfig=dict(target="torch.nn.Identity")))
    with torch.device("meta"):
        ldm = LatentDiffusion(unet_config=dict(target="mobi_b200.openaimodel.UNetModel", params=cfg),
                              first_stage_config=vae(False) if with_vae else None,
                              linear_start=0.00085, linear_end=0.0120, timesteps=1000, first_stage_key="inpaint",
                              image_size=latent, channels=4, conditioning_key="crossattn", scale_factor=0.18215,
                              lidar_scale_factor=0.18215, use_camera=True, use_lidar=use_lidar)
    ldm = ldm.to_empty(device="cuda")
    ldm.register_schedule(linear_start=0.00085, linear_end=0.0120, timesteps=1000)
    ldm = ldm.to("cuda").eval()
    ldm.model.diffusion_model.load_state_dict(sd, strict=True)
    return ldm, cfg, {k: v.cuda() for k, v in sd.items()}


def oracle_unet(sd, cfg, chunk=8):
    """apply_model of the oracle, evaluated `chunk` rows at a time (rows are independent except the camera / lidar pairs,
    which an even chunk never splits): bounds the materialised T x T scores of the eager attention to a few GB."""
    from oracle import unet_oracle as uo

    def apply_model(x, t, c):
        with torch.no_grad():
            return torch.cat([uo.unet_forward(sd, cfg, x[i:i + chunk], t[i:i + chunk], c[i:i + chunk])
                              for i in range(0, x.shape[0], chunk)])
    return apply_model


# ----------------------------------------------------------------------------------------------- (i) one UNet call
def test_unet_call_latent64_32rows_vs_oracle():
    """Exactly bench.py's per-GPU shape at 8 joint samples: 32 rows (8 joint samples x CFG) at latent 64, t = 981: the
    T = 4096 / d = 40 fused attention, the CTA-pair convs and the cluster GroupNorm chunking all at their bench sizes."""
    from oracle import unet_oracle as uo
    _no_tf32()
    ldm, cfg, sd = build_ldm(64)
    net = ldm.model.diffusion_model
    inp = uo.synth_inputs(8, 64, seed=1, device="cuda")
    x = torch.cat([inp["x_T"], inp["inpaint_image"], inp["inpaint_mask"]], 1)
    x_in = torch.cat([x, x])
    t_in = torch.full((32,), 981, device="cuda", dtype=torch.long)
    c_in = torch.cat([inp["uc"], inp["cond"]]).contiguous()
    from oracle.precision import bf16_operand_floor
    eps = net(x_in, t_in, context=c_in)
    ref = oracle_unet(sd, cfg)(x_in, t_in, c_in)
    check_eps(eps, ref, bf16_operand_floor(sd, cfg, x_in, t_in, c_in, ref=ref), "UNet call, 32 rows at latent 64, t=981")
    # the CFG shared-prefix path (what the samplers run) gives the same bits
    net.cfg_shared_halves = True
    try:
        eps2 = net(x_in, t_in, context=c_in)
    finally:
        net.cfg_shared_halves = False
    assert torch.equal(eps, eps2)


def test_unet_call_64_and_128_rows_run_and_match_32_row_slices():
    """Config 3 at G = 1 puts 64 joint samples on one GPU; bench.py micro-batches them.  Rows are independent, so a
    64- / 128-row call must reproduce the 32-row call on its first rows (grid-limit regressions show up here)."""
    from oracle import unet_oracle as uo
    ldm, cfg, _ = build_ldm(64)
    net = ldm.model.diffusion_model
    inp = uo.synth_inputs(32, 64, seed=1, device="cuda")
    x = torch.cat([inp["x_T"], inp["inpaint_image"], inp["inpaint_mask"]], 1)
    t = torch.full((64,), 481, device="cuda", dtype=torch.long)
    e32 = net(x[:32].contiguous(), t[:32], context=inp["cond"][:32].contiguous())
    e64 = net(x, t, context=inp["cond"])
    d = relerr(e64[:32], e32)
    print("64-row call vs 32-row call on the shared rows: %.3e" % d)
    assert torch.isfinite(e64).all() and d < 2e-3   # split-K / tile schedules may differ with M: fp32 summation order only
    x128, t128, c128 = torch.cat([x, x]), torch.cat([t, t]), torch.cat([inp["cond"], inp["cond"]])
    e128 = net(x128, t128, context=c128)
    assert torch.isfinite(e128).all() and relerr(e128[:32], e32) < 2e-3 and relerr(e128[64:96], e32) < 2e-3


@pytest.mark.parametrize("graph", [False, True])
def test_sampler_single_step_32rows_graph_on_and_off(graph):
    """One DDIM step (S = 1: t = 1) of 8 joint samples with CFG 5 through DDIMSampler, with and without the CUDA graph."""
    from mobi_b200.ddim import DDIMSampler
    from oracle import sampler_oracle as so
    from oracle import unet_oracle as uo
    _no_tf32()
    ldm, cfg, sd = build_ldm(64)
    inp = uo.synth_inputs(8, 64, seed=1, device="cuda")
    smp = DDIMSampler(ldm, use_cuda_graph=graph)
    got, _ = smp.sample(S=1, conditioning=inp["cond"], batch_size=16, shape=[4, 64, 64], verbose=False,
                        unconditional_guidance_scale=5.0, unconditional_conditioning=inp["uc"], eta=0.0, x_T=inp["x_T"],
                        test_model_kwargs=dict(inpaint_image=inp["inpaint_image"], inpaint_mask=inp["inpaint_mask"]))
    ref, trace = so.ddim_sample(oracle_unet(sd, cfg), so.register_schedule(), 1, inp["x_T"], inp["cond"], inp["uc"], 5.0,
                                inp["inpaint_image"], inp["inpaint_mask"])
    e = relerr(got, ref)
    print("one CFG-5 DDIM step, 32 rows, graph=%s: x_prev max-abs-rel %.3e" % (graph, e))
    assert e < 1e-2


# ----------------------------------------------------------------------------------------------- (ii) full sampler runs
def _full_run(sampler_name, n_joint, use_lidar=True):
    from mobi_b200.ddim import DDIMSampler
    from mobi_b200.plms import PLMSSampler
    from oracle import sampler_oracle as so
    from oracle import unet_oracle as uo
    _no_tf32()
    ldm, cfg, sd = build_ldm(64, use_lidar=use_lidar)
    rows = 2 * n_joint if use_lidar else n_joint
    inp = uo.synth_inputs(n_joint, 64, seed=1, device="cuda")
    if not use_lidar:   # pbe.yaml: camera rows only, one context token
        inp = {k: v[::2].contiguous() for k, v in inp.items()}
        inp["cond"], inp["uc"] = inp["cond"][:, :1].contiguous(), inp["uc"][:, :1].contiguous()
    calls = []
    base = oracle_unet(sd, cfg)

    def recording(x, t, c):
        out = base(x, t, c)
        calls.append((x, t, c, out))
        return out

    sched = so.register_schedule()
    if sampler_name == "ddim":
        ref, trace = so.ddim_sample(recording, sched, 50, inp["x_T"], inp["cond"], inp["uc"], 5.0, inp["inpaint_image"],
                                    inp["inpaint_mask"])
        smp = DDIMSampler(ldm)
        got, _ = smp.sample(S=50, conditioning=inp["cond"], batch_size=rows, shape=[4, 64, 64], verbose=False,
                            unconditional_guidance_scale=5.0, unconditional_conditioning=inp["uc"], eta=0.0,
                            x_T=inp["x_T"], test_model_kwargs=dict(inpaint_image=inp["inpaint_image"],
                                                                   inpaint_mask=inp["inpaint_mask"]))
        assert smp.launches == 50 and len(calls) == 50
    else:
        ref, trace = so.plms_sample(recording, sched, 50, inp["x_T"], inp["cond"], inp["uc"], 5.0, inp["inpaint_image"],
                                    inp["inpaint_mask"])
        smp = PLMSSampler(ldm)
        got, _ = smp.sample(S=50, conditioning=inp["cond"], batch_size=rows, shape=[4, 64, 64], verbose=False,
                            unconditional_guidance_scale=5.0, unconditional_conditioning=inp["uc"], eta=0.0,
                            x_T=inp["x_T"], inpaint_image=inp["inpaint_image"], inpaint_mask=inp["inpaint_mask"])
        assert smp.launches == 51 and len(calls) == 51          # S + 1 UNet evaluations (plms.py:223-226)
    # per-step eps, teacher-forced: our UNet on the oracle's own inputs of every step (no drift in the comparison)
    from oracle.precision import bf16_operand_floor
    curve, rms = [], []
    for x, t, c, out in calls:
        eps = ldm.apply_model(x, t, c)
        curve.append(relerr(eps, out))
        rms.append(relrms(eps, out))
    worst = int(np.argmax(curve))
    floors = {i: bf16_operand_floor(sd, cfg, calls[i][0], calls[i][1], calls[i][2], ref=calls[i][3])
              for i in sorted({0, worst, len(calls) - 1})}
    e_final, c_final = relerr(got, ref), cosine(got, ref)
    print("%s 50 steps CFG 5 at latent 64 (%d rows x CFG, lidar=%s): per-step eps max-abs-rel max %.3e (step %d) mean "
          "%.3e, rel-rms max %.3e; final latents max-abs-rel %.3e cosine %.6f"
          % (sampler_name, rows, use_lidar, max(curve), worst, float(np.mean(curve)), max(rms), e_final, c_final))
    print("  bf16-operand floor (max-abs-rel / rel-rms) at steps " +
          ", ".join("%d: %.2e / %.2e" % (i, f[0], f[1]) for i, f in floors.items()))
    print("  eps error curve: " + " ".join("%.1e" % v for v in curve))
    assert max(curve) < TOL_EPS and max(rms) < TOL_EPS_RMS
    for i, f in floors.items():
        assert curve[i] < 1.5 * f[0] and rms[i] < 1.5 * f[1], (i, curve[i], rms[i], f)
    assert e_final < TOL_FINAL and c_final > 0.9995
    return ldm, got, ref


def test_ddim_50_steps_cfg5_latent64_vs_oracle():
    """The north star's parity target: 50-step DDIM, CFG 5, mobi_nusc_512 shape (2 joint samples = 8 UNet rows per call)."""
    _full_run("ddim", 2)


def test_plms_50_steps_cfg5_latent64_vs_oracle():
    """scripts/realism_test_bench.sh's sampler: PLMS, 50 steps = 51 UNet evaluations, CFG 5 (1 joint sample)."""
    _full_run("plms", 1)


def test_pbe_50_steps_and_camera_decode_vs_oracle():
    """BASELINE.json config 4: pbe.yaml camera-only UNet, 50-step DDIM + CFG, then the 512-px first-stage decode."""
    from mobi_b200.autoencoder import AutoencoderKL
    from oracle import unet_oracle as uo
    from oracle import vae_oracle as vo
    ldm, got, ref = _full_run("ddim", 2, use_lidar=False)
    del ldm
    torch.cuda.empty_cache()
    dd = vo.default_ddconfig(False, resolution=512)
    vsd = uo.synth_state_dict(vo.state_dict_shapes(dd), seed=30)
    vae = AutoencoderKL(ddconfig=dd, lossconfig=dict(target="torch.nn.Identity"), embed_dim=4).cuda().eval()
    vae.load_state_dict(vsd, strict=True)
    vsd = {k: v.cuda() for k, v in vsd.items()}
    z = got / 0.18215                                            # decode_first_stage, ddpm.py:846-849
    img = vae.decode(z)
    with torch.no_grad():
        want_same_z = vo.vae_decode(vsd, dd, z)                  # decoder error alone
        want_e2e = vo.vae_decode(vsd, dd, ref / 0.18215)         # oracle sampler -> oracle decoder
    from oracle.precision import vae_bf16_operand_floor
    floor = vae_bf16_operand_floor(vsd, dd, z, want_same_z)
    e_dec, e_e2e = relerr(img, want_same_z), relerr(img, want_e2e)
    print("pbe: 512-px decode of the sampled latents max-abs-rel %.3e (bf16-operand floor %.3e); sampler+decode end to "
          "end %.3e cosine %.6f" % (e_dec, floor[0], e_e2e, cosine(img, want_e2e)))
    assert img.shape == (2, 3, 512, 512)
    assert e_dec < TOL_VAE and e_dec < 1.5 * floor[0]
    assert e_e2e < 2 * TOL_FINAL and cosine(img, want_e2e) > 0.995


# ----------------------------------------------------------------------------------------------- (iv) VAE at 512 px
@pytest.mark.parametrize("lidar", [False, True])
def test_vae_decode_512px_vs_oracle(lidar):
    """configs/mobi_nusc_512.yaml first_stage / lidar_stage decoders at the resolution bench.py's e2e leg runs: latent
    64 x 64 -> 512 x 512 (two-pass GroupNorm path, T = 4096 mid-block attention)."""
    from mobi_b200.autoencoder import AutoencoderKL
    from oracle import unet_oracle as uo
    from oracle import vae_oracle as vo
    _no_tf32()
    cfg = vo.default_ddconfig(lidar, resolution=512)
    sd = uo.synth_state_dict(vo.state_dict_shapes(cfg), seed=30 + int(lidar))
    vae = AutoencoderKL(ddconfig=cfg, lossconfig=dict(target="torch.nn.Identity"), embed_dim=4).cuda().eval()
    vae.load_state_dict(sd, strict=True)
    sd = {k: v.cuda() for k, v in sd.items()}
    gen = torch.Generator(device="cpu").manual_seed(5)
    z = torch.randn(2, 4, 64, 64, generator=gen).cuda()
    out = vae.decode(z)
    with torch.no_grad():
        ref = vo.vae_decode(sd, cfg, z)
    from oracle.precision import vae_bf16_operand_floor
    floor = vae_bf16_operand_floor(sd, cfg, z, ref)
    e = relerr(out, ref)
    print("VAE decode at 512 px (lidar=%s): max-abs-rel %.3e rel-rms %.3e cosine %.6f | bf16-operand floor %.3e / %.3e"
          % (lidar, e, relrms(out, ref), cosine(out, ref), floor[0], floor[1]))
    assert out.shape == (2, 2 if lidar else 3, 512, 512)
    assert e < TOL_VAE and e < 1.5 * floor[0]

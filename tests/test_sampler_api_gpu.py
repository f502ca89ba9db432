"""API surface of the drop-in samplers beyond the golden runs: guidance off (scale 1 / no
unconditional conditioning), the `rest=` keyword instead of `test_model_kwargs`, callbacks, intermediates, device RNG
for a missing x_T, ONE joint sample (2 UNet rows), a batch larger than the captured graph (re-capture), and the
stateless `p_sample_ddim` step, each against the fp32 oracle on the same inputs (tolerance as tests/test_unet_gpu.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 2.5e-2


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


@pytest.fixture(scope="module")
def setup():
    from oracle import sampler_oracle as so
    from oracle import unet_oracle as uo
    from test_unet_gpu import build_ldm   # tests/ is on sys.path under pytest's rootdir conftest
    cfg = uo.tiny_unet_config()
    ldm, sd = build_ldm(cfg)
    apply_ref = lambda x, t, c: uo.unet_forward(sd, cfg, x, t, c)
    return ldm, cfg, apply_ref, so.register_schedule(), uo


def test_no_guidance_and_rest_keyword(setup):
    from mobi_b200.ddim import DDIMSampler
    from oracle import sampler_oracle as so
    ldm, cfg, apply_ref, sched, uo = setup
    inp = uo.synth_inputs(2, 16, context_dim=cfg["context_dim"], seed=3, device="cuda")
    rest = torch.cat([inp["inpaint_image"], inp["inpaint_mask"]], 1)           # ddim.py:173-176: kwargs['rest']
    calls, imgs = [], []   # (S must divide 1000: like the reference, range(0, 1000, 1000 // S) + 1 may index step 1000)
    smp = DDIMSampler(ldm)
    samples, inter = smp.sample(S=4, conditioning=inp["cond"], batch_size=4, shape=[4, 16, 16], verbose=False,
                                unconditional_guidance_scale=1.0, eta=0.0, x_T=inp["x_T"], rest=rest, log_every_t=1,
                                callback=calls.append, img_callback=lambda p, i: imgs.append((i, tuple(p.shape))))
    with torch.no_grad():
        ref, trace = so.ddim_sample(apply_ref, sched, 4, inp["x_T"], inp["cond"], None, 1.0, inp["inpaint_image"],
                                    inp["inpaint_mask"])
    assert rel(samples, ref) < TOL
    assert calls == [0, 1, 2, 3] and imgs == [(i, (4, 4, 16, 16)) for i in range(4)]
    assert len(inter["x_inter"]) == 5 and len(inter["pred_x0"]) == 5           # start + every step (log_every_t = 1)
    assert rel(inter["pred_x0"][-1], trace[-1][1]) < TOL
    assert smp.launches == 4                                                    # no CFG doubling, one UNet call per step


def test_single_joint_sample_then_larger_batch_recaptures(setup):
    from mobi_b200.ddim import DDIMSampler
    from oracle import sampler_oracle as so
    ldm, cfg, apply_ref, sched, uo = setup
    smp = DDIMSampler(ldm)
    for n_joint in (1, 3):                                                      # 2 rows, then 6 rows with the same sampler
        inp = uo.synth_inputs(n_joint, 16, context_dim=cfg["context_dim"], seed=4 + n_joint, device="cuda")
        kw = dict(test_model_kwargs=dict(inpaint_image=inp["inpaint_image"], inpaint_mask=inp["inpaint_mask"]))
        samples, _ = smp.sample(S=2, conditioning=inp["cond"], batch_size=2 * n_joint, shape=[4, 16, 16], verbose=False,
                                unconditional_guidance_scale=4.0, unconditional_conditioning=inp["uc"], eta=0.0,
                                x_T=inp["x_T"], **kw)
        with torch.no_grad():
            ref, _ = so.ddim_sample(apply_ref, sched, 2, inp["x_T"], inp["cond"], inp["uc"], 4.0, inp["inpaint_image"],
                                    inp["inpaint_mask"])
        assert samples.shape == (2 * n_joint, 4, 16, 16) and rel(samples, ref) < TOL


def test_missing_x_T_uses_device_rng_and_eta_noise_path_runs(setup):
    from mobi_b200.ddim import DDIMSampler
    ldm, cfg, apply_ref, sched, uo = setup
    inp = uo.synth_inputs(1, 16, context_dim=cfg["context_dim"], seed=9, device="cuda")
    kw = dict(test_model_kwargs=dict(inpaint_image=inp["inpaint_image"], inpaint_mask=inp["inpaint_mask"]))
    outs = []
    for _ in range(2):
        torch.manual_seed(1234)                                                 # x_T and the eta noise come from torch's RNG
        s, _ = DDIMSampler(ldm).sample(S=4, conditioning=inp["cond"], batch_size=2, shape=[4, 16, 16], verbose=False,
                                       unconditional_guidance_scale=2.0, unconditional_conditioning=inp["uc"], eta=0.5, **kw)
        outs.append(s)
    assert torch.isfinite(outs[0]).all() and torch.equal(outs[0], outs[1])      # deterministic under a fixed seed
    s0, _ = DDIMSampler(ldm).sample(S=4, conditioning=inp["cond"], batch_size=2, shape=[4, 16, 16], verbose=False,
                                    unconditional_guidance_scale=2.0, unconditional_conditioning=inp["uc"], eta=0.0,
                                    x_T=inp["x_T"], **kw)
    assert not torch.equal(s0, outs[0])


def test_p_sample_ddim_single_step(setup):
    from mobi_b200.ddim import DDIMSampler
    from oracle import sampler_oracle as so
    ldm, cfg, apply_ref, sched, uo = setup
    inp = uo.synth_inputs(2, 16, context_dim=cfg["context_dim"], seed=5, device="cuda")
    smp = DDIMSampler(ldm)
    smp.make_schedule(ddim_num_steps=10, ddim_eta=0.0, verbose=False)
    index = 9
    t = torch.full((4,), int(smp.ddim_timesteps[index]), device="cuda", dtype=torch.long)
    x_prev, pred = smp.p_sample_ddim(inp["x_T"], inp["cond"], t, index, unconditional_guidance_scale=3.0,
                                     unconditional_conditioning=inp["uc"],
                                     test_model_kwargs=dict(inpaint_image=inp["inpaint_image"], inpaint_mask=inp["inpaint_mask"]))
    with torch.no_grad():
        _, trace = so.ddim_sample(apply_ref, sched, 10, inp["x_T"], inp["cond"], inp["uc"], 3.0, inp["inpaint_image"],
                                  inp["inpaint_mask"], steps_to_run=1)
    assert rel(x_prev, trace[0][0]) < TOL and rel(pred, trace[0][1]) < TOL


def test_wrong_device_and_missing_kwargs_fail_loudly(setup):
    from mobi_b200.ddim import DDIMSampler
    ldm, cfg, apply_ref, sched, uo = setup
    inp = uo.synth_inputs(1, 16, context_dim=cfg["context_dim"], seed=6, device="cuda")
    with pytest.raises(Exception, match="test_model_kwargs|rest"):            # ddim.py:176
        DDIMSampler(ldm).sample(S=2, conditioning=inp["cond"], batch_size=2, shape=[4, 16, 16], verbose=False,
                                x_T=inp["x_T"])
    with pytest.raises(RuntimeError):
        ldm.model.diffusion_model(torch.zeros(2, 9, 16, 16), torch.zeros(2, dtype=torch.long), context=torch.zeros(2, 2, 32))


def test_cfg_shared_prefix_is_exact(setup):
    """Under classifier-free guidance the two halves of the UNet batch are identical copies up to the first context
    injection; computing that prefix once (UNetModel.cfg_shared_halves) must not change a single bit."""
    ldm, cfg, apply_ref, sched, uo = setup
    inp = uo.synth_inputs(2, 16, context_dim=cfg["context_dim"], seed=8, device="cuda")
    x = torch.cat([inp["x_T"], inp["inpaint_image"], inp["inpaint_mask"]], 1)
    x_in, t_in = torch.cat([x, x]), torch.full((8,), 481, device="cuda", dtype=torch.long)
    c_in = torch.cat([inp["uc"], inp["cond"]]).contiguous()
    unet = ldm.model.diffusion_model
    ref = unet(x_in, t_in, context=c_in)
    unet.cfg_shared_halves = True
    try:
        got = unet(x_in, t_in, context=c_in)
    finally:
        unet.cfg_shared_halves = False
    assert torch.equal(got, ref)
    assert not torch.equal(got[:4], got[4:])      # the halves do differ after the context enters


def test_new_conditioning_at_a_recycled_address_is_not_served_from_a_cache(setup):
    """forward() used to key its context tables on (data_ptr, _version, shape).  mobi_b200 ops write
    through raw pointers (no version bump) and the caching allocator recycles addresses, so new conditioning could be
    served the OLD tables.  Two different contexts through the SAME storage, written by an op kernel, must differ."""
    from mobi_b200 import ops
    ldm, cfg, apply_ref, sched, uo = setup
    unet = ldm.model.diffusion_model
    inp = uo.synth_inputs(2, 16, context_dim=cfg["context_dim"], seed=11, device="cuda")
    x = torch.cat([inp["x_T"], inp["inpaint_image"], inp["inpaint_mask"]], 1)
    t = torch.full((4,), 500, device="cuda", dtype=torch.long)
    c = inp["cond"].clone()
    version = c._version
    e1 = unet(x, t, context=c).clone()
    ops.scale_f32(inp["uc"].contiguous(), 1.0, out=c)          # new contents, same pointer, same version counter
    assert c._version == version
    e2 = unet(x, t, context=c)
    with torch.no_grad():
        r1, r2 = apply_ref(x, t, inp["cond"]), apply_ref(x, t, inp["uc"])
    assert rel(e1, r1) < TOL and rel(e2, r2) < TOL and rel(e2, r1) > 4 * TOL
    # del + realloc at the same address
    ptr = c.data_ptr()
    del c
    c3 = inp["cond"].clone()
    e3 = unet(x, t, context=c3)
    print("recycled address: %s" % (c3.data_ptr() == ptr))
    assert rel(e3, r1) < TOL


def test_logged_intermediates_are_snapshots_on_the_blend_path(setup):
    """with mask / x0 the step kernel blends the known latent into the running sample in place; the tensors
    already appended to intermediates['x_inter'] must not change afterwards (the reference builds a new tensor per step)."""
    from mobi_b200.ddim import DDIMSampler
    ldm, cfg, apply_ref, sched, uo = setup
    inp = uo.synth_inputs(2, 16, context_dim=cfg["context_dim"], seed=21, device="cuda")
    mask = (torch.rand(4, 1, 16, 16, device="cuda") > 0.5).float()
    x0 = torch.randn(4, 4, 16, 16, device="cuda")
    seen = []
    smp = DDIMSampler(ldm)
    _, inter = smp.sample(S=4, conditioning=inp["cond"], batch_size=4, shape=[4, 16, 16], verbose=False,
                          unconditional_guidance_scale=3.0, unconditional_conditioning=inp["uc"], eta=0.0, x_T=inp["x_T"],
                          mask=mask, x0=x0, log_every_t=1,
                          img_callback=lambda p, i: seen.append(None),
                          test_model_kwargs=dict(inpaint_image=inp["inpaint_image"], inpaint_mask=inp["inpaint_mask"]))
    assert len(inter["x_inter"]) == 5 and len(seen) == 4
    assert torch.equal(inter["x_inter"][0], inp["x_T"])                       # the start noise was not blended over
    assert all(not torch.equal(a, b) for a, b in zip(inter["x_inter"][:-1], inter["x_inter"][1:]))

"""Parity of the CUDA first-stage / range autoencoder (AutoencoderKL.decode / encode through the drop-in classes and the
C ABI) against (a) golden vectors produced by the unmodified reference (tests/golden/vae_tiny.npz) and (b) the fp32
oracle on the GPU at the real model width.  Tolerance: max-abs-rel <= 2e-2 (bf16 operands, fp32 accumulation) and, at
full width, <= 1.5 x the bf16-operand floor of the same decode (oracle/precision.py: the fp32 oracle with only the conv /
attention operands rounded to bf16 is itself ~1.2-1.5e-2 away from the fp32 oracle over the decoder's ~30 convolutions)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-2


def relerr(a, b):
    a, b = a.double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


def build_vae(cfg, seed):
    from mobi_b200.autoencoder import AutoencoderKL
    from oracle import unet_oracle as uo
    from oracle import vae_oracle as vo
    sd = uo.synth_state_dict(vo.state_dict_shapes(cfg), seed=seed)
    vae = AutoencoderKL(ddconfig=cfg, lossconfig=dict(target="torch.nn.Identity"), embed_dim=4).cuda().eval()
    vae.load_state_dict(sd, strict=True)
    return vae, {k: v.cuda() for k, v in sd.items()}


def test_vae_tiny_vs_reference_golden():
    from mobi_b200.ddpm import LatentDiffusion
    from oracle import vae_oracle as vo
    g = np.load(os.path.join(GOLDEN, "vae_tiny.npz"))
    cam, _ = build_vae(vo.tiny_ddconfig(False), 10)
    lid, _ = build_vae(vo.tiny_ddconfig(True), 20)
    z = torch.from_numpy(g["z"]).cuda()
    h_cam, h_lid = z[::2], z[1::2]                       # decode_sample, ddpm.py:1420-1447
    img = cam.decode((1.0 / 0.18215 * h_cam)[:, :4])     # decode_first_stage, ddpm.py:846-849, 893-899
    rng = lid.decode((1.0 / 0.18215 * h_lid)[:, :4])
    torch.cuda.synchronize()
    e_img, e_rng = relerr(img, g["image"]), relerr(rng, g["range"])
    print("tiny vae decode max-abs-rel vs reference golden: camera %.3e lidar %.3e" % (e_img, e_rng))
    assert e_img < 2e-2 and e_rng < 2e-2   # 64-channel model: bf16 operand rounding averages over fewer terms
    m_cam = cam.encode(torch.from_numpy(g["cam_in"]).cuda()).parameters
    m_lid = lid.encode(torch.from_numpy(g["lid_in"]).cuda()).parameters
    e_mc, e_ml = relerr(m_cam, g["cam_moments"]), relerr(m_lid, g["lid_moments"])
    print("tiny vae encode moments max-abs-rel vs reference golden: camera %.3e lidar %.3e" % (e_mc, e_ml))
    assert e_mc < 2e-2 and e_ml < 2e-2


@pytest.mark.parametrize("lidar", [False, True])
def test_vae_fullwidth_decode_vs_oracle(lidar):
    """configs/mobi_nusc_256.yaml first_stage / lidar_stage ddconfig (ch 128, mult 1-2-4-4) at 256 px: latent 32x32."""
    from oracle import vae_oracle as vo
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = vo.default_ddconfig(lidar, resolution=256)
    vae, sd = build_vae(cfg, 30 + int(lidar))
    gen = torch.Generator(device="cpu").manual_seed(5)
    z = torch.randn(2, 4, 32, 32, generator=gen).cuda()
    out = vae.decode(z)
    with torch.no_grad():
        ref = vo.vae_decode(sd, cfg, z)
    torch.cuda.synchronize()
    from oracle.precision import vae_bf16_operand_floor
    floor = vae_bf16_operand_floor(sd, cfg, z, ref)
    e = relerr(out, ref)
    cos = torch.nn.functional.cosine_similarity(out.flatten().double(), ref.flatten().double(), dim=0).item()
    print("full-width vae decode (lidar=%s) max-abs-rel %.3e cosine %.6f | bf16-operand floor %.3e" % (lidar, e, cos, floor[0]))
    assert out.shape == (2, 2 if lidar else 3, 256, 256)
    assert e < TOL and e < 1.5 * floor[0]


def test_vae_fullwidth_encode_vs_oracle():
    from oracle import vae_oracle as vo
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = vo.default_ddconfig(True, resolution=256)
    vae, sd = build_vae(cfg, 41)
    gen = torch.Generator(device="cpu").manual_seed(6)
    x = torch.randn(2, 2, 256, 256, generator=gen).cuda()
    post = vae.encode(x)
    with torch.no_grad():
        ref = vo.vae_encode_moments(sd, cfg, x)
    e = relerr(post.parameters, ref)
    print("full-width lidar vae encode moments max-abs-rel %.3e" % e)
    assert post.mode().shape == (2, 4, 32, 32)
    assert e < TOL

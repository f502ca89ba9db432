"""Parity of the conditioning encoders (mapper Transformer + final_ln + proj_out, BBoxEmbedder)
against tests/golden/cond_full.npz, produced at FULL size by the unmodified reference modules
(ldm/modules/encoders/modules.py, xf.py) in oracle/make_golden.py.  bf16 GEMM operands, fp32 accumulation and residual
stream: tolerance 1e-2 of the output's max, like every other tensor-core path."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class _StubClip(torch.nn.Module):
    """Stands for the frozen CLIP vision tower: returns the pooled features it is given as `pixel_values`."""

    def forward(self, pixel_values):
        class Out:
            pooler_output = pixel_values
        return Out()


def test_learned_conditioning_vs_reference_golden():
    from mobi_b200 import encoders
    from oracle import cond_oracle as co
    from oracle import unet_oracle as uo
    g = np.load(os.path.join(GOLDEN, "cond_full.npz"))
    sd = uo.synth_state_dict(co.shapes(), seed=30)
    model = encoders.FrozenCLIPImageEmbedder(["ref_image", "ref_bbox"], transformer=_StubClip())
    proj_out = torch.nn.Linear(1024, 768)
    pre = "cond_stage_model."
    missing = model.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, strict=True)
    proj_out.load_state_dict({"weight": sd["proj_out.weight"], "bias": sd["proj_out.bias"]})
    model, proj_out = model.cuda(), proj_out.cuda()
    cond = dict(ref_image=torch.from_numpy(g["pooled"]).cuda(), ref_bbox=torch.from_numpy(g["bbox"]).cuda())
    out = encoders.learned_conditioning(model, proj_out, cond)
    torch.cuda.synchronize()
    ref = torch.from_numpy(g["cond"]).cuda()
    assert out.shape == ref.shape == (6, 2, 768)
    e_img = ((out[:, 0] - ref[:, 0]).abs().max() / ref[:, 0].abs().max()).item()
    e_box = ((out[:, 1] - ref[:, 1]).abs().max() / ref[:, 1].abs().max()).item()
    print("conditioning tokens max-abs-rel vs reference golden: image %.3e bbox %.3e" % (e_img, e_box))
    assert e_img < 1e-2 and e_box < 1e-2
    # module-level signatures of the reference
    tok = model.bbox_embedder(cond["ref_bbox"])
    assert tok.shape == (6, 1, 768)
    x = model.mapper(cond["ref_image"].unsqueeze(1))
    assert x.shape == (6, 1, 1024)


def test_fourier_embed_matches_reference_order():
    from mobi_b200 import ops
    from oracle import cond_oracle as co
    x = torch.rand(5, 8, 3, device="cuda") * 2 - 1
    got = ops.fourier_embed(x, 4).reshape(5, 8, 27).float()
    ref = co.fourier_embed(x.cpu(), 4).cuda()
    assert (got - ref).abs().max().item() < 8e-3   # bf16 output

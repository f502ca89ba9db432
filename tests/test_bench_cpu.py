"""Host-side logic of the measurement and of the test-bench caller, on CPU: how bench.py cuts BASELINE.json config 3 (64
joint samples) over N GPUs and into micro-batches, the layout of the synthetic dataset batch, and the nested-batch helpers of
mobi_b200.pipeline.  No compute call is made (the product path has no CPU fallback)."""
import importlib.util
import os
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _args(**kw):
    base = dict(total_samples=64, samples_per_gpu=0, micro_batch=32, pbe=False, latent=64, ddim_steps=50, gpus=1)
    base.update(kw)
    return types.SimpleNamespace(**base)


def test_config3_is_cut_as_written():
    b = _bench()
    # 64 joint samples in total, 64 / N per GPU, strong scaling; micro-batches never exceed the shard and divide it
    assert b.job_shape(_args(), 1) == (64, 64, 32, "strong")
    assert b.job_shape(_args(), 2) == (32, 64, 32, "strong")
    assert b.job_shape(_args(), 4) == (16, 64, 16, "strong")
    assert b.job_shape(_args(), 8) == (8, 64, 8, "strong")
    assert b.job_shape(_args(micro_batch=24), 1) == (64, 64, 16, "strong")     # largest divisor of the shard <= 24
    # the round-1 weak-scaling mode stays available
    assert b.job_shape(_args(samples_per_gpu=8), 8) == (8, 64, 8, "weak")
    with pytest.raises(SystemExit):
        b.job_shape(_args(total_samples=10), 4)
    name = b.workload_name(_args(), 8)
    assert "batch 64 joint samples over 8 GPU(s) = 8/GPU" in name and "mobi_nusc_512" in name
    assert "pbe.yaml" in b.workload_name(_args(pbe=True), 1)


def test_reference_flop_table_matches_the_survey():
    b = _bench()
    assert b.UNET_FLOPS_PER_JOINT[64] == 2043895808000 and b.UNET_FLOPS_PER_JOINT[32] == 419359047680
    assert b.UNET_FLOPS_PBE_ROW[64] == 835382476800


def test_synthetic_dataset_batch_has_the_reference_layout():
    from mobi_b200 import pipeline, synth
    batch = synth.synthetic_dataset_batch(3, px=64, seed=1, pin=False)
    img, lid = batch["image"], batch["lidar"]
    assert img["GT"].shape == (3, 3, 64, 64) and img["inpaint_mask"].shape == (3, 1, 64, 64)
    assert lid["range_data"].shape == (3, 2, 64, 64) and lid["range_depth_orig"].shape == (3, 32, 1096)
    assert img["cond"]["ref_bbox"].shape == lid["cond"]["ref_bbox"].shape == (3, 8, 3) and batch["bbox_3d"].shape == (3, 8, 3)
    assert torch.equal(img["inpaint_image"], img["GT"] * img["inpaint_mask"])          # the hole is blanked (1 = keep)
    hole = 1.0 - img["inpaint_mask"].mean().item()
    assert 0.15 < hole < 0.25                                                           # object_area_crop: 0.2
    # nested helpers used by the e2e leg and the CLI
    part = pipeline.batch_slice(batch, 1, 3)
    assert part["image"]["GT"].shape[0] == 2 and part["lidar"]["cond"]["ref_image"].shape == (2, 1024)
    assert pipeline.batch_bytes(batch) == sum(
        t.numel() * t.element_size() for t in list(img.values())[:3] + list(img["cond"].values()) + [batch["bbox_3d"]] +
        [v for v in lid.values() if torch.is_tensor(v)] + list(lid["cond"].values()))
    moved = pipeline.batch_to_device(part, "cpu")
    assert moved["bbox_3d"].shape == (2, 8, 3)


def test_unconditional_conditioning_follows_the_reference_script():
    """inference_test_bench.py:423-428: [learnable_vector, bbox_uncond_vector] per row, the second only with ref_bbox."""
    from mobi_b200 import pipeline
    m = types.SimpleNamespace(learnable_vector=torch.full((1, 1, 4), 1.0), bbox_uncond_vector=torch.full((1, 1, 4), 2.0),
                              cond_stage_key=["ref_image", "ref_bbox"])
    uc = pipeline.unconditional_conditioning(m, 6)
    assert uc.shape == (6, 2, 4) and uc[:, 0].eq(1).all() and uc[:, 1].eq(2).all()
    m.cond_stage_key = ["ref_image"]
    assert pipeline.unconditional_conditioning(m, 6).shape == (6, 1, 4)

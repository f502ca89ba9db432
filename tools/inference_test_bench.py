#!/usr/bin/env python
"""Counterpart of the reference's scripts/inference_test_bench.py for the drop-in package: same model-side flags
(--config, --ckpt, --plms, --ddim_steps, --ddim_eta, --scale, --n_samples, --seed, --fixed_code, --skip_save, --outdir),
the reference YAML schema (configs/mobi_nusc_512.yaml) and the reference checkpoint format.  The dataset side is out of
scope (no nuScenes here): batches come from --synthetic N (mobi_b200.synth.synthetic_dataset_batch), a stand-in with the
dataset's exact layout, and the CLIP tower is entered at its pooler_output.

  python tools/inference_test_bench.py --config configs/mobi_nusc_512.yaml --ckpt model.ckpt --synthetic 8 --n_batches 2
  python tools/inference_test_bench.py --config configs/mobi_nusc_512.yaml --random-init --synthetic 8   # no checkpoint

Per batch (mobi_b200.pipeline.inpaint_batch = inference_test_bench.py:403-464 + 567-629): host -> device, get_input,
conditioning, 50-step sampling with CFG, decode_sample, both first-stage decodes, range-view post-processing, device ->
host of what the reference writes to disk (decoded camera patch, edited sweep, edited point cloud).  Unless --skip_save,
the arrays are written as .npy under --outdir."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=os.path.join(ROOT, "configs", "mobi_nusc_512.yaml"))
    ap.add_argument("--ckpt", default=None, help="reference-format checkpoint ({'state_dict': ...})")
    ap.add_argument("--random-init", action="store_true", help="synthetic weights instead of a checkpoint")
    ap.add_argument("--plms", action="store_true")
    ap.add_argument("--ddim_steps", type=int, default=50)
    ap.add_argument("--ddim_eta", type=float, default=0.0)
    ap.add_argument("--scale", type=float, default=5.0)
    ap.add_argument("--n_samples", type=int, default=8, help="batch size")
    ap.add_argument("--synthetic", type=int, default=8, help="joint samples in the synthetic stand-in dataset")
    ap.add_argument("--n_batches", type=int, default=0, help="0 = the whole synthetic dataset")
    ap.add_argument("--seed", type=int, default=321)
    ap.add_argument("--fixed_code", action="store_true")
    ap.add_argument("--skip_save", action="store_true", help="for speed measurements (as in the reference)")
    ap.add_argument("--outdir", default=os.path.join(ROOT, "gpurun_out", "test_bench"))
    ap.add_argument("opts", nargs="*", help="key=value overrides of top-level config variables (dot-list style)")
    return ap.parse_args()


def main():
    opt = parse()
    from mobi_b200 import config, pipeline, synth
    if not torch.cuda.is_available():
        raise SystemExit("inference_test_bench: needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", 0)
    torch.manual_seed(opt.seed)
    cfg = config.load_config(opt.config)
    if opt.opts:   # use_lidar=True ref_mode=... style overrides: re-resolve the document with them applied
        import yaml
        raw = yaml.safe_load(open(opt.config))
        for kv in opt.opts:
            k, v = kv.split("=", 1)
            raw[k] = yaml.safe_load(v)
        cfg = config._resolve(raw, raw)
    tower = synth.PooledFeatureTower()
    if opt.ckpt:
        model = config.build_model(cfg, ckpt=opt.ckpt, device=dev, clip_tower=tower, meta_init=True, verbose=True)
    else:
        if not opt.random_init:
            raise SystemExit("inference_test_bench: give --ckpt or --random-init")
        model = config.build_model(cfg, device=dev, clip_tower=tower)
        for i, m in enumerate((model.model.diffusion_model, model.first_stage_model, model.lidar_stage_model,
                               model.cond_stage_model, model.proj_out)):
            synth.init_synthetic_(m, seed=i)
    sampler = pipeline.make_sampler(model, plms=opt.plms)
    px = int(cfg.get("image_height", 8 * model.image_size))
    data = synth.synthetic_dataset_batch(opt.synthetic, px=px, seed=opt.seed)
    bs = opt.n_samples
    n_batches = (opt.synthetic + bs - 1) // bs if opt.n_batches <= 0 else opt.n_batches
    os.makedirs(opt.outdir, exist_ok=True)
    start_code = None
    if opt.fixed_code:
        start_code = torch.randn([2 * bs, model.channels, model.image_size, model.image_size], device=dev)
    times = []
    for b in range(n_batches):
        lo = (b * bs) % opt.synthetic
        host = pipeline.batch_slice(data, lo, min(lo + bs, opt.synthetic))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        batch = pipeline.batch_to_device(host, dev)
        n = batch["bbox_3d"].shape[0]
        out = pipeline.inpaint_batch(model, sampler, batch, ddim_steps=opt.ddim_steps, scale=opt.scale,
                                     ddim_eta=opt.ddim_eta, start_code=None if start_code is None else start_code[:2 * n])
        res = {k: out[k].cpu() for k in ("image_sample", "range_pred", "pred_instance_mask", "n_points")}
        pts = out["pred_points"].cpu()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        if not opt.skip_save:
            np.save(os.path.join(opt.outdir, "batch%03d_image_sample.npy" % b), res["image_sample"].numpy())
            np.save(os.path.join(opt.outdir, "batch%03d_range_pred.npy" % b), res["range_pred"].numpy())
            np.save(os.path.join(opt.outdir, "batch%03d_pred_points.npy" % b), pts.numpy())
        print("batch %d: %d joint samples in %.2f s (%.2f samples/s), %d edited points/sample"
              % (b, n, times[-1], n / times[-1], int(res["n_points"].float().mean())), flush=True)
    steady = times[1:] or times
    print(json.dumps({"batches": n_batches, "batch_size": bs, "sampler": "plms" if opt.plms else "ddim",
                      "ddim_steps": opt.ddim_steps, "scale": opt.scale,
                      "samples_per_s_steady": bs / (sum(steady) / len(steady)), "first_batch_s": times[0]}))


if __name__ == "__main__":
    main()

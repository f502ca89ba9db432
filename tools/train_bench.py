#!/usr/bin/env python
"""Training-step timing (BASELINE.json config 5): mobi_nusc_512 UNet forward + backward + gradient all-reduce + AdamW on
synthetic latents, one process per GPU (torchrun for N > 1, NCCL).  Prints one JSON line with the step time, the
per-kernel-class breakdown of one step and the achieved tensor-core rate against SURVEY.md §8(d)'s FLOP counts.

  python tools/train_bench.py [--samples-per-gpu 2] [--latent 64] [--steps 5] [--warmup 2]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FWD_FLOPS_PER_JOINT = {64: 2043895808000, 32: 419359047680}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples-per-gpu", type=int, default=2)   # configs/mobi_nusc_512.yaml:11 batch_size
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="one blocking all-reduce after the backward (round-1 behaviour)")
    args = ap.parse_args()
    import torch.distributed as dist
    from mobi_b200 import ops, synth
    from mobi_b200.training import UNetTrainer
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.samples_per_gpu
    ldm = synth.build_synthetic_ldm(latent=args.latent, device=dev, seed=0)
    tr = UNetTrainer(ldm, use_cuda_graph=not args.no_graph, overlap_allreduce=not args.no_overlap)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    R, h = 2 * n, args.latent
    x_start = torch.randn(R, 9, h, h, device=dev, generator=g)
    noise = torch.randn(R, 4, h, h, device=dev, generator=g)
    cond = torch.randn(R, 2, 768, device=dev, generator=g)
    t = torch.randint(0, 1000, (R,), device=dev, generator=g)

    def step():
        loss = tr.forward_backward(x_start, t, noise, cond)
        tr.step()
        return loss

    for _ in range(args.warmup):
        loss = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ops.Stats.launches
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    launches = (ops.Stats.launches - l0) // args.steps
    p_l2, p_sum = tr.flat.params.double().norm().item(), tr.flat.params.double().sum().item()   # same on every rank after DDP steps
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    tr.use_cuda_graph = False          # per-kernel CUDA events need eager launches
    step()
    ops.Stats.begin_profile()
    step()
    prof = ops.Stats.end_profile()
    if rank == 0:
        fwd = FWD_FLOPS_PER_JOINT[args.latent] * n
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        tc = {k: v for k, v in prof.items() if k.split(" ")[0] in ("gemm", "conv", "attention")}
        tc_ms, tc_fl = sum(v["ms"] for v in tc.values()), sum(v["flops"] for v in tc.values())
        line = {"metric": "training step (UNet fwd+bwd, adapter grads, all-reduce, AdamW)", "ms_per_step": ms.item(),
                "samples_per_s": n * world / (ms.item() / 1e3), "n_gpus": world, "joint_samples_per_gpu": n,
                "latent": args.latent, "cuda_graph": not args.no_graph, "overlap_allreduce": not args.no_overlap,
                "loss": loss.item(), "trainable_params": tr.flat.numel, "params_l2_after": p_l2, "params_sum_after": p_sum,
                "kernel_launches_per_step": launches,
                "nominal_tflops_3x_forward": 3 * fwd / (ms.item() / 1e3) / 1e12,
                "tensor_core_launch_tflops": tc_fl / (tc_ms / 1e3) / 1e12 if tc_ms else None, "peak_tflops": peak,
                "by_kernel_ms": {k: round(v["ms"], 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
                "by_kernel_launches": {k: v["launches"] for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}}
        if os.environ.get("MOBI_GEMM_SHAPES", "0") == "1":   # per-shape table on stderr: kind, launches, ms, TFLOP/s
            for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:45]:
                tf = v["flops"] / (v["ms"] / 1e3) / 1e12 if v["ms"] and v["flops"] else 0.0
                print("%-44s %5d %8.3f ms %8.1f TF/s" % (k, v["launches"], v["ms"], tf), file=sys.stderr)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

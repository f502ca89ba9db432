#!/usr/bin/env python
"""bench.py — MObI joint camera+lidar inpainting throughput on B200 (contract: see the task brief).

Workload = BASELINE.json config 3 AS WRITTEN: mobi_nusc_512.yaml, 50-step DDIM joint camera+lidar inpainting with
classifier-free guidance 5, batch 64 joint samples sharded over the N GPUs of one box (64 / N per GPU: 64 on one GPU,
8 each on eight), i.e. STRONG scaling of a fixed job.  One "step" = one full 50-step DDIM sampling pass over the rank's
shard (latent 64x64; 2 UNet rows per joint sample, 4 with CFG), run in micro-batches of --micro-batch joint samples
(default 32 = 128 UNet rows per call).  No collective on the data path.  The weak-scaling figure of round 1 (8 joint
samples per GPU whatever N) is kept as the secondary key `weak_8_per_gpu`.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--total-samples 64] [--micro-batch 32] [--latent 64]
  python bench.py --samples-per-gpu 8 ...  # fixed per-GPU shard instead of total / N (weak scaling)
  python bench.py --impl reference ...     # the reference's own CPU path (unmodified reference modules where
                                           # /root/reference exists, else the oracle port) on the host cores
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# SURVEY.md §8(d) / BASELINE.md §2: algorithmic FLOPs of the REFERENCE forward per joint sample (2 rows), no CFG
UNET_FLOPS_PER_JOINT = {64: 2043895808000, 32: 419359047680}
UNET_FLOPS_PBE_ROW = {64: 835382476800}   # camera-only UNet (pbe.yaml), per row at latent 64
CFG_SCALE = 5.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--total-samples", type=int, default=64, help="joint samples of the whole job (config 3: 64)")
    ap.add_argument("--samples-per-gpu", type=int, default=0,
                    help="fixed per-GPU shard (weak scaling) instead of total-samples / gpus")
    ap.add_argument("--micro-batch", type=int, default=32, help="joint samples per sampler call on one GPU")
    ap.add_argument("--budget-s", type=float, default=760.0,
                    help="wall-clock guard: the driver kills a run at 870 s; if W + K full steps would not fit, fewer "
                         "timed steps are run and the line says so (steps_requested)")
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--pbe", action="store_true",
                    help="BASELINE.json config 4: camera-only Paint-by-Example UNet (pbe.yaml), 1 row per sample, 1 context "
                         "token, camera VAE decode only (implies --no-train: that model has no adapters to train)")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step (config 5) measurement")
    ap.add_argument("--train-samples-per-gpu", type=int, default=2)  # configs/mobi_nusc_512.yaml:11 batch_size
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def job_shape(args, world):
    """(joint samples per GPU, total joint samples, micro-batch, scaling) of this run."""
    if args.samples_per_gpu > 0:
        n, total, scaling = args.samples_per_gpu, args.samples_per_gpu * world, "weak"
    else:
        if args.total_samples % world:
            raise SystemExit("bench.py: --total-samples %d is not divisible by %d GPUs" % (args.total_samples, world))
        n, total, scaling = args.total_samples // world, args.total_samples, "strong"
    mb = max(1, min(args.micro_batch, n))
    while n % mb:
        mb -= 1
    return n, total, mb, scaling


def workload_name(args, world=None):
    world = world or args.gpus
    n, total, mb, scaling = job_shape(args, world)
    if args.pbe:
        return "pbe.yaml %d-step DDIM + CFG %.1f camera-only inpainting at latent %d, batch %d over %d GPU(s) = %d " \
               "samples/GPU in micro-batches of %d" % (args.ddim_steps, CFG_SCALE, args.latent, total, world, n, mb)
    return ("mobi_nusc_512" if args.latent == 64 else "mobi_nusc_256") + \
        " %d-step DDIM + CFG %.1f joint camera+lidar inpainting, batch %d joint samples over %d GPU(s) = %d/GPU in " \
        "micro-batches of %d (%d UNet rows per call)" % (args.ddim_steps, CFG_SCALE, total, world, n, mb, 4 * mb)


# ---------------------------------------------------------------------------------------------- reference arm
class CpuReference:
    """The reference's CPU implementation of the path, fp32, all host cores: ONE DDIM step with CFG (one p_sample_ddim
    call, ddim.py:165-213 = the config-1 unit of BASELINE.json) of `n_joint` joint samples at the workload's latent size.

    kind "reference": the UNMODIFIED reference modules (LatentDiffusion.apply_model + DDIMSampler.p_sample_ddim) imported
    from MOBI_REFERENCE_ROOT through oracle/ref_shims.py, where that tree exists (the build container);
    kind "port": oracle/unet_oracle.py + sampler_oracle.py (the GPU box has no reference tree).
    A whole 50-step sample of one joint sample takes ~6 minutes of CPU time, so the bounded sample is one step and the
    samples/s figure is that step x 50 ("extrapolated": true); `seconds` are the measured ones."""

    def __init__(self, latent, threads, n_joint=1, state_dict=None):
        import torch
        from oracle import ref_shims
        from oracle import sampler_oracle as so
        from oracle import unet_oracle as uo
        torch.set_num_threads(threads)
        self.threads, self.latent, self.n_joint = threads, latent, n_joint
        self.cfg = uo.default_unet_config(image_size=latent)
        self.sd = state_dict if state_dict is not None else uo.synth_state_dict(uo.state_dict_shapes(self.cfg), seed=0)
        self.inp = uo.synth_inputs(n_joint, latent, seed=1)
        self.kind = "port"
        self.so, self.uo = so, uo
        self.sched = so.register_schedule()
        if ref_shims.reference_available():
            try:
                ref_shims.install()
                from oracle.make_golden import build_reference_ldm
                from ldm.models.diffusion.ddim import DDIMSampler
                model = build_reference_ldm(dict(self.cfg))
                model.model.diffusion_model.load_state_dict(self.sd, strict=True)
                self.sampler = DDIMSampler(model)
                self.sampler.make_schedule(ddim_num_steps=50, ddim_eta=0.0, verbose=False)
                self.kind = "reference"
            except Exception as exc:   # an incomplete tree: fall back to the port and say why on stderr
                print("bench.py: reference tree present but not usable (%s: %s); timing the oracle port"
                      % (type(exc).__name__, str(exc)[:200]), file=sys.stderr)

    def step(self):
        """Returns (seconds, x_prev) of one DDIM + CFG step at index 49 (t = 981) from x_T."""
        import torch
        inp = self.inp
        t0 = time.perf_counter()
        with torch.no_grad():
            if self.kind == "reference":
                ts = torch.full((inp["x_T"].shape[0],), 981, dtype=torch.long)
                x_prev, _ = self.sampler.p_sample_ddim(
                    inp["x_T"], inp["cond"], ts, index=49, unconditional_guidance_scale=CFG_SCALE,
                    unconditional_conditioning=inp["uc"],
                    test_model_kwargs=dict(inpaint_image=inp["inpaint_image"], inpaint_mask=inp["inpaint_mask"]))
            else:
                x_prev, _ = self.so.ddim_sample(lambda x, t, c: self.uo.unet_forward(self.sd, self.cfg, x, t, c),
                                                self.sched, 50, inp["x_T"], inp["cond"], inp["uc"], CFG_SCALE,
                                                inp["inpaint_image"], inp["inpaint_mask"], steps_to_run=1)
        return time.perf_counter() - t0, x_prev

    def describe(self, seconds, ddim_steps):
        return {"value": 1.0 * self.n_joint / (ddim_steps * seconds), "unit": "samples/s", "cores": self.threads,
                "kind": self.kind, "extrapolated": True, "measured_seconds": seconds,
                "sample": "ONE DDIM step (one CFG UNet evaluation, %d rows) of %d joint sample(s) at latent %d, fp32, "
                          "%.1f s measured; samples/s = 1 / (%d steps x that)"
                          % (4 * self.n_joint, self.n_joint, self.latent, seconds, ddim_steps)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ref = CpuReference(args.latent, threads, 1)
    # bounded sample: each bench "step" is ONE DDIM step (one CFG UNet evaluation) of ONE joint sample at the workload's
    # latent size; ms_per_step is what was measured, `value` extrapolates it to whole 50-step samples
    times = [ref.step()[0] for _ in range(args.warmup + args.steps)][args.warmup:]
    sec = sum(times) / len(times)
    base = ref.describe(sec, args.ddim_steps)
    line = {"impl": "reference", "metric": "inpainted joint samples/sec (50-step DDIM + CFG)", "value": base["value"],
            "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "step_is": "one DDIM step of one joint sample (bounded sample of the workload)",
            "higher_is_better": True, "scaling": "strong" if args.samples_per_gpu <= 0 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "inputs": "larger than L2 (4.2 GB fp32 weights)"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- training step
def measure_train_step(ldm, dev, world, rank, n, latent, steps=3, warmup=1):
    """BASELINE.json config 5: UNet forward + backward (adapter gradients) + ONE NCCL all-reduce of the flat gradient
    buffer + AdamW, `n` joint samples per GPU, timed with CUDA events (max over ranks).  Runs after every inference
    measurement because it re-homes the trainable parameters into the flat buffer."""
    import torch
    import torch.distributed as dist
    from mobi_b200 import ops
    from mobi_b200.training import UNetTrainer
    tr = UNetTrainer(ldm)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    R = 2 * n
    x_start = torch.randn(R, 9, latent, latent, device=dev, generator=g)
    noise = torch.randn(R, 4, latent, latent, device=dev, generator=g)
    cond = torch.randn(R, 2, 768, device=dev, generator=g)
    t = torch.randint(0, 1000, (R,), device=dev, generator=g)

    def step():
        loss = tr.forward_backward(x_start, t, noise, cond)
        tr.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ops.Stats.launches
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    fwd = UNET_FLOPS_PER_JOINT.get(latent, 0) * n
    return {"workload": "mobi_nusc_512 training step: UNet fwd+bwd bf16, %d trainable (adapter) parameters, "
                        "flat-buffer all-reduce%s, AdamW" % (tr.flat.numel, " over NCCL" if world > 1 else " (1 rank: none)"),
            "joint_samples_per_gpu": n, "ms_per_step": ms, "samples_per_s": n * world / (ms / 1e3),
            "gpu_launches_per_step": (ops.Stats.launches - l0) // steps, "loss": float(loss.item()),
            "nominal_tflops_3x_forward_per_gpu": 3 * fwd / (ms / 1e3) / 1e12 if fwd else None}


# ---------------------------------------------------------------------------------------------- range-view post-processing
def measure_range_post(dev, n, px, cpu_samples=2, iters=20):
    """decoded lidar image -> un-cropped sweep -> instance mask -> paste -> edited point cloud for `n`
    samples (mobi_b200.lidar.postprocess_lidar_samples: two launches), against the NumPy port of the reference's per-sample
    host code (scripts/inference_test_bench.py:567-629) timed on `cpu_samples` of the same inputs."""
    import numpy as np
    import torch
    from mobi_b200 import lidar, synth
    batch, dec, bbox = synth.synthetic_lidar_batch(n, px=px, device=dev)
    out = lidar.postprocess_lidar_samples(dec, batch, bbox)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in e:
        flush.zero_()                                             # evict the inputs from L2 between timed iterations
        a.record()
        out = lidar.postprocess_lidar_samples(dec, batch, bbox)
        b.record()
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in e]))
    P = out["range_pred"].shape[-1] * out["range_pred"].shape[-2]
    n_pts = out["n_points"].sum().item()
    algo = n * (2 * px * px * 4 + P * (2 * 4 + 2 * 4 + 2 * 4 + 7 * 4 + 16 + 1)) + n_pts * 20
    # CPU leg: the oracle port on the first samples of the SAME inputs (this is the one place bench may run oracle/)
    from oracle import range_oracle as ro
    k = min(cpu_samples, n)
    host = {key: v[:k].cpu().numpy() for key, v in batch.items()}
    inp = dict(range_depth=dec[:k, [0]].cpu().numpy(), range_int=dec[:k, [1]].cpu().numpy(),
               range_depth_orig=host["range_depth_orig"], range_int_orig=host["range_int_orig"],
               range_pitch=host["range_pitch"], range_yaw=host["range_yaw"],
               range_instance_mask_orig=host["range_instance_mask_orig"], crop_left=host["range_shift_left"],
               width_crop=host["width_crop"], min_depth_obj=host["min_depth_obj"], max_depth_obj=host["max_depth_obj"])
    t0 = time.perf_counter()
    want = ro.run_pipeline(inp, bbox[:k].cpu().numpy())
    cpu_ms = (time.perf_counter() - t0) * 1e3 / k
    same = bool(np.array_equal(out["pred_instance_mask"][:k].cpu().numpy(), want["pred_instance_mask"].astype(np.uint8))
                and np.array_equal(out["range_pred"][:k].cpu().numpy(), want["range_pred"].astype(np.float32))
                and out["n_points"][:k].cpu().tolist() == [len(p) for p in want["pred_points"]])
    return {"workload": "%d decoded %dx%d range images -> 32x1096 sweeps, instance masks, edited clouds" % (n, px, px),
            "ms_per_batch": ms, "samples_per_s": n / (ms / 1e3), "gpu_launches_per_batch": 2,
            "algorithmic_bytes": algo, "achieved_GBps": algo / (ms / 1e3) / 1e9, "l2": "256 MB flush between iterations",
            "points_per_sample": n_pts / n, "cpu_port_ms_per_sample": cpu_ms, "cpu_port_cores": 1,
            "matches_cpu_port_bit_exact": same}


# ---------------------------------------------------------------------------------------------- VAE decode roofline
VAE_DECODE_FLOPS = {64: (2.515e12, 2.685e12), 32: (0.622e12, 0.665e12)}   # per sample: camera, lidar


def measure_vae_decode(ldm, dev, n, latent, peak_sustained, peak_burst, iters=3):
    """decode_first_stage of `n` latents for both modalities, CUDA events on the launching stream; FLOPs = the reference's
    own count per sample."""
    import torch
    z = torch.randn(n, 4, latent, latent, device=dev)
    out = {}
    for name, module_name, fl in (("camera", "first_stage_model", VAE_DECODE_FLOPS[latent][0]),
                                  ("lidar", "lidar_stage_model", VAE_DECODE_FLOPS[latent][1])):
        if getattr(ldm, module_name, None) is None:
            continue
        ldm.decode_first_stage(z, module_name=module_name)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ldm.decode_first_stage(z, module_name=module_name)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        tf = n * fl / (ms / 1e3) / 1e12
        out[name] = {"ms_per_%d_samples" % n: ms, "tflops": tf, "frac_of_sustained_peak": tf / peak_sustained,
                     "frac_of_burst_peak": tf / peak_burst}
    return out


# ---------------------------------------------------------------------------------------------- native arm
def run_native(args):
    import torch
    import torch.distributed as dist
    from mobi_b200 import ops, sharding, synth
    from mobi_b200.ddim import DDIMSampler

    t_start = time.perf_counter()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the native arm has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, total, mb, scaling = job_shape(args, world)
    latent = args.latent
    rps = 1 if args.pbe else 2                 # UNet rows per sample: camera only, or (camera, lidar)
    ldm = synth.build_synthetic_ldm(latent=latent, use_lidar=not args.pbe, device=dev, seed=0, with_vae=True,
                                    with_cond=not args.pbe)
    sampler = DDIMSampler(ldm, use_cuda_graph=not args.no_graph)
    # the job is `total` joint samples cut into per-rank shards between samples (mobi_b200/sharding.py); every sample's
    # inputs and noise depend on its GLOBAL index only, so results do not depend on the number of GPUs
    lo, hi = sharding.shard_bounds(total, world, rank)
    full = synth.synthetic_inputs(total, latent, seed=1, rows_per_sample=rps, n_ctx=1 if args.pbe else 2)
    host = {k: sharding.shard_rows(v, world, rank, rows_per_sample=rps).contiguous() for k, v in full.items()}
    host["x_T"] = sharding.sample_noise((4, latent, latent), lo, hi, base_seed=1, rows_per_sample=rps)
    host = {k: v.pin_memory() for k, v in host.items()}
    devin = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    px = 8 * latent
    host_img = torch.empty((n, 3, px, px), dtype=torch.float32).pin_memory()
    host_rng = torch.empty((n, 2, px, px), dtype=torch.float32).pin_memory()
    chunks = [(i * mb * rps, (i + 1) * mb * rps) for i in range(n // mb)]   # micro-batches: row ranges of the shard

    def sample_rows(inp, a, b):
        return sampler.sample(S=args.ddim_steps, conditioning=inp["cond"][a:b], batch_size=b - a,
                              shape=[4, latent, latent], verbose=False, unconditional_guidance_scale=CFG_SCALE,
                              unconditional_conditioning=inp["uc"][a:b], eta=0.0, x_T=inp["x_T"][a:b],
                              test_model_kwargs=dict(inpaint_image=inp["inpaint_image"][a:b],
                                                     inpaint_mask=inp["inpaint_mask"][a:b]))[0]

    def step_device():
        return [sample_rows(devin, a, b) for a, b in chunks]

    # the e2e leg of the joint model is the reference test bench's per-batch body (mobi_b200/pipeline.py): a batch with
    # the DATASET's layout in pinned host memory (camera / range images, masks, conditioning, sweep tensors)
    from mobi_b200 import pipeline
    host_batch = None if args.pbe else synth.synthetic_dataset_batch(n, px=px, seed=100 + rank, pin=True)
    host_out = {}

    def to_host(name, t, s0, s1):
        if name not in host_out:
            host_out[name] = torch.empty((n,) + tuple(t.shape[1:]), dtype=t.dtype).pin_memory()
        host_out[name][s0:s1].copy_(t, non_blocking=True)                                   # D2H of the results

    def step_e2e():
        """What a user of the reference runs per batch (scripts/inference_test_bench.py:414-464, 567-629): the batch from
        pinned host memory, get_input (4 VAE encodes + latent assembly), conditioning tokens, sampler.sample, decode_sample,
        decode_first_stage for both modalities, range-view post-processing, results back to the host."""
        outs = []
        for a, b in chunks:
            s0, s1 = a // rps, b // rps
            if args.pbe:
                inp = {k: v[a:b].to(dev, non_blocking=True) for k, v in host.items()}      # H2D from pinned memory
                out = sample_rows(inp, 0, b - a)
                host_img[s0:s1].copy_(ldm.decode_first_stage(out), non_blocking=True)
                outs.append(out)
                continue
            batch = pipeline.batch_to_device(pipeline.batch_slice(host_batch, s0, s1), dev)  # H2D from pinned memory
            out = pipeline.inpaint_batch(ldm, sampler, batch, ddim_steps=args.ddim_steps, scale=CFG_SCALE)
            for name in ("image_sample", "range_pred", "pred_instance_mask", "pred_points", "n_points"):
                to_host(name, out[name], s0, s1)
            outs.append(out["samples"])
        return outs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- warm-up (first step also captures the CUDA graphs), then the wall-clock guard
    steps, warmup = args.steps, max(args.warmup, 1)
    tw0 = time.perf_counter()
    step_device()
    torch.cuda.synchronize()
    t_first = time.perf_counter() - tw0
    tw1 = time.perf_counter()
    if warmup > 1:
        step_device()
        torch.cuda.synchronize()
    t_step = (time.perf_counter() - tw1) if warmup > 1 else t_first
    e2e_steps = min(steps, 3)
    tail_s = 60.0 + (e2e_steps + 1) * t_step * 1.08          # e2e leg + roofline / train / CPU legs
    projected = (time.perf_counter() - t_start) + (max(0, warmup - 2) + steps) * t_step + tail_s
    steps_requested = steps
    if projected > args.budget_s:
        g = torch.tensor([max(3, int((args.budget_s - (time.perf_counter() - t_start) - tail_s) / t_step) -
                              max(0, warmup - 2))], device=dev)
        if world > 1:
            dist.all_reduce(g, op=dist.ReduceOp.MIN)
        steps = min(steps, int(g.item()))
        print("bench.py: %d + %d steps of %.1f s would not fit the %.0f s wall-clock budget: timing %d steps"
              % (warmup, steps_requested, t_step, args.budget_s, steps), file=sys.stderr)
    for _ in range(max(0, warmup - 2)):
        step_device()
    clocks = ClockSampler(local)
    clocks.start()
    launches0, evals0 = ops.Stats.launches, sampler.launches
    ms = timed(step_device, steps)
    gpu_launches = ops.Stats.launches - launches0
    unet_evals = sampler.launches - evals0
    clock_info = clocks.finish()
    first = step_e2e()[0]
    ms_e2e = timed(step_e2e, e2e_steps)

    value = n * world * steps / (ms / 1e3)
    e2e_value = n * world * e2e_steps / (ms_e2e / 1e3)
    if args.pbe:
        h2d, d2h = sum(v.numel() * v.element_size() for v in host.values()), host_img.numel() * 4
    else:
        h2d = pipeline.batch_bytes(host_batch)
        d2h = sum(v.numel() * v.element_size() for v in host_out.values())

    # ---- weak-scaling figure of round 1: 8 joint samples per GPU whatever N (secondary)
    weak = None
    if not args.pbe and args.samples_per_gpu <= 0 and n >= 8:
        w8 = lambda: sample_rows(devin, 0, 8 * rps)                                            # noqa: E731
        w8()
        ms_w = timed(w8, 3)
        weak = {"joint_samples_per_gpu": 8, "value": 8 * world * 3 / (ms_w / 1e3), "unit": "samples/s",
                "ms_per_step": ms_w / 3, "steps": 3, "unet_step_ms": ms_w / 3 / args.ddim_steps}

    # ---- roofline of the dominant kernel class (tcgen05 GEMM / implicit conv), timed live with CUDA events on
    # the launching stream over one eager UNet evaluation of one micro-batch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_burst = peaks.get("bf16_tflops", 1590.0)
    peak_src = "of measured (MEASURED_PEAKS.json bf16_tflops_sustained; burst figure beside it)" if peaks else \
        "of fallback (1.4 PFLOP/s sustained under the power cap; B200_PROFILING.md; burst 1.59)"
    roofline = vae = None
    unet = ldm.model.diffusion_model
    if rank == 0:
        rows = 2 * rps * mb
        x_half = torch.randn(rows // 2, 9, latent, latent, device=dev)
        x_in = torch.cat([x_half, x_half])                   # the two CFG halves see the same x (ddim.py:180)
        t_in = torch.full((rows,), 481, device=dev, dtype=torch.long)
        c_in = torch.cat([devin["uc"][:rps * mb], devin["cond"][:rps * mb]]).contiguous()
        unet.pin_context(c_in, unet.prepare_context(c_in))   # context-only work runs once per sampling run, not per step
        unet.cfg_shared_halves = True                        # what the samplers run under CFG
        try:
            unet(x_in, t_in, context=c_in)
            torch.cuda.synchronize()
            ops.Stats.begin_profile()
            unet(x_in, t_in, context=c_in)
            prof = ops.Stats.end_profile()
        finally:
            unet.cfg_shared_halves = False
        tc_ms = sum(prof[k]["ms"] for k in ("gemm", "conv") if k in prof)
        tc_fl = sum(prof[k]["flops"] for k in ("gemm", "conv") if k in prof)
        tc_n = sum(prof[k]["launches"] for k in ("gemm", "conv") if k in prof)
        all_ms = sum(v["ms"] for v in prof.values())
        achieved = tc_fl / (tc_ms / 1e3) / 1e12
        per_sample = UNET_FLOPS_PBE_ROW.get(latent, 0) if args.pbe else UNET_FLOPS_PER_JOINT.get(latent, 0)
        step_flops = per_sample * 2 * mb   # x2: CFG doubles the rows
        ms_unet = ms / steps / max(1, unet_evals // steps)
        step_tf = step_flops / (ms_unet / 1e3) / 1e12 if step_flops else None
        # DRAM traffic of the dominant kernel: a STATIC figure from the committed `ncu --set full` capture of one
        # representative launch (the level-0 conv3x3), not measured in this run
        traffic, traffic_note = None, None
        for cand in ("r02", "r01"):
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", cand, "ncu_traffic.json")))
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                traffic_note = {"static": True, "launch": tj["kernel"], "algorithmic_bytes": tj["algorithmic_bytes"],
                                "source": tj["source"]}
                break
            except Exception:
                continue
        roofline = {"bound": "tensor", "kernel": "gemm2_kernel (persistent tcgen05 GEMM + implicit-GEMM conv3x3)",
                    "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    "frac_of_burst_peak": achieved / peak_burst, "peak_burst": peak_burst,
                    "peak_source": peak_src, "traffic": traffic, "traffic_note": traffic_note,
                    "launches_per_unet_call": tc_n, "flops_per_launch_avg": tc_fl / max(1, tc_n),
                    "ms_per_launch_avg": tc_ms / max(1, tc_n),
                    "share_of_unet_time": tc_ms / all_ms if all_ms else None,
                    "by_kernel_ms": {k: round(v["ms"], 3) for k, v in sorted(prof.items())},
                    "attention_tflops": (prof["attention"]["flops"] / (prof["attention"]["ms"] / 1e3) / 1e12)
                    if "attention" in prof else None,
                    "unet_rows_per_call": rows, "unet_step_ms": ms_unet,
                    "unet_step_algorithmic_tflops": step_tf,
                    "unet_step_frac_of_peak": step_tf / peak_tf if step_tf else None,
                    "unet_step_frac_of_burst_peak": step_tf / peak_burst if step_tf else None}
        try:
            vae = measure_vae_decode(ldm, dev, min(mb, 8), latent, peak_tf, peak_burst)
        except Exception as exc:
            vae = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}

    # ---- parity check of this very model outside the timed region: one CFG DDIM step of joint sample 0 on the GPU
    # path vs the CPU leg below (the reference / oracle port on the SAME weights and inputs)
    check_in = None
    if rank == 0 and not args.no_cpu_baseline and not args.pbe:
        from oracle import unet_oracle as uo
        ci = uo.synth_inputs(1, latent, seed=1)
        smp1 = DDIMSampler(ldm, use_cuda_graph=False)
        smp1.make_schedule(ddim_num_steps=50, ddim_eta=0.0, verbose=False)
        cg = {k: v.to(dev) for k, v in ci.items()}
        x_prev_gpu, _ = smp1.p_sample_ddim(cg["x_T"], cg["cond"], torch.full((2,), 981, device=dev, dtype=torch.long), 49,
                                           unconditional_guidance_scale=CFG_SCALE, unconditional_conditioning=cg["uc"],
                                           test_model_kwargs=dict(inpaint_image=cg["inpaint_image"],
                                                                  inpaint_mask=cg["inpaint_mask"]))
        check_in = (x_prev_gpu.float().cpu(), {k: v.detach().float().cpu() for k, v in unet.state_dict().items()})

    train = None
    if not args.no_train and not args.pbe:
        try:
            train = measure_train_step(ldm, dev, world, rank, args.train_samples_per_gpu, latent)
        except Exception as exc:  # the headline line must survive a failure of this secondary measurement
            train = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
            if world > 1:
                raise
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    range_post = None
    if not args.pbe and not args.no_cpu_baseline:
        try:
            range_post = measure_range_post(dev, min(n, 32), px)
        except Exception as exc:
            range_post = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
    cpu_baseline = check = None
    if not args.no_cpu_baseline:          # rank 0 at every N
        threads = os.cpu_count() or 1
        ref = CpuReference(latent, threads, 1, state_dict=check_in[1] if check_in else None)
        sec, x_prev_cpu = ref.step()
        cpu_baseline = ref.describe(sec, args.ddim_steps)
        if check_in is not None:
            err = ((check_in[0].double() - x_prev_cpu.double()).abs().max() / x_prev_cpu.double().abs().max()).item()
            check = {"what": "x_prev after one CFG-5 DDIM step (t=981) of joint sample 0: GPU path vs the CPU %s on the "
                             "same weights and inputs, outside the timed region" % ref.kind,
                     "max_abs_rel": err, "tolerance": 1e-2, "ok": bool(err < 1e-2)}
            if not check["ok"]:
                raise RuntimeError("bench.py: GPU result differs from the CPU %s: %r" % (ref.kind, check))
    line = {"metric": "inpainted joint samples/sec (50-step DDIM + CFG)", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "bf16 operands, fp32 accumulate/norms/residuals", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "latent": latent, "total_joint_samples": total,
                       "joint_samples_per_gpu": n, "micro_batch": mb, "rows_per_unet_call": 2 * rps * mb,
                       "ddim_steps": args.ddim_steps, "cfg_scale": CFG_SCALE,
                       "sharding": "%d samples / %d GPUs, no collective" % (total, world),
                       "l2": "inputs larger than L2 (2.1 GB bf16 weights streamed per UNet call)",
                       "cuda_graph": not args.no_graph,
                       "e2e_includes": ("pinned H2D of latents/conditioning, sampling, camera VAE decode to %dx%d, D2H of "
                                        "decoded images; %d timed steps" % (px, px, e2e_steps)) if args.pbe else
                                       ("the reference test bench's per-batch body (mobi_b200.pipeline.inpaint_batch): "
                                        "pinned H2D of the dataset-layout batch (%dx%d camera + range images, masks, "
                                        "conditioning, sweeps), get_input = 4 VAE encodes + latent assembly, conditioning "
                                        "tokens after the CLIP tower, sampling, decode_sample, camera + range-view VAE "
                                        "decode, range-view post-processing, D2H of decoded image, edited sweep, instance "
                                        "mask and point cloud; %d timed steps" % (px, px, e2e_steps))},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
            "gpu_launches": gpu_launches, "unet_evals": unet_evals, "clocks": clock_info,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "check": check, "weak_8_per_gpu": weak,
            "vae_decode": vae, "train_step": train, "range_post": range_post,
            "first_step_s": t_first, "wall_s": time.perf_counter() - t_start}
    if steps != steps_requested:
        line["steps_requested"] = steps_requested
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: anything libraries print there (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved, "w")
    sys.stdout = real_stdout_proxy = _Tee(real_stdout)
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_native(args)
    finally:
        real_stdout_proxy.flush()


class _Tee:
    """print() target that writes to the saved stdout descriptor (fd 1 itself is pointed at stderr)."""

    def __init__(self, f):
        self.f = f

    def write(self, s):
        return self.f.write(s)

    def flush(self):
        self.f.flush()


if __name__ == "__main__":
    main()

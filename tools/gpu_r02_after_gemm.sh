#!/bin/bash
# round 2: after GEMM changes: kernel + UNet + headline + VAE tests, micro-benchmarks, then a short config-3 bench
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
python tools/kbench.py gemm conv 2>&1 | grep "^gemm\|^conv"
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_headline_gpu.py tests/test_vae_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --total-samples 32 --steps 2 --warmup 3 --budget-s 10000 > gpurun_out/r02_bench_after_gemm.json 2> gpurun_out/r02_bench_after_gemm.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_after_gemm.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value", d["value"], "e2e", d["e2e"]["value"], "unet_step_ms", r["unet_step_ms"], "frac", r["unet_step_frac_of_peak"], "by_kernel", r["by_kernel_ms"])
print("weak8", d["weak_8_per_gpu"]["unet_step_ms"], "vae", d["vae_decode"]["camera"]["ms_per_8_samples"], d["vae_decode"]["lidar"]["ms_per_8_samples"], "train", d["train_step"]["ms_per_step"])
PY

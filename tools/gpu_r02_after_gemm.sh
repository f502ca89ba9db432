#!/bin/bash
# round 2: after the GEMM epilogue / tile-policy changes: kernel + UNet + headline tests, then a short config-3 bench
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
python tools/kbench.py gemm 2>&1 | grep "^gemm"
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_headline_gpu.py tests/test_vae_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --total-samples 32 --steps 2 --warmup 3 --budget-s 10000 > gpurun_out/r02_bench_after_gemm.json 2> gpurun_out/r02_bench_after_gemm.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/r02_bench_after_gemm.json

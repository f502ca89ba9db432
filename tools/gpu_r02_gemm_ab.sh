#!/bin/bash
# round 2: GEMM / conv micro-benchmarks + the GEMM kernel tests after an epilogue change
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
python tools/kbench.py gemm conv 2>&1 | tail -23
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm or conv or colstats or producer" 2>&1 | tail -5

#!/bin/bash
# round 2: 12 epilogue warps for the GEGLU class: GEMM tests + micro-benchmark
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm or conv" 2>&1 | tail -4
timeout 200 python tools/kbench.py gemm 2>&1 | grep "^gemm"

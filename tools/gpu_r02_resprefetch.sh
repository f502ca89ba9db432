#!/bin/bash
# round 2: residual L2 prefetch A/B (MOBI_RES_PREFETCH=0 is the old behaviour) + GEMM / conv tests
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm or conv" 2>&1 | tail -4
for v in 0 1; do
  echo "=== MOBI_RES_PREFETCH=$v"
  MOBI_RES_PREFETCH=$v timeout 200 python tools/kbench.py gemm conv 2>&1 | grep "res\|^conv"
done

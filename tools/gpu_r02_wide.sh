#!/bin/bash
# round 2: 320-column tiles on CTA pairs (pair = 3): correctness under a short timeout first, then the A/B
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "wide" 2>&1 | tail -15
rc=${PIPESTATUS[0]}; echo "wide tests rc=$rc"
if [ "$rc" != "0" ]; then exit 0; fi
for cfg in "0 0" "160 3"; do
  set -- $cfg
  echo "=== tile_n=$1 pair=$2"
  KB_GEMM_TILE=$1 KB_GEMM_PAIR=$2 timeout 120 python tools/kbench.py conv 2>&1 | grep "^conv"
done
echo "=== gemm, tile 160 pair 3"
KB_GEMM_TILE=160 KB_GEMM_PAIR=3 timeout 120 python tools/kbench.py gemm 2>&1 | grep "^gemm"

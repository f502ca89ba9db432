#!/bin/bash
# round 2: 4-CTA clusters with multicast A tiles (pair = 2): correctness under a short timeout first, then the A/B
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "quad" 2>&1 | tail -15
rc=${PIPESTATUS[0]}; echo "quad tests rc=$rc"
if [ "$rc" != "0" ]; then exit 0; fi
for cfg in "0 0" "160 2" "128 2" "256 2"; do
  set -- $cfg
  echo "=== tile_n=$1 pair=$2"
  KB_GEMM_TILE=$1 KB_GEMM_PAIR=$2 timeout 120 python tools/kbench.py conv 2>&1 | grep "^conv"
done
echo "=== gemm PLAIN shapes, tile 160 pair 2 (only the f32 / residual rows are meaningful)"
KB_GEMM_TILE=160 KB_GEMM_PAIR=2 timeout 120 python tools/kbench.py gemm 2>&1 | grep "res\|proj_in\|Error\|error" | head -12

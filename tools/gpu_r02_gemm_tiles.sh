#!/bin/bash
# round 2: tile width x CTA-pair matrix of the transformer GEMMs (L2 -> SM traffic per FLOP is what bounds K = 320)
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
for cfg in "0 0" "160 1" "256 -1" "256 1" "128 1"; do
  set -- $cfg
  echo "=== tile_n=$1 pair=$2"
  KB_GEMM_TILE=$1 KB_GEMM_PAIR=$2 python tools/kbench.py gemm 2>&1 | grep "^gemm"
done

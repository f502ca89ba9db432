#!/bin/bash
# round 2: ln_adapter with difference tables: rows-per-warp / occupancy variants (libraries built with -DAD_R3/6/12, -DAD_MINB)
mkdir -p gpurun_out
for v in 5 6 7 8 9; do
  echo "=== variant $v"
  export MOBI_B200_LIB=$PWD/mobi_b200/_build/libmobi_ln_v$v.so
  timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "ln_adapter" 2>&1 | tail -2
  timeout 100 python tools/kbench.py ln 2>&1 | grep "ln_adapter"
done

#!/bin/bash
# round 2: `ncu --set full` of the five level-0 transformer GEMMs (K = 320: qkv, to_out + residual, ff1 GEGLU, ff2 + residual,
# proj_in) -- one launch per shape (KB_WARMUP=0 KB_REPS=1), so the first five gemm2 launches are those shapes in order.
# Usage: tools/gpu_r02_ncu_gemm.sh [tag]
tag=${1:-l0}
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
KB_WARMUP=0 KB_REPS=1 timeout 600 ncu --set full --import-source on --clock-control none -f -k regex:gemm2_kernel -c 5 \
    -o gpurun_out/ncu_r02_gemm_$tag python tools/kbench.py gemm > gpurun_out/ncu_r02_gemm_$tag.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_r02_gemm_$tag.log
ls -la gpurun_out | grep ncu_r02_gemm

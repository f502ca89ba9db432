#!/bin/bash
# `ncu --set full` captures of the three dominant kernels at their mobi_nusc_512 shapes (32 rows), driven by the kernel
# micro-benchmark so that each capture costs one short process instead of a replay of the whole sampler.
# Usage: tools/gpu_ncu_kbench.sh <tag>      -> gpurun_out/ncu_<tag>_{conv,ff1,attn}.ncu-rep
tag=${1:-r01}
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none -f"
# conv3x3 320->320 @64x64 (first conv shape of kbench: launches 0..12 of gemm2_kernel)
timeout 400 $NCU -k regex:gemm2_kernel -s 6 -c 1 -o gpurun_out/ncu_${tag}_conv python tools/kbench.py conv > gpurun_out/ncu_${tag}_conv.log 2>&1
echo "conv rc=$?"
# ff1 GEGLU C=320 (third GEMM shape: launches 26..38)
timeout 400 $NCU -k regex:gemm2_kernel -s 32 -c 1 -o gpurun_out/ncu_${tag}_ff1 python tools/kbench.py gemm > gpurun_out/ncu_${tag}_ff1.log 2>&1
echo "ff1 rc=$?"
# L0 self-attention d=40 T=4096 (attention3, row-major V)
timeout 400 $NCU -k regex:attention3 -s 6 -c 1 -o gpurun_out/ncu_${tag}_attn python tools/kbench.py attn > gpurun_out/ncu_${tag}_attn.log 2>&1
echo "attn rc=$?"
ls -la gpurun_out | grep ncu_${tag}

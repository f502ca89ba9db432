#!/bin/bash
# Runs the kernel parity tests on the GPU box, one pytest process per kernel family so that a trap in one
# kernel (sticky CUDA error) does not poison the others.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
status=0
for grp in "gemm_plain" "gemm_residual or gemm_geglu or gemm_head" "conv_implicit" "conv_im2col" "attention and not ctx" \
           "groupnorm or layernorm or ln_adapter or small or ctx_attention or sampler"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$grp" --tb=short -p no:cacheprovider \
      > "gpurun_out/kern_${name}.log" 2>&1
  rc=$?
  echo "== $grp -> rc=$rc"; tail -n 12 "gpurun_out/kern_${name}.log"
  [ $rc -ne 0 ] && status=1
done
exit $status

#!/bin/bash
# round 2, first GPU call: the whole -m gpu suite (incl. the new headline-shape parity tests) and the micro-batch sweep
# of the config-3 bench (16 / 32 / 64 joint samples per sampler call on one GPU)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_pytest_gpu_first.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_first.log
tail -5 gpurun_out/r02_pytest_gpu_first.log
for mb in 16 32 64; do
  timeout 400 python bench.py --total-samples $mb --micro-batch $mb --steps 1 --warmup 1 --no-train --no-cpu-baseline \
      --budget-s 10000 > gpurun_out/r02_sweep_mb$mb.json 2> gpurun_out/r02_sweep_mb$mb.err
  echo "mb=$mb rc=$?"; tail -c 600 gpurun_out/r02_sweep_mb$mb.json
done

#!/bin/bash
# 2 GPUs: training step with the bucketed all-reduce overlapped with the backward vs one blocking all-reduce
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/train_bench.py --steps 10 --warmup 3 > gpurun_out/r02_train_2gpu_overlap.json 2> gpurun_out/r02_train_2gpu_overlap.err
echo "overlap rc=$?"; tail -c 900 gpurun_out/r02_train_2gpu_overlap.json; tail -3 gpurun_out/r02_train_2gpu_overlap.err
timeout 600 $TR tools/train_bench.py --steps 10 --warmup 3 --no-overlap > gpurun_out/r02_train_2gpu_blocking.json 2> gpurun_out/r02_train_2gpu_blocking.err
echo "blocking rc=$?"; tail -c 900 gpurun_out/r02_train_2gpu_blocking.json
timeout 400 python tools/train_bench.py --steps 10 --warmup 3 > gpurun_out/r02_train_1gpu.json 2>/dev/null; tail -c 600 gpurun_out/r02_train_1gpu.json

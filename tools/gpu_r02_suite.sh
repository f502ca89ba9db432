#!/bin/bash
# round 2: the whole -m gpu suite with timings, then a short config-3 bench (16 joint samples, one timed step)
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
timeout 1700 python -m pytest tests -m gpu -q -s --durations=15 > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r02_pytest_gpu.log; grep -n "max-abs-rel\|floor\|pipeline vs" gpurun_out/r02_pytest_gpu.log | cut -c1-330 | tail -40
timeout 600 python bench.py --total-samples ${BENCH_TOTAL:-16} --micro-batch 16 --steps 1 --warmup 2 --budget-s 10000 \
    > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r02_bench_short.json; tail -5 gpurun_out/r02_bench_short.err

#!/bin/bash
# round 2: the whole -m gpu suite with timings, smoke(), then (BENCH=1) a short config-3 bench
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
timeout 1700 python -m pytest tests -m gpu -q -s --durations=12 > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -22 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "${BENCH:-0}" = "1" ]; then
timeout 600 python bench.py --total-samples ${BENCH_TOTAL:-32} --steps 1 --warmup 2 --budget-s 10000 > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_short.json
fi

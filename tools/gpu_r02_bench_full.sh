#!/bin/bash
# the driver's own command line: config 3 as written on one GPU (64 joint samples), 20 timed + 5 warm-up steps
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
( nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 1000 > gpurun_out/r02_bench_full_clocks.csv ) &
SMI=$!
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err
echo "bench rc=$?"; kill $SMI
tail -c 2500 gpurun_out/r02_bench_full.json; grep "bench.py:" gpurun_out/r02_bench_full.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; cat gpurun_out/r02_bench_reference_arm.json | cut -c1-600

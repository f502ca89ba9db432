#!/bin/bash
# ncu evidence for one round: (1) launch list with per-launch device time of a shortened bench run,
# (2) --set full captures of the heaviest kernels.  Usage: tools/gpu_profile.sh <tag> [kernel-regex ...]
tag=${1:-r01}; shift
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --ddim-steps 2 --no-cpu-baseline --no-graph"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${tag}.csv $BENCH > gpurun_out/launches_${tag}.log 2>&1
echo "launch list rc=$?"; tail -n 2 gpurun_out/launches_${tag}.log
for k in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 10 -c 3 \
      -f -o gpurun_out/prof_${tag}_${k} $BENCH > gpurun_out/prof_${tag}_${k}.log 2>&1
  echo "ncu full $k rc=$?"
done
ls -la gpurun_out | tail -n 20

#!/bin/bash
# round-2 ncu evidence: (1) launch list with per-launch device time of a shortened bench run (eager, 2 DDIM steps of 8
# joint samples), (2) --set full captures of the three dominant kernels at their mobi_nusc_512 shapes (32 rows).
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
BENCH="python bench.py --total-samples 8 --micro-batch 8 --steps 1 --warmup 1 --ddim-steps 2 --no-cpu-baseline --no-graph --no-train --budget-s 10000"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv $BENCH > gpurun_out/launches_r02.log 2>&1
echo "launch list rc=$?"; tail -n 1 gpurun_out/launches_r02.log | cut -c1-200
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 400 $NCU -k regex:attention4 -s 2 -c 1 -o gpurun_out/ncu_r02_attention4 python tools/attn_prof.py 32 40 4096 0 > gpurun_out/ncu_r02_attention4.log 2>&1; echo "attn rc=$?"
timeout 400 $NCU -k regex:gemm2_kernel -s 6 -c 1 -o gpurun_out/ncu_r02_conv python tools/kbench.py conv > gpurun_out/ncu_r02_conv.log 2>&1; echo "conv rc=$?"
timeout 400 $NCU -k regex:gn_stream -s 3 -c 1 -o gpurun_out/ncu_r02_gn_stream python tools/kbench.py gn > gpurun_out/ncu_r02_gn_stream.log 2>&1; echo "gn rc=$?"
ls -la gpurun_out | grep -E "r02.*(ncu-rep|csv)"

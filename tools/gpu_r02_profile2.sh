#!/bin/bash
# round 2 (final kernels): (1) ncu launch list of two eager DDIM steps of 8 joint samples, (2) --set full capture of the
# level-0 conv3x3 on wide pair tiles (refreshes profiles/r02/ncu_traffic.json), (3) the same for the level-0 attention.
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02b.csv \
    python tools/launch_list_run.py > gpurun_out/launches_r02b.log 2>&1
echo "launch list rc=$?"; tail -n 1 gpurun_out/launches_r02b.log | cut -c1-200; wc -l gpurun_out/launches_r02b.csv
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 400 $NCU -k regex:gemm2_kernel -s 6 -c 1 -o gpurun_out/ncu_r02_conv_wide python tools/kbench.py conv > gpurun_out/ncu_r02_conv_wide.log 2>&1; echo "conv rc=$?"
ls -la gpurun_out | grep -E "r02b|conv_wide"

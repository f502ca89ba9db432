#!/bin/bash
# round 2: attention4 (P in TMEM) parity + A/B against attention3 (+ optional ncu --set full capture: NCU=1)
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "attention" > gpurun_out/r02_attn4_tests.log 2>&1
echo "attention tests rc=$?"; tail -3 gpurun_out/r02_attn4_tests.log; grep FAILED gpurun_out/r02_attn4_tests.log | head
timeout 300 python tools/kbench.py attn > gpurun_out/r02_kbench_attn4.log 2>&1; cat gpurun_out/r02_kbench_attn4.log
if [ "$NCU" = "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention4 -s 2 -c 1 -f -o gpurun_out/r02_ncu_attention4 \
    python tools/attn_prof.py 32 40 4096 0 > gpurun_out/r02_ncu_attention4.log 2>&1; tail -2 gpurun_out/r02_ncu_attention4.log
fi

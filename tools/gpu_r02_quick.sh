#!/bin/bash
# quick GPU loop: selected tests ($TESTS, pytest -k $K) and a short config-3 bench
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
timeout 1200 python -m pytest ${TESTS:-tests/test_kernels_gpu.py} -m gpu -q -x -s ${K:+-k "$K"} > gpurun_out/r02_quick_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r02_quick_tests.log; grep -n "max-abs-rel\|FAILED\|Error" gpurun_out/r02_quick_tests.log | cut -c1-300 | tail -30
if [ "${BENCH:-1}" = "1" ]; then
timeout 600 python bench.py --total-samples ${BENCH_TOTAL:-16} --micro-batch ${BENCH_MB:-16} --steps 1 --warmup 2 --budget-s 10000 ${BENCH_ARGS} \
    > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_short.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02_bench_short.json').read().strip().splitlines()[-1])
r=l['roofline']
print('value',l['value'],'e2e',l['e2e']['value'],'weak8',l['weak_8_per_gpu'])
print({k:r[k] for k in ('achieved','frac','unet_step_ms','unet_step_frac_of_peak','by_kernel_ms','attention_tflops')})
print('vae',l['vae_decode']); print('train',l['train_step']); print('check',l['check'])
PY
fi

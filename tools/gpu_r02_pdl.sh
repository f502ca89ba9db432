#!/bin/bash
# programmatic dependent launch A/B: parity tests with PDL on, then the short config-3 bench with MOBI_PDL=1 and =0
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_sampler_api_gpu.py tests/test_vae_gpu.py -m gpu -q -x > gpurun_out/r02_pdl_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02_pdl_tests.log
for pdl in 1 0; do
  MOBI_PDL=$pdl timeout 600 python bench.py --total-samples 16 --micro-batch 16 --steps 2 --warmup 2 --budget-s 10000 --no-train --no-cpu-baseline \
      > gpurun_out/r02_bench_pdl$pdl.json 2> gpurun_out/r02_bench_pdl$pdl.err
  echo "pdl=$pdl rc=$?"
  python - <<PY
import json
l=json.loads(open('gpurun_out/r02_bench_pdl$pdl.json').read().strip().splitlines()[-1])
r=l['roofline']
print('pdl=$pdl value',l['value'],'unet_step_ms',r['unet_step_ms'],'frac',r['unet_step_frac_of_peak'],'weak8',l['weak_8_per_gpu']['unet_step_ms'], 'sum_kernels', sum(r['by_kernel_ms'].values()), 'vae', l['vae_decode']['camera'])
PY
done

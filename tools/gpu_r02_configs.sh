#!/bin/bash
# BASELINE.json configs 2 and 4 through the same bench: mobi_nusc_256 (16 joint samples, latent 32) and pbe.yaml (camera-only, 64 samples)
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
timeout 600 python bench.py --latent 32 --total-samples 16 --micro-batch 16 --steps 3 --warmup 2 --no-train > gpurun_out/r02_bench_mobi_nusc_256.json 2> gpurun_out/r02_bench_256.err
echo "256 rc=$?"; tail -2 gpurun_out/r02_bench_256.err | cut -c1-300
timeout 900 python bench.py --pbe --total-samples 64 --micro-batch 32 --steps 2 --warmup 2 > gpurun_out/r02_bench_pbe_512.json 2> gpurun_out/r02_bench_pbe.err
echo "pbe rc=$?"; tail -2 gpurun_out/r02_bench_pbe.err | cut -c1-300
python - <<'PY'
import json
for f in ('r02_bench_mobi_nusc_256','r02_bench_pbe_512'):
    try:
        l=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); r=l['roofline']
        print(f, {k:l[k] for k in ('value','ms_per_step')}, 'e2e', l['e2e']['value'], {k:r[k] for k in ('unet_step_ms','unet_step_frac_of_peak','unet_rows_per_call')}, l.get('check'))
    except Exception as e: print(f, 'ERR', e)
PY

#!/bin/bash
# config 3 on 8 GPUs exactly like the driver launches it (torchrun, one rank per GPU), short: 2 timed steps
N=${N:-8}
mkdir -p gpurun_out
timeout ${LIMIT:-170} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 2 --warmup 1 \
    > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
echo "bench N=$N rc=$?"; tail -3 gpurun_out/r02_bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
l=json.loads(open('gpurun_out/r02_bench_${N}gpu.json').read().strip().splitlines()[-1])
r=l['roofline']
print({k:l[k] for k in ('value','n_gpus','steps','ms_per_step','scaling')}, 'e2e',l['e2e']['value'],'weak8',l['weak_8_per_gpu'])
print(l['config']['workload']); print({k:r[k] for k in ('unet_step_ms','unet_step_frac_of_peak','unet_rows_per_call')}); print('train',l['train_step']); print('wall',l['wall_s'])
PY

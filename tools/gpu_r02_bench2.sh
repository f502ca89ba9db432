#!/bin/bash
# config 3 on N GPUs exactly like the driver launches it (torchrun, one rank per GPU), short
N=${N:-2}
mkdir -p gpurun_out
python -c "from mobi_b200 import build; build.build()" || exit 1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps ${STEPS:-2} --warmup ${WARMUP:-1} \
    > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
echo "bench N=$N rc=$?"; tail -3 gpurun_out/r02_bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
l=json.loads(open('gpurun_out/r02_bench_${N}gpu.json').read().strip().splitlines()[-1])
r=l['roofline']
print({k:l[k] for k in ('value','n_gpus','steps','ms_per_step','scaling')}, 'e2e',l['e2e']['value'],'weak8',l['weak_8_per_gpu'])
print(l['config']['workload']); print({k:r[k] for k in ('unet_step_ms','unet_step_frac_of_peak','unet_rows_per_call')}); print('train',l['train_step']); print('cpu',l['cpu_baseline']); print('check',l['check']); print('wall',l['wall_s'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 1 --warmup 0 2>/dev/null | cut -c1-300

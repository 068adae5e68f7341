#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo gpus=$N
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29645 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2c45_bench$N.json 2> gpurun_out/r2c45_bench$N.err
python -c "
import json; d=json.loads(open('gpurun_out/r2c45_bench$N.json').read().strip().splitlines()[-1])
print('bench$N', d['ms_per_step'], d['value'], d['final_loss'], d.get('phases_unoverlapped'), d['e2e']['value'])" || tail -12 gpurun_out/r2c45_bench$N.err

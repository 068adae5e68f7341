#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2c21_bench8.json 2> gpurun_out/r2c21_bench8.err
python -c "
import json; d=json.loads(open('gpurun_out/r2c21_bench8.json').read().strip().splitlines()[-1])
print('bench8', d['ms_per_step'], d['value'], d['final_loss'], d.get('phases_unoverlapped'), d['launch'][:70], d['e2e']['value'])" || tail -12 gpurun_out/r2c21_bench8.err

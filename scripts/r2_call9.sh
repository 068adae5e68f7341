#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_fused_gather_gpu.py -q -x > gpurun_out/r2c9_pytest_2gpu.log 2>&1
tail -25 gpurun_out/r2c9_pytest_2gpu.log
for flag in "" "--no-overlap"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-rooflines $flag > gpurun_out/r2c9_bench2$flag.json 2> gpurun_out/r2c9_bench2$flag.err
python -c "
import json; d=json.loads(open('gpurun_out/r2c9_bench2$flag.json').read().strip().splitlines()[-1])
print('bench2 $flag', d['ms_per_step'], d['value'], d['final_loss'], d.get('allreduce_overlapped'), d.get('phases_unoverlapped'), d['launch'][:60])" || tail -8 "gpurun_out/r2c9_bench2$flag.err"
done

#!/bin/bash
# 2-GPU validation of the data-parallel paths: fused normalise+all-gather tests, graph-mode bench, eager DDP bench
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_fused_gather_gpu.py -q -x 2>&1 | tail -3
for g in "" "--no-graph"; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 5 --no-cpu-baseline --no-kernel-rooflines $g > gpurun_out/dp2$g.log 2>&1
  echo "rc=$? $g"; tail -1 gpurun_out/dp2$g.log | cut -c1-1200
done

#!/bin/bash
# ncu evidence for the native kernels (one GPU, short commands).
set -x
mkdir -p gpurun_out
# 1. launch list of one full bench step window (shares, cold cache)
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 4000 --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/ncu_step.log 2>&1
# 2. full captures of the native kernels in isolation
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"sa_mlp|sa_sample|gemm_kernel" -s 12 -c 8 -o gpurun_out/prof_pointnet -f python scripts/time_pointnet.py 5120 > gpurun_out/ncu_pn.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"attention_fwd|ce_fwd_bwd|pairwise_locs" -c 6 -o gpurun_out/prof_attn -f python scripts/time_attention.py > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_step.csv
tail -3 gpurun_out/ncu_step.log gpurun_out/ncu_pn.log gpurun_out/ncu_attn.log

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:sa_mlp -s 6 -c 2 -o gpurun_out/prof_sa_mlp -f python scripts/time_pointnet.py 5120 > gpurun_out/ncu_mlp.log 2>&1
tail -3 gpurun_out/ncu_mlp.log
ls -la gpurun_out/*.ncu-rep

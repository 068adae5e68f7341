#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:sa_mlp -s 14 -c 2 -o gpurun_out/prof_sa2 -f python scripts/time_pointnet.py 5120 > gpurun_out/ncu_sa2.log 2>&1
ls -la gpurun_out/prof_sa2.ncu-rep; tail -2 gpurun_out/ncu_sa2.log

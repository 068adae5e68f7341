#!/bin/bash
# ncu --set full capture of the native GEMM at one shape of the step (default: SA3 layer 0)
mkdir -p gpurun_out
SHAPE=${1:-sa3_l0}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 1 -o gpurun_out/prof_gemm_$SHAPE -f python scripts/time_gemm_shapes.py quick $SHAPE > gpurun_out/prof_gemm.log 2>&1
tail -2 gpurun_out/prof_gemm.log

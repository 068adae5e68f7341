#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 10 -o gpurun_out/r2_prof_gemm -f python scripts/prof_gemm_r2.py > gpurun_out/r2c4_ncu.log 2>&1
tail -3 gpurun_out/r2c4_ncu.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/r2c4_pytest_full.log
tail -12 gpurun_out/r2c4_pytest_full.log

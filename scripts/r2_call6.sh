#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/mma_bench.py > gpurun_out/r2_mma_bench.json 2>&1
cat gpurun_out/r2_mma_bench.json | tr -d '\n' | head -c 3000; echo
timeout 300 python scripts/debug_relu.py > gpurun_out/r2c6_debug_relu.log 2>&1
cat gpurun_out/r2c6_debug_relu.log | tail -40

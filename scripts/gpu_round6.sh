#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_pointnet_module.py tests/test_gps_modules.py tests/test_gemm_gpu.py -q -m gpu -s > gpurun_out/pytest_r6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r6.log
grep -E "passed|failed|rel err|fused bf16|Error|assert" gpurun_out/pytest_r6.log | tail -12 | cut -c1-400
timeout -s KILL 300 python scripts/time_pointnet.py > gpurun_out/time_pointnet.json 2> gpurun_out/time_pointnet.err; cat gpurun_out/time_pointnet.json; tail -3 gpurun_out/time_pointnet.err

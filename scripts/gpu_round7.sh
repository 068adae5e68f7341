#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_attention_gpu.py tests/test_gps_modules.py -q -m gpu -s > gpurun_out/pytest_r7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r7.log
grep -E "passed|failed|fused bf16|Error|assert|^E " gpurun_out/pytest_r7.log | tail -14 | cut -c1-400
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-1200 gpurun_out/bench.json; tail -3 gpurun_out/bench.err

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gps_modules.py tests/test_pointnet_module.py -q -m gpu -s > gpurun_out/pytest_gps.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gps.log
tail -25 gpurun_out/pytest_gps.log | cut -c1-700
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-3000; tail -5 gpurun_out/bench.err

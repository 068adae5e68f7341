#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log
tail -6 gpurun_out/pytest_full.log | cut -c1-300
timeout -s KILL 300 python scripts/time_pointnet.py > gpurun_out/time_pointnet.json 2>/dev/null; cat gpurun_out/time_pointnet.json
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-2800 gpurun_out/bench.json; tail -3 gpurun_out/bench.err

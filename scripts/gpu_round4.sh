#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_pointnet_module.py tests/test_tc05_gpu.py -q -m gpu -s > gpurun_out/pytest_pn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pn.log
tail -40 gpurun_out/pytest_pn.log | cut -c1-600
timeout -s KILL 300 python scripts/time_pointnet.py > gpurun_out/time_pointnet.json 2> gpurun_out/time_pointnet.err; cat gpurun_out/time_pointnet.json; tail -5 gpurun_out/time_pointnet.err

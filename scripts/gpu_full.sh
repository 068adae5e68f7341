#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log
tail -5 gpurun_out/pytest_full.log | cut -c1-300
timeout -s KILL 200 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' 2>&1 | tail -2
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-700 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout -s KILL 600 python bench.py --workload pointops_sweep > gpurun_out/bench_sweep.json 2>> gpurun_out/bench.err; cut -c1-1500 gpurun_out/bench_sweep.json
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_ref.json

#!/bin/bash
# Round 2, GPU call 1: baseline of everything the round works on.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_gpu.txt
timeout 600 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -E "PARITY|passed|failed|Error|error|assert" | tail -30 > gpurun_out/r2c1_pytest.log
tail -5 gpurun_out/r2c1_pytest.log
timeout 300 python scripts/time_gemm_shapes.py 2>&1 | tail -1 > gpurun_out/r2c1_gemm_shapes.json
timeout 300 python scripts/time_linear_bwd.py 2>&1 | tail -1 > gpurun_out/r2c1_linear_bwd.json
for f in 0 1; do
  SVB200_NATIVE_BWD_GEMM=$f timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-rooflines 2>&1 | tail -1 > gpurun_out/r2c1_bench_bwd$f.json
  python -c "import json; d=json.load(open('gpurun_out/r2c1_bench_bwd$f.json')); print('native_bwd_gemm=$f', d['ms_per_step'], d['value'], d['final_loss'])"
done
timeout 400 python scripts/step_launch_list.py 2>&1 | tail -45 > gpurun_out/r2c1_cupti.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_step_launches_ncu.csv python scripts/step_launch_list.py --ncu > gpurun_out/r2c1_ncu.log 2>&1
echo ncu_rc=$?
wc -l gpurun_out/r2_step_launches_ncu.csv

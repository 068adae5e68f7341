#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_relu.py > gpurun_out/r2c5_debug_relu.log 2>&1
cat gpurun_out/r2c5_debug_relu.log | tail -30
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c5_pytest_full.log 2>&1
tail -12 gpurun_out/r2c5_pytest_full.log
for c in 0 1 2; do
SVB200_GEMM_CTAS=$c timeout 300 python scripts/time_gemm_shapes.py 2>&1 | tail -1 > gpurun_out/r2c5_gemm_shapes_cta$c.json
SVB200_GEMM_CTAS=$c timeout 300 python scripts/time_linear_bwd.py 2>&1 | tail -1 > gpurun_out/r2c5_linear_bwd_cta$c.json
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-rooflines 2>&1 | tail -1 > gpurun_out/r2c5_bench.json
python -c "import json; d=json.load(open('gpurun_out/r2c5_bench.json')); print('bench', d['ms_per_step'], d['value'], d['final_loss'], d['gpu_launches'])"
timeout 400 python scripts/step_launch_list.py 2>&1 | tail -45 > gpurun_out/r2c5_cupti.log
head -3 gpurun_out/r2c5_cupti.log

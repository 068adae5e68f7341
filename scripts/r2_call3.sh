#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linear_gpu.py tests/test_gemm_gpu.py -q -x 2>&1 | tail -15 > gpurun_out/r2c3_linear.log
tail -4 gpurun_out/r2c3_linear.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2c3_pytest.log
tail -12 gpurun_out/r2c3_pytest.log
for c in 1 2; do
SVB200_GEMM_CTAS=$c timeout 300 python scripts/time_gemm_shapes.py 2>&1 | tail -1 > gpurun_out/r2c3_gemm_shapes_cta$c.json
SVB200_GEMM_CTAS=$c timeout 300 python scripts/time_linear_bwd.py 2>&1 | tail -1 > gpurun_out/r2c3_linear_bwd_cta$c.json
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-rooflines 2>&1 | tail -1 > gpurun_out/r2c3_bench.json
python -c "import json; d=json.load(open('gpurun_out/r2c3_bench.json')); print('bench', d['ms_per_step'], d['value'], d['final_loss'], d['gpu_launches'])"
timeout 400 python scripts/step_launch_list.py 2>&1 | tail -45 > gpurun_out/r2c3_cupti.log
head -3 gpurun_out/r2c3_cupti.log

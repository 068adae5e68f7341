#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c7_pytest_full.log 2>&1
tail -8 gpurun_out/r2c7_pytest_full.log
timeout 300 python scripts/time_gemm_shapes.py 2>&1 | tail -1 > gpurun_out/r2c7_gemm_shapes.json
timeout 300 python scripts/time_linear_bwd.py 2>&1 | tail -1 > gpurun_out/r2c7_linear_bwd.json
python -c "
import json
d=json.load(open('gpurun_out/r2c7_gemm_shapes.json')); print({k:(v.get('native_ms'), v.get('cublas_ms')) for k,v in d.items()})
d=json.load(open('gpurun_out/r2c7_linear_bwd.json')); print({k:(v['dgrad_native_ms'], v['dgrad_cublas_ms'], v['wgrad_native_ms'], v['wgrad_cublas_ms']) for k,v in d['shapes'].items()}); print(d['sum_ms'])"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-rooflines 2>&1 | tail -1 > gpurun_out/r2c7_bench.json
python -c "import json; d=json.load(open('gpurun_out/r2c7_bench.json')); print('bench', d['ms_per_step'], d['value'], d['final_loss'], d['gpu_launches'])"
timeout 400 python scripts/step_launch_list.py 2>&1 | tail -45 > gpurun_out/r2c7_cupti.log
head -3 gpurun_out/r2c7_cupti.log

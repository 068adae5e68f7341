#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c20_pytest_full.log 2>&1
tail -3 gpurun_out/r2c20_pytest_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('__SMOKE_OK__')" 2>&1 | tail -3
timeout 700 python bench.py > gpurun_out/r2c20_bench_default.json 2> gpurun_out/r2c20_bench_default.err
python -c "
import json; d=json.loads(open('gpurun_out/r2c20_bench_default.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','scaling','dtype','gpu_launches')})
print('e2e', d['e2e'], 'clocks', d['clocks'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
" || tail -5 gpurun_out/r2c20_bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c13_pytest_full.log 2>&1
tail -4 gpurun_out/r2c13_pytest_full.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c13_bench_default.json 2> gpurun_out/r2c13_bench_default.err
python -c "
import json; d=json.loads(open('gpurun_out/r2c13_bench_default.json').read().strip().splitlines()[-1])
print('bench', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'equal_work', d.get('equal_work_n1',{}).get('ms_per_step'))
print('pointops', {k:d['roofline_pointops'][k] for k in ('ms','frac','ref_cuda_ms')})
a=d['roofline_attention']; print({k:(round(a[k]['fwd_ms'],4), round(a[k].get('bwd_ms',0),4)) for k in a if isinstance(a[k],dict)})
" || tail -5 gpurun_out/r2c13_bench_default.err
timeout 400 python scripts/step_launch_list.py 2>&1 | tail -45 > gpurun_out/r2c13_cupti.log
head -3 gpurun_out/r2c13_cupti.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_step_launches_ncu.csv python scripts/step_launch_list.py --ncu > gpurun_out/r2c13_ncu.log 2>&1
echo ncu_rc=$? $(wc -l < gpurun_out/r2_step_launches_ncu.csv)

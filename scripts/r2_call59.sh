#!/bin/bash
# final single-GPU evidence pass of round 2 (re-run after the mask-view change)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c59_pytest_full.log 2>&1
tail -3 gpurun_out/r2c59_pytest_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('__SMOKE_OK__')" 2>&1 | tail -2
timeout 700 python bench.py > gpurun_out/r2c59_bench_default.json 2> gpurun_out/r2c59_bench_default.err
python -c "
import json; d=json.loads(open('gpurun_out/r2c59_bench_default.json').read().strip().splitlines()[-1])
print('bench', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'equal_work', d.get('equal_work_n1',{}).get('ms_per_step'), 'cpu', d['cpu_baseline']['value'])
print('roofline', d['roofline'])
print('pointops', {k:d['roofline_pointops'][k] for k in ('ms','frac','ref_cuda_ms')})
a=d['roofline_attention']; print({k:(round(a[k]['fwd_ms'],4), round(a[k].get('bwd_ms',0),4)) for k in a if isinstance(a[k],dict)})
" || tail -5 gpurun_out/r2c59_bench_default.err
for w in scanrefer objcls; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2c59_bench_$w.json 2> gpurun_out/r2c59_bench_$w.err
  python -c "
import json,sys; d=json.loads(open('gpurun_out/r2c59_bench_$w.json').read().strip().splitlines()[-1]); print('$w', d['ms_per_step'], d['value'], d['unit'], d['e2e']['value'])" || tail -3 gpurun_out/r2c59_bench_$w.err
done
timeout 400 python scripts/step_launch_list.py 2>&1 | tail -45 > gpurun_out/r2c59_cupti.log
head -3 gpurun_out/r2c59_cupti.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_step_launches_ncu.csv python scripts/step_launch_list.py --ncu > gpurun_out/r2c59_ncu.log 2>&1
echo ncu_rc=$? $(wc -l < gpurun_out/r2_step_launches_ncu.csv)

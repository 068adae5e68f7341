#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c8_pytest_full.log 2>&1
tail -4 gpurun_out/r2c8_pytest_full.log
timeout 600 python bench.py > gpurun_out/r2c8_bench_default.json 2> gpurun_out/r2c8_bench_default.err
python -c "
import json; d=json.loads(open('gpurun_out/r2c8_bench_default.json').read().strip().splitlines()[-1])
print('bench', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'equal_work', d.get('equal_work_n1'))
print('roofline', {k:d['roofline'][k] for k in ('ms','frac')}, 'pointops', {k:d['roofline_pointops'][k] for k in ('ms','frac','ref_cuda_ms')})
print('attention', json.dumps(d.get('roofline_attention'))[:1500])
print('cpu', d.get('cpu_baseline'))
" || tail -5 gpurun_out/r2c8_bench_default.err
for w in scanrefer objcls pointops_sweep; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2c8_bench_$w.json 2> gpurun_out/r2c8_bench_$w.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2c8_bench_$w.json').read().strip().splitlines()[-1])
print('$w', d['metric'], d['value'], d['unit'], d['ms_per_step'], d.get('final_loss'), d.get('gpu_launches'))" || tail -5 gpurun_out/r2c8_bench_$w.err
done

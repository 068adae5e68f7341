#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pointnet_module.py tests/test_pointops_gpu.py tests/test_attention_gpu.py::test_rerouted_bert_matches_huggingface_bert_gpu -q -s 2>&1 | grep -E "TRAINABLE|passed|failed|Error|error|assert|E  " | tail -20
timeout 400 python bench.py --workload objcls --no-cpu-baseline --no-kernel-rooflines > gpurun_out/r2c14_bench_objcls.json 2> gpurun_out/r2c14_bench_objcls.err
python -c "
import json; d=json.loads(open('gpurun_out/r2c14_bench_objcls.json').read().strip().splitlines()[-1])
print('objcls', d['value'], d['unit'], d['ms_per_step'], d.get('final_loss'), d.get('gpu_launches'))" || tail -8 gpurun_out/r2c14_bench_objcls.err

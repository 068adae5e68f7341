#!/bin/bash
# First GPU call of the next round: (1) the GPU replay of the reference-gradient fixture, (2) backward-GEMM shapes native vs
# cuBLAS, (3) the step with and without SVB200_NATIVE_BWD_GEMM.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_zz_reference_gradients_gpu.py -q -rxX 2>&1 | tail -5
timeout 200 python scripts/time_linear_bwd.py 2>&1 | tail -1 | tee gpurun_out/linear_bwd_shapes.json
for f in 0 1; do
  SVB200_NATIVE_BWD_GEMM=$f timeout 250 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-rooflines 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('native_bwd_gemm=$f', d['ms_per_step'], d['value'], d['final_loss'])"
done

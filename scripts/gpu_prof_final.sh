#!/bin/bash
# Round-end evidence: timings + one ncu --set full capture of each new-generation kernel (short single-GPU commands)
mkdir -p gpurun_out
timeout 120 python scripts/time_attention_train.py > gpurun_out/r1_attention_times.json 2>gpurun_out/err1.log; cat gpurun_out/r1_attention_times.json
timeout 120 python scripts/time_gemm_shapes.py > gpurun_out/r1_gemm_shapes.json 2>gpurun_out/err2.log; cat gpurun_out/r1_gemm_shapes.json
timeout 120 python scripts/time_layer_norm.py > gpurun_out/r1_layer_norm_times.json 2>gpurun_out/err3.log; cat gpurun_out/r1_layer_norm_times.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 3 -o gpurun_out/prof2_attn -f python scripts/time_attention_train.py quick > gpurun_out/prof2_attn.log 2>&1; tail -1 gpurun_out/prof2_attn.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 1 -o gpurun_out/prof2_gemm_lmhead -f python scripts/time_gemm_shapes.py quick lm_head > gpurun_out/prof2_gemm.log 2>&1; tail -1 gpurun_out/prof2_gemm.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_ -c 4 -o gpurun_out/prof2_ln -f python scripts/time_layer_norm.py quick > gpurun_out/prof2_ln.log 2>&1; tail -1 gpurun_out/prof2_ln.log

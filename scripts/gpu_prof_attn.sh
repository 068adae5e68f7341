#!/bin/bash
# attention fwd/bwd timing + one ncu --set full capture of each attention kernel at the joint-130 shape
mkdir -p gpurun_out
timeout 200 python scripts/time_attention_train.py > gpurun_out/attn_train_times.json 2> gpurun_out/attn_train_times.err
cat gpurun_out/attn_train_times.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention -c 3 -o gpurun_out/prof_attn -f python scripts/time_attention_train.py quick > gpurun_out/prof_attn.log 2>&1
tail -3 gpurun_out/prof_attn.log

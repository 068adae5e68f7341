#!/bin/bash
# ncu --set full of the GEMM family AFTER the issue-loop / TMA-store changes (heuristic tile pick; 5 launches)
mkdir -p gpurun_out
cat > /tmp/gemm5.py <<'P'
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from sceneverse_b200 import native
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return (torch.randn(*s, device="cuda", generator=g) * 0.5).bfloat16()
x, w, b = rnd(19200, 3072), rnd(768, 3072), torch.zeros(768, device="cuda")
gy, wq, xq = rnd(19200, 3072), rnd(3072, 768), rnd(19200, 768)
dw, db = torch.zeros(3072, 768, device="cuda"), torch.zeros(3072, device="cuda")
gj, wj = rnd(8320, 2304), rnd(2304, 768)
for _ in range(2):
    native.linear_fwd(x, w, b)                                    # bert_ffn2 forward
    native.linear_fwd(xq, wq, torch.zeros(3072, device="cuda"))   # bert_ffn1 forward (K = 768)
    native.linear_fwd(xq, wq, torch.zeros(3072, device="cuda"), act="gelu")   # + GELU epilogue
    native.linear_dgrad(gj, wj)                                   # joint in_proj dgrad
    native.linear_wgrad(gy, xq, dw=dw, db=db, accumulate=True)    # bert_ffn1 wgrad + bias gradient
torch.cuda.synchronize()
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel --launch-skip 5 -c 5 -o gpurun_out/r2_prof_gemm_final -f python /tmp/gemm5.py > gpurun_out/r2c54_ncu.log 2>&1
echo rc=$?; tail -2 gpurun_out/r2c54_ncu.log

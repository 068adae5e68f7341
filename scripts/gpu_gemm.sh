#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gemm_gpu.py -q -x > gpurun_out/pytest_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gemm.log
tail -30 gpurun_out/pytest_gemm.log | cut -c1-500
timeout -s KILL 200 python - <<'P' 2>&1 | tail -12
import sys; sys.path.insert(0,'.')
import torch
from sceneverse_b200 import native
def t(fn,n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(True); b=torch.cuda.Event(True); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n
for (M,N,K) in [(8320,2304,768),(8320,2048,768),(8320,768,2048),(3200,30522,768),(81920,256,272),(81920,512,256),(81920,768,512),(16384,4096,4096)]:
    a=torch.randn(M,K,device='cuda').bfloat16(); w=torch.randn(N,K,device='cuda').bfloat16()
    tm=t(lambda: native.gemm(a,w)); tt=t(lambda: a@w.t())
    print(f"{M}x{N}x{K}: mine {tm:.3f} ms {2*M*N*K/tm/1e9:.0f} TF/s | cublas {tt:.3f} ms {2*M*N*K/tt/1e9:.0f} TF/s")
P

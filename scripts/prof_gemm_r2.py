"""One launch per (shape, direction, CTA mode) of the native GEMM family, for `ncu --set full -k regex:gemm_kernel`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sceneverse_b200 import native
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return (torch.randn(*s, device="cuda", generator=g) * 0.5).bfloat16()
M, N, K = 19200, 768, 3072
x, w, b = rnd(M, K), rnd(N, K), torch.zeros(N, device="cuda")
gy, wq, xq = rnd(19200, 3072), rnd(3072, 768), rnd(19200, 768)
dw, db = torch.zeros(3072, 768, device="cuda"), torch.zeros(3072, device="cuda")
gj, wj = rnd(8320, 2304), rnd(2304, 768)
for c in (1, 2):
    native.gemm_force_ctas(c)
    native.linear_fwd(x, w, b)                                    # bert_ffn2 forward (A = 118 MB)
    native.linear_fwd(xq, wq, torch.zeros(3072, device="cuda"))   # bert_ffn1 forward
    native.linear_dgrad(gj, wj)                                   # joint in_proj dgrad (MN-major B)
    native.linear_wgrad(gy, xq, dw=dw, db=db, accumulate=True)    # bert_ffn1 wgrad (both MN-major)
    native.linear_wgrad(gy, xq, dw=dw, db=None, accumulate=True)  # same without the bias-gradient MMA
torch.cuda.synchronize()
print("done")

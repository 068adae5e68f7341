"""Native (sv_gemm_bf16_ex, transposed operands) vs cuBLAS for the two backward GEMMs of every linear shape of the step.
One JSON line; decides whether SVB200_NATIVE_BWD_GEMM should become the default."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sceneverse_b200 import native
native.gemm_force_ctas(int(os.environ.get("SVB200_GEMM_CTAS", "0")))
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return (torch.randn(*s, device="cuda", generator=g) * 0.5).bfloat16()
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(True); c=torch.cuda.Event(True); a.record()
    for _ in range(n): fn()
    c.record(); torch.cuda.synchronize(); return a.elapsed_time(c)/n
# (name, tokens M, out features N, in features Kin)
shapes = [("bert300_qkv", 19200, 768, 768), ("bert300_ffn1", 19200, 3072, 768), ("bert300_ffn2", 19200, 768, 3072),
          ("bert50_ffn1", 3200, 3072, 768), ("joint_in_proj", 8320, 2304, 768), ("joint_ffn1", 8320, 2048, 768),
          ("joint_ffn2", 8320, 768, 2048), ("spatial_ffn1", 5120, 2048, 768), ("spatial_qkv", 5120, 768, 768),
          ("lm_head", 3200, 30528, 768)]
res = {}
for name, M, N, Kin in shapes:
    gg, W, x = rnd(M, N), rnd(N, Kin), rnd(M, Kin)
    dw, db = torch.zeros(N, Kin, device="cuda"), torch.zeros(N, device="cuda")
    r = {"dgrad_native_ms": t(lambda: native.linear_dgrad(gg, W)), "dgrad_cublas_ms": t(lambda: gg @ W),
         "wgrad_native_ms": t(lambda: native.linear_wgrad(gg, x, dw=dw, db=db, accumulate=True)),
         "wgrad_cublas_ms": t(lambda: torch.mm(gg.t(), x, out_dtype=torch.float32))}
    res[name] = {k: round(v, 4) for k, v in r.items()}
tot = {k: round(sum(v[k] for v in res.values()), 3) for k in next(iter(res.values()))}
print(json.dumps({"shapes": res, "sum_ms": tot}))

"""native.gemm vs cuBLAS at the shapes of the step (LM head, SA3 chain, fc)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sceneverse_b200 import native
native.gemm_force_ctas(int(os.environ.get("SVB200_GEMM_CTAS", "0")))
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
def t(fn, n=10):
    if quick:
        fn(); torch.cuda.synchronize(); return 0.0
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(True); c=torch.cuda.Event(True); a.record()
    for _ in range(n): fn()
    c.record(); torch.cuda.synchronize(); return a.elapsed_time(c)/n
g = torch.Generator(device="cuda").manual_seed(0)
res = {}
ALL = [
    ("lm_head", 3200, 30528, 768, None, 0), ("sa3_l0", 81920, 256, 272, "relu", 0), ("sa3_l1", 81920, 512, 256, "relu", 0),
    ("sa3_l2", 81920, 768, 512, "relu", 16), ("fc", 5120, 768, 768, None, 0), ("ffn1", 8320, 2048, 768, "relu", 0),
    ("bert_ffn1", 19200, 3072, 768, "gelu", 0), ("bert_ffn1_noact", 19200, 3072, 768, None, 0),
    ("bert_ffn2", 19200, 768, 3072, None, 0), ("bert_qkv", 19200, 2304, 768, None, 0), ("joint_qkv", 8320, 2304, 768, None, 0),
    ("spatial_q", 5120, 768, 768, None, 0), ("ffn2", 8320, 768, 2048, None, 0)]
shapes = [x for x in ALL if x[0] == (sys.argv[2] if len(sys.argv) > 2 else "lm_head")] if quick else ALL
for name, M, N, K, act, rowmax in shapes:
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    ms = t(lambda: native.gemm(a, w, bias=b, act=act, rowmax=rowmax))
    res[name] = {"native_ms": round(ms, 4), "native_tflops": round(2 * M * N * K / max(ms, 1e-9) / 1e9, 1)}
    if not quick:
        ms2 = t(lambda: torch.nn.functional.linear(a, w, b.bfloat16()))
        res[name].update(cublas_ms=round(ms2, 4), cublas_tflops=round(2 * M * N * K / ms2 / 1e9, 1))
        if act:   # the library path pays a second kernel for the activation the native epilogue fuses
            fa = torch.nn.functional.relu if act == "relu" else torch.nn.functional.gelu
            res[name]["cublas_plus_act_ms"] = round(t(lambda: fa(torch.nn.functional.linear(a, w, b.bfloat16()))), 4)
print(json.dumps(res))

"""One-shot check of the transposed-operand (MN-major) GEMM variants against torch, plus a timing at the BERT FFN
backward shapes.  Prints one JSON line."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sceneverse_b200 import native
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return (torch.randn(*s, device="cuda", generator=g) * 0.5).bfloat16()
res = {"cases": []}
ok = True
for (M, N, K) in [(128, 64, 64), (200, 136, 72), (384, 768, 3072), (777, 320, 1000), (64, 2048, 130)]:
    for at, bt in [(False, True), (True, True), (True, False)]:
        A = rnd(K, M) if at else rnd(M, K)
        B = rnd(K, N) if bt else rnd(N, K)
        want = (A.float().t() if at else A.float()) @ (B.float() if bt else B.float().t())
        try:
            got = native.gemm_ex(A, B, a_transposed=at, b_transposed=bt, out_dtype=torch.float32)
            err = (got - want).abs().max().item() / (want.abs().max().item() + 1e-9)
        except Exception as e:
            err = repr(e)[:80]
        good = isinstance(err, float) and err < 2e-2
        ok = ok and good
        res["cases"].append([M, N, K, int(at), int(bt), err if not isinstance(err, float) else round(err, 5)])
res["ok"] = ok
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(True); c=torch.cuda.Event(True); a.record()
    for _ in range(n): fn()
    c.record(); torch.cuda.synchronize(); return a.elapsed_time(c)/n
if ok:
    M, N, K = 19200, 3072, 768        # BERT FFN1: dgrad (M x K) = g (M x N) . W (N x K);  wgrad (N x K) = g^T . x
    gg, W, x = rnd(M, N), rnd(N, K), rnd(M, K)
    res["dgrad_native_ms"] = round(t(lambda: native.gemm_ex(gg, W, b_transposed=True)), 4)
    res["dgrad_cublas_ms"] = round(t(lambda: gg @ W), 4)
    res["wgrad_native_ms"] = round(t(lambda: native.gemm_ex(gg, x, a_transposed=True, b_transposed=True, out_dtype=torch.float32)), 4)
    res["wgrad_cublas_ms"] = round(t(lambda: torch.mm(gg.t(), x, out_dtype=torch.float32)), 4)
print(json.dumps(res))

"""Correctness + timing of the multicast-cluster GEMM variant (sv_gemm_force_ctas(4)): CTA pairs in clusters of two
sharing the B tile, against the plain CTA-pair kernel (2) and an fp32 reference."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sceneverse_b200 import native
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return (torch.randn(*s, device="cuda", generator=g) * 0.5).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, n=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]
out = {}
for (M, N, K) in ((1024, 512, 512), (1000, 600, 520), (19200, 768, 3072), (19200, 3072, 768), (8320, 2304, 768)):
    x, w, b = rnd(M, K), rnd(N, K), torch.randn(N, device="cuda", generator=g)
    gy = rnd(M, N)
    ref_y = (x.float() @ w.float().t() + b)
    ref_dx = gy.float() @ w.float()
    ref_dw = gy.float().t() @ x.float()
    row = {}
    for force in (2, 4):
        native.gemm_force_ctas(force)
        y = native.linear_fwd(x, w, b); torch.cuda.synchronize()
        dx = native.linear_dgrad(gy, w); torch.cuda.synchronize()
        dw = torch.zeros(N, K, device="cuda"); db = torch.zeros(N, device="cuda")
        native.linear_wgrad(gy, x, dw=dw, db=db, accumulate=True); torch.cuda.synchronize()
        e = lambda a, r: float((a.float() - r).abs().max() / r.abs().max())
        row[f"err{force}"] = (e(y, ref_y), e(dx, ref_dx), e(dw, ref_dw), e(db, gy.float().sum(0)))
        print(M, N, K, "force", force, "rel err fwd/dgrad/wgrad/dbias", row[f"err{force}"], flush=True)
    if M >= 8000:
        for force in (2, 4):
            native.gemm_force_ctas(force)
            dwz = torch.zeros(N, K, device="cuda")
            row[f"ms{force}"] = (round(timed(lambda: native.linear_fwd(x, w, b)), 4), round(timed(lambda: native.linear_dgrad(gy, w)), 4),
                                 round(timed(lambda: native.linear_wgrad(gy, x, dw=dwz, db=None, accumulate=True)), 4))
            print(M, N, K, "force", force, "ms fwd/dgrad/wgrad", row[f"ms{force}"], flush=True)
    out[f"{M}x{N}x{K}"] = row
native.gemm_force_ctas(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r2_gemm_multicast.json", "w"), indent=1)

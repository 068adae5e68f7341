"""Fused dropout+residual+LayerNorm vs the ATen sequence (dropout, add, layer_norm) at the shapes of the step."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from sceneverse_b200 import ops
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
def t(fn, n=20):
    if quick:
        fn(); torch.cuda.synchronize(); return 0.0
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(True); c=torch.cuda.Event(True); a.record()
    for _ in range(n): fn()
    c.record(); torch.cuda.synchronize(); return a.elapsed_time(c)/n
g = torch.Generator(device="cuda").manual_seed(0)
res = {}
for name, R in ([("bert300", 19200)] if quick else [("bert300", 19200), ("joint130", 8320), ("obj80", 5120), ("txt50", 3200)]):
    D = 768
    x = torch.randn(R, D, device="cuda", generator=g).bfloat16().requires_grad_(True)
    r = torch.randn(R, D, device="cuda", generator=g).bfloat16().requires_grad_(True)
    w = torch.ones(D, device="cuda", requires_grad=True); b = torch.zeros(D, device="cuda", requires_grad=True)
    go = torch.randn(R, D, device="cuda", generator=g).bfloat16()
    def nat():
        y = ops._LayerNormFn.apply(x, r, w, b, 1e-5, 0.1, 7)
        y.backward(go)
    def ref():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = F.layer_norm(r + F.dropout(x, 0.1, True), (D,), w, b, 1e-5)
        y.backward(go.float())
    fwd = t(lambda: ops._LayerNormFn.apply(x.detach(), r.detach(), w.detach(), b.detach(), 1e-5, 0.1, 7))
    both = t(nat)
    res[name] = {"native_fwd_ms": round(fwd, 4), "native_fwd_bwd_ms": round(both, 4),
                 "fwd_GBps_algorithmic": round(8 * D * R / max(fwd, 1e-9) / 1e6, 1),
                 "bwd_GBps_algorithmic": round(8 * D * R / max(both - fwd, 1e-9) / 1e6, 1)}
    if not quick:
        res[name]["aten_fwd_bwd_ms"] = round(t(ref), 4)
print(json.dumps(res))

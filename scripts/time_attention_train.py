"""Forward+backward timing of the native attention kernels against SDPA at the GPS training shapes (B = 64 scenes)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from sceneverse_b200 import native, ops
B, H, E = 64, 12, 768
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
def t(fn, n=20):
    if quick:
        fn(); torch.cuda.synchronize(); return 0.0
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(True); c=torch.cuda.Event(True); a.record()
    for _ in range(n): fn()
    c.record(); torch.cuda.synchronize(); return a.elapsed_time(c)/n
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return torch.randn(*s, device="cuda", generator=g)
res = {}
for name, Lq, Lk in [("joint130", 130, 130), ("self80", 80, 80), ("cross80x50", 80, 50)]:
    q = rnd(B, Lq, E).bfloat16(); k = rnd(B, Lk, E).bfloat16(); v = rnd(B, Lk, E).bfloat16(); go = rnd(B, Lq, E).bfloat16()
    for p in (0.0, 0.1):
        res[f"{name}_fwd_p{p}"] = t(lambda: native.attention(q, k, v, H, return_lse=True, dropout_p=p, seed=5))
        out, lse = native.attention(q, k, v, H, return_lse=True, dropout_p=p, seed=5)
        res[f"{name}_bwd_p{p}"] = t(lambda: native.attention_backward(q, k, v, out, go, lse, H, dropout_p=p, seed=5))
        if quick: continue
        qh, kh, vh = (x.view(B, -1, H, 64).transpose(1, 2).detach().requires_grad_(True) for x in (q, k, v))
        goh = go.view(B, Lq, H, 64).transpose(1, 2)
        res[f"{name}_sdpa_fwd_p{p}"] = t(lambda: F.scaled_dot_product_attention(qh, kh, vh, dropout_p=p))
        def fb():
            o = F.scaled_dot_product_attention(qh, kh, vh, dropout_p=p)
            o.backward(goh)
        res[f"{name}_sdpa_fwdbwd_p{p}"] = t(fb)
print(json.dumps({k: round(v, 4) for k, v in res.items()}))

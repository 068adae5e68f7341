import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from sceneverse_b200 import ops, native
def rnd(*s, seed=0, scale=0.5):
    g = torch.Generator(device="cuda").manual_seed(seed + sum(s)); return torch.randn(*s, device="cuda", generator=g) * scale
def rel(a, b): return ((a.float()-b.float()).abs().max()/(b.float().abs().max()+1e-9)).item()
for M in (512, 520, 640, 1000):
    g2 = rnd(M, 2048).bfloat16(); w = rnd(2048, 768, scale=0.03).bfloat16()
    for force in (0, 1, 2):
        native.gemm_force_ctas(force)
        try:
            dx = native.linear_dgrad(g2, w, out_dtype=torch.float32)
            print("dgrad M", M, "force", force, "err", rel(dx, g2.float() @ w.float()))
        except Exception as e:
            print("dgrad M", M, "force", force, "EXC", e)
native.gemm_force_ctas(0)
for force in (0, 1):
    native.gemm_force_ctas(force)
    x = rnd(4, 130, 768, seed=1).requires_grad_(True); W = (rnd(2048, 768, seed=2) * 0.1).requires_grad_(True); b = (rnd(2048, seed=3) * 0.2).requires_grad_(True)
    go = rnd(4, 130, 2048, seed=4).bfloat16()
    for act in (None, "relu", "gelu"):
        try:
            for t in (x, W, b): t.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = ops.linear(x, W, b, activation=act)
            y.backward(go)
            got = [y.detach().float(), x.grad.clone(), W.grad.clone(), b.grad.clone()]
            xf = x.detach().bfloat16().float().requires_grad_(True); wf = W.detach().bfloat16().float().requires_grad_(True); bf = b.detach().clone().requires_grad_(True)
            pre = F.linear(xf, wf, bf); yr = pre if act is None else (torch.relu(pre) if act == "relu" else F.gelu(pre))
            yr.backward(go.float())
            print("force", force, "act", act, "y", rel(got[0], yr), "dx", rel(got[1], xf.grad), "dW", rel(got[2], wf.grad), "db", rel(got[3], bf.grad))
            if act == "relu":
                yb = y.detach().reshape(-1, 2048)
                print("   mask agree", ((yb > 0) == (yr.reshape(-1, 2048) > 0)).float().mean().item())
                g2 = native.act_bwd(go.reshape(-1, 2048).contiguous(), yb.contiguous(), "relu")
                print("   act_bwd vs torch", rel(g2, go.reshape(-1, 2048).float() * (yb > 0)))
                dx2 = native.linear_dgrad(g2, W.detach().bfloat16(), out_dtype=torch.float32)
                print("   dgrad(act_bwd) vs fp32", rel(dx2, g2.float() @ W.detach().bfloat16().float()), "vs ref", rel(dx2, xf.grad.reshape(-1, 768)))
        except Exception as e:
            traceback.print_exc()

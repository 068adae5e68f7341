import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from sceneverse_b200 import ops, native
def rnd(*s, seed=0, scale=0.5):
    g = torch.Generator(device="cuda").manual_seed(seed + sum(s)); return torch.randn(*s, device="cuda", generator=g) * scale
def rel(a, b): return ((a.float()-b.float()).abs().max()/(b.float().abs().max()+1e-9)).item()
for force in (0, 1, 2):
    native.gemm_force_ctas(force)
    x = rnd(4, 130, 768, seed=1).requires_grad_(True); W = (rnd(2048, 768, seed=2) * 0.1).requires_grad_(True); b = (rnd(2048, seed=3) * 0.2).requires_grad_(True)
    go = rnd(4, 130, 2048, seed=4).bfloat16()
    for act in (None, "relu", "gelu"):
        for t in (x, W, b): t.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = ops.linear(x, W, b, activation=act)
        y.backward(go)
        got = [y.detach().float(), x.grad.clone(), W.grad.clone(), b.grad.clone()]
        for t in (x, W, b): t.grad = None
        xf = x.detach().bfloat16().float().requires_grad_(True); wf = W.detach().bfloat16().float().requires_grad_(True); bf = b.detach().clone().requires_grad_(True)
        pre = F.linear(xf, wf, bf); yr = pre if act is None else (torch.relu(pre) if act == "relu" else F.gelu(pre))
        yr.backward(go.float())
        # the same with the mask taken from the bf16 output (what the kernel sees)
        print("force", force, "act", act, "y", rel(got[0], yr), "dx", rel(got[1], xf.grad), "dW", rel(got[2], wf.grad), "db", rel(got[3], bf.grad))
        if act == "relu":
            gm = go.float() * (yr > 0)
            print("   mask agree", ((got[0] > 0) == (yr > 0)).float().mean().item(), "dx(fp32 ref with masked g)", rel(got[1], gm.reshape(-1, 2048) @ wf.detach()))
            g2 = native.act_bwd(go.reshape(-1, 2048).contiguous(), y.detach().reshape(-1, 2048).contiguous(), "relu")
            print("   act_bwd vs torch", rel(g2, go.reshape(-1,2048).float() * (y.detach().reshape(-1,2048) > 0)))
            dx2 = native.linear_dgrad(g2, W.detach().bfloat16(), out_dtype=torch.float32)
            print("   dgrad of act_bwd output vs fp32", rel(dx2, g2.float() @ W.detach().bfloat16().float()))

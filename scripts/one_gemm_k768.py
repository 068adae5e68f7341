import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sceneverse_b200 import native
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return (torch.randn(*s, device="cuda", generator=g) * 0.5).bfloat16()
x1, w1, b1 = rnd(19200, 768), rnd(3072, 768), torch.zeros(3072, device="cuda")
native.gemm_force_ctas(2)
for _ in range(3): native.linear_fwd(x1, w1, b1)
torch.cuda.synchronize()

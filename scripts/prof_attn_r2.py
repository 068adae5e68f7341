import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sceneverse_b200 import native
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return (torch.randn(*s, device="cuda", generator=g) * 0.5).bfloat16()
B, H, D = 64, 12, 768
for L in (300, 130):
    q, k, v, go = rnd(B, L, D), rnd(B, L, D), rnd(B, L, D), rnd(B, L, D)
    o, lse = native.attention(q, k, v, H, return_lse=True, dropout_p=0.1, seed=5)
    native.attention_backward(q, k, v, o, go, lse, H, dropout_p=0.1, seed=5)
torch.cuda.synchronize()
print("done")

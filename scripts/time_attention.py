"""Timing of the fused attention / CE / pairwise kernels at the GPS shapes (B = 64 scenes)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sceneverse_b200 import native, ops
B, H, E = 64, 12, 768
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(True); c=torch.cuda.Event(True); a.record()
    for _ in range(n): fn()
    c.record(); torch.cuda.synchronize(); return a.elapsed_time(c)/n
res = {}
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return torch.randn(*s, device="cuda", generator=g)
# spatial self-attention over 80 objects
q, k, v = (rnd(B, 80, E).bfloat16() for _ in range(3))
sw = rnd(B, 80, 72); locs = ops.calc_pairwise_locs(rnd(B, 80, 3), None); mask = torch.zeros(B, 80, dtype=torch.bool, device="cuda"); mask[:, 60:] = True
res["spatial_attn_80_ms"] = t(lambda: native.attention(q, k, v, H, key_padding_mask=mask, spatial_w=sw, spatial_heads=H, pairwise_locs=locs))
res["spatial_attn_80_torch_ms"] = t(lambda: ops._spatial_attention_torch(q, k, v, sw.bfloat16(), locs.bfloat16(), H, H, key_padding_mask=mask))
# joint self-attention over 130 tokens, cross attention 80x50
qkv = rnd(B, 130, 3 * E).bfloat16(); qj, kj, vj = qkv.split(E, dim=-1); m130 = torch.zeros(B, 130, dtype=torch.bool, device="cuda")
res["joint_attn_130_ms"] = t(lambda: native.attention(qj, kj, vj, H, key_padding_mask=m130))
res["joint_attn_130_sdpa_ms"] = t(lambda: torch.nn.functional.scaled_dot_product_attention(qj.reshape(B,130,H,64).transpose(1,2), kj.reshape(B,130,H,64).transpose(1,2), vj.reshape(B,130,H,64).transpose(1,2)))
kc, vc = rnd(B, 50, E).bfloat16(), rnd(B, 50, E).bfloat16()
res["cross_attn_80x50_ms"] = t(lambda: native.attention(q, kc, vc, H))
flops = lambda Lq, Lk: 4 * B * H * Lq * Lk * 64
res["cross_attn_80x50_tflops"] = flops(80, 50) / res["cross_attn_80x50_ms"] / 1e9
res["joint_attn_130_tflops"] = flops(130, 130) / res["joint_attn_130_ms"] / 1e9
res["pairwise_locs_ms"] = t(lambda: ops.calc_pairwise_locs(rnd(B, 80, 3), None))
logits = rnd(3200, 30522).bfloat16().requires_grad_(True); labels = torch.randint(0, 30522, (3200,), device="cuda"); labels[torch.rand(3200, device="cuda") < 0.85] = -1
res["fused_ce_fwd_ms"] = t(lambda: ops.cross_entropy(logits, labels, ignore_index=-1))
res["torch_ce_fwd_ms"] = t(lambda: torch.nn.functional.cross_entropy(logits.float(), labels, ignore_index=-1))
print(json.dumps(res))

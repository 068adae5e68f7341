"""Experiment: whole-step CUDA graph capture at N=1 (fwd + loss + bwd + clip + AdamW), timing vs eager."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sceneverse_b200 import model as M, train, weights
dev = torch.device("cuda", 0)
b = bench.make_scene_batches(1, bench.SCENES, 42)[0]
static = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
tf = weights.synthetic_tensor("text_features", (607, 768))
ps = train.PretrainStep(M.pretrain_config(1, text_features=tf), dev, dtype=torch.bfloat16, seed=1)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(True); c=torch.cuda.Event(True); a.record()
    for _ in range(n): fn()
    c.record(); torch.cuda.synchronize(); return a.elapsed_time(c)/n
print("eager ms/step", t(lambda: ps.step(dict(static))))
# capturable optimizer: rebuild AdamW with capturable=True and tensor lr
groups = [{k: v for k, v in g.items() if k != 'params'} | {'params': g['params']} for g in ps.optimizer.param_groups]
for g in groups:
    g['lr'] = torch.tensor(float(g['lr']), device=dev)
    g.pop('initial_lr', None); g.pop('fused', None); g.pop('capturable', None)
opt = torch.optim.AdamW(groups, betas=(0.9, 0.98), fused=True, capturable=True)
params = ps.parameters()
def raw_step():
    opt.zero_grad(set_to_none=False)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        total, _ = ps.module(dict(static))
    total.backward()
    torch.nn.utils.clip_grad_norm_(params, 5.0, foreach=True)
    opt.step()
    return total.detach()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): raw_step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
print("eager (capturable opt) ms/step", t(raw_step))
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        out = raw_step()
    torch.cuda.synchronize()
    print("graph ms/step", t(lambda: g.replay()), "loss", float(out))
except Exception as e:
    print("capture failed:", repr(e)[:600])

"""Which Python lines issue the ATen elementwise / copy ops of one eager GPS pre-training step (TorchDispatchMode + the
innermost sceneverse_b200 frames); complements aten_attribution.py, which has the device times per (op, shapes)."""
import os, sys, collections, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.utils._python_dispatch import TorchDispatchMode
import bench
from sceneverse_b200 import model as M, train, weights
dev = torch.device("cuda", 0)
b = bench.make_scene_batches(1, bench.SCENES, 42)[0]
batch = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
tf = weights.synthetic_tensor("text_features", (607, 768))
ps = train.PretrainStep(M.pretrain_config(1, text_features=tf), dev, dtype=torch.bfloat16, seed=1, cuda_graph=True)
for _ in range(2):
    ps.step(dict(batch))
torch.cuda.synchronize()
agg = collections.Counter()
class Log(TorchDispatchMode):
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        name = str(func).replace("aten.", "")
        t = next((a for a in args if isinstance(a, torch.Tensor)), None)
        if t is not None and t.numel() >= 64 * 50 * 768 and any(k in name for k in ("copy", "add", "div", "mul", "clone", "cat", "sum", "fill", "zero")):
            fr = [f for f in traceback.extract_stack() if "sceneverse_b200" in f.filename and "scripts" not in f.filename]
            where = " < ".join(f"{os.path.basename(f.filename)}:{f.lineno}:{f.name}" for f in fr[-3:][::-1]) or "<no python frame>"
            agg[(name, tuple(t.shape), str(t.dtype).replace("torch.", ""), where)] += 1
        return out
with Log():
    ps._raw_step()
torch.cuda.synchronize()
for (name, shape, dt, where), n in sorted(agg.items(), key=lambda kv: (-kv[1] * (kv[0][1][0] * kv[0][1][-1] if len(kv[0][1]) else 1))):
    print(f"x{n:3d} {name:22s} {str(shape):22s} {dt:9s} {where}")

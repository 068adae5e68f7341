"""Which ATen ops (by CPU op name and input shapes) own the library kernel time of one eager GPS pre-training step."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sceneverse_b200 import model as M, train, weights
dev = torch.device("cuda", 0)
b = bench.make_scene_batches(1, bench.SCENES, 42)[0]
batch = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
tf = weights.synthetic_tensor("text_features", (607, 768))
ps = train.PretrainStep(M.pretrain_config(1, text_features=tf), dev, dtype=torch.bfloat16, seed=1, cuda_graph=True)
for _ in range(2):
    ps.step(dict(batch))          # capture happened: flat state, packs, shadows are live
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    ps._raw_step()                # the same step eagerly (what the graph recorded)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.key.startswith("aten::") and e.self_device_time_total > 0:
        rows.append((e.self_device_time_total / 1e3, e.count, e.key, str(e.input_shapes)[:140]))
rows.sort(reverse=True)
print("total aten self cuda ms", sum(r[0] for r in rows))
for r in rows[:45]:
    print(f"{r[0]:7.3f} ms x{r[1]:4d} {r[2]:28s} {r[3]}")

"""torch.profiler breakdown of one GPS pre-training step (CUDA kernel time by name)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sceneverse_b200 import model as M, train, weights
dev = torch.device("cuda", 0)
b = bench.make_scene_batches(1, bench.SCENES, 42)[0]
batch = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
tf = weights.synthetic_tensor("text_features", (607, 768))
ps = train.PretrainStep(M.pretrain_config(1, text_features=tf), dev, dtype=torch.bfloat16, seed=1)
for _ in range(3): ps.step(dict(batch))
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity, record_function
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2): ps.step(dict(batch))
    torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted([(e.device_time_total / 2e3, e.count // 2, e.key) for e in ev if e.device_time_total > 0], reverse=True)
tot = sum(r[0] for r in rows if not r[2].startswith("aten::") and not r[2].startswith("autograd"))
print("total kernel ms/step (sum of leaf kernels):", tot)
for r in rows[:60]: print(f"{r[0]:9.3f} ms  x{r[1]:5d}  {r[2][:110]}")
# per-phase timing with events
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(True); c=torch.cuda.Event(True); a.record()
    for _ in range(n): fn()
    c.record(); torch.cuda.synchronize(); return a.elapsed_time(c)/n
m = ps.module.model
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    print("bert50 fwd ms", t(lambda: m.lang_encoder(batch['txt_ids'], batch['txt_masks'])))
    print("bert300 fwd ms", t(lambda: m.lang_encoder(batch['scene_txt_ids'], batch['scene_txt_masks'])))
    print("point_encoder fwd ms", t(lambda: m.point_encoder(batch['obj_fts'], batch['obj_locs'], batch['obj_masks'], batch['obj_sem_masks'])))
    print("pointnet fused fwd ms", t(lambda: m.point_encoder.point_feature_extractor(batch['obj_fts'].view(-1,1024,6))))

"""Kernel-only launch list of the graph-replayed GPS pre-training step (bench.py's default workload).
  ncu mode   : ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \\
                   --log-file gpurun_out/r2_step_launches_ncu.csv python scripts/step_launch_list.py --ncu
               (one replay between cudaProfilerStart/Stop; per-launch times are cold-cache and serialised)
  cupti mode : python scripts/step_launch_list.py   -> gpurun_out/r2_step_launches_cupti.json
               (torch.profiler kernel activities of 3 warm replays: per-kernel-name count and mean time per step)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from sceneverse_b200 import model as M, train, weights

ncu = "--ncu" in sys.argv
dev = torch.device("cuda", 0)
b = bench.make_scene_batches(1, bench.SCENES, 42)[0]
batch = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
tf = weights.synthetic_tensor("text_features", (607, 768))
ps = train.PretrainStep(M.pretrain_config(1, text_features=tf), dev, dtype=torch.bfloat16, seed=1, cuda_graph=True)
for _ in range(3):
    ps.step(dict(batch))
torch.cuda.synchronize()
assert ps.graph is not None
if ncu:
    torch.cuda.profiler.start()
    ps.step(dict(batch))
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ps.step(dict(batch))
e1.record()
torch.cuda.synchronize()
ms_step = e0.elapsed_time(e1) / 10
from torch.profiler import ProfilerActivity, profile
R = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(R):
        ps.step(dict(batch))
    torch.cuda.synchronize()
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA and ev.device_time_total > 0:
        r = rows.setdefault(ev.name, [0, 0.0])
        r[0] += 1
        r[1] += ev.device_time_total
out = sorted(({"kernel": k, "launches_per_step": v[0] / R, "ms_per_step": v[1] / R / 1e3} for k, v in rows.items()),
             key=lambda r: -r["ms_per_step"])
json.dump({"ms_per_step_events": ms_step, "sum_kernel_ms_per_step": sum(r["ms_per_step"] for r in out), "kernels": out},
          open("gpurun_out/r2_step_launches_cupti.json", "w"), indent=1)
print("ms/step", ms_step, "sum of kernels", sum(r["ms_per_step"] for r in out), "distinct", len(out))
for r in out[:40]:
    print(f"{r['ms_per_step']:8.3f} ms x{r['launches_per_step']:7.1f}  {r['kernel'][:120]}")

"""Is the step host-bound?  Host enqueue time per step (no sync inside) vs device time per step."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sceneverse_b200 import model as M, train, weights
dev = torch.device("cuda", 0)
b = bench.make_scene_batches(1, bench.SCENES, 42)[0]
batch = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
tf = weights.synthetic_tensor("text_features", (607, 768))
ps = train.PretrainStep(M.pretrain_config(1, text_features=tf), dev, dtype=torch.bfloat16, seed=1)
for _ in range(5): ps.step(dict(batch))
torch.cuda.synchronize()
N = 10
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
host = []
e0.record()
t_all = time.perf_counter()
for _ in range(N):
    t0 = time.perf_counter(); ps.step(dict(batch)); host.append(time.perf_counter() - t0)
e1.record(); torch.cuda.synchronize()
wall = (time.perf_counter() - t_all) / N * 1e3
print(json.dumps({"host_enqueue_ms_per_step": [round(h * 1e3, 2) for h in host], "device_ms_per_step": round(e0.elapsed_time(e1) / N, 2),
                  "wall_ms_per_step": round(wall, 2)}))
# phase split of the host time
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable()
for _ in range(3): ps.step(dict(batch))
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35); print(s.getvalue()[:6000])

"""torchrun --nproc-per-node 2 scripts/dp_overlap_2rank.py : the data-parallel step with the gradient all-reduce split in two and
overlapped with the text encoder's backward (both collectives captured in the CUDA graph) must produce the same parameters as
the plain graph | one all-reduce | graph step, and every rank must hold the same parameters."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
from sceneverse_b200 import model as M, synthetic, train, weights, ops
tf = weights.synthetic_tensor("text_features", (607, 768))
d = synthetic.scene_batch(100 + rank, B=4, O=80, P=1024, L=50, Ls=300)
batch = {k: torch.from_numpy(v).to(dev) for k, v in d.items()}

def run(overlap):
    cfg = M.pretrain_config(world, text_features=tf)
    cfg["solver"]["sched"]["args"]["warmup_steps"] = 0
    ps = train.PretrainStep(cfg, dev, dtype=torch.bfloat16, seed=7, cuda_graph=True, overlap_allreduce=overlap)
    for m in ps.module.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if hasattr(m, "dropout") and isinstance(m.dropout, float):
            m.dropout = 0.0
    # the overlapped capture runs one more eager warm-up step than the plain one: align the two runs on the optimisation step
    losses = [float(ps.step(dict(batch))) for _ in range(3 if overlap else 4)]
    losses = losses[-3:]
    params = {n: p.detach().float().clone() for n, p in ps.module.named_parameters() if p.requires_grad}
    info = (ps.overlapped, ps.dp_graph)
    ps.close(); ops.clear_shadows()
    return losses, params, info

l0, p0, i0 = run(False)
l1, p1, i1 = run(True)
worst = 0.0
for n in p0:
    den = p0[n].abs().max().item() + 1e-12
    worst = max(worst, (p0[n] - p1[n]).abs().max().item() / den)
# replicas agree
chk = torch.stack([p.double().sum() for p in p1.values()]).sum().reshape(1)
both = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(both, chk)
same = all(abs(float(b) - float(both[0])) <= 1e-6 * abs(float(both[0])) for b in both)
if rank == 0:
    print(f"plain {i0} losses {l0}\noverlap {i1} losses {l1}\nworst relative parameter difference {worst:.3e}  replicas_equal {same}")
    print("DP_OVERLAP_OK=%s" % (i1[0] and worst < 2e-2 and same and all(abs(a - b) < 5e-2 for a, b in zip(l0, l1))))
dist.barrier(); dist.destroy_process_group()

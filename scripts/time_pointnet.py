"""Per-stage CUDA-event timing of the fused PointNet++ path at the model shape (5120 clouds)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sceneverse_b200 import synthetic, weights, _lib
from sceneverse_b200.modules.pointnet import PointNetPP, GPS_SPEC
from sceneverse_b200.pointnet2 import _ext

B = int(sys.argv[1]) if len(sys.argv) > 1 else 5120
net = PointNetPP(**GPS_SPEC).eval()
net.load_state_dict(weights.synthetic_state_dict(net, 0)); net = net.cuda()
base = synthetic.object_batch(5, 256, 1024, 0.3)
x = torch.from_numpy(np.tile(base, (B // 256 + 1, 1, 1))[:B]).cuda()
pk = net._pack(); lib = _lib.gps(); st = torch.cuda.current_stream().cuda_stream
xyz = x[..., :3].contiguous()

def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

res = {}
res["xyz_copy_ms"] = t(lambda: x[..., :3].contiguous())
out = _ext.sa_sample2(xyz, 32, 0.2, 32, 16, 0.4, 32)
res["sa_sample2_ms"] = t(lambda: _ext.sa_sample2(xyz, 32, 0.2, 32, 16, 0.4, 32))
fi1, nx1, bi1, fi2, nx2, bi2 = out
feat1 = torch.empty((B, 32, 128), dtype=torch.bfloat16, device="cuda")
feat2 = torch.empty((B, 16, 256), dtype=torch.bfloat16, device="cuda")
f1 = lambda: _lib.check(lib, lib.sv_sa1_mlp_bf16(x.data_ptr(), nx1.data_ptr(), bi1.data_ptr(), pk["sa1"].data_ptr(), B, 1024, 32, feat1.data_ptr(), st), "sa1")
f2 = lambda: _lib.check(lib, lib.sv_sa2_mlp_bf16(nx1.data_ptr(), feat1.data_ptr(), nx2.data_ptr(), bi2.data_ptr(), pk["sa2"].data_ptr(), B, 32, 32, feat2.data_ptr(), st), "sa2")
res["sa1_mlp_ms"] = t(f1); res["sa2_mlp_ms"] = t(f2)
res["sa1_mlp_tflops"] = 2 * B * 1024 * (16 * 64 + 64 * 64 + 64 * 128) / res["sa1_mlp_ms"] / 1e9
res["sa2_mlp_tflops"] = 2 * B * 512 * (144 * 128 + 128 * 128 + 128 * 256) / res["sa2_mlp_ms"] / 1e9
res["forward_fused_ms"] = t(lambda: net.forward_fused(x))
res["clouds"] = B
print(json.dumps(res))

import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sceneverse_b200 import _lib
lib = _lib.gps()
out = torch.zeros(148, dtype=torch.int64, device="cuda")
res = {}
for blocks in (1, 148):
    for N in (64, 128, 256):
        for a_mn, b_mn in ((0, 0), (0, 1), (1, 0), (1, 1)):
            iters = 2000
            for _ in range(2):
                _lib.check(lib, lib.sv_mma_bench(N, a_mn, b_mn, iters, blocks, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "bench")
                torch.cuda.synchronize()
            cyc = float(out[:blocks].double().mean()) / (iters * 4)
            res[f"blocks{blocks}_N{N}_a{'MN' if a_mn else 'K'}_b{'MN' if b_mn else 'K'}"] = {"cycles_per_mma": round(cyc, 1), "ideal_math_cycles": N / 2}
print(json.dumps(res, indent=1))

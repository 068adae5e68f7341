"""tcgen05.mma issue-rate micro-benchmark (csrc/mma_bench.cu): SM cycles per MMA by operand layout, and — with all 148
SMs busy — the SM clock the chip actually holds (cycles / globaltimer ns) with zero vs random operand bits and with the
GEMM main loop's commit-per-stage pattern."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sceneverse_b200 import _lib
lib = _lib.gps()
out = torch.zeros(2 * 148, dtype=torch.int64, device="cuda")
res = {}
def run(N, a_mn, b_mn, iters, blocks, flags):
    for _ in range(2):
        _lib.check(lib, lib.sv_mma_bench(N, a_mn | (flags << 1), b_mn, iters, blocks, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "bench")
        torch.cuda.synchronize()
    cyc = float(out[:blocks].double().mean())
    ns = float(out[blocks:2 * blocks].double().mean())
    return cyc, ns
for blocks in (1, 148):
    for N in (64, 128, 256):
        for a_mn, b_mn in ((0, 0), (0, 1), (1, 0), (1, 1)):
            iters = 2000
            cyc, ns = run(N, a_mn, b_mn, iters, blocks, 0)
            res[f"blocks{blocks}_N{N}_a{'MN' if a_mn else 'K'}_b{'MN' if b_mn else 'K'}"] = {"cycles_per_mma": round(cyc / (iters * 4), 1), "ideal_math_cycles": N / 2}
for flags, tag, mn in ((0, "zeros", 0), (1, "random", 0), (2, "zeros_commit_ring", 0), (3, "random_commit_ring", 0),
                       (3 | 4, "random_ring_4stages", 0), (3 | 4, "random_ring_4stages_MN", 1), (3 | 8, "random_ring_pollers", 0),
                       (1 | 16, "random_producer_handshake", 0), (1 | 4 | 8 | 16, "random_all", 0), (1 | 4 | 8 | 16, "random_all_MN", 1),
                       (1 | 4 | 32, "random_4stages_converged_elect", 0), (1 | 4 | 8 | 32, "random_4stages_pollers_converged_elect_MN", 1)):
    iters = 20000   # 80 K MMAs of 128x256x16 = ~5 ms: long enough for the power management to react
    cyc, ns = run(256, mn, mn, iters, 148, flags)
    res[f"sustained_148_N256_{tag}"] = {"cycles_per_mma": round(cyc / (iters * 4), 1), "sm_ghz": round(cyc / ns, 3),
                                        "tflops": round(148 * iters * 4 * 2 * 128 * 256 * 16 / ns / 1e3, 1)}
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/r2_mma_bench.json", "w"), indent=1)

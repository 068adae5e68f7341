"""gpurun_out/r2_step_launches_cupti.json (+ the ncu launch list) -> profiles/<tag>_step_launch_list.csv and a
native / library split.  Usage: python scripts/summarize_launch_list.py <tag> [cupti.json] [ncu.csv]"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
cupti = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "r2_step_launches_cupti.json")
ncu = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "r2_step_launches_ncu.csv")
NATIVE = re.compile(r"(sa_sample|sa_mlp|gemm_kernel|attn_fwd|attn_bwd|fused_attn|ln_fwd|ln_bwd|l2norm|colsum|pairwise_locs|ce_fwd_bwd|norm_allgather|"
                    r"group_points|gather_points|ball_query|fps_|three_|adamw_|sqnorm_partial|scale_inplace|act_bwd|embedding_bwd|scene_prep|token_mask|coin_mask|mma_bench)")


def family(name):
    if NATIVE.search(name) and "at::native" not in name:
        return "native"
    if name.startswith("nvjet") or "cublas" in name or "cutlass" in name or "gemv" in name or "splitKreduce" in name:
        return "library:cuBLAS"
    if "nccl" in name.lower():
        return "library:NCCL"
    if "Memcpy" in name or "Memset" in name:
        return "memcpy/memset"
    return "library:ATen"


d = json.load(open(cupti))
rows = d["kernels"]
tot = sum(r["ms_per_step"] for r in rows)
fam = {}
for r in rows:
    r["family"] = family(r["kernel"])
    f = fam.setdefault(r["family"], [0.0, 0.0])
    f[0] += r["ms_per_step"]
    f[1] += r["launches_per_step"]
out = os.path.join(ROOT, "profiles", f"{tag}_step_launch_list.csv")
with open(out, "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["# graph-replayed GPS pre-training step, 1 x B200, CUPTI kernel activities (warm, mean of 3 replays); "
                f"ms/step by CUDA events = {d['ms_per_step_events']:.3f}, sum of kernels = {tot:.3f}"])
    w.writerow(["family", "ms_per_step", "share", "launches_per_step", "kernel"])
    for r in rows:
        w.writerow([r["family"], f"{r['ms_per_step']:.4f}", f"{r['ms_per_step'] / tot:.4f}", f"{r['launches_per_step']:.1f}", r["kernel"][:160]])
summary = {"ms_per_step_events": d["ms_per_step_events"], "sum_kernel_ms": tot,
           "families": {k: {"ms": round(v[0], 3), "share": round(v[0] / tot, 4), "launches": round(v[1], 1)} for k, v in sorted(fam.items())}}
if os.path.exists(ncu):
    n = {}
    for row in csv.reader(open(ncu)):
        if len(row) > 14 and row[12] == "gpu__time_duration.sum":
            f = n.setdefault(family(row[4]), [0.0, 0])
            f[0] += float(row[14].replace(",", "")) / 1e6
            f[1] += 1
    nt = sum(v[0] for v in n.values())
    summary["ncu_cold_serialised"] = {"sum_ms": round(nt, 3), "families": {k: {"ms": round(v[0], 3), "share": round(v[0] / nt, 4), "launches": v[1]}
                                                                            for k, v in sorted(n.items())}}
json.dump(summary, open(os.path.join(ROOT, "profiles", f"{tag}_step_launch_summary.json"), "w"), indent=1)
print(json.dumps(summary, indent=1))

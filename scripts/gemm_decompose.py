"""Which part of the native GEMM bounds it: time each shape with the epilogue body and / or the TMA loads switched off
(sv_gemm_force_ctas bits 8..: 1 = no epilogue body, 2 = no loads).  Results are garbage with a switch on — timing only."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sceneverse_b200 import native
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return (torch.randn(*s, device="cuda", generator=g) * 0.5).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, n=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]
x2, w2, b2 = rnd(19200, 3072), rnd(768, 3072), torch.zeros(768, device="cuda")
x1, w1, b1 = rnd(19200, 768), rnd(3072, 768), torch.zeros(3072, device="cuda")
gj, wj = rnd(8320, 2304), rnd(2304, 768)
dw = torch.zeros(3072, 768, device="cuda")
cases = {
    "fwd_19200x768x3072": (lambda: native.linear_fwd(x2, w2, b2), 2 * 19200 * 768 * 3072),
    "fwd_19200x3072x768": (lambda: native.linear_fwd(x1, w1, b1), 2 * 19200 * 768 * 3072),
    "dgrad_8320x768x2304": (lambda: native.linear_dgrad(gj, wj), 2 * 8320 * 768 * 2304),
    "wgrad_3072x768x19200": (lambda: native.linear_wgrad(x2, x1, dw=dw, db=None, accumulate=True), 2 * 19200 * 768 * 3072),
}
out = {}
for name, (fn, flops) in cases.items():
    for c in (1, 2):
        row = {}
        for dbg, tag in ((0, "full"), (1, "no_epilogue"), (2, "no_loads"), (3, "mma_only")):
            native.gemm_force_ctas(c | (dbg << 8))
            ms = timed(fn)
            row[tag] = {"ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1)}
        out[f"{name}_ctas{c}"] = row
        print(name, c, {k: v["ms"] for k, v in row.items()}, flush=True)
native.gemm_force_ctas(0)
ref = {"fwd_19200x768x3072": timed(lambda: torch.nn.functional.linear(x2, w2)),
       "fwd_19200x3072x768": timed(lambda: torch.nn.functional.linear(x1, w1)),
       "dgrad_8320x768x2304": timed(lambda: gj @ wj),
       "wgrad_3072x768x19200": timed(lambda: x2.t() @ x1)}
out["cublas_ms"] = {k: round(v, 4) for k, v in ref.items()}
print("cublas", out["cublas_ms"])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r2_gemm_decompose.json", "w"), indent=1)

# where the MMA-issuing thread spends its cycles (sv_gemm_profile), per shape, full kernel and loads-only / MMA-only modes
from sceneverse_b200 import _lib
lib = _lib.gps()
prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
lib.sv_gemm_profile(prof.data_ptr())
issue = {}
for name, (fn, flops) in cases.items():
    for c in (1, 2):
        for dbg, tag in ((0, "full"), (1, "no_epilogue"), (3, "mma_only")):
            native.gemm_force_ctas(c | (dbg << 8))
            prof.zero_(); fn(); torch.cuda.synchronize()
            p = prof.view(148, 16).double()
            act = p[:, 3] > 0
            loop, wfull, wacc, ks = (float(p[act, i].mean()) for i in range(4))
            span = float(p[act, 5].max() - p[act, 4].min())
            issue[f"{name}_ctas{c}_{tag}"] = {"ctas_issuing": int(act.sum()), "loop_cycles": loop, "wait_operands": round(wfull / loop, 3),
                                              "wait_accumulator": round(wacc / loop, 3), "k_steps": ks, "cycles_per_k_step": round(loop / ks, 1),
                                              "loop_us": round(float(p[act, 6].mean()) / 1e3, 2),
                                              "epi_warp_cycles_per_tile": {k: round(float((p[:, 8 + i] / p[:, 11].clamp(min=1)).mean()), 0) for i, k in enumerate(("bias_barrier", "wait_accumulator", "body"))}, "sm_ghz": round(loop / float(p[act, 6].mean()), 3)}
            print(name, c, tag, issue[f"{name}_ctas{c}_{tag}"], flush=True)
lib.sv_gemm_profile(None)
native.gemm_force_ctas(0)
out["issue_thread"] = issue
json.dump(out, open("gpurun_out/r2_gemm_decompose.json", "w"), indent=1)

"""bf16 parity gate: the WHOLE forward chain (fused tcgen05 PointNet++ -> native spatial attention / LayerNorm / GEMMs -> V2
joint layers -> GroundHeadV1 (the ScanRefer head, hidden 384) / OVPretrainHead -> losses) under bf16 autocast with every
native kernel live, against the goldens of the UNMODIFIED reference run in fp32 (tests/golden/model_gps_stack.npz).

Error measure: max |got - ref| / max |ref| per output.  Two gates per output:
  (1) TOL: 1.5 x the error measured on the B200 (raw numbers: profiles/r2_parity_bf16.json, table in DESIGN.md §2);
  (2) the yardstick tests/golden/model_gps_stack_bf16_autocast_dev.json: how far the reference's OWN bf16 path (the
      unmodified modules under torch.autocast(bfloat16), fp32 residual stream, fp32 backbone) sits from its fp32 run — the
      B200 path must stay within 2.5x of that on every output.
north_star asks for grounding logits within 1e-3: that holds on the fp32 path (test_gps_modules.py, 1e-4 on the GPU, 1e-5 on
the CPU).  With bf16 GEMM operands (2^-9 relative rounding per operand) no implementation gets there — the reference's own
autocast run deviates by 1.7e-2 on og3d, this path by 1.6e-2."""
import json
import os

import numpy as np
import pytest
import torch

from sceneverse_b200 import _lib, weights
from sceneverse_b200.modules import grounding, heads, losses, vision

from .test_gps_modules import Z, inputs, load

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# output -> asserted bound (measured value on the B200 in the comment)
TOL = {                                  # measured on the B200 (round 2)
    "vis_obj_pre": 9.5e-3,               # 6.2e-3
    "vis_obj": 1.75e-2,                  # 1.16e-2
    "v2_txt": 1.9e-2,                    # 1.27e-2
    "v2_obj": 1.65e-2,                   # 1.09e-2
    "gh_txt_cls": 1.75e-2,               # 1.17e-2
    "gh_obj_cls": 1.5e-2,                # 9.8e-3
    "gh_obj_cls_pre": 1.3e-2,            # 8.5e-3
    "og3d": 2.4e-2,                      # 1.62e-2 (reference's own bf16 autocast: 1.71e-2)
    "lm": 2.45e-2,                       # 1.63e-2
    "obj_lm": 1.9e-2,                    # 1.27e-2
    "loss_within": 2e-3, "loss_obj_between": 5e-3, "loss_scene_between": 2e-3, "loss_og3d": 2.5e-3, "loss_lm": 1e-3,
    # measured 6.6e-4, 2.8e-3, 4.0e-4, 1.1e-3, 2.4e-4 (reference's own bf16 autocast: 5.4e-4, 1.7e-3, 1.3e-3, 1.4e-3, 1.5e-4)
}
YARD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "model_gps_stack_bf16_autocast_dev.json")))["dev"]


def rel_err(got, want):
    got = got.detach().float().cpu().numpy()
    fin = np.isfinite(want)
    assert (np.isfinite(got) == fin).all()
    return float(np.abs(got[fin] - want[fin]).max() / (np.abs(want[fin]).max() + 1e-12))


def test_native_bf16_chain_vs_reference_goldens():
    t, txt, scene_txt = inputs("cuda")
    tf = weights.synthetic_tensor("text_features", (607, 768))
    enc = load(vision.PointOpenVocabEncoder(None, freeze=True, text_features=tf), 0).cuda()
    v2 = load(grounding.UnifiedSpatialCrossEncoderV2(None), 1).cuda()
    gh = load(heads.GroundHeadV1(None, input_size=768, hidden_size=384, sem_cls_size=607), 4).cuda()
    ph = load(heads.OVPretrainHead(None), 5).cuda()
    err = {}
    n0 = _lib.launch_count()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        assert enc.point_feature_extractor.fused_available(t["obj_fts"].view(-1, 1024, 6))
        obj, obj_pre, sem = enc(t["obj_fts"], t["obj_locs"], t["obj_masks"], t["obj_sem_masks"], t["obj_labels"], 1, 1)
        err["vis_obj_pre"] = rel_err(obj_pre, Z["vis_obj_pre"])
        err["vis_obj"] = rel_err(obj, Z["vis_obj"])
        t2, o2 = v2(txt, t["txt_masks"], obj, t["obj_locs"], t["obj_masks"])
        err["v2_txt"], err["v2_obj"] = rel_err(t2, Z["v2_txt"]), rel_err(o2, Z["v2_obj"])
        a, b, c, og = gh(t2, o2, obj_pre, t["obj_masks"])
        err["gh_txt_cls"] = rel_err(a, Z["gh_txt_cls"])
        err["gh_obj_cls"] = rel_err(b[:, :, :64], Z["gh_obj_cls"])
        err["gh_obj_cls_pre"] = rel_err(c[:, :, :64], Z["gh_obj_cls_pre"])
        err["og3d"] = rel_err(og, Z["gh_og3d"])
        lm, ol = ph(t2, o2)
        err["lm"] = rel_err(lm[:, :, :128], Z["ph_txt_lm_slice"])
        err["obj_lm"] = rel_err(ol[:, :, :64], Z["ph_obj_lm"])
        dd = dict(t)
        dd.update(intra_obj_embeds=o2, intra_text_embed=t2[:, 0], inter_obj_embeds=obj, inter_text_embed=txt[:, 0],
                  scene_embed=obj.mean(dim=1), scene_text_embed=scene_txt, og3d_logits=og, txt_lm_cls_logits=lm)
        cfg = {"num_gpu": 1}
        for key, fn in [("loss_within", losses.TextObjWithinBatch(cfg)), ("loss_obj_between", losses.TextObjBetweenBatch(cfg)),
                        ("loss_scene_between", losses.TextSceneBetweenBatch(cfg)), ("loss_og3d", losses.og3d_loss),
                        ("loss_lm", losses.lm_cls_loss)]:
            if isinstance(fn, torch.nn.Module):
                fn = fn.cuda()
            want = float(Z[key])
            err[key] = abs(float(fn(dd)) - want) / max(1.0, abs(want))
    native = _lib.launch_count() - n0
    # sampling, 2 SA-MLP, 4 GEMM, pairwise, 4 spatial + 4 joint attention, 16+ LayerNorm, CE: nothing fell back to a library
    assert native >= 40, native
    assert (sem.argmax(-1).cpu().numpy() == Z["vis_sem_cls_argmax"]).mean() > 0.95
    print("PARITY_BF16 " + json.dumps({"errors": err, "native_launches": native}))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        json.dump({"errors": err, "tolerances": TOL, "native_launches": native,
                   "reference_bf16_autocast_dev": {k: v["max_rel"] for k, v in YARD.items()},
                   "measure": "max|got-ref| / max|ref| vs the fp32 reference goldens (losses: |got-ref| / max(1,|ref|))"},
                  open(os.path.join(out, "r2_parity_bf16.json"), "w"), indent=1)
    bad = {k: (v, TOL[k]) for k, v in err.items() if not v < TOL[k]}
    assert not bad, bad
    # yardstick: never further from the fp32 reference than 2.5x the reference's own bf16-autocast run (losses: + 2e-3 floor)
    far = {k: (v, YARD[k]["max_rel"]) for k, v in err.items()
           if YARD[k]["max_rel"] > 0 and v > 2.5 * YARD[k]["max_rel"] + (2e-3 if k.startswith("loss_") else 0.0)}
    assert not far, far

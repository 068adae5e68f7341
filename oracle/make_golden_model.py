"""ORACLE — TEST INFRASTRUCTURE ONLY.  Run IN THE BUILD CONTAINER (needs /root/reference):
    python oracle/make_golden_model.py
Runs the UNMODIFIED reference Python modules on CPU (fp32) through oracle/ref_shims.py, with the CPU
restatement of the CUDA ops as `_ext`, on seeded synthetic inputs and name-keyed synthetic weights
(sceneverse_b200/weights.py — reproducible anywhere without the reference), and writes small
fixtures to tests/golden/model_*.npz / *.json.
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402
from sceneverse_b200 import synthetic, weights  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def golden_pointnetpp():
    from modules.layers.pointnet import PointNetPP
    net = PointNetPP(sa_n_points=[32, 16, None], sa_n_samples=[32, 32, None], sa_radii=[0.2, 0.4, None],
                     sa_mlps=[[3, 64, 64, 128], [128, 128, 128, 256], [256, 256, 512, 768]]).eval()
    net.load_state_dict(weights.synthetic_state_dict(net, seed=0))
    x = torch.from_numpy(synthetic.object_batch(3, 8, 1024, pad_fraction=0.25))
    inter = {}
    hooks = [net.encoder[i].register_forward_hook(lambda m, a, o, i=i: inter.__setitem__(i, o)) for i in range(3)]
    with torch.no_grad():
        y = net(x)
    for h in hooks:
        h.remove()
    np.savez_compressed(os.path.join(OUT, "model_pointnetpp.npz"),
                        input_seed=3, n_clouds=8, pad_fraction=0.25, weight_seed=0,
                        new_xyz1=inter[0][0].numpy(), feat1=inter[0][1].numpy(),
                        new_xyz2=inter[1][0].numpy(), feat2=inter[1][1].numpy(),
                        feat3=inter[2][1].numpy(), out=y.numpy())
    return {k: list(v.shape) for k, v in net.state_dict().items()}


def main():
    ref_shims.install()
    torch.manual_seed(0)
    shapes = {"PointNetPP": golden_pointnetpp()}
    # state_dict contracts of the registry classes: key -> shape
    from modules.build import GROUNDING_REGISTRY, HEADS_REGISTRY, VISION_REGISTRY
    import modules.vision.pcd_openvocab_encoder  # noqa: F401
    import modules.grounding.unified_encoder  # noqa: F401
    import modules.heads.grounding_head  # noqa: F401
    import modules.heads.pretrain_head  # noqa: F401
    with tempfile.TemporaryDirectory() as d:
        ref_shims.write_text_features(d, weights.synthetic_tensor("text_features", (607, 768)))
        enc = VISION_REGISTRY.get("PointOpenVocabEncoder")(None, lang_path=d, freeze=True)
    shapes["PointOpenVocabEncoder"] = {k: list(v.shape) for k, v in enc.state_dict().items()}
    for name in ["UnifiedSpatialCrossEncoderV2", "UnifiedSpatialCrossEncoderV1", "EntitySpatialCrossEncoder"]:
        m = GROUNDING_REGISTRY.get(name)(None)
        shapes[name] = {k: list(v.shape) for k, v in m.state_dict().items()}
    for name, kw in [("GroundHeadV1", dict(input_size=768, hidden_size=384, sem_cls_size=607)), ("GroundHead", {}),
                     ("PretrainHeadV1", {}), ("OVPretrainHead", {})]:
        m = HEADS_REGISTRY.get(name)(None, **kw)
        shapes[name] = {k: list(v.shape) for k, v in m.state_dict().items()}
    json.dump(shapes, open(os.path.join(OUT, "state_dict_shapes.json"), "w"), indent=0, sort_keys=True)
    print({k: len(v) for k, v in shapes.items()})


if __name__ == "__main__":
    main()

"""ORACLE — TEST INFRASTRUCTURE ONLY.  Run ON THE GPU BOX:  python oracle/make_golden_gpu.py

Runs the REFERENCE's own CUDA kernels (oracle/_ref/_ext_ref.so = /root/reference's
_ext_src/src/*.cu compiled for sm_100a by oracle/build_ref_ext.py) on seeded inputs and writes
small .npz fixtures (inputs + reference outputs) to gpurun_out/golden/; they are then committed
under tests/golden/ and pin both the CPU oracle (tests/test_oracle.py) and the CUDA library
(tests/test_pointops_gpu.py).  It also cross-checks the CPU oracle on every case it writes.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref_ext, pointops_ref  # noqa: E402
from tests import cases  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def main():
    ref = build_ref_ext.load_prebuilt()
    assert ref is not None, "oracle/_ref/_ext_ref.so missing"
    out = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    bad = 0
    fps = cases.fps_cases()
    keep = ["sa1_objects", "sa2_shape", "ball_n3", "ball_n80", "ball_n100", "ball_n513", "ball_n1000",
            "adversarial_n32", "adversarial_n80", "adversarial_n1024", "m_gt_n"]
    for name in keep:
        xyz, m = fps[name]
        idx = ref.furthest_point_sampling(dev(xyz), m).cpu().numpy()
        np.savez_compressed(os.path.join(out, f"pointops_fps_{name}.npz"), op="fps", xyz=xyz, m=m, idx=idx)
        ok = np.array_equal(pointops_ref.furthest_point_sampling(xyz, m), idx)
        bad += not ok
        print("fps", name, "oracle==reference:", ok)
    big = cases.fps_cases_large()
    for name in ["ball_n2048", "adversarial_n2048"]:
        xyz, m = big[name]
        idx = ref.furthest_point_sampling(dev(xyz), m).cpu().numpy()
        np.savez_compressed(os.path.join(out, f"pointops_fps_{name}.npz"), op="fps", xyz=xyz, m=m, idx=idx)
        ok = np.array_equal(pointops_ref.furthest_point_sampling(xyz, m), idx)
        bad += not ok
        print("fps", name, "oracle==reference:", ok)
    for name, (new_xyz, xyz, r, ns) in cases.bq_cases().items():
        if name in ("many_centres",):
            continue
        idx = ref.ball_query(dev(new_xyz), dev(xyz), r, ns).cpu().numpy()
        np.savez_compressed(os.path.join(out, f"pointops_bq_{name}.npz"), op="ball_query", new_xyz=new_xyz, xyz=xyz,
                            radius=r, nsample=ns, idx=idx)
        ok = np.array_equal(pointops_ref.ball_query(new_xyz, xyz, r, ns), idx)
        bad += not ok
        print("ball_query", name, "oracle==reference:", ok)
    rng = np.random.default_rng(1)
    u = rng.standard_normal((2, 200, 3)).astype(np.float32)
    k = rng.standard_normal((2, 48, 3)).astype(np.float32)
    d, i = ref.three_nn(dev(u), dev(k))
    d, i = d.cpu().numpy(), i.cpu().numpy()
    np.savez_compressed(os.path.join(out, "pointops_three_nn.npz"), op="three_nn", unknown=u, known=k, dist2=d, idx=i)
    od, oi = pointops_ref.three_nn(u, k)
    ok = np.array_equal(od, d) and np.array_equal(oi, i)
    bad += not ok
    print("three_nn oracle==reference:", ok)
    feats = rng.standard_normal((2, 6, 48)).astype(np.float32)
    w = rng.random((2, 200, 3)).astype(np.float32)
    o = ref.three_interpolate(dev(feats), dev(i), dev(w)).cpu().numpy()
    np.savez_compressed(os.path.join(out, "pointops_three_interpolate.npz"), op="three_interpolate", points=feats,
                        idx=i, weight=w, out=o)
    ok = np.array_equal(pointops_ref.three_interpolate(feats, i, w), o)
    bad += not ok
    print("three_interpolate oracle==reference:", ok)
    pts = rng.standard_normal((2, 6, 256)).astype(np.float32)
    gi = rng.integers(0, 256, size=(2, 8, 16)).astype(np.int32)
    o = ref.group_points(dev(pts), dev(gi)).cpu().numpy()
    np.savez_compressed(os.path.join(out, "pointops_group_points.npz"), op="group_points", points=pts, idx=gi, out=o)
    ok = np.array_equal(pointops_ref.group_points(pts, gi), o)
    bad += not ok
    print("group_points oracle==reference:", ok)
    print("GOLDEN_MISMATCHES", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)

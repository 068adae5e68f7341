"""PointNetPP module: state_dict contract (CPU) and parity of the fused tcgen05 path (GPU) against
(a) the reference's own modules run on CPU in fp32 (tests/golden/model_pointnetpp.npz, made by
oracle/make_golden_model.py) and (b) the generic path (reference operator sequence on the native
point ops + cuDNN fp32) on the same GPU."""
import json
import os

import numpy as np
import pytest
import torch

from sceneverse_b200 import synthetic, weights

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def make_net():
    from sceneverse_b200.modules.pointnet import GPS_SPEC, PointNetPP
    net = PointNetPP(**GPS_SPEC).eval()
    net.load_state_dict(weights.synthetic_state_dict(net, seed=0))
    return net


def test_state_dict_matches_reference_contract():
    want = json.load(open(os.path.join(GOLDEN, "state_dict_shapes.json")))["PointNetPP"]
    got = {k: list(v.shape) for k, v in make_net().state_dict().items()}
    assert got == want


def rel_err(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / (np.abs(want).max() + 1e-12)), \
        float(np.abs(got - want).mean() / (np.abs(want).mean() + 1e-12))


@pytest.mark.gpu
def test_generic_path_matches_reference_fp32():
    """Reference operator sequence on the native kernels, fp32 (TF32 off): 1e-5-level parity with the reference on CPU."""
    z = np.load(os.path.join(GOLDEN, "model_pointnetpp.npz"))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = make_net().cuda()
    x = torch.from_numpy(synthetic.object_batch(int(z["input_seed"]), int(z["n_clouds"]), 1024,
                                                float(z["pad_fraction"]))).cuda()
    with torch.no_grad():
        y = net.forward_generic(x)
    mx, mean = rel_err(y.cpu().numpy(), z["out"])
    assert mx < 1e-4 and mean < 1e-5, (mx, mean)  # fp32, different summation order only


@pytest.mark.gpu
def test_fused_path_indices_and_features():
    z = np.load(os.path.join(GOLDEN, "model_pointnetpp.npz"))
    net = make_net().cuda()
    x = torch.from_numpy(synthetic.object_batch(int(z["input_seed"]), int(z["n_clouds"]), 1024,
                                                float(z["pad_fraction"]))).cuda()
    with torch.no_grad():
        assert net.fused_available(x)
    for p in net.parameters():  # the frozen-backbone configuration of all_pretrain.yaml
        p.requires_grad = False
    assert net.fused_available(x)
    y, inter = net.forward_fused(x, return_intermediates=True)
    # sampling is index work: the sampled centres must be bit-identical to the reference's
    np.testing.assert_array_equal(inter["new_xyz"].cpu().numpy(), z["new_xyz1"])
    np.testing.assert_array_equal(inter["new_xyz2"].cpu().numpy(), z["new_xyz2"])
    # features: bf16 tensor-core math vs the fp32 reference (tolerances: bf16 operand rounding through 3/6/9 layers)
    f1 = inter["feat1"].float().transpose(1, 2).cpu().numpy()  # (B,128,32) like the reference
    f2 = inter["feat2"].float().transpose(1, 2).cpu().numpy()
    e1, e2, e3 = rel_err(f1, z["feat1"]), rel_err(f2, z["feat2"]), rel_err(y.cpu().numpy(), z["out"])
    print("rel err (max/max, mean/mean): SA1", e1, "SA2", e2, "out", e3)
    assert e1[0] < 2e-2 and e1[1] < 5e-3, e1
    assert e2[0] < 3e-2 and e2[1] < 8e-3, e2
    assert e3[0] < 4e-2 and e3[1] < 1e-2, e3


@pytest.mark.gpu
def test_sa_kernels_against_bf16_emulation():
    """Tight check of the two tcgen05 kernels: same bf16-rounded operands, fp32 accumulation in torch."""
    import torch.nn.functional as F
    from sceneverse_b200.modules.pointnet import fold_bn
    net = make_net().cuda()
    x = torch.from_numpy(synthetic.object_batch(11, 37, 1024, 0.2)).cuda()  # 37 clouds: partial last super-tile
    y, it = net.forward_fused(x, return_intermediates=True)
    bf = lambda t: t.to(torch.bfloat16).float()

    def emulate(rows, mlp):  # rows (..., K) f32 already bf16-rounded
        h = rows
        for j in range(3):
            layer = getattr(mlp, f"layer{j}")
            w, s = fold_bn(layer.conv.weight, layer.bn.bn)
            h = torch.relu(h @ bf(w).t() + s)
            if j < 2:
                h = bf(h)
        return bf(h.max(dim=-2).values)

    B = x.shape[0]
    idx = it["ball_idx"].long()                                               # (B,32,32)
    g = torch.gather(x[:, None].expand(B, 32, 1024, 6), 2, idx[..., None].expand(B, 32, 32, 6))
    g = torch.cat([g[..., :3] - it["new_xyz"][:, :, None], g[..., 3:]], -1)
    want1 = emulate(bf(g), net.encoder[0].mlps[0])                            # (B,32,128)
    got1 = it["feat1"].float()
    assert (got1 - want1).abs().max().item() <= 2e-2 * want1.abs().max().item()
    idx2 = it["ball_idx2"].long()                                             # (B,16,32) into the 32 level-1 points
    gx = torch.gather(it["new_xyz"][:, None].expand(B, 16, 32, 3), 2, idx2[..., None].expand(B, 16, 32, 3))
    gf = torch.gather(got1[:, None].expand(B, 16, 32, 128), 2, idx2[..., None].expand(B, 16, 32, 128))
    rows2 = torch.cat([bf(gx - it["new_xyz2"][:, :, None]), gf], -1)          # reference order [xyz | feat]
    want2 = emulate(rows2, net.encoder[1].mlps[0])
    got2 = it["feat2"].float()
    assert (got2 - want2).abs().max().item() <= 2e-2 * want2.abs().max().item()


@pytest.mark.gpu
def test_objcls_step_trains_pointnet_through_native_grads():
    """BASELINE.json configs[1]: 64 objects x 1024 points, PointNet++ trainable (train-mode BN), bf16 autocast,
    607-way open-vocabulary CE with label smoothing: one optimisation step must produce finite gradients for every
    PointNet++ parameter through group_points_grad, and match the same step computed with torch gathers."""
    from sceneverse_b200 import model as M
    from sceneverse_b200.modules import losses
    d = synthetic.scene_batch(5, B=1, O=64, P=1024, all_valid=True)
    batch = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    tf = weights.synthetic_tensor("text_features", (607, 768)).cuda()
    net = M.ObjCls({"num_gpu": 1, "solver": {"lr": 1e-3}}, text_embeds=tf).cuda().train()
    net.point_feature_extractor.load_state_dict(weights.synthetic_state_dict(net.point_feature_extractor, 0))
    net.dropout.p = 0.0
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = net(dict(batch))
        loss = losses.obj_cls_loss(out)
    loss.backward()
    grads = {n: p.grad for n, p in net.named_parameters()}
    assert all(g is not None and torch.isfinite(g).all() for g in grads.values())
    assert float(grads["point_feature_extractor.encoder.0.mlps.0.layer0.conv.weight"].abs().sum()) > 0
    # same computation with torch.gather instead of the native group/gather kernels (identical indices)
    pn = net.point_feature_extractor
    x = batch["obj_fts"].view(64, 1024, 6).float()
    xyz, feats = x[..., :3].contiguous(), x[..., 3:].transpose(1, 2).contiguous()
    for p in net.parameters():
        p.grad = None
    from sceneverse_b200.pointnet2 import _ext
    with torch.autocast("cuda", dtype=torch.bfloat16):
        for sa in pn.encoder:
            if sa.npoint is not None:
                fi = _ext.furthest_point_sampling(xyz, sa.npoint).long()
                new_xyz = torch.gather(xyz, 1, fi[..., None].expand(-1, -1, 3))
                bi = _ext.ball_query(new_xyz.contiguous(), xyz, sa.radius, sa.nsample).long()
                Bn, npnt, ns = bi.shape
                gx = torch.gather(xyz[:, None].expand(-1, npnt, -1, -1), 2, bi[..., None].expand(-1, -1, -1, 3))
                gx = (gx - new_xyz[:, :, None]).permute(0, 3, 1, 2)
                C = feats.shape[1]
                gf = torch.gather(feats[:, :, None].expand(-1, -1, npnt, -1), 3, bi[:, None].expand(-1, C, -1, -1))
                g = torch.cat([gx, gf], 1)
            else:
                new_xyz = None
                g = torch.cat([xyz.transpose(1, 2).unsqueeze(2), feats.unsqueeze(2)], 1)
            h = sa.mlps[0](g)
            feats = torch.nn.functional.max_pool2d(h, kernel_size=[1, h.size(3)]).squeeze(-1)
            xyz = new_xyz
        emb = pn.fc(feats.view(64, -1))
        out2 = {"obj_logits": (emb @ tf.t().to(emb.dtype)).view(1, 64, -1), "obj_labels": batch["obj_labels"], "obj_masks": batch["obj_masks"]}
        loss2 = losses.obj_cls_loss(out2)
    loss2.backward()
    assert abs(float(loss) - float(loss2)) < 2e-2 * abs(float(loss2))
    g1 = grads["point_feature_extractor.encoder.1.mlps.0.layer0.conv.weight"].float()
    g2 = pn.encoder[1].mlps[0].layer0.conv.weight.grad.float()
    assert (g1 - g2).abs().max().item() <= 0.1 * g2.abs().max().item() + 1e-6


@pytest.mark.gpu
def test_trainable_backbone_native_path_matches_the_generic_operator_sequence():
    """Config C2 (ObjCls, trainable PointNet++, train-mode BatchNorm): the channels-last native path (tcgen05 GEMMs for the 1x1
    convolutions + csrc/pn_train.cu batch-statistic BatchNorm / max kernels, bf16) against the reference operator sequence in
    fp32 (torch Conv2d / BatchNorm2d / max_pool2d on the same native point ops): output, every parameter gradient and the
    BatchNorm running statistics."""
    import copy
    from sceneverse_b200 import pn_train
    from sceneverse_b200.modules.pointnet import GPS_SPEC, PointNetPP
    torch.manual_seed(0)
    net = PointNetPP(**GPS_SPEC)
    net.load_state_dict(weights.synthetic_state_dict(net, 0))
    net = net.cuda().train()
    ref = copy.deepcopy(net)
    x = torch.from_numpy(synthetic.object_batch(3, 48, 1024, pad_fraction=0.0)).cuda()
    go = torch.randn(48, 768, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert pn_train.available(x, net)
        y = net(x)
    y.float().backward(go)
    yr = ref.forward_generic(x)                      # fp32, no autocast: the reference operator sequence
    yr.backward(go)
    err = (y.float() - yr).abs().max().item() / yr.abs().max().item()
    assert err < 4e-2, err
    worst = {}
    for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, n
        e = (p.grad.float() - q.grad).abs().max().item() / (q.grad.abs().max().item() + 1e-8)
        cos = torch.nn.functional.cosine_similarity(p.grad.float().flatten(), q.grad.flatten(), dim=0).item()
        worst[n] = (round(e, 4), round(cos, 5))
    print("TRAINABLE_PN grads", worst)
    # behind the last neighbourhood max the two paths route the gradient identically: tight agreement there; in front of a max
    # taken over bf16 values ties are frequent and "first maximum" picks a different (equally valid) point than the fp32 run,
    # so the earlier layers are only required to point the same way
    for n in ("fc.weight", "fc.bias", "encoder.2.mlps.0.layer2.bn.bn.weight", "encoder.2.mlps.0.layer2.bn.bn.bias"):
        assert worst[n][1] > 0.999 and worst[n][0] < 0.1, (n, worst[n])
    for n, (e, cos) in worst.items():
        assert cos > 0.8, (n, e, cos)
    for (n, b1), (_, b2) in zip(net.named_buffers(), ref.named_buffers()):
        if "running" in n:
            assert (b1 - b2).abs().max().item() <= 3e-2 * (b2.abs().max().item() + 1e-3), n
        elif "num_batches_tracked" in n:
            assert int(b1) == int(b2) == 1
    print("TRAINABLE_PN output err", err)


@pytest.mark.gpu
def test_trainable_backbone_building_blocks_match_torch_on_the_same_operands():
    """Unit parity of the three autograd Functions of the train-mode path against torch formulations fed with the SAME bf16
    operands (so that rounding / tie-breaking of the inputs is shared): 1x1 conv + batch-statistic BatchNorm + ReLU, the
    neighbourhood max, and the channels-last grouping with its scatter-add gradient."""
    import torch.nn.functional as F
    from sceneverse_b200 import pn_train
    g = torch.Generator(device="cuda").manual_seed(0)

    def rel(a, b):
        return (a.float() - b.float()).abs().max().item() / (b.float().abs().max().item() + 1e-9)
    # --- conv1x1 + BN(batch statistics) + ReLU -----------------------------------------------------------------------------
    R, Cin, Cout = 8192, 67, 128
    Xb = torch.zeros(R, 72, device="cuda", dtype=torch.bfloat16)
    Xb[:, :Cin] = torch.randn(R, Cin, device="cuda", generator=g).bfloat16()
    X = Xb.clone().requires_grad_(True)
    W = (torch.randn(Cout, Cin, 1, 1, device="cuda", generator=g) * 0.2).requires_grad_(True)
    gam = (1 + 0.2 * torch.randn(Cout, device="cuda", generator=g)).requires_grad_(True)
    bet = (0.2 * torch.randn(Cout, device="cuda", generator=g)).requires_grad_(True)
    go = torch.randn(R, Cout, device="cuda", generator=g).bfloat16()
    out, stats = pn_train._ConvBNReLUFn.apply(X, W, gam, bet, 1e-5)
    out.backward(go)
    Xr = Xb[:, :Cin].float().requires_grad_(True)
    Wr = W.detach().bfloat16().float().reshape(Cout, Cin).requires_grad_(True)
    gr, br = gam.detach().clone().requires_grad_(True), bet.detach().clone().requires_grad_(True)
    y0 = Xr @ Wr.t()
    yr = y0 + (y0.detach().bfloat16().float() - y0.detach())      # the kernel stores Y in bf16 (straight-through rounding)
    ref = torch.relu(F.batch_norm(yr, None, None, gr, br, True, 0.0, 1e-5))
    ref.backward(go.float())
    assert rel(out, ref) < 1e-2
    assert rel(stats[0], yr.mean(0)) < 1e-3 and rel(stats[2], yr.var(0, unbiased=True)) < 2e-3
    assert rel(X.grad[:, :Cin], Xr.grad) < 2.5e-2 and rel(W.grad.reshape(Cout, Cin), Wr.grad) < 2.5e-2
    assert rel(gam.grad, gr.grad) < 1e-2 and rel(bet.grad, br.grad) < 1e-2
    # --- neighbourhood max (tie-free values) ---------------------------------------------------------------------------------
    G, ns, C = 300, 32, 64
    base = torch.rand(G * ns, C, device="cuda", generator=g).argsort(0).float() / (G * ns)      # all distinct per column
    xv = base.bfloat16().float()
    keep = torch.ones_like(xv, dtype=torch.bool)
    x = xv.bfloat16().requires_grad_(True)
    o = pn_train._RowGroupMaxFn.apply(x, ns)
    gm = torch.randn(G, C, device="cuda", generator=g).bfloat16()
    o.backward(gm)
    xr = x.detach().float().requires_grad_(True)
    orr = xr.view(G, ns, C).amax(1)
    assert torch.equal(o.float(), orr)
    # gradient: the whole group gradient lands on maximal entries (ties share rows in torch; here exactly one row gets it)
    assert torch.allclose(x.grad.float().view(G, ns, C).sum(1), gm.float(), atol=1e-6)
    assert ((x.grad.float() != 0) <= (x.detach().float().view(G, ns, C) == orr[:, None]).view(G * ns, C)).all()
    # --- channels-last grouping + scatter-add gradient ---------------------------------------------------------------------------
    B, N, Cf, np_, ns = 5, 200, 13, 7, 9
    xyz = torch.randn(B, N, 3, device="cuda", generator=g)
    cen = torch.randn(B, np_, 3, device="cuda", generator=g)
    feat = torch.randn(B, N, Cf, device="cuda", generator=g).requires_grad_(True)
    idx = torch.randint(0, N, (B, np_, ns), device="cuda", generator=g, dtype=torch.int32)
    Xg = pn_train._GroupRowsFn.apply(xyz, cen, feat, idx, np_, ns)
    assert Xg.shape == (B * np_ * ns, 16)
    il = idx.long()
    gx = torch.gather(xyz[:, None].expand(-1, np_, -1, -1), 2, il[..., None].expand(-1, -1, -1, 3)) - cen[:, :, None]
    gf = torch.gather(feat[:, None].expand(-1, np_, -1, -1), 2, il[..., None].expand(-1, -1, -1, Cf))
    want = torch.cat([gx, gf], -1).reshape(-1, 3 + Cf)
    assert rel(Xg[:, :16].float()[:, :3 + Cf], want.detach()) < 8e-3 and (Xg[:, 3 + Cf:] == 0).all()
    gX = torch.randn(Xg.shape, device="cuda", generator=g).bfloat16()
    Xg.backward(gX, retain_graph=False)
    got = feat.grad.clone()
    feat.grad = None
    gf.backward(gX[:, 3:3 + Cf].float().reshape(B, np_, ns, Cf))
    assert rel(got, feat.grad) < 1e-5

"""Fused tcgen05 attention forward vs the reference formulas (ops.attention / ops.spatial_attention in fp32 torch)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g) * scale


@pytest.mark.parametrize("Lq,Lk", [(80, 80), (130, 130), (80, 50), (50, 80), (50, 50), (32, 32), (1, 7), (129, 160), (300, 256), (17, 200)])
def test_plain_attention(Lq, Lk):
    from sceneverse_b200 import native, ops
    B, H, E = 3, 12, 768
    q, k, v = rand(B, Lq, E, seed=1), rand(B, Lk, E, seed=2), rand(B, Lk, E, seed=3)
    mask = torch.zeros(B, Lk, dtype=torch.bool, device="cuda")
    mask[1, Lk // 2:] = True
    mask[2, -1] = True
    qb, kb, vb = q.bfloat16(), k.bfloat16(), v.bfloat16()
    got = native.attention(qb, kb, vb, H, key_padding_mask=mask).float()
    want = ops.attention(qb.float(), kb.float(), vb.float(), H, key_padding_mask=mask)
    err = (got - want).abs().max().item() / want.abs().max().item()
    assert err < 2e-2, err


def test_packed_qkv_views():
    from sceneverse_b200 import native, ops
    B, L, H, E = 2, 130, 12, 768
    qkv = rand(B, L, 3 * E, seed=5).bfloat16()
    q, k, v = qkv.split(E, dim=-1)  # strided views, as produced by the packed in-projection
    got = native.attention(q, k, v, H).float()
    want = ops.attention(q.float(), k.float(), v.float(), H)
    assert (got - want).abs().max().item() / want.abs().max().item() < 2e-2


@pytest.mark.parametrize("L,spatial_heads", [(80, 12), (32, 12), (80, 1)])
def test_spatial_attention(L, spatial_heads):
    from sceneverse_b200 import native, ops
    B, H, E = 4, 12, 768
    q, k, v = rand(B, L, E, seed=1).bfloat16(), rand(B, L, E, seed=2).bfloat16(), rand(B, L, E, seed=3).bfloat16()
    sw = rand(B, L, spatial_heads * 6, seed=4, scale=2.0)
    centers = rand(B, L, 3, seed=6, scale=2.0)
    locs = ops.calc_pairwise_locs(centers, None)
    mask = torch.zeros(B, L, dtype=torch.bool, device="cuda")
    mask[0, L - 7:] = True
    mask[3, 5:] = True
    got = native.attention(q, k, v, H, key_padding_mask=mask, spatial_w=sw, spatial_heads=spatial_heads,
                           pairwise_locs=locs).float()
    want, _ = ops._spatial_attention_torch(q.float(), k.float(), v.float(), sw, locs, H, spatial_heads, key_padding_mask=mask)
    err = (got - want).abs().max().item() / want.abs().max().item()
    assert err < 2e-2, err


def test_spatial_attention_backward_matches_torch():
    from sceneverse_b200 import ops
    B, L, H, E = 2, 80, 12, 768
    q, k, v = (rand(B, L, E, seed=s).bfloat16().requires_grad_(True) for s in (1, 2, 3))
    sw = rand(B, L, 72, seed=4).requires_grad_(True)
    locs = ops.calc_pairwise_locs(rand(B, L, 3, seed=6, scale=2.0), None)
    out, attn = ops.spatial_attention(q, k, v, sw, locs, H, H)
    assert attn is None and out.dtype == torch.bfloat16
    assert isinstance(out.grad_fn, ops._AttentionFn._backward_cls)  # native backward kernels
    out.float().square().sum().backward()
    g_native = [t.grad.clone() for t in (q, k, v, sw)]
    for t in (q, k, v, sw):
        t.grad = None
    out2, _ = ops._spatial_attention_torch(q.float(), k.float(), v.float(), sw, locs, H, H)
    out2.float().square().sum().backward()
    for a, b in zip(g_native, (q.grad, k.grad, v.grad, sw.grad)):
        assert (a.float() - b.float()).abs().max().item() <= 5e-2 * b.float().abs().max().item() + 1e-6


def test_pairwise_locs_kernel_matches_torch():
    from sceneverse_b200 import ops
    locs6 = rand(5, 80, 6, seed=9, scale=3.0)
    locs6[2, 40:] = 0.0  # padded objects sit at the origin and take part in the max distance
    got = ops.calc_pairwise_locs(locs6[:, :, :3], locs6[:, :, 3:])            # strided view -> native kernel
    c = locs6[:, :, :3].cpu()
    want = ops.calc_pairwise_locs(c, None)                                     # torch formulation on CPU
    assert got.shape == (5, 80, 80, 5)
    assert torch.allclose(got.cpu(), want, rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_fused_cross_entropy_matches_torch(dtype):
    from sceneverse_b200 import ops
    R, V = 300, 30522
    logits = (rand(R, V, seed=11) * 3).to(dtype).requires_grad_(True)
    labels = torch.randint(0, V, (R,), device="cuda")
    labels[torch.rand(R, device="cuda") < 0.8] = -1
    labels[0] = 5
    loss = ops.cross_entropy(logits.view(6, 50, V), labels.view(6, 50), ignore_index=-1)
    loss.backward()
    ref_logits = logits.detach().float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ref_logits, labels, ignore_index=-1)
    ref.backward()
    assert abs(float(loss) - float(ref)) < (2e-3 if dtype == torch.bfloat16 else 1e-5) * max(1.0, abs(float(ref)))
    g, gr = logits.grad.float(), ref_logits.grad
    assert (g - gr).abs().max().item() <= (1e-2 if dtype == torch.bfloat16 else 1e-6) * gr.abs().max().item() + 1e-9
    # -inf logits (masked objects in og3d) give zero probability, not NaN
    og = rand(4, 80, seed=12)
    og[:, 60:] = float("-inf")
    og.requires_grad_(True)
    tgt = torch.tensor([[3], [10], [59], [0]], device="cuda")
    l2 = ops.cross_entropy(og, tgt.squeeze(1))
    l2.backward()
    want = torch.nn.functional.cross_entropy(og.detach(), tgt.squeeze(1))
    assert abs(float(l2) - float(want)) < 1e-5 and torch.isfinite(og.grad).all()


@pytest.mark.parametrize("Lq,Lk", [(80, 80), (130, 130), (50, 80), (80, 50), (33, 7), (256, 200), (129, 65)])
def test_plain_attention_backward_native(Lq, Lk):
    from sceneverse_b200 import ops
    B, H, E = 3, 12, 768
    q = rand(B, Lq, E, seed=1).bfloat16().requires_grad_(True)
    k = rand(B, Lk, E, seed=2).bfloat16().requires_grad_(True)
    v = rand(B, Lk, E, seed=3).bfloat16().requires_grad_(True)
    mask = torch.zeros(B, Lk, dtype=torch.bool, device="cuda")
    mask[1, Lk // 2:] = True
    go = rand(B, Lq, E, seed=4)
    out = ops.attention(q, k, v, H, key_padding_mask=mask)              # native forward + native backward
    assert isinstance(out.grad_fn, ops._AttentionFn._backward_cls)
    out.backward(go.bfloat16())
    got = [t.grad.float().clone() for t in (q, k, v)]
    qf, kf, vf = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    B_, hd = B, 64
    qh = qf.view(B, Lq, H, hd).transpose(1, 2); kh = kf.view(B, Lk, H, hd).transpose(1, 2); vh = vf.view(B, Lk, H, hd).transpose(1, 2)
    att = (qh @ kh.transpose(-1, -2)) * 0.125
    att = att.masked_fill(mask[:, None, None, :], float("-inf")).softmax(-1)
    want_out = (att @ vh).transpose(1, 2).reshape(B, Lq, E)
    want_out.backward(go.bfloat16().float())
    for g, w, name in zip(got, (qf.grad, kf.grad, vf.grad), "qkv"):
        err = (g - w).abs().max().item() / (w.abs().max().item() + 1e-9)
        assert err < 3e-2, (name, err)


def _dropout_keep(seed, B, H, Lq, Lk, p):
    """numpy restatement of csrc/attn_common.cuh drop_row_key / drop_pair_hash / drop_keep."""
    import numpy as np
    M64, M32 = np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        rows = np.arange(B * H * Lq, dtype=np.uint64)
        x = (np.uint64(seed) + rows * np.uint64(0x9E3779B97F4A7C15)) & M64
        x ^= x >> np.uint64(33)
        x = (x * np.uint64(0xff51afd7ed558ccd)) & M64
        x ^= x >> np.uint64(33)
        x = (x * np.uint64(0xc4ceb9fe1a85ec53)) & M64
        x ^= x >> np.uint64(33)
        rk = (x & M32)[:, None]                                   # (rows, 1)
        j = np.arange(Lk, dtype=np.uint64)[None, :]
        h = rk ^ (((j >> np.uint64(1)) * np.uint64(0x9E3779B1)) & M32)
        h ^= h >> np.uint64(16)
        h = (h * np.uint64(0x85EBCA6B)) & M32
        h ^= h >> np.uint64(13)
        h = (h * np.uint64(0xC2B2AE35)) & M32
        h ^= h >> np.uint64(16)
        u16 = (h >> ((j & np.uint64(1)) * np.uint64(16))) & np.uint64(0xFFFF)
    return (u16 >= np.uint64(int(p * 65536.0 + 0.5))).reshape(B, H, Lq, Lk)


@pytest.mark.parametrize("Lq,Lk", [(130, 130), (80, 50)])
def test_attention_dropout_forward_backward(Lq, Lk):
    """Attention-weight dropout inside the kernels: the mask is a pure function of (seed, b, h, i, j), so the torch
    reference can be fed exactly the same mask."""
    import numpy as np
    from sceneverse_b200 import native
    B, H, E, p, seed = 2, 12, 768, 0.1, 123456789
    q = rand(B, Lq, E, seed=1).bfloat16(); k = rand(B, Lk, E, seed=2).bfloat16(); v = rand(B, Lk, E, seed=3).bfloat16()
    go = rand(B, Lq, E, seed=4).bfloat16()
    out, lse = native.attention(q, k, v, H, return_lse=True, dropout_p=p, seed=seed)
    dq, dk, dv, _ = native.attention_backward(q, k, v, out, go, lse, H, dropout_p=p, seed=seed)
    keep = torch.from_numpy(_dropout_keep(seed, B, H, Lq, Lk, p)).cuda()
    assert abs(keep.float().mean().item() - (1 - p)) < 0.01
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    qh = qf.view(B, Lq, H, 64).transpose(1, 2); kh = kf.view(B, Lk, H, 64).transpose(1, 2); vh = vf.view(B, Lk, H, 64).transpose(1, 2)
    att = ((qh @ kh.transpose(-1, -2)) * 0.125).softmax(-1) * keep / (1 - p)
    want = (att @ vh).transpose(1, 2).reshape(B, Lq, E)
    want.backward(go.float())
    assert (out.float() - want).abs().max().item() / want.abs().max().item() < 2e-2
    for g, w, name in zip((dq, dk, dv), (qf.grad, kf.grad, vf.grad), "qkv"):
        err = (g.float() - w).abs().max().item() / (w.abs().max().item() + 1e-9)
        assert err < 3e-2, (name, err)

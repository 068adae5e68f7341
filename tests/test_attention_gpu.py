"""Fused tcgen05 attention forward vs the reference formulas (ops.attention / ops.spatial_attention in fp32 torch)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g) * scale


@pytest.mark.parametrize("Lq,Lk", [(80, 80), (130, 130), (80, 50), (50, 80), (50, 50), (32, 32), (1, 7), (129, 160), (300, 256), (17, 200), (300, 300), (40, 384)])
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


@pytest.mark.parametrize("Lq,Lk", [(80, 80), (130, 130), (50, 80), (80, 50), (33, 7), (256, 200), (129, 65), (300, 300), (384, 50)])
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


@pytest.mark.parametrize("Lq,Lk", [(130, 130), (80, 50), (300, 300)])
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


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("R,D", [(4000, 768), (37, 384), (5, 1024), (260, 8), (1, 256)])
def test_layer_norm_native_matches_torch(dtype, R, D):
    """Fused residual + LayerNorm (no dropout): forward and all four gradients against F.layer_norm in fp32."""
    from sceneverse_b200 import ops
    import torch.nn.functional as F
    x = rand(R, D, seed=1).to(dtype).requires_grad_(True)
    r = rand(R, D, seed=2).to(dtype).requires_grad_(True)
    w = (1.0 + 0.1 * rand(D, seed=3)).requires_grad_(True)
    b = (0.1 * rand(D, seed=4)).requires_grad_(True)
    go = rand(R, D, seed=5).to(dtype)
    for use_res in (True, False):
        for t in (x, r, w, b):
            t.grad = None
        y = ops.layer_norm(x, w, b, 1e-5, residual=r if use_res else None)
        assert y.dtype == dtype and isinstance(y.grad_fn, ops._LayerNormFn._backward_cls)
        y.backward(go)
        got = [y.detach().float()] + [t.grad.float().clone() for t in ((x, r, w, b) if use_res else (x, w, b))]
        xf, rf, wf, bf = (t.detach().float().requires_grad_(True) for t in (x, r, w, b))
        s = (xf + rf) if use_res else xf
        if dtype == torch.bfloat16:
            s = s + (s.detach().bfloat16().float() - s.detach())   # the kernel normalises the bf16-rounded sum
        yr = F.layer_norm(s, (D,), wf, bf, 1e-5)
        yr.backward(go.float())
        want = [yr.detach()] + [t.grad for t in ((xf, rf, wf, bf) if use_res else (xf, wf, bf))]
        tol = 2e-2 if dtype == torch.bfloat16 else 2e-5
        for i, (g, wv) in enumerate(zip(got, want)):
            err = (g - wv).abs().max().item() / (wv.abs().max().item() + 1e-9)
            assert err < tol, (use_res, i, err)


def test_layer_norm_dropout_mask_and_backward():
    """Dropout inside the fused LayerNorm: the mask is the documented counter hash, so torch can be fed the same mask."""
    from sceneverse_b200 import ops
    import torch.nn.functional as F
    R, D, p, seed = 300, 768, 0.1, 987654321
    x = rand(R, D, seed=1).bfloat16().requires_grad_(True)
    r = rand(R, D, seed=2).bfloat16().requires_grad_(True)
    w = (1.0 + 0.1 * rand(D, seed=3)).requires_grad_(True)
    b = (0.1 * rand(D, seed=4)).requires_grad_(True)
    go = rand(R, D, seed=5).bfloat16()
    y = ops._LayerNormFn.apply(x, r, w, b, 1e-5, p, seed)
    y.backward(go)
    keep = torch.from_numpy(_dropout_keep(seed, 1, 1, R, D, p)).cuda().view(R, D)
    assert abs(keep.float().mean().item() - (1 - p)) < 0.01
    xf, rf, wf, bf = (t.detach().float().requires_grad_(True) for t in (x, r, w, b))
    s = rf + xf * keep / (1 - p)
    s = s + (s.detach().bfloat16().float() - s.detach())
    yr = F.layer_norm(s, (D,), wf, bf, 1e-5)
    yr.backward(go.float())
    for g, wv, name in zip((y, x.grad, r.grad, w.grad, b.grad), (yr, xf.grad, rf.grad, wf.grad, bf.grad), "yxrwb"):
        err = (g.detach().float() - wv.detach()).abs().max().item() / (wv.abs().max().item() + 1e-9)
        assert err < 2e-2, (name, err)
    assert ((x.grad == 0) == ~keep).float().mean().item() > 0.999   # dropped positions get exactly zero gradient


def test_padded_vocab_head_and_cross_entropy():
    """LM-head decoder on the native GEMM with the vocabulary padded to 16-byte rows + CE on the padded logits:
    loss and the gradients of the hidden states, decoder weight and bias against F.linear + F.cross_entropy (fp32)."""
    from sceneverse_b200 import ops
    import torch.nn.functional as F
    R, K, V = 384, 768, 3001                      # odd class count, like BERT's 30522
    h = rand(R, K, seed=1).bfloat16().requires_grad_(True)
    W = (rand(V, K, seed=2) * 0.05).requires_grad_(True)
    b = (rand(V, seed=3) * 0.1).requires_grad_(True)
    labels = torch.randint(0, V, (R,), device="cuda")
    labels[torch.rand(R, device="cuda") < 0.7] = -1
    logits = ops.padded_vocab_linear(h, W, b)
    assert logits.shape == (R, V) and logits._sv_padded.shape == (R, 3008)
    assert float(logits._sv_padded[:, V:].abs().max()) == 0.0
    loss = ops.cross_entropy(logits, labels, ignore_index=-1)
    loss.backward()
    hf, Wf, bf = (t.detach().float().requires_grad_(True) for t in (h, W, b))
    ref_logits = F.linear(hf, Wf.bfloat16().float(), bf)
    ref = F.cross_entropy(ref_logits, labels, ignore_index=-1)
    ref.backward()
    assert (logits.float() - ref_logits).abs().max().item() < 2e-2 * ref_logits.abs().max().item()
    assert abs(float(loss) - float(ref)) < 5e-3 * max(1.0, abs(float(ref)))
    for g, w, name in zip((h.grad, W.grad, b.grad), (hf.grad, Wf.grad, bf.grad), "hWb"):
        assert g.dtype == (torch.bfloat16 if name == "h" else torch.float32)
        err = (g.float() - w).abs().max().item() / (w.abs().max().item() + 1e-12)
        assert err < 3e-2, (name, err)
    # the plain slice view still works for consumers that do not know about the padding
    loss2 = F.cross_entropy(ops.padded_vocab_linear(h.detach(), W.detach(), b.detach()).float(), labels, ignore_index=-1)
    assert abs(float(loss2) - float(ref)) < 5e-3 * max(1.0, abs(float(ref)))


def test_dropout_seed_offset_counter_changes_masks():
    """sv_dropout_seed_offset: with a registered device counter the same (frozen) seed gives a new mask whenever the
    counter changes and the same mask when it does not — what a replayed CUDA graph relies on."""
    from sceneverse_b200 import native, _lib
    B, H, L, E = 2, 12, 64, 768
    q, k, v = (rand(B, L, E, seed=i).bfloat16() for i in (1, 2, 3))
    counter = torch.zeros(1, dtype=torch.int64, device="cuda")
    lib = _lib.gps()
    try:
        assert lib.sv_dropout_seed_offset(counter.data_ptr()) == 0
        a = native.attention(q, k, v, H, dropout_p=0.2, seed=77).clone()
        b = native.attention(q, k, v, H, dropout_p=0.2, seed=77).clone()
        counter.add_(1)
        c = native.attention(q, k, v, H, dropout_p=0.2, seed=77).clone()
        assert torch.equal(a, b) and not torch.equal(a, c)
    finally:
        lib.sv_dropout_seed_offset(None)
    d = native.attention(q, k, v, H, dropout_p=0.2, seed=77)
    assert torch.equal(a, d)                                               # counter 0 == no counter


@pytest.mark.parametrize("R,N,dtype", [(19200, 768, torch.bfloat16), (3200, 3072, torch.bfloat16), (77, 8, torch.float32),
                                       (5000, 30528, torch.bfloat16), (1, 64, torch.bfloat16)])
def test_colsum_kernel(R, N, dtype):
    from sceneverse_b200 import native
    x = rand(R, N, seed=R + N).to(dtype)
    got = native.colsum(x)
    want = x.double().sum(0)
    assert got.dtype == torch.float32
    assert (got.double() - want).abs().max().item() <= 1e-5 * max(1.0, x.double().abs().sum(0).max().item())
    assert torch.equal(got, native.colsum(x))                       # deterministic
    if N >= 16:                                                       # strided rows (a column slice of a wider matrix)
        assert torch.allclose(native.colsum(x[:, : N // 2]), got[: N // 2], rtol=1e-6, atol=1e-6)


def test_linear_fn_matches_autocast_linear():
    """ops.linear's training path (native GEMMs, fused bias gradient, fp32 weight gradient) vs F.linear under autocast."""
    from sceneverse_b200 import ops
    import torch.nn.functional as F
    x = rand(4, 130, 768, seed=1).requires_grad_(True)
    W = (rand(2048, 768, seed=2) * 0.05).requires_grad_(True)
    b = (rand(2048, seed=3) * 0.1).requires_grad_(True)
    go = rand(4, 130, 2048, seed=4)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = ops.linear(x, W, b, activation="relu")
    assert isinstance(y.grad_fn, ops._LinearFn._backward_cls)      # bias + ReLU live in the GEMM epilogue: one autograd node
    y.backward(go.bfloat16())
    got = [y.detach().float(), x.grad.clone(), W.grad.clone(), b.grad.clone()]
    assert W.grad.dtype == torch.float32 and b.grad.dtype == torch.float32 and x.grad.dtype == torch.float32
    for t in (x, W, b):
        t.grad = None
    # fp32 reference on the bf16-rounded operands (torch's own bf16 GEMMs may reduce in bf16:
    # torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction)
    xr = x.detach().bfloat16().float().requires_grad_(True)
    Wr = W.detach().bfloat16().float().requires_grad_(True)
    br = b.detach().clone().requires_grad_(True)
    yr = F.relu(F.linear(xr, Wr, br))
    yr.backward(go.bfloat16().float())
    for g, w, name in zip(got, (yr.detach().float(), xr.grad, Wr.grad, br.grad), "yxWb"):
        err = (g - w).abs().max().item() / (w.abs().max().item() + 1e-9)
        assert err < 1e-2, (name, err)


def test_rerouted_bert_matches_huggingface_bert_gpu():
    """The text encoder (reference: modules/language/bert.py:8-26, HF BertModel with 4 layers) with its linears, self-attention,
    feed-forward and LayerNorms re-routed onto the native kernels vs the UNPATCHED HuggingFace module with the same weights:
    eval mode, bf16 autocast for the native path, fp32 for HF; also one training-mode backward to the embeddings."""
    from transformers import BertConfig, BertModel
    from sceneverse_b200 import model as M
    from sceneverse_b200.modules.layers import route_linears
    torch.manual_seed(0)
    enc = route_linears(M.BERTLanguageEncoder(None)).cuda().eval()
    ref = BertModel(BertConfig(hidden_size=768, num_hidden_layers=4, num_attention_heads=12, type_vocab_size=2)).cuda().eval()
    ref.load_state_dict(enc.model.state_dict())
    ids = torch.randint(1000, 30000, (8, 50), device="cuda")
    mask = (torch.arange(50, device="cuda")[None, :] < torch.tensor([50, 12, 30, 50, 8, 44, 50, 25], device="cuda")[:, None]).long()
    with torch.no_grad():
        want = ref(ids, mask).last_hidden_state
        with torch.autocast("cuda", dtype=torch.bfloat16):
            got = enc(ids, mask)
    valid = mask.bool()
    err = (got.float() - want)[valid].abs().max().item() / want[valid].abs().max().item()
    assert err < 3e-2, err
    # training mode, dropout off: gradients reach every parameter the HF module trains (except the unused pooler)
    enc.train()
    for m in enc.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    # a random read-out: the plain sum of a LayerNorm output is constant, its gradient would be rounding noise
    proj = torch.randn(50, 768, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = enc(ids, mask)
    (out.float() * proj * valid[..., None]).sum().backward()
    missing = [n for n, p in enc.named_parameters() if p.grad is None and "pooler" not in n]
    assert not missing, missing
    ref.train()
    for m in ref.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    (ref(ids, mask).last_hidden_state * proj * valid[..., None]).sum().backward()
    # the key bias gradient is mathematically zero (softmax is invariant to a per-query shift): compare against a floor
    scale = max(q.grad.abs().max().item() for n, q in ref.named_parameters() if "pooler" not in n)
    for (n, p), (_, q) in zip(enc.model.named_parameters(), ref.named_parameters()):
        if "pooler" in n:
            continue
        rel = (p.grad.float() - q.grad).abs().max().item() / (q.grad.abs().max().item() + 2e-3 * scale)
        assert rel < 8e-2, (n, rel)


@pytest.mark.parametrize("shape,dtype", [((64, 80, 768), torch.float32), ((64, 768), torch.float32), ((37, 768), torch.bfloat16),
                                          ((5, 3, 264), torch.float32)])
def test_l2_normalize_matches_torch_forward_and_backward(shape, dtype):
    """ops.l2_normalize (csrc/layer_norm.cu l2norm_*) against F.normalize(dim=-1, p=2) and its autograd, incl. an all-zero row
    (clamped by eps: y = 0, dx = g / eps)."""
    import torch.nn.functional as F
    from sceneverse_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(*shape, device="cuda", generator=g).to(dtype)
    x.view(-1, shape[-1])[1].zero_()
    go = torch.randn(*shape, device="cuda", generator=g).to(dtype)
    x1 = x.clone().requires_grad_()
    y1 = ops.l2_normalize(x1)
    y1.backward(go)
    x2 = x.float().clone().requires_grad_()
    y2 = F.normalize(x2, dim=-1, p=2)
    y2.backward(go.float())
    tol = 1e-5 if dtype == torch.float32 else 1.2e-2
    assert y1.dtype == dtype and torch.allclose(y1.float(), y2, atol=tol, rtol=tol)
    live = torch.ones(x.view(-1, shape[-1]).shape[0], dtype=torch.bool, device="cuda")
    live[1] = False
    d1, d2 = x1.grad.float().view(-1, shape[-1]), x2.grad.view(-1, shape[-1])
    assert torch.allclose(d1[live], d2[live], atol=tol * 4, rtol=tol * 4)
    # the clamped row: torch gives g / eps as well (the norm's own gradient at 0 is 0)
    assert torch.allclose(d1[1], d2[1], rtol=1e-2 if dtype == torch.bfloat16 else 1e-5, atol=0)

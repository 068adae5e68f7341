"""The three GEMMs of a linear layer on the native tcgen05 kernel family (csrc/gemm.cu: sv_linear_fwd / dgrad / wgrad) and
the autograd Functions built on them (ops.linear, ops.ffn) against torch's fp32 formulation of the same bf16 operands
(reference call sites: F.linear + its AddmmBackward in every layer of modules/layers/transformers.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from .test_attention_gpu import _dropout_keep

pytestmark = pytest.mark.gpu


def rnd(*s, seed=0, scale=0.5):
    g = torch.Generator(device="cuda").manual_seed(seed + sum(s))
    return torch.randn(*s, device="cuda", generator=g) * scale


@pytest.fixture(params=[0, 1, 2, 4], ids=["auto", "cta1", "cta2", "cta2x2_multicast"])
def ctas(request):
    """Every shape runs on the single-CTA kernels, on the CTA-pair (cta_group::2) kernels and under the heuristic."""
    from sceneverse_b200 import native
    native.gemm_force_ctas(request.param)
    yield request.param
    native.gemm_force_ctas(0)


def rel(got, want):
    return (got.float() - want.float()).abs().max().item() / (want.float().abs().max().item() + 1e-9)


@pytest.mark.parametrize("M,N,K", [(8320, 2304, 768), (19200, 768, 3072), (5120, 72, 768), (130, 768, 768), (3, 768, 768),
                                   (1000, 600, 136), (3200, 3072, 768)])
@pytest.mark.parametrize("act", [None, "relu", "gelu"])
def test_linear_fwd_epilogues(M, N, K, act, ctas):
    from sceneverse_b200 import native
    x, w, b = rnd(M, K).bfloat16(), rnd(N, K, scale=K ** -0.5).bfloat16(), rnd(N, seed=1)
    out, pre = native.linear_fwd(x, w, b, act, want_pre=True)
    want_pre = x.float() @ w.float().t() + b
    want = want_pre if act is None else (torch.relu(want_pre) if act == "relu" else F.gelu(want_pre))
    assert rel(pre, want_pre) < 8e-3 and rel(out, want) < 8e-3
    out32 = native.linear_fwd(x, w, b, act, out_dtype=torch.float32)
    assert rel(out32, want) < 2e-3          # fp32 store: only the accumulation order and the erf approximation differ


@pytest.mark.parametrize("act", ["relu", "gelu"])
def test_linear_fwd_dropout_mask_is_the_documented_hash(act, ctas):
    from sceneverse_b200 import native
    M, N, K, p, seed = 777, 2048, 768, 0.1, 123456789
    x, w, b = rnd(M, K).bfloat16(), rnd(N, K, scale=K ** -0.5).bfloat16(), rnd(N, seed=1)
    out = native.linear_fwd(x, w, b, act, dropout_p=p, seed=seed, out_dtype=torch.float32)
    keep = torch.from_numpy(_dropout_keep(seed, 1, 1, M, N, p)).cuda().view(M, N)
    pre = x.float() @ w.float().t() + b
    want = (torch.relu(pre) if act == "relu" else F.gelu(pre)) * keep / (1 - p)
    assert abs(keep.float().mean().item() - (1 - p)) < 0.01
    assert rel(out, want) < 2e-3
    assert ((out == 0) | keep).all()


@pytest.mark.parametrize("M,N,Kin", [(8320, 768, 2048), (19200, 3072, 768), (5120, 2304, 768), (100, 8, 384), (3200, 30528, 768)])
def test_linear_dgrad_plain(M, N, Kin, ctas):
    from sceneverse_b200 import native
    g, w = rnd(M, N).bfloat16(), rnd(N, Kin, scale=N ** -0.5).bfloat16()
    got = native.linear_dgrad(g, w, out_dtype=torch.float32)
    assert rel(got, g.float() @ w.float()) < 2e-3


@pytest.mark.parametrize("act,p", [("relu", 0.0), ("relu", 0.1), ("gelu", 0.0), ("gelu", 0.1)])
def test_linear_dgrad_activation_derivative_epilogue(act, p, ctas):
    """dgrad of layer 2 of an FFN with d/dpre of dropout(act(pre)) fused: against autograd of the same expression."""
    from sceneverse_b200 import native
    M, N, H, seed = 1000, 768, 2048, 424242
    g, w2 = rnd(M, N).bfloat16(), rnd(N, H, scale=N ** -0.5).bfloat16()
    pre = rnd(M, H, seed=3, scale=1.5).bfloat16()
    keep = torch.from_numpy(_dropout_keep(seed, 1, 1, M, H, p)).cuda().view(M, H) if p > 0 else torch.ones(M, H, device="cuda", dtype=torch.bool)
    pf = pre.float().requires_grad_(True)
    h = (torch.relu(pf) if act == "relu" else F.gelu(pf)) * keep / (1 - p)
    (h @ w2.float().t()).backward(g.float())
    aux = h.detach().bfloat16() if act == "relu" else pre
    got = native.linear_dgrad(g, w2, act=act, aux=aux, dropout_p=p, seed=seed, out_dtype=torch.float32)
    assert rel(got, pf.grad) < 3e-3


@pytest.mark.parametrize("M,N,Kin", [(19200, 768, 768), (8320, 2304, 768), (19200, 3072, 768), (8320, 768, 2048), (5120, 72, 768),
                                     (2, 768, 768), (333, 607, 384), (3200, 30522, 768)])
def test_linear_wgrad_split_k_bias_and_accumulate(M, N, Kin, ctas):
    from sceneverse_b200 import native
    Np = (N + 7) // 8 * 8
    g = torch.zeros(M, Np, device="cuda", dtype=torch.bfloat16)
    g[:, :N] = rnd(M, N).bfloat16()
    x = rnd(M, Kin, seed=5).bfloat16()
    want_w = g[:, :N].float().t() @ x.float()
    want_b = g[:, :N].float().sum(0)
    dw, db = native.linear_wgrad(g, x, n_out=N, want_db=True)
    assert rel(dw, want_w) < 2e-3 and rel(db, want_b) < 2e-3
    base_w, base_b = rnd(N, Kin, seed=9), rnd(N, seed=10)        # accumulate into an existing gradient
    dw2, db2 = base_w.clone(), base_b.clone()
    native.linear_wgrad(g, x, n_out=N, dw=dw2, db=db2, accumulate=True)
    assert rel(dw2, base_w + want_w) < 2e-3 and rel(db2, base_b + want_b) < 2e-3


@pytest.mark.parametrize("shape,N,act", [((64, 130, 768), 2304, None), ((5120, 768), 607, None), ((64, 80, 6), 768, None),
                                         ((64, 80, 384), 1, None), ((640, 768), 768, "gelu"), ((640, 384), 384, "relu"),
                                         ((2, 768), 607, None)])
def test_ops_linear_autograd_matches_torch(shape, N, act):
    """ops.linear (padded N / padded K / tiny M included) forward + all three gradients vs torch autograd in fp32."""
    from sceneverse_b200 import ops
    K = shape[-1]
    x = rnd(*shape).bfloat16().requires_grad_(True)
    w = rnd(N, K, scale=K ** -0.5).requires_grad_(True)
    b = rnd(N, seed=2).requires_grad_(True)
    go = rnd(*shape[:-1], N, seed=3).bfloat16()
    y = ops.linear(x, w, b, activation=act)
    assert y.shape == (*shape[:-1], N) and y.dtype == torch.bfloat16
    y.backward(go)
    xf = x.detach().float().requires_grad_(True)
    wf = w.detach().bfloat16().float().requires_grad_(True)
    bf = b.detach().clone().requires_grad_(True)
    yr = F.linear(xf, wf, bf)
    yr = yr if act is None else (torch.relu(yr) if act == "relu" else F.gelu(yr))
    yr.backward(go.float())
    assert rel(y, yr) < 8e-3
    assert rel(x.grad, xf.grad) < 1e-2 and rel(w.grad, wf.grad) < 1e-2 and rel(b.grad, bf.grad) < 1e-2


@pytest.mark.parametrize("act,p,H", [("relu", 0.1, 2048), ("gelu", 0.1, 2048), ("gelu", 0.0, 3072)])
def test_ops_ffn_autograd_matches_torch(act, p, H):
    from sceneverse_b200 import ops
    M, D = 2000, 768
    x = rnd(M, D).bfloat16().requires_grad_(True)
    w1, b1 = rnd(H, D, scale=D ** -0.5).requires_grad_(True), rnd(H, seed=1).requires_grad_(True)
    w2, b2 = rnd(D, H, scale=H ** -0.5).requires_grad_(True), rnd(D, seed=2).requires_grad_(True)
    go = rnd(M, D, seed=3).bfloat16()
    ops._dropout_calls[0] = 41
    seed = (torch.initial_seed() * 0x9E3779B1 + 42 * 0x85EBCA6B) & 0x7FFFFFFFFFFFFFFF     # what _next_dropout_seed will hand out
    y = ops.ffn(x, w1, b1, w2, b2, activation=act, dropout_p=p)
    y.backward(go)
    keep = torch.from_numpy(_dropout_keep(seed, 1, 1, M, H, p)).cuda().view(M, H) if p > 0 else 1.0
    xf = x.detach().float().requires_grad_(True)
    ps = [t.detach().clone().requires_grad_(True) for t in (w1, b1, w2, b2)]
    pre = F.linear(xf, ps[0].bfloat16().float(), ps[1])
    h = (torch.relu(pre) if act == "relu" else F.gelu(pre)) * keep / (1 - p)
    h = h + (h.detach().bfloat16().float() - h.detach())          # the kernel stores h in bf16
    yr = F.linear(h, ps[2].bfloat16().float(), ps[3])
    yr.backward(go.float())
    assert rel(y, yr) < 1e-2
    assert rel(x.grad, xf.grad) < 3e-2       # dL/dpre passes through bf16 between the two dgrad GEMMs
    for got, want, name in zip((w1.grad, b1.grad, w2.grad, b2.grad), ps, ("w1", "b1", "w2", "b2")):
        assert rel(got, want.grad) < 2.5e-2, name


def test_direct_gradient_accumulation_into_existing_grad():
    """ops.DIRECT_GRAD: the wgrad kernel adds into weight.grad / bias.grad (the flat buffer of train.PretrainStep) and
    autograd receives no parameter gradient; a parameter used twice accumulates both contributions."""
    from sceneverse_b200 import ops
    x1, x2 = rnd(300, 768).bfloat16(), rnd(500, 768, seed=4).bfloat16()
    w, b = rnd(768, 768, scale=0.03).requires_grad_(True), rnd(768, seed=1).requires_grad_(True)
    w.grad, b.grad = torch.ones_like(w), torch.ones_like(b)
    ops.DIRECT_GRAD[0] = True
    try:
        (ops.linear(x1, w, b).float().sum() + 2 * ops.linear(x2, w, b).float().sum()).backward()
    finally:
        ops.DIRECT_GRAD[0] = False
    want_w = 1 + torch.ones(300, 768, device="cuda").t() @ x1.float() + 2 * torch.ones(500, 768, device="cuda").t() @ x2.float()
    assert rel(w.grad, want_w) < 2e-3 and rel(b.grad, torch.full_like(b, 1 + 300 + 1000)) < 2e-3


def test_embedding_backward_scatter_add():
    from sceneverse_b200 import native
    ids = torch.randint(0, 1000, (64, 50), device="cuda")
    ids[:, -3:] = 0
    g = rnd(64, 50, 768).bfloat16()
    dw = torch.zeros(1000, 768, device="cuda")
    native.embedding_bwd(g, ids, dw, padding_idx=0)
    emb = torch.nn.Embedding(1000, 768, padding_idx=0).cuda()
    emb(ids).backward(g.float())
    assert rel(dw, emb.weight.grad) < 1e-5


def test_linear_packed_and_attention_packed_match_the_unpacked_path():
    """q | k | v as ONE projection GEMM + attention on the packed tensor (strided TMA maps, dq | dk | dv written into one packed
    gradient) must equal three separate linears + ops.attention, forward and backward."""
    from sceneverse_b200 import ops
    B, L, E, H = 8, 130, 768, 12
    x = rnd(B, L, E).bfloat16().requires_grad_(True)
    lins = [torch.nn.Linear(E, E).cuda() for _ in range(3)]
    kpm = torch.zeros(B, L, dtype=torch.bool, device="cuda")
    kpm[:, 100:] = True
    go = rnd(B, L, E, seed=7).bfloat16()
    out = ops.attention_packed(ops.linear_packed(x, lins), H, key_padding_mask=kpm)
    out.backward(go)
    got = [out.detach(), x.grad.clone()] + [p.grad.clone() for l in lins for p in (l.weight, l.bias)]
    x.grad = None
    for l in lins:
        l.weight.grad = l.bias.grad = None
    q, k, v = (ops.linear(x, l.weight, l.bias) for l in lins)
    ref = ops.attention(q, k, v, H, key_padding_mask=kpm)
    ref.backward(go)
    want = [ref.detach(), x.grad] + [p.grad for l in lins for p in (l.weight, l.bias)]
    for g, w, name in zip(got, want, ["out", "dx", "dWq", "dbq", "dWk", "dbk", "dWv", "dbv"]):
        assert rel(g, w) < 1e-2, name


@pytest.mark.parametrize("V,pad", [(1000, 0), (2, None), (512, None)])
def test_ops_embedding_autograd_matches_torch(V, pad):
    from sceneverse_b200 import ops
    ids = torch.randint(0, V, (64, 50), device="cuda")
    w = rnd(V, 768).requires_grad_(True)
    go = rnd(64, 50, 768, seed=3)
    ops.embedding(ids, w, pad).backward(go)
    wr = w.detach().clone().requires_grad_(True)
    F.embedding(ids, wr, pad).backward(go)
    assert rel(w.grad, wr.grad) < 1e-5


def test_gemm_profile_counters_and_timing_switches():
    """sv_gemm_profile: every launch writes the issuing thread's cycle counters per CTA; the k-steps it reports are the
    work list of the launch.  The timing-only switches (bits 8.. of sv_gemm_force_ctas) must leave the kernel runnable."""
    from sceneverse_b200 import _lib, native
    lib = _lib.gps()
    M, N, K = 2048, 768, 512                      # 8 x 3 pair tiles of 256 x 256, 8 k-steps each
    x, w = rnd(M, K).bfloat16(), rnd(N, K, seed=1, scale=0.05).bfloat16()
    b = torch.zeros(N, device="cuda")
    prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    try:
        native.gemm_force_ctas(2)
        assert lib.sv_gemm_profile(prof.data_ptr()) == 0
        want = native.linear_fwd(x, w, b).float()
        torch.cuda.synchronize()
        p = prof.view(148, 16)
        total = int(p[:, 3].sum())                                             # k-steps issued over all leaders:
        assert total > 0 and total % ((M // 256) * (K // 64)) == 0            # row tiles x k-steps x (column tiles of the pick)
        act = p[:, 3] > 0
        assert bool((p[act, 0] > 0).all()) and bool((p[act, 1] <= p[act, 0]).all())   # loop cycles; waiting is part of them
        assert bool((p[act, 6] > 0).all())                                    # globaltimer ns of the loop
        lib.sv_gemm_profile(None)
        for dbg in (1, 2, 3):                     # no epilogue body / no loads / neither: garbage out, but it must finish
            native.gemm_force_ctas(2 | (dbg << 8))
            native.linear_fwd(x, w, b)
            torch.cuda.synchronize()
        native.gemm_force_ctas(2)
        again = native.linear_fwd(x, w, b).float()
        assert torch.equal(again, want)           # the switches leave no state behind
    finally:
        lib.sv_gemm_profile(None)
        native.gemm_force_ctas(0)

"""Operator dispatch of the attention stack.  Each op has ONE implementation per device class:
CUDA tensors go to the native library where a kernel exists (no silent fallback: if the library is
missing the call raises) and to cuBLAS/ATen for the rest; CPU tensors are only accepted for
shape/contract tests of the host logic.  `NATIVE` lists which ops are hand-written kernels.
"""
import math

import torch
import torch.nn.functional as F

NATIVE = {"linear": False,              # plain GEMM -> cuBLAS (library GEMM); fused GEMMs (SA3, fc) use native.gemm
          "attention": True,            # nn.MultiheadAttention core incl. attention dropout: native tcgen05 forward + backward
          "spatial_attention": True,    # MultiHeadAttentionSpatial core: native tcgen05 forward and backward
          "calc_pairwise_locs": True,
          "cross_entropy": True,        # masked-LM / grounding CE: fused native forward+gradient
          "layer_norm": True}           # dropout + residual add + LayerNorm: one native pass per direction


_dropout_calls = [0]


def _next_dropout_seed():
    """Deterministic per-call seed of the in-kernel attention dropout: torch's global seed + a call counter (no host
    sync, CUDA-graph safe); each call then hashes (seed, b, h, i, j)."""
    _dropout_calls[0] += 1
    return (torch.initial_seed() * 0x9E3779B1 + _dropout_calls[0] * 0x85EBCA6B) & 0x7FFFFFFFFFFFFFFF


def _native_ok(*tensors):
    return all(t.is_cuda and t.dtype == torch.bfloat16 for t in tensors)


class _AttentionFn(torch.autograd.Function):
    """Fused tcgen05 attention: native forward (keeps the per-query log-sum-exp) and native backward (two kernels that
    recompute P).  spatial_n_head must equal n_head when a gate is given (one weight set per head)."""

    @staticmethod
    def forward(ctx, q, k, v, sw, locs, kpm, n_head, spatial_n_head, dropout_p=0.0, seed=0):
        from . import native
        out, lse = native.attention(q, k, v, n_head, key_padding_mask=kpm, spatial_w=sw, spatial_heads=spatial_n_head,
                                    pairwise_locs=locs, return_lse=True, dropout_p=dropout_p, seed=seed)
        ctx.save_for_backward(q, k, v, out, lse, sw, locs, kpm)
        ctx.n_head, ctx.drop = n_head, (dropout_p, seed)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        from . import native
        q, k, v, out, lse, sw, locs, kpm = ctx.saved_tensors
        dq, dk, dv, dsw = native.attention_backward(q, k, v, out, grad_out, lse, ctx.n_head, key_padding_mask=kpm,
                                                    spatial_w=sw, pairwise_locs=locs, dropout_p=ctx.drop[0], seed=ctx.drop[1])
        return dq, dk, dv, dsw, None, None, None, None, None, None


class _SpatialAttentionRecomputeFn(torch.autograd.Function):
    """Native forward, torch-recompute backward: only for a gate shared across heads (spatial_multihead=False)."""

    @staticmethod
    def forward(ctx, q, k, v, sw, locs, kpm, n_head, spatial_n_head):
        from . import native
        ctx.save_for_backward(q, k, v, sw, locs, kpm)
        ctx.heads = (n_head, spatial_n_head)
        return native.attention(q, k, v, n_head, key_padding_mask=kpm, spatial_w=sw, spatial_heads=spatial_n_head,
                                pairwise_locs=locs)

    @staticmethod
    def backward(ctx, grad_out):
        q, k, v, sw, locs, kpm = ctx.saved_tensors
        n_head, spatial_n_head = ctx.heads
        with torch.enable_grad():
            qq, kk, vv, ss = (t.detach().float().requires_grad_(True) for t in (q, k, v, sw))
            out, _ = _spatial_attention_torch(qq, kk, vv, ss, locs, n_head, spatial_n_head, kpm)
            gq, gk, gv, gs = torch.autograd.grad(out, (qq, kk, vv, ss), grad_out.float())
        return gq.to(q.dtype), gk.to(k.dtype), gv.to(v.dtype), gs.to(sw.dtype), None, None, None, None


import os as _os

# opt-in: run the two backward GEMMs of every linear on the native tcgen05 GEMM with transposed (MN-major) operands
# instead of cuBLAS.  The kernel path is parity-tested (tests/test_gemm_gpu.py::test_gemm_transposed_operands); the
# whole-step timing with it has not been measured yet, so the default stays on the library GEMMs.
_NATIVE_BWD_GEMM = _os.environ.get("SVB200_NATIVE_BWD_GEMM", "0") == "1"

_MM_OUT_DTYPE = [None]  # does torch.mm accept out_dtype on this build? (decided at first use)


def _mm_f32(a, b):
    """a @ b with bf16 operands and an fp32 result (weight gradients accumulate into fp32 parameters)."""
    if _MM_OUT_DTYPE[0] is None:
        try:
            torch.mm(a[:8, :8].contiguous(), b[:8, :8].contiguous(), out_dtype=torch.float32)
            _MM_OUT_DTYPE[0] = True
        except Exception:
            _MM_OUT_DTYPE[0] = False
    if _MM_OUT_DTYPE[0]:
        return torch.mm(a, b, out_dtype=torch.float32)
    return torch.mm(a, b).float()


# bf16 shadows of fp32 parameters, refreshed once per step with ONE multi-tensor copy (train.PretrainStep) instead of one
# cast kernel per weight and per bias inside every linear (autocast's behaviour: ~200 tiny launches per step)
_SHADOW = {}
_SHADOW_LISTS = [[], []]   # (bf16 destinations, fp32 sources)


def register_shadows(module):
    """Create a bf16 shadow for every floating-point parameter of `module` that the bf16 linears read."""
    for p in module.parameters():
        if p.is_cuda and p.dtype == torch.float32 and p.dim() >= 1 and id(p) not in _SHADOW:
            s = torch.empty_like(p, dtype=torch.bfloat16)
            _SHADOW[id(p)] = (s, p)        # the strong reference to p keeps its id from being reused
            _SHADOW_LISTS[0].append(s)
            _SHADOW_LISTS[1].append(p.detach())
    refresh_shadows()


def refresh_shadows():
    """Must run after every parameter update and before the next forward that should use the shadows."""
    if _SHADOW_LISTS[0]:
        torch._foreach_copy_(_SHADOW_LISTS[0], _SHADOW_LISTS[1])


def clear_shadows():
    _SHADOW.clear()
    _SHADOW_LISTS[0].clear()
    _SHADOW_LISTS[1].clear()


def _bf16_of(p):
    e = _SHADOW.get(id(p))
    return e[0] if e is not None and e[1] is p else p.to(torch.bfloat16)


class _LinearFn(torch.autograd.Function):
    """Training-path linear in bf16: y = x W^T + b.  The three GEMMs are plain library GEMMs (cuBLAS, the weight is cast
    to bf16 once per call exactly as autocast would); the bias gradient — a strided ATen reduction in the reference's
    AddmmBackward, ~2 ms per step over ~100 layers — is the native column-sum kernel; the weight gradient is produced
    directly in fp32."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        wb = _bf16_of(weight)
        x2 = x.reshape(-1, x.shape[-1])
        if x2.dtype != torch.bfloat16:
            x2 = x2.to(torch.bfloat16)
        ctx.save_for_backward(x2, wb)
        ctx.in_shape, ctx.in_dtype = x.shape, x.dtype
        y = F.linear(x2, wb, _bf16_of(bias) if bias is not None else None)
        return y.view(*x.shape[:-1], -1)

    @staticmethod
    def backward(ctx, g):
        from . import native
        x2, wb = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        if g2.dtype != torch.bfloat16:
            g2 = g2.to(torch.bfloat16)
        g2 = g2.contiguous()
        dx = dw = db = None
        native_ok = _NATIVE_BWD_GEMM and x2.is_contiguous() and wb.is_contiguous()
        if ctx.needs_input_grad[0]:
            # dgrad: (M,N) . (N,Kin) — the weight is the transposed (MN-major) B operand
            dx = native.gemm_ex(g2, wb, b_transposed=True) if native_ok else torch.mm(g2, wb)
            dx = dx.view(ctx.in_shape).to(ctx.in_dtype)
        if ctx.needs_input_grad[1]:
            # wgrad: g^T (N,M) . x (M,Kin) — both operands transposed in memory, fp32 result
            dw = native.gemm_ex(g2, x2, a_transposed=True, b_transposed=True, out_dtype=torch.float32) if native_ok \
                else _mm_f32(g2.t(), x2)
        if ctx.needs_input_grad[2]:
            db = native.colsum(g2)
        return dx, dw, db


def linear(x, weight, bias=None, activation=None):
    if x.is_cuda and weight.dtype == torch.float32 and (x.dtype == torch.bfloat16 or _autocast_on()) and \
            torch.is_grad_enabled() and (weight.requires_grad or x.requires_grad) and weight.shape[0] % 8 == 0 and \
            weight.shape[1] % 8 == 0 and x.numel() > 0:
        y = _LinearFn.apply(x, weight, bias)
    else:
        y = F.linear(x, weight, bias)
    if activation == "relu":
        y = F.relu(y)
    elif activation == "gelu":
        y = F.gelu(y)
    elif activation is not None:
        raise ValueError(activation)
    return y


def calc_pairwise_locs(obj_centers, obj_whls, eps=1e-10, pairwise_rel_type='center', spatial_dist_norm=True,
                       spatial_dim=5):
    """modules/utils.py:38-87, 'center' relation (the one GPS uses): (B,O,3) -> (B,O,O,5)
    [dist/max_dist, dz/dist, dist2d/dist, dy/dist2d, dx/dist2d]; the max-distance normaliser includes padded objects."""
    if pairwise_rel_type != 'center':
        raise NotImplementedError(pairwise_rel_type)
    if obj_centers.is_cuda and obj_centers.dtype == torch.float32 and spatial_dim == 5 and not obj_centers.requires_grad \
            and obj_centers.stride(2) == 1 and obj_centers.stride(0) == obj_centers.size(1) * obj_centers.stride(1):
        from . import _lib
        B, O, _ = obj_centers.shape
        out = torch.empty((B, O, O, 5), dtype=torch.float32, device=obj_centers.device)
        lib = _lib.gps()
        with torch.cuda.device(obj_centers.device):
            st = lib.sv_pairwise_locs_f32(obj_centers.data_ptr(), obj_centers.stride(1), B, O, float(eps),
                                          1 if spatial_dist_norm else 0, out.data_ptr(),
                                          torch.cuda.current_stream(obj_centers.device).cuda_stream)
        _lib.check(lib, st, "sv_pairwise_locs_f32")
        return out
    d = obj_centers[:, :, None, :] - obj_centers[:, None, :, :]
    dist = torch.sqrt((d ** 2).sum(3) + eps)
    if spatial_dist_norm:
        max_d = dist.reshape(dist.size(0), -1).max(dim=1)[0]
        norm = dist / max_d[:, None, None]
    else:
        norm = dist
    if spatial_dim == 1:
        return norm.unsqueeze(3)
    dist2d = torch.sqrt((d[..., :2] ** 2).sum(3) + eps)
    locs = torch.stack([norm, d[..., 2] / dist, dist2d / dist, d[..., 1] / dist2d, d[..., 0] / dist2d], dim=3)
    return locs[..., 1:] if spatial_dim == 4 else locs


def attention(q, k, v, num_heads, key_padding_mask=None, dropout_p=0.0):
    """Scaled-dot-product attention on packed heads: q (B,Lq,E), k/v (B,Lk,E) -> (B,Lq,E).
    key_padding_mask (B,Lk) bool, True = ignore (nn.MultiheadAttention convention)."""
    B, Lq, E = q.shape
    Lk, hd = k.shape[1], E // num_heads
    if hd == 64 and Lk <= 384 and _native_ok(q, k, v):
        needs_grad = torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad)
        if not needs_grad and dropout_p == 0.0:
            from . import native
            return native.attention(q, k, v, num_heads, key_padding_mask=key_padding_mask)
        if Lq <= 384:
            return _AttentionFn.apply(q, k, v, None, None, key_padding_mask, num_heads, 0, float(dropout_p),
                                      _next_dropout_seed() if dropout_p > 0.0 else 0)
    qh = q.view(B, Lq, num_heads, hd).transpose(1, 2)
    kh = k.view(B, Lk, num_heads, hd).transpose(1, 2)
    vh = v.view(B, Lk, num_heads, hd).transpose(1, 2)
    mask = None
    if key_padding_mask is not None:
        mask = key_padding_mask.logical_not()[:, None, None, :]
    out = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=mask, dropout_p=dropout_p)
    return out.transpose(1, 2).reshape(B, Lq, E)


def spatial_attention(q, k, v, spatial_weights, pairwise_locs, n_head, spatial_n_head, key_padding_mask=None):
    """Dispatch: bf16 CUDA tensors with head dim 64 -> fused tcgen05 kernel (attention map not returned: every reference
    caller discards it); anything else -> the torch formulation below."""
    if _native_ok(q, k, v) and q.shape[-1] // n_head == 64 and k.shape[1] <= 384:
        fn = _AttentionFn if (spatial_n_head == n_head and q.shape[1] <= 384) else _SpatialAttentionRecomputeFn
        out = fn.apply(q, k, v, spatial_weights.float(), pairwise_locs.float(), key_padding_mask, n_head, spatial_n_head)
        return out, None
    return _spatial_attention_torch(q, k, v, spatial_weights, pairwise_locs, n_head, spatial_n_head, key_padding_mask)


def _spatial_attention_torch(q, k, v, spatial_weights, pairwise_locs, n_head, spatial_n_head, key_padding_mask=None):
    """Core of MultiHeadAttentionSpatial 'cond' (transformers.py:188-237):
    softmax(log(clamp(sigmoid(w . loc + b), 1e-6)) + q k^T / sqrt(dh)) v, masked keys excluded.
    q,k,v (B,L,E) already projected; spatial_weights (B,L,spatial_n_head*(d+1)), per head [bias, w_1..w_d];
    pairwise_locs (B,L,T,d).  Returns (out (B,L,E), attn (H,B,L,T))."""
    B, L, E = q.shape
    T, hd = k.shape[1], E // n_head
    d = pairwise_locs.shape[-1]
    qh = q.view(B, L, n_head, hd).permute(2, 0, 1, 3)
    kh = k.view(B, T, n_head, hd).permute(2, 0, 1, 3)
    vh = v.view(B, T, n_head, hd).permute(2, 0, 1, 3)
    attn = torch.einsum('hblk,hbtk->hblt', qh, kh) / math.sqrt(hd)
    sw = spatial_weights.view(B, L, spatial_n_head, d + 1).permute(2, 0, 1, 3)
    if spatial_n_head == 1:
        sw = sw.expand(n_head, -1, -1, -1)
    loc = torch.sigmoid(torch.einsum('hbld,bltd->hblt', sw[..., 1:], pairwise_locs) + sw[..., :1])
    if key_padding_mask is not None:
        m = key_padding_mask[None, :, None, :]
        attn = attn.masked_fill(m, float('-inf'))
        loc = loc.masked_fill(m, 0)
    fused = torch.softmax(torch.log(torch.clamp(loc, min=1e-6)) + attn, 3)
    out = torch.einsum('hblt,hbtv->hblv', fused, vh).permute(1, 2, 0, 3).reshape(B, L, E)
    return out, fused


class _CrossEntropyFn(torch.autograd.Function):
    """mean cross-entropy over the rows whose label != ignore_index; forward and gradient come from ONE native launch.
    n_classes < logits.shape[1]: the trailing columns are padding (padded-vocabulary logits): ignored by the loss, zero in
    the gradient, which keeps the padded row stride so the backward GEMMs stay 16-byte aligned."""

    @staticmethod
    def forward(ctx, logits2d, labels, ignore_index, n_classes):
        from . import _lib
        R, W = logits2d.shape
        V = n_classes if n_classes else W
        loss_rows = torch.empty(R, dtype=torch.float32, device=logits2d.device)
        need_grad = logits2d.requires_grad
        grad = torch.empty((R, W), dtype=logits2d.dtype, device=logits2d.device) if need_grad else None
        lib = _lib.gps()
        with torch.cuda.device(logits2d.device):
            st = lib.sv_cross_entropy_fwd_bwd_strided(
                logits2d.data_ptr(), logits2d.stride(0), 1 if logits2d.dtype == torch.bfloat16 else 0, labels.data_ptr(), R,
                V, int(ignore_index), loss_rows.data_ptr(), grad.data_ptr() if grad is not None else None, W,
                torch.cuda.current_stream(logits2d.device).cuda_stream)
        _lib.check(lib, st, "sv_cross_entropy_fwd_bwd_strided")
        count = (labels != ignore_index).sum().clamp(min=1).float()
        ctx.save_for_backward(grad, count)
        return loss_rows.sum() / count

    @staticmethod
    def backward(ctx, gout):
        grad, count = ctx.saved_tensors
        return grad.mul_((gout / count).to(grad.dtype)), None, None, None


class _PaddedVocabLinearFn(torch.autograd.Function):
    """logits_p = h @ Wp^T + bp with the class dimension zero-padded to a multiple of 64, so that every row of the logits
    and of their gradient is 16-byte aligned: the forward runs on the native tcgen05 GEMM (bias fused), the two backward
    GEMMs on aligned library kernels (an odd class count such as BERT's 30522 otherwise drops cuBLAS to a 4x slower
    legacy kernel).  Reference: modules/heads/pretrain_head.py:22-32 (decoder + bias)."""

    @staticmethod
    def forward(ctx, h2, weight, bias):
        from . import native
        V, K = weight.shape
        Vp = (V + 63) // 64 * 64
        wp = torch.empty((Vp, K), dtype=torch.bfloat16, device=h2.device)
        wp[:V].copy_(weight)
        wp[V:].zero_()
        bp = torch.zeros(Vp, dtype=torch.float32, device=h2.device)
        bp[:V].copy_(bias)
        ctx.save_for_backward(h2, wp)
        ctx.V = V
        return native.gemm(h2, wp, bias=bp)

    @staticmethod
    def backward(ctx, gp):
        h2, wp = ctx.saved_tensors
        V = ctx.V
        gp = gp.contiguous()
        dh = gp @ wp
        dw = (gp.t() @ h2)[:V].float()
        db = gp.sum(0, dtype=torch.float32)[:V]
        return dh, dw, db


def padded_vocab_linear(h, weight, bias):
    """(..., K) -> (..., V) logits of a wide classifier.  CUDA bf16: computed as padded logits (see _PaddedVocabLinearFn);
    the returned tensor is the (..., :V) view and carries the padded parent as `_sv_padded`, which ops.cross_entropy uses
    to produce the padded gradient directly.  Otherwise: ops.linear."""
    V, K = weight.shape
    if h.is_cuda and (h.dtype == torch.bfloat16 or _autocast_on()) and K % 8 == 0 and V % 64 != 0 and bias is not None:
        h2 = h.reshape(-1, K).to(torch.bfloat16).contiguous()
        lp = _PaddedVocabLinearFn.apply(h2, weight, bias).view(*h.shape[:-1], -1)
        out = lp[..., :V]
        out._sv_padded = lp
        return out
    return linear(h, weight, bias)


class _LayerNormFn(torch.autograd.Function):
    """y = LayerNorm(residual + dropout(x)) in one native pass; the backward regenerates the dropout mask and produces
    d(x), d(residual), d(gamma), d(beta) (gamma / beta partial sums reduced in a fixed order)."""

    @staticmethod
    def forward(ctx, x, residual, gamma, beta, eps, p, seed):
        from . import _lib
        D = x.shape[-1]
        x2 = x.reshape(-1, D)
        R = x2.shape[0]
        r2 = residual.reshape(-1, D) if residual is not None else None
        need_s = r2 is not None or p > 0.0
        y = torch.empty_like(x2)
        s = torch.empty_like(x2) if need_s else None
        stats = torch.empty((2, R), dtype=torch.float32, device=x.device)
        lib = _lib.gps()
        with torch.cuda.device(x.device):
            st = lib.sv_layer_norm_fwd(x2.data_ptr(), r2.data_ptr() if r2 is not None else None,
                                       1 if x.dtype == torch.bfloat16 else 0, R, D, gamma.data_ptr(), beta.data_ptr(),
                                       float(eps), float(p), int(seed), y.data_ptr(), s.data_ptr() if need_s else None,
                                       stats[0].data_ptr(), stats[1].data_ptr(),
                                       torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(lib, st, "sv_layer_norm_fwd")
        ctx.save_for_backward(s if need_s else x2, gamma, stats)
        ctx.meta = (float(p), int(seed), residual is not None, x.shape)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        s, gamma, stats = ctx.saved_tensors
        p, seed, has_res, shape = ctx.meta
        R, D = s.shape
        g2 = g.reshape(R, D).to(s.dtype).contiguous()
        ds = torch.empty_like(s)
        dx = torch.empty_like(s) if p > 0.0 else None
        dgb = torch.empty((2, D), dtype=torch.float32, device=s.device)
        lib = _lib.gps()
        scratch = torch.empty(lib.sv_layer_norm_scratch_floats(D), dtype=torch.float32, device=s.device)
        with torch.cuda.device(s.device):
            st = lib.sv_layer_norm_bwd(g2.data_ptr(), s.data_ptr(), 1 if s.dtype == torch.bfloat16 else 0, R, D,
                                       gamma.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), p, seed, ds.data_ptr(),
                                       dx.data_ptr() if dx is not None else None, dgb[0].data_ptr(), dgb[1].data_ptr(),
                                       scratch.data_ptr(), torch.cuda.current_stream(s.device).cuda_stream)
        _lib.check(lib, st, "sv_layer_norm_bwd")
        d_x = (dx if dx is not None else ds).view(shape)
        return d_x, (ds.view(shape) if has_res else None), dgb[0], dgb[1], None, None, None


def _autocast_on():
    try:
        return torch.is_autocast_enabled("cuda")
    except TypeError:
        return torch.is_autocast_enabled()


def layer_norm(x, weight, bias, eps=1e-5, residual=None, dropout_p=0.0):
    """LayerNorm(residual + dropout(x)) * weight + bias over the last dimension (post-norm block tail of
    transformers.py:145-154,311-315; plain LayerNorm with residual=None, dropout_p=0).  The caller passes dropout_p = 0
    outside training.  CUDA: one native kernel, I/O in bf16 under autocast / for bf16 inputs (statistics in fp32), else
    fp32; CPU: the torch formulation."""
    D = x.shape[-1]
    if x.is_cuda and D % 8 == 0 and 8 <= D <= 1024 and weight is not None and bias is not None and x.numel() > 0 and \
            x.dtype in (torch.bfloat16, torch.float32, torch.float16):
        low = x.dtype != torch.float32 or (residual is not None and residual.dtype != torch.float32) or _autocast_on()
        dt = torch.bfloat16 if low else torch.float32
        x_ = x.to(dt).contiguous()
        r_ = None
        if residual is not None:
            r_ = residual.to(dt).expand_as(x_).contiguous()
        return _LayerNormFn.apply(x_, r_, weight.float().contiguous(), bias.float().contiguous(), eps, float(dropout_p),
                                  _next_dropout_seed() if dropout_p > 0.0 else 0)
    s = x if dropout_p == 0.0 else F.dropout(x, dropout_p, True)
    if residual is not None:
        s = residual + s
    return F.layer_norm(s, (D,), weight, bias, eps)


def cross_entropy(logits, labels, ignore_index=-100):
    """F.cross_entropy(logits.permute(0, 2, 1) / logits, labels, ignore_index=...) with the class dimension LAST in
    `logits` ((..., V) with (...) matching labels).  CUDA bf16/f32 -> fused native kernel; otherwise torch."""
    V = logits.shape[-1]
    padded = getattr(logits, "_sv_padded", None)
    if padded is not None and padded.is_cuda and padded.shape[:-1] == logits.shape[:-1] and padded.is_contiguous():
        return _CrossEntropyFn.apply(padded.view(-1, padded.shape[-1]), labels.reshape(-1).contiguous(), ignore_index, V)
    if logits.is_cuda and logits.dtype in (torch.bfloat16, torch.float32) and logits.stride(-1) == 1:
        l2 = logits.reshape(-1, V)
        if l2.stride(1) == 1:
            return _CrossEntropyFn.apply(l2, labels.reshape(-1).contiguous(), ignore_index, 0)
    return F.cross_entropy(logits.reshape(-1, V).float(), labels.reshape(-1), ignore_index=ignore_index)

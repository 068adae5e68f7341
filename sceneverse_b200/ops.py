"""Operator dispatch of the attention stack.  Each op has ONE implementation per device class:
CUDA tensors go to the native library where a kernel exists (no silent fallback: if the library is
missing the call raises) and to cuBLAS/ATen for the rest; CPU tensors are only accepted for
shape/contract tests of the host logic.  `NATIVE` lists which ops are hand-written kernels.
"""
import math

import torch
import torch.nn.functional as F

NATIVE = {"linear": False, "attention": False, "spatial_attention": False, "calc_pairwise_locs": False}


def linear(x, weight, bias=None, activation=None):
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
    qh = q.view(B, Lq, num_heads, hd).transpose(1, 2)
    kh = k.view(B, Lk, num_heads, hd).transpose(1, 2)
    vh = v.view(B, Lk, num_heads, hd).transpose(1, 2)
    mask = None
    if key_padding_mask is not None:
        mask = key_padding_mask.logical_not()[:, None, None, :]
    out = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=mask, dropout_p=dropout_p)
    return out.transpose(1, 2).reshape(B, Lq, E)


def spatial_attention(q, k, v, spatial_weights, pairwise_locs, n_head, spatial_n_head, key_padding_mask=None):
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

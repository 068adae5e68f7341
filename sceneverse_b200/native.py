"""Thin torch-tensor front-ends of the libsvgps kernels (allocation, stream, status checks)."""
import torch

from . import _lib

_ACT = {None: 0, "none": 0, "relu": 1, "gelu": 2}


def gemm(a, w, bias=None, act=None, residual=None, out_dtype=torch.bfloat16, rowmax=0):
    """epilogue(a @ w.T): a (M,K) bf16, w (N,K) bf16 (nn.Linear layout), bias (N) f32 -> (M,N) (or (M/16,N) with rowmax=16).
    Rows may be strided (last dim contiguous, stride % 8 == 0)."""
    assert a.is_cuda and a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M // rowmax if rowmax else M, N), dtype=out_dtype, device=a.device)
    if bias is not None:
        bias = bias.float().contiguous()
    if residual is not None:
        residual = residual.to(out_dtype).contiguous()
    lib = _lib.gps()
    with torch.cuda.device(a.device):
        st = lib.sv_gemm_bf16(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), M, N, K,
                              bias.data_ptr() if bias is not None else None, _ACT[act],
                              residual.data_ptr() if residual is not None else None, out.data_ptr(), N,
                              1 if out_dtype == torch.float32 else 0, rowmax,
                              torch.cuda.current_stream(a.device).cuda_stream)
    _lib.check(lib, st, "sv_gemm_bf16")
    return out


def attention(q, k, v, num_heads, key_padding_mask=None, spatial_w=None, spatial_heads=0, pairwise_locs=None,
              return_lse=False, dropout_p=0.0, seed=0):
    """Fused attention forward: q (B,Lq,E), k/v (B,Lk,E) bf16 (views with a contiguous last dim are fine), head dim 64.
    key_padding_mask (B,Lk) bool, True = ignore.  spatial_w (B,Lq,spatial_heads*6) + pairwise_locs (B,Lq,Lk,5) switch on the
    MultiHeadAttentionSpatial 'cond' gate.  Returns (B,Lq,E) bf16."""
    B, Lq, E = q.shape
    Lk = k.shape[1]
    assert E == num_heads * 64 and q.dtype == k.dtype == v.dtype == torch.bfloat16
    assert q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1
    out = torch.empty((B, Lq, E), dtype=torch.bfloat16, device=q.device)
    kpm = key_padding_mask.to(torch.uint8).contiguous() if key_padding_mask is not None else None
    sw = locs = None
    if spatial_w is not None:
        sw = spatial_w.float().contiguous()
        locs = pairwise_locs.float().contiguous()
    lse = torch.empty((B, num_heads, Lq), dtype=torch.float32, device=q.device) if return_lse else None
    lib = _lib.gps()
    with torch.cuda.device(q.device):
        st = lib.sv_attention_fwd_dropout_bf16(
            q.data_ptr(), q.stride(0), q.stride(1), k.data_ptr(), k.stride(0), k.stride(1), v.data_ptr(), v.stride(0),
            v.stride(1), out.data_ptr(), out.stride(0), out.stride(1), kpm.data_ptr() if kpm is not None else None,
            sw.data_ptr() if sw is not None else None, int(spatial_heads), locs.data_ptr() if locs is not None else None,
            B, num_heads, Lq, Lk, 0.125, lse.data_ptr() if lse is not None else None, float(dropout_p), int(seed),
            torch.cuda.current_stream(q.device).cuda_stream)
    _lib.check(lib, st, "sv_attention_fwd_dropout_bf16")
    return (out, lse) if return_lse else out


def attention_backward(q, k, v, out, grad_out, lse, num_heads, key_padding_mask=None, spatial_w=None, pairwise_locs=None,
                       dropout_p=0.0, seed=0):
    """Gradients of `attention` (gate requires one weight set per head).  Returns (dq, dk, dv, d_spatial_w or None)."""
    B, Lq, E = q.shape
    Lk = k.shape[1]
    dev = q.device
    grad_out = grad_out.to(torch.bfloat16).contiguous()
    out = out.contiguous()
    dq = torch.empty((B, Lq, E), dtype=torch.bfloat16, device=dev)
    dk = torch.empty((B, Lk, E), dtype=torch.bfloat16, device=dev)
    dv = torch.empty((B, Lk, E), dtype=torch.bfloat16, device=dev)
    dvec = torch.empty((B, num_heads, Lq), dtype=torch.float32, device=dev)
    kpm = key_padding_mask.to(torch.uint8).contiguous() if key_padding_mask is not None else None
    sw = locs = dsw = None
    if spatial_w is not None:
        sw, locs = spatial_w.float().contiguous(), pairwise_locs.float().contiguous()
        dsw = torch.empty_like(sw)
    lib = _lib.gps()
    with torch.cuda.device(dev):
        st = lib.sv_attention_bwd_dropout_bf16(
            q.data_ptr(), q.stride(0), q.stride(1), k.data_ptr(), k.stride(0), k.stride(1), v.data_ptr(), v.stride(0),
            v.stride(1), out.data_ptr(), grad_out.data_ptr(), kpm.data_ptr() if kpm is not None else None,
            sw.data_ptr() if sw is not None else None, locs.data_ptr() if locs is not None else None, lse.data_ptr(),
            B, num_heads, Lq, Lk, 0.125, dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
            dsw.data_ptr() if dsw is not None else None, dvec.data_ptr(), float(dropout_p), int(seed),
            torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(lib, st, "sv_attention_bwd_dropout_bf16")
    return dq, dk, dv, dsw


def colsum(x):
    """Column sums of a (R, N) bf16 / f32 matrix (rows may be strided) -> (N,) f32; deterministic native kernel."""
    assert x.is_cuda and x.dim() == 2 and x.stride(1) == 1 and x.dtype in (torch.bfloat16, torch.float32)
    R, N = x.shape
    lib = _lib.gps()
    out = torch.empty(N, dtype=torch.float32, device=x.device)
    scratch = torch.empty(lib.sv_colsum_scratch_floats(N), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        st = lib.sv_colsum(x.data_ptr(), x.stride(0), 1 if x.dtype == torch.bfloat16 else 0, R, N, out.data_ptr(),
                           scratch.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(lib, st, "sv_colsum")
    return out


def gemm_ex(a, b, a_transposed=False, b_transposed=False, bias=None, act=None, out_dtype=torch.bfloat16):
    """epilogue(A . B^T) with operands optionally given transposed in memory (no copies):
    a: (M,K), or (K,M) when a_transposed;  b: (N,K), or (K,N) when b_transposed.  bf16 in, bf16 / f32 out (M,N)."""
    assert a.is_cuda and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if a_transposed else a.shape
    N, Kb = (b.shape[1], b.shape[0]) if b_transposed else b.shape
    assert K == Kb, (a.shape, b.shape)
    out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    if bias is not None:
        bias = bias.float().contiguous()
    lib = _lib.gps()
    with torch.cuda.device(a.device):
        st = lib.sv_gemm_bf16_ex(a.data_ptr(), a.stride(0), 1 if a_transposed else 0, b.data_ptr(), b.stride(0),
                                 1 if b_transposed else 0, M, N, K, bias.data_ptr() if bias is not None else None, _ACT[act],
                                 None, out.data_ptr(), N, 1 if out_dtype == torch.float32 else 0, 0,
                                 torch.cuda.current_stream(a.device).cuda_stream)
    _lib.check(lib, st, "sv_gemm_bf16_ex")
    return out

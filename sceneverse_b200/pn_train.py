"""Train-mode PointNet++ set abstraction on the native kernels (config C2, `ObjCls` pre-training with a trainable backbone;
reference: pointnet2_modules.py:34-75, pointnet2_utils.py:291-419, pytorch_utils.py:11-36,67-120).

Channels-LAST formulation: the grouped tensor is a bf16 row matrix X[(b, centre, sample)][channel]; each SharedMLP layer
(Conv2d 1x1 without bias -> BatchNorm2d with BATCH statistics -> ReLU) is one native GEMM (csrc/gemm.cu, forward / dgrad / wgrad)
plus the column-statistic kernels of csrc/pn_train.cu; the neighbourhood max is a max over `nsample` consecutive rows.  The
parameters are the reference's own (`conv.weight (Cout,Cin,1,1)`, `bn.bn.{weight,bias,running_mean,running_var}`): checkpoints and
state_dicts are untouched, running statistics are updated exactly like nn.BatchNorm2d (momentum, unbiased variance)."""
import torch
import torch.nn.functional as F

from . import _lib, native


def _st(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _round8(n):
    return (n + 7) // 8 * 8


class _GroupRowsFn(torch.autograd.Function):
    """(xyz (B,N,3) f32, centre (B,np,3) f32 | None, feat (B,N,C) f32 | bf16 | None, idx (B,np,ns) i32 | None) -> X (B*np*ns, Cp) bf16."""

    @staticmethod
    def forward(ctx, xyz, centre, feat, idx, np_, ns):
        B, N, _ = xyz.shape
        C = 0 if feat is None else feat.shape[-1]
        Cp = _round8(3 + C)
        X = torch.empty((B * np_ * ns, Cp), dtype=torch.bfloat16, device=xyz.device)
        lib = _lib.gps()
        xyz, feat = xyz.contiguous(), (feat.contiguous() if feat is not None else None)
        with torch.cuda.device(xyz.device):
            st = lib.sv_pn_group_rows(xyz.data_ptr(), centre.contiguous().data_ptr() if centre is not None else None,
                                      feat.data_ptr() if feat is not None else None, 1 if (feat is not None and feat.dtype == torch.bfloat16) else 0,
                                      idx.data_ptr() if idx is not None else None, B, N, C, np_, ns, Cp, X.data_ptr(), _st(xyz))
        _lib.check(lib, st, "sv_pn_group_rows")
        ctx.save_for_backward(idx)
        ctx.meta = (B, N, C, np_, ns, Cp, None if feat is None else feat.dtype)
        return X

    @staticmethod
    def backward(ctx, dX):
        (idx,) = ctx.saved_tensors
        B, N, C, np_, ns, Cp, fdt = ctx.meta
        if C == 0 or not ctx.needs_input_grad[2]:
            return None, None, None, None, None, None
        dX = dX.contiguous()
        dfeat = torch.empty((B, N, C), dtype=torch.float32, device=dX.device)
        lib = _lib.gps()
        with torch.cuda.device(dX.device):
            st = lib.sv_pn_group_rows_grad(dX.data_ptr(), idx.data_ptr() if idx is not None else None, B, N, C, np_, ns, Cp,
                                           dfeat.data_ptr(), _st(dX))
        _lib.check(lib, st, "sv_pn_group_rows_grad")
        return None, None, dfeat.to(fdt), None, None, None


class _ConvBNReLUFn(torch.autograd.Function):
    """out = relu(BatchNorm_batchstats(X W^T)): X (R,Kp) bf16 (zero-padded columns), W = conv.weight (Cout,Cin,1,1) f32."""

    @staticmethod
    def forward(ctx, X, weight, gamma, beta, eps):
        Cout, Cin = weight.shape[0], weight.shape[1]
        R, Kp = X.shape
        wb = weight.detach().reshape(Cout, Cin).to(torch.bfloat16)
        if Kp != Cin:
            wb = F.pad(wb, (0, Kp - Cin))
        y = native.linear_fwd(X, wb.contiguous())
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        stats = torch.empty((3, Cout), dtype=torch.float32, device=X.device)     # mean, rstd, unbiased variance
        out = torch.empty_like(y)
        lib = _lib.gps()
        scratch = torch.empty(lib.sv_pn_scratch_floats(Cout), dtype=torch.float32, device=X.device)
        with torch.cuda.device(X.device):
            st = lib.sv_pn_bn_relu_fwd(y.data_ptr(), R, Cout, g32.data_ptr(), b32.data_ptr(), float(eps), stats[0].data_ptr(),
                                       stats[1].data_ptr(), stats[2].data_ptr(), out.data_ptr(), scratch.data_ptr(), _st(X))
        _lib.check(lib, st, "sv_pn_bn_relu_fwd")
        ctx.save_for_backward(X, wb, y, stats, g32, b32)
        ctx.cin = Cin
        ctx.mark_non_differentiable(stats)
        return out, stats

    @staticmethod
    def backward(ctx, dout, _dstats):
        X, wb, y, stats, g32, b32 = ctx.saved_tensors
        R, Cout = y.shape
        dout = dout.contiguous()
        dy = torch.empty_like(y)
        dgb = torch.empty((2, Cout), dtype=torch.float32, device=y.device)
        lib = _lib.gps()
        scratch = torch.empty(lib.sv_pn_scratch_floats(Cout), dtype=torch.float32, device=y.device)
        with torch.cuda.device(y.device):
            st = lib.sv_pn_bn_relu_bwd(y.data_ptr(), dout.data_ptr(), R, Cout, g32.data_ptr(), b32.data_ptr(), stats[0].data_ptr(),
                                       stats[1].data_ptr(), dy.data_ptr(), dgb[0].data_ptr(), dgb[1].data_ptr(), 0,
                                       scratch.data_ptr(), _st(y))
        _lib.check(lib, st, "sv_pn_bn_relu_bwd")
        dX = native.linear_dgrad(dy, wb) if ctx.needs_input_grad[0] else None
        dw, _ = native.linear_wgrad(dy, X)
        dw = dw[:, :ctx.cin].reshape(Cout, ctx.cin, 1, 1)
        return dX, dw, dgb[0], dgb[1], None


class _RowGroupMaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ns):
        R, C = x.shape
        G = R // ns
        out = torch.empty((G, C), dtype=torch.bfloat16, device=x.device)
        arg = torch.empty((G, C), dtype=torch.uint8, device=x.device)
        lib = _lib.gps()
        with torch.cuda.device(x.device):
            st = lib.sv_pn_rowgroup_max(x.data_ptr(), G, ns, C, out.data_ptr(), arg.data_ptr(), _st(x))
        _lib.check(lib, st, "sv_pn_rowgroup_max")
        ctx.save_for_backward(arg)
        ctx.ns = ns
        return out

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        G, C = arg.shape
        g = g.contiguous().to(torch.bfloat16)
        gx = torch.empty((G * ctx.ns, C), dtype=torch.bfloat16, device=g.device)
        lib = _lib.gps()
        with torch.cuda.device(g.device):
            st = lib.sv_pn_rowgroup_max_grad(g.data_ptr(), arg.data_ptr(), G, ctx.ns, C, gx.data_ptr(), _st(g))
        _lib.check(lib, st, "sv_pn_rowgroup_max_grad")
        return gx, None


def _update_running_stats(bn, stats):
    """nn.BatchNorm2d's training-mode bookkeeping (momentum update with the UNBIASED batch variance)."""
    if bn.track_running_stats and bn.running_mean is not None:
        with torch.no_grad():
            bn.num_batches_tracked += 1
            m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
            bn.running_mean.mul_(1 - m).add_(stats[0].to(bn.running_mean.dtype), alpha=m)
            bn.running_var.mul_(1 - m).add_(stats[2].to(bn.running_var.dtype), alpha=m)


def available(pc, net):
    """Train-mode BatchNorm (or a backbone that needs gradients) on CUDA under bf16 autocast."""
    if not (pc.is_cuda and pc.dtype == torch.float32 and torch.is_autocast_enabled("cuda")
            and torch.get_autocast_dtype("cuda") == torch.bfloat16):
        return False
    bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    return all(type(m) is torch.nn.BatchNorm2d and m.training and m.affine for m in bns)


def sa_forward(sa, xyz, feat):
    """One PointnetSAModule in training mode: xyz (B,N,3) f32, feat (B,N,C) point-major (f32 | bf16) -> (new_xyz, (B,np,Cout) bf16)."""
    from . import pointnet2_utils
    from .pointnet2 import _ext
    B, N, _ = xyz.shape
    if sa.npoint is not None:
        idx = _ext.furthest_point_sampling(xyz, sa.npoint)
        new_xyz = torch.gather(xyz, 1, idx.long()[..., None].expand(-1, -1, 3)).contiguous()
        bidx = _ext.ball_query(new_xyz, xyz, sa.radius, sa.nsample)
        np_, ns = sa.npoint, sa.nsample
        X = _GroupRowsFn.apply(xyz, new_xyz, feat, bidx, np_, ns)
    else:                                       # GroupAll: one group of all N points, xyz NOT centred (pointnet2_utils.py:389-419)
        new_xyz, np_, ns = None, 1, N
        X = _GroupRowsFn.apply(xyz, None, feat, None, 1, N)
    mlp = sa.mlps[0]
    for j in range(len(sa.mlp_spec) - 1):
        layer = getattr(mlp, f"layer{j}")
        bn = layer.bn.bn
        X, stats = _ConvBNReLUFn.apply(X, layer.conv.weight, bn.weight, bn.bias, bn.eps)
        _update_running_stats(bn, stats)
    out = _RowGroupMaxFn.apply(X, ns)           # (B*np, Cout)
    return new_xyz, out.view(B, np_, -1)


def forward(net, pc):
    """PointNetPP.forward for a trainable backbone: (n, P, 3 + C) f32 -> (n, D)."""
    from . import ops
    xyz = pc[..., 0:3].contiguous()
    feat = pc[..., 3:].contiguous() if pc.size(-1) > 3 else None
    for sa in net.encoder:
        new_xyz, feat = sa_forward(sa, xyz, feat)
        xyz = new_xyz if new_xyz is not None else xyz
    return ops.linear(feat.reshape(feat.shape[0], -1), net.fc.weight, net.fc.bias)

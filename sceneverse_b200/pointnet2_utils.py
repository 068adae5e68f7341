"""Host-side mirror of the reference's operator layer
(/root/reference/modules/third_party/pointnet2/pointnet2_utils.py:48-419): the same public names,
argument order and autograd behaviour, calling the CUDA library through
sceneverse_b200.pointnet2._ext.  Index outputs are non-differentiable; feature gathers have the
scatter-add backward of the reference (pointnet2_utils.py:107-111,246-251).
"""
import torch
from torch import nn
from torch.autograd import Function

from .pointnet2 import _ext


class _FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        inds = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, grad=None):
        return None, None


furthest_point_sample = _FurthestPointSampling.apply


class _Gather(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.idx, ctx.n = idx, features.size(2)
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        return _ext.gather_points_grad(grad_out.contiguous(), ctx.idx, ctx.n), None


def gather_operation(features, idx):
    """pointnet2_utils.py:80-114 (fp32 kernel; other dtypes are widened)."""
    if features.dtype != torch.float32:
        features = features.float()
    return _Gather.apply(features.contiguous(), idx)


class _ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, known):
        dist2, idx = _ext.three_nn(unknown, known)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx  # reference returns sqrt(dist2), pointnet2_utils.py:138-140

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = _ThreeNN.apply


class _ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.saved = (idx, weight, features.size(2))
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.saved
        return _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, m), None, None


three_interpolate = _ThreeInterpolate.apply


class _Grouping(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.idx, ctx.n = idx, features.size(2)
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        return _ext.group_points_grad(grad_out.contiguous(), ctx.idx, ctx.n), None


def grouping_operation(features, idx):
    """pointnet2_utils.py:206-254.  The native kernel is fp32-only like the reference's (utils.h:21-25); reduced-precision
    features (bf16 autocast with a trainable backbone, where the reference itself would assert) are widened first."""
    if features.dtype != torch.float32:
        features = features.float()
    return _Grouping.apply(features.contiguous(), idx)


class _BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        inds = _ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = _BallQuery.apply


class QueryAndGroup(nn.Module):
    """ball query -> group xyz (minus centre) and features -> concat (pointnet2_utils.py:291-373).
    `sample_uniformly`/`ret_unique_cnt` (a host-side torch.unique loop in the reference, unused by
    the GPS path) are not provided."""

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz, self.normalize_xyz = ret_grouped_xyz, normalize_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz = grouped_xyz / self.radius
        if features is not None:
            grouped = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz
        return (new_features, grouped_xyz) if self.ret_grouped_xyz else new_features


class GroupAll(nn.Module):
    """pointnet2_utils.py:376-419."""

    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped = features.unsqueeze(2)
        return torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped

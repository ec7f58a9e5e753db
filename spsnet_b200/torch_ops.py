"""`torch.ops.spsk.*` -- the thin torch custom-op layer over the C-ABI (north_star "thin C-ABI torch
custom-op layer").

Each op checks device / dtype / contiguity, allocates its outputs, passes torch's current stream and calls ONE libspsk entry
point (include/spsk.h); a fake (meta) implementation gives the dispatcher the output shapes, so the ops are visible to
`torch.compile`, `torch.export`, FakeTensor tracing and `torch.library.opcheck`; the three differentiable ops carry
`register_autograd` formulas built from the `*_grad` ops (reference pointnet2_utils.py:90-98,166-178,207-222).  The
`autograd.Function` aliases of `pointnet2_utils` (the reference's own public names) and these ops share the same launch code.

    import spsnet_b200.torch_ops                      # registers the library
    idx = torch.ops.spsk.furthest_point_sample(xyz, 4096)
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import pointnet2_utils as pu

_lib_def = torch.library.custom_op


class _Ctx:
    """Stand-in for the autograd ctx of the pointnet2_utils Functions (their forward bodies ARE the launch code)."""


# ---- sampling ----------------------------------------------------------------------------------------------------------
@_lib_def("spsk::furthest_point_sample", mutates_args=())
def furthest_point_sample(xyz: Tensor, npoint: int) -> Tensor:
    return pu.FarthestPointSampling.forward(_Ctx(), xyz, npoint)


@furthest_point_sample.register_fake
def _(xyz, npoint):
    return xyz.new_empty((xyz.shape[0], npoint), dtype=torch.int32)


@_lib_def("spsk::furthest_point_sample_with_dist", mutates_args=())
def furthest_point_sample_with_dist(dist: Tensor, npoint: int) -> Tensor:
    return pu.FurthestPointSamplingWithDist.forward(_Ctx(), dist, npoint)


@furthest_point_sample_with_dist.register_fake
def _(dist, npoint):
    return dist.new_empty((dist.shape[0], npoint), dtype=torch.int32)


@_lib_def("spsk::score_topk", mutates_args=())
def score_topk(cls_features: Tensor, npoint: int, stds: Optional[Tensor] = None) -> Tensor:
    return pu.score_topk(cls_features, npoint, stds=stds)


@score_topk.register_fake
def _(cls_features, npoint, stds=None):
    return cls_features.new_empty((cls_features.shape[0], npoint), dtype=torch.int32)


# ---- gather ------------------------------------------------------------------------------------------------------------
@_lib_def("spsk::gather_points", mutates_args=())
def gather_points(features: Tensor, idx: Tensor) -> Tensor:
    return pu.GatherOperation.forward(_Ctx(), features, idx)


@gather_points.register_fake
def _(features, idx):
    return features.new_empty((features.shape[0], features.shape[1], idx.shape[1]))


@_lib_def("spsk::gather_points_grad", mutates_args=())
def gather_points_grad(grad_out: Tensor, idx: Tensor, n: int) -> Tensor:
    ctx = _Ctx()
    ctx.for_backwards = (idx, grad_out.shape[1], n)
    return pu.GatherOperation.backward(ctx, grad_out)[0]


@gather_points_grad.register_fake
def _(grad_out, idx, n):
    return grad_out.new_empty((grad_out.shape[0], grad_out.shape[1], n))


def _gather_setup(ctx, inputs, output):
    features, idx = inputs
    ctx.save_for_backward(idx)
    ctx.n = features.shape[2]


def _gather_bwd(ctx, grad):
    (idx,) = ctx.saved_tensors
    return torch.ops.spsk.gather_points_grad(grad.contiguous(), idx, ctx.n), None


gather_points.register_autograd(_gather_bwd, setup_context=_gather_setup)


@_lib_def("spsk::gather_rows", mutates_args=())
def gather_rows(points: Tensor, idx: Tensor) -> Tensor:
    return pu.gather_rows(points, idx)


@gather_rows.register_fake
def _(points, idx):
    return points.new_empty((points.shape[0], idx.shape[1], points.shape[2]))


# ---- ball query / grouping ---------------------------------------------------------------------------------------------
@_lib_def("spsk::ball_query", mutates_args=())
def ball_query(radius: float, nsample: int, xyz: Tensor, new_xyz: Tensor) -> Tensor:
    return pu.BallQuery.forward(_Ctx(), radius, nsample, xyz, new_xyz)


@ball_query.register_fake
def _(radius, nsample, xyz, new_xyz):
    return xyz.new_empty((xyz.shape[0], new_xyz.shape[1], nsample), dtype=torch.int32)


@_lib_def("spsk::ball_query_dilated", mutates_args=())
def ball_query_dilated(max_radius: float, min_radius: float, nsample: int, xyz: Tensor, new_xyz: Tensor) -> Tensor:
    return pu.BallQueryDilated.forward(_Ctx(), max_radius, min_radius, nsample, xyz, new_xyz)


@ball_query_dilated.register_fake
def _(max_radius, min_radius, nsample, xyz, new_xyz):
    return xyz.new_empty((xyz.shape[0], new_xyz.shape[1], nsample), dtype=torch.int32)


@_lib_def("spsk::group_points", mutates_args=())
def group_points(features: Tensor, idx: Tensor) -> Tensor:
    return pu.GroupingOperation.forward(_Ctx(), features, idx)


@group_points.register_fake
def _(features, idx):
    return features.new_empty((features.shape[0], features.shape[1], idx.shape[1], idx.shape[2]))


@_lib_def("spsk::group_points_grad", mutates_args=())
def group_points_grad(grad_out: Tensor, idx: Tensor, n: int) -> Tensor:
    ctx = _Ctx()
    ctx.for_backwards = (idx, n)
    return pu.GroupingOperation.backward(ctx, grad_out)[0]


@group_points_grad.register_fake
def _(grad_out, idx, n):
    return grad_out.new_empty((grad_out.shape[0], grad_out.shape[1], n))


def _group_setup(ctx, inputs, output):
    features, idx = inputs
    ctx.save_for_backward(idx)
    ctx.n = features.shape[2]


def _group_bwd(ctx, grad):
    (idx,) = ctx.saved_tensors
    return torch.ops.spsk.group_points_grad(grad.contiguous(), idx, ctx.n), None


group_points.register_autograd(_group_bwd, setup_context=_group_setup)


# ---- interpolation -----------------------------------------------------------------------------------------------------
@_lib_def("spsk::three_nn", mutates_args=())
def three_nn(unknown: Tensor, known: Tensor) -> Tuple[Tensor, Tensor]:
    class C(_Ctx):
        def mark_non_differentiable(self, *a):
            pass

    return pu.ThreeNN.forward(C(), unknown, known)


@three_nn.register_fake
def _(unknown, known):
    shape = (unknown.shape[0], unknown.shape[1], 3)
    return unknown.new_empty(shape), unknown.new_empty(shape, dtype=torch.int32)


@_lib_def("spsk::three_interpolate", mutates_args=())
def three_interpolate(features: Tensor, idx: Tensor, weight: Tensor) -> Tensor:
    return pu.ThreeInterpolate.forward(_Ctx(), features, idx, weight)


@three_interpolate.register_fake
def _(features, idx, weight):
    return features.new_empty((features.shape[0], features.shape[1], idx.shape[1]))


@_lib_def("spsk::three_interpolate_grad", mutates_args=())
def three_interpolate_grad(grad_out: Tensor, idx: Tensor, weight: Tensor, m: int) -> Tensor:
    ctx = _Ctx()
    ctx.three_interpolate_for_backward = (idx, weight, m)
    return pu.ThreeInterpolate.backward(ctx, grad_out)[0]


@three_interpolate_grad.register_fake
def _(grad_out, idx, weight, m):
    return grad_out.new_empty((grad_out.shape[0], grad_out.shape[1], m))


def _interp_setup(ctx, inputs, output):
    features, idx, weight = inputs
    ctx.save_for_backward(idx, weight)
    ctx.m = features.shape[2]


def _interp_bwd(ctx, grad):
    idx, weight = ctx.saved_tensors
    return torch.ops.spsk.three_interpolate_grad(grad.contiguous(), idx, weight, ctx.m), None, None


three_interpolate.register_autograd(_interp_bwd, setup_context=_interp_setup)

OPS = ("furthest_point_sample", "furthest_point_sample_with_dist", "score_topk", "gather_points", "gather_points_grad", "gather_rows",
       "ball_query", "ball_query_dilated", "group_points", "group_points_grad", "three_nn", "three_interpolate", "three_interpolate_grad")

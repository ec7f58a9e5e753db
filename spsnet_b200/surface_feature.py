"""Drop-in for the reference's `pcdet/ops/pointnet2/pointnet2_batch/surface_feature.py`:
SPSNet's surface-feature extractor, switched on by `USE_SURFACE: True` (tools/cfgs/kitti_models/SPSNet.yaml:48) and
used by `PAGNet_Backbone` (PAGNet_backbone.py:29-31,151-162).

Same classes, constructor arguments and `state_dict` layout (FCLayer, Aggregator, DenseEdgeConv, FeatureExtraction:
reference surface_feature.py:7-187).  At inference on the GPU one unit (transform FC + DenseEdgeConv) is
    spsk_edge_conv_point -> spsk_ball_query -> spsk_edge_conv_aggregate          (csrc/edge_conv.cu)
instead of ~15 torch kernels that materialise (B, N, K, 72) tensors; with autograd enabled the same ops run as plain
torch modules on top of the drop-in `pointnet2_utils` (so gradients flow exactly as in the reference).

Reference quirk kept on purpose (SURVEY.md §8f rank 4 "latent stride bug"): in dynamic-graph mode DenseEdgeConv passes the
24-channel FEATURES as `xyz` to QueryAndGroup (surface_feature.py:79, :170-173).  The reference ball-query kernel assumes
3 floats per point (ball_query_gpu.cu:18-20), so scene b's "coordinates" are the floats [3 N b, 3 N (b+1)) of the flat
(B, N, 24) buffer.  `_as_ball_query_coords` reproduces exactly that reinterpretation so neighbour lists are bit-identical.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import pointnet2_utils as pu
from ._lib import EdgeAggrWeights, EdgePointWeights, check, lib

__all__ = ["FCLayer", "Aggregator", "DenseEdgeConv", "FeatureExtraction"]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _as_ball_query_coords(pos: torch.Tensor) -> torch.Tensor:
    """What the reference ball-query kernel reads when handed a contiguous (B, N, d) tensor as `xyz`: the first B*N*3
    floats of the buffer, viewed as (B, N, 3).  Identity for d = 3."""
    B, N, _ = pos.shape
    return pos.contiguous().view(-1)[: B * N * 3].view(B, N, 3)


_ACTIVATIONS = {None: nn.Identity, "relu": nn.ReLU, "elu": lambda: nn.ELU(alpha=1.0), "lrelu": lambda: nn.LeakyReLU(0.1)}


class FCLayer(nn.Module):
    """Linear + optional activation; parameters `linear.weight` / `linear.bias` (reference surface_feature.py:7-26)."""

    def __init__(self, in_features, out_features, bias=True, activation=None):
        super().__init__()
        if activation not in _ACTIVATIONS:
            raise ValueError(f"unknown activation {activation!r}")
        self.linear = nn.Linear(in_features, out_features, bias=bias)
        self.activation = _ACTIVATIONS[activation]()

    def forward(self, x):
        return self.activation(self.linear(x))


class Aggregator(nn.Module):
    """Reduction over the neighbour axis: 'mean' | 'sum' | 'max' (reference surface_feature.py:28-43)."""

    def __init__(self, oper):
        super().__init__()
        if oper not in ("mean", "sum", "max"):
            raise AssertionError(oper)
        self.oper = oper

    def forward(self, x, dim=2):
        if self.oper == "max":
            return x.max(dim=dim)[0]   # (not amax: ties must send the gradient to one element, like the reference)
        return {"mean": torch.mean, "sum": torch.sum}[self.oper](x, dim=dim)


class DenseEdgeConv(nn.Module):
    """reference surface_feature.py:45-115."""

    def __init__(self, in_channels, num_fc_layers, growth_rate, radius=0.8, knn=32, aggr="max", activation="relu",
                 relative_feat_only=False):
        super().__init__()
        self.in_channels = in_channels
        self.knn = knn
        self.radius = radius
        assert num_fc_layers > 2
        self.num_fc_layers = num_fc_layers
        self.growth_rate = growth_rate
        self.relative_feat_only = relative_feat_only
        self._activation = activation
        self.group = pu.QueryAndGroup(radius, knn, use_xyz=False)
        if relative_feat_only:
            self.layer_first = FCLayer(in_channels, growth_rate, bias=True, activation=activation)
        else:
            self.layer_first = FCLayer(3 * in_channels, growth_rate, bias=True, activation=activation)
        self.layer_last = FCLayer(in_channels + (num_fc_layers - 1) * growth_rate, growth_rate, bias=True, activation=None)
        self.layers = nn.ModuleList()
        for i in range(1, num_fc_layers - 1):
            self.layers.append(FCLayer(in_channels + i * growth_rate, growth_rate, bias=True, activation=activation))
        self.aggr = Aggregator(aggr)

    @property
    def out_channels(self):
        return self.in_channels + self.num_fc_layers * self.growth_rate

    def neighbours(self, pos: torch.Tensor) -> torch.Tensor:
        """(B, N, knn) int32 neighbour lists the reference's QueryAndGroup(xyz=pos, new_xyz=pos) produces."""
        coords = _as_ball_query_coords(pos)
        return pu.ball_query_msg([self.radius], [self.knn], coords, coords)[0]  # grid kernel for large N, same lists

    def fusable(self) -> bool:
        return (self.num_fc_layers == 3 and self.in_channels == 24 and self.growth_rate == 12 and self._activation == "relu"
                and self.aggr.oper == "max")

    def get_edge_feature(self, x, pos, idx=None):
        """(B, N, d) -> (B, N, K, 3 d) (or d when relative_feat_only); reference :73-87."""
        idx = self.neighbours(pos) if idx is None else idx
        knn_feat = pu.grouping_operation(x.permute(0, 2, 1).contiguous(), idx).permute(0, 2, 3, 1).contiguous()
        x_tiled = x.unsqueeze(-2).expand_as(knn_feat)
        if self.relative_feat_only:
            return knn_feat - x_tiled
        return torch.cat([x_tiled, knn_feat, knn_feat - x_tiled], dim=3)

    def forward(self, x, pos, idx=None):
        """(B, N, d), (B, N, d') -> (B, N, d + L c); reference :89-115 (plain torch modules: the autograd path)."""
        edge_feat = self.get_edge_feature(x, pos, idx)
        y = torch.cat([self.layer_first(edge_feat), x.unsqueeze(-2).repeat(1, 1, self.knn, 1)], dim=-1)
        for layer in self.layers:
            y = torch.cat([layer(y), y], dim=-1)
        y = torch.cat([self.layer_last(y), y], dim=-1)
        return self.aggr(y, dim=-2)


def _pack_unit(trans: FCLayer, conv: DenseEdgeConv):
    """(EdgePointWeights, EdgeAggrWeights) of one unit; see csrc/edge_conv.cu for the algebra."""
    d, g = conv.in_channels, conv.growth_rate
    Wt = trans.linear.weight.detach().double().cpu()
    bt = trans.linear.bias.detach().double().cpu() if trans.linear.bias is not None else torch.zeros(d, dtype=torch.float64)
    W1, b1 = conv.layer_first.linear.weight.detach().double().cpu(), conv.layer_first.linear.bias.detach().double().cpu()
    W2, b2 = conv.layers[0].linear.weight.detach().double().cpu(), conv.layers[0].linear.bias.detach().double().cpu()
    W3, b3 = conv.layer_last.linear.weight.detach().double().cpu(), conv.layer_last.linear.bias.detach().double().cpu()
    if conv.relative_feat_only:
        WP, WQ = -W1, W1
    else:
        W1a, W1b, W1c = W1[:, :d], W1[:, d:2 * d], W1[:, 2 * d:]
        WP, WQ = W1a - W1c, W1b + W1c
    W2a, W2b = W2[:, :g], W2[:, g:]
    W3a, W3b, W3c = W3[:, :g], W3[:, g:2 * g], W3[:, 2 * g:]
    M = torch.cat([WP, WQ, W2b, W3c], dim=0)                       # (48, 24)
    c = torch.cat([b1, torch.zeros(g, dtype=torch.float64), b2, b3])
    pw = EdgePointWeights()
    pw.cin = Wt.shape[1]
    pw.relu = 1 if isinstance(trans.activation, nn.ReLU) else 0
    if not isinstance(trans.activation, (nn.ReLU, nn.Identity)):
        raise NotImplementedError("fused surface features support ReLU / identity transforms")
    flat = Wt.t().contiguous().view(-1).float().tolist()          # [k*24 + o]
    pw.wt[: len(flat)] = flat
    pw.bt[:] = bt.float().tolist()
    pw.m[:] = M.t().contiguous().view(-1).float().tolist()        # [k*48 + o]
    pw.c[:] = c.float().tolist()
    aw = EdgeAggrWeights()
    aw.w2a[:] = W2a.t().contiguous().view(-1).float().tolist()
    aw.w3a[:] = W3a.t().contiguous().view(-1).float().tolist()
    aw.w3b[:] = W3b.t().contiguous().view(-1).float().tolist()
    return pw, aw


class FeatureExtraction(nn.Module):
    """reference surface_feature.py:118-187."""

    def __init__(self, in_channels=3, dynamic_graph=True, conv_channels=24, num_convs=4, conv_num_fc_layers=3,
                 conv_growth_rate=12, conv_knn=16, conv_aggr="max", activation="relu"):
        super().__init__()
        self.in_channels = in_channels
        self.dynamic_graph = dynamic_graph
        self.num_convs = num_convs
        self.transforms = nn.ModuleList()
        self.convs = nn.ModuleList()
        for i in range(num_convs):
            trans = FCLayer(in_channels, conv_channels, bias=True, activation=None if i == 0 else activation)
            conv = DenseEdgeConv(conv_channels, num_fc_layers=conv_num_fc_layers, growth_rate=conv_growth_rate, knn=conv_knn,
                                 aggr=conv_aggr, activation=activation, relative_feat_only=(i == 0))
            self.transforms.append(trans)
            self.convs.append(conv)
            in_channels = conv.out_channels

    @property
    def out_channels(self):
        return self.convs[-1].out_channels

    # ---- fused inference path -----------------------------------------------------------------------------------
    def invalidate_cache(self) -> None:
        self.__dict__.pop("_packed", None)

    def load_state_dict(self, *a, **k):
        self.invalidate_cache()
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.invalidate_cache()
        return super()._apply(fn, *a, **k)

    def _fused_ok(self, x: torch.Tensor) -> bool:
        return (x.is_cuda and x.dtype == torch.float32 and not self.training and not torch.is_grad_enabled()
                and all(c.fusable() for c in self.convs)
                and all(t.linear.in_features <= 64 and isinstance(t.activation, (nn.ReLU, nn.Identity)) for t in self.transforms))

    def fused_forward(self, x: torch.Tensor, forced_idx=None, return_idx: bool = False):
        """(B, N, C) -> (B, N, 60).  forced_idx: optional list of per-unit (B, N, K) int32 neighbour lists (tests);
        return_idx: also return the per-unit neighbour lists and transformed features t."""
        packed = self.__dict__.get("_packed")
        if packed is None:
            packed = [_pack_unit(t, c) for t, c in zip(self.transforms, self.convs)]
            self.__dict__["_packed"] = packed
        B, N, _ = x.shape
        pos0 = x
        cur = x.contiguous()
        idx_list, t_list = [], []
        with torch.cuda.device(x.device):
            for i, (pw, aw) in enumerate(packed):
                conv = self.convs[i]
                rows = B * N
                t = torch.empty((B, N, 24), dtype=torch.float32, device=x.device)
                u = torch.empty((B, N, 48), dtype=torch.float32, device=x.device)
                check(lib.spsk_edge_conv_point(C.byref(pw), rows, cur.data_ptr(), cur.shape[2], t.data_ptr(), u.data_ptr(),
                                               _stream()), "edge_conv_point")
                if forced_idx is not None:
                    idx = forced_idx[i]
                else:
                    idx = conv.neighbours(t if self.dynamic_graph else pos0)
                idx_list.append(idx)
                t_list.append(t)
                out = torch.empty((B, N, 60), dtype=torch.float32, device=x.device)
                check(lib.spsk_edge_conv_aggregate(C.byref(aw), B, N, idx.shape[2], idx.data_ptr(), t.data_ptr(), u.data_ptr(),
                                                   out.data_ptr(), 60, _stream()), "edge_conv_aggregate")
                cur = out
        return (cur, idx_list, t_list) if return_idx else cur

    # ---- reference structure (autograd path) -----------------------------------------------------------------------
    def dynamic_graph_forward(self, x):
        for i in range(self.num_convs):
            x = self.transforms[i](x)
            x = self.convs[i](x, x)
        return x

    def static_graph_forward(self, pos):
        x = pos
        for i in range(self.num_convs):
            x = self.transforms[i](x)
            x = self.convs[i](x, pos)
        return x

    def forward(self, x):
        if self._fused_ok(x):
            return self.fused_forward(x)
        return self.dynamic_graph_forward(x) if self.dynamic_graph else self.static_graph_forward(x)

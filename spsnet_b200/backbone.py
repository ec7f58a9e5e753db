"""The callers of the hot path: IA-SSD's `IASSD_Backbone` and SPSNet-IA's `PAGNet_Backbone`.

Mirrors reference pcdet/models/backbones_3d/IASSD_backbone.py:10-212 and PAGNet_backbone.py:10-237: same
constructor `(model_cfg, num_class, input_channels)`, same `SA_modules` ModuleList (so backbone
checkpoints load), same `batch_dict` keys on output.  `model_cfg` may be the reference's EasyDict or a
plain dict (wrapped in `Cfg`).  Differences: the per-scene point-count loop + host-sync assert
(IASSD_backbone.py:116-120) is replaced by a shape check, and PAGNet's optional DenseEdgeConv surface
feature (`USE_SURFACE`) is not built.
"""
from __future__ import annotations

import copy
from typing import Any

import torch
import torch.nn as nn

from . import pointnet2_modules, pointnet2_utils


from .configs import (Cfg, KITTI_IASSD_SA_CONFIG, kitti_iassd_cfg, kitti_spsnet_cfg, kitti_spsnet_surface_cfg,  # noqa: F401  (re-exported)
                      randomize_bn_stats, waymo_iassd_cfg)


class LazyList(list):
    """A list whose entries are built on first access.  `encoder_coords` and `sa_ins_preds` (reference IASSD_backbone.py:
    113,139-148) are read by the heads' target assignment and losses only -- training code -- yet cost ten small cat / cast
    kernels per forward; at inference they are materialised only if somebody looks.  Entries are either values or zero-argument
    callables; everything the callables capture (encoder_xyz, the class logits) is a regular output of the forward, so a
    later access -- also after a CUDA-graph replay -- sees the current contents."""

    def _get(self, i):
        v = list.__getitem__(self, i)
        if callable(v):
            v = v()
            list.__setitem__(self, i, v)
        return v

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._get(j) for j in range(*i.indices(len(self)))]
        return self._get(i if i >= 0 else len(self) + i)

    def __iter__(self):
        return (self._get(i) for i in range(len(self)))


class IASSD_Backbone(nn.Module):
    """Backbone for IA-SSD (reference IASSD_backbone.py:7-212)."""

    _pass_stds = False

    def __init__(self, model_cfg: Any, num_class: int, input_channels: int, **kwargs):
        super().__init__()
        if isinstance(model_cfg, dict) and not isinstance(model_cfg, Cfg):
            model_cfg = Cfg(model_cfg)
        self.model_cfg = model_cfg
        self.num_class = num_class
        sa = model_cfg.SA_CONFIG
        self.layer_types = sa.LAYER_TYPE
        self.ctr_idx_list = sa.CTR_INDEX
        self.layer_inputs = sa.LAYER_INPUT
        self.aggregation_mlps = sa.get("AGGREGATION_MLPS", None)
        self.confidence_mlps = sa.get("CONFIDENCE_MLPS", None)
        self.max_translate_range = sa.get("MAX_TRANSLATE_RANGE", None)

        self.SA_modules = nn.ModuleList()
        # SPSNet's surface feature (reference PAGNet_backbone.py:29-31; USE_SURFACE: True in SPSNet.yaml:48)
        self._use_surface = bool(self._pass_stds and sa.get("USE_SURFACE", False))
        if self._use_surface:
            from . import surface_feature

            self.SF_extract = surface_feature.FeatureExtraction()
        channel_in = input_channels - 3
        channel_out_list = [channel_in]
        channel_out = channel_in
        for k in range(len(sa.NSAMPLE_LIST)):
            src = self.layer_inputs[k][-1] if isinstance(self.layer_inputs[k], list) else self.layer_inputs[k]
            channel_in = channel_out_list[src]
            if self.layer_types[k] == "SA_Layer":
                mlps = [[channel_in] + list(spec) for spec in sa.MLPS[k]]
                channel_out = sum(spec[-1] for spec in mlps)
                agg = list(self.aggregation_mlps[k]) if self.aggregation_mlps and self.aggregation_mlps[k] else None
                if agg:
                    channel_out = agg[-1]
                conf = list(self.confidence_mlps[k]) if self.confidence_mlps and self.confidence_mlps[k] else None
                extra = {}
                if sa.get("SS_RADIUS_LIST", None) is not None:
                    extra = {"ss_radii": sa.SS_RADIUS_LIST[k], "ss_nsamples": sa.SS_NSAMPLE_LIST[k]}
                self.SA_modules.append(pointnet2_modules.PointnetSAModuleMSG_WithSampling(
                    npoint_list=sa.NPOINT_LIST[k], sample_range_list=sa.SAMPLE_RANGE_LIST[k],
                    sample_type_list=sa.SAMPLE_METHOD_LIST[k], radii=sa.RADIUS_LIST[k], nsamples=sa.NSAMPLE_LIST[k],
                    mlps=mlps, use_xyz=True, dilated_group=sa.DILATED_GROUP[k], aggregation_mlp=agg,
                    confidence_mlp=conf, num_class=self.num_class, **extra))
            elif self.layer_types[k] == "Vote_Layer":
                self.SA_modules.append(pointnet2_modules.Vote_layer(
                    mlp_list=sa.MLPS[k], pre_channel=channel_out_list[self.layer_inputs[k]],
                    max_translate_range=self.max_translate_range))
            if self._use_surface and k == 3:
                channel_out += 60  # the vote layer also sees the 60 surface channels (PAGNet_backbone.py:91-92)
            channel_out_list.append(channel_out)
        self.num_point_features = channel_out
        for k, m in enumerate(self.SA_modules):   # fp16 range guard: which module a flagged tensor-core launch belongs to
            m._spsk_tag = (k % 31) + 1

    # ---- FPS prefetch ----------------------------------------------------------------------------------------------
    # The D-FPS of layer i+1 needs only the CENTRES of layer i, which exist long before layer i's ball query, MLPs and
    # aggregation have run.  It is a latency chain on 16 of 148 SMs, so it is launched on a side stream at that moment and
    # joined when layer i+1 starts (also inside CUDA-graph capture: the fork/join becomes graph edges).  Same kernels, same
    # inputs, same results; SPSK_FPS_PREFETCH=0 disables it.
    def _prefetchable(self, i: int) -> int:
        """npoint of layer i+1's D-FPS if it can be started from layer i's centres, else 0."""
        import os

        if os.environ.get("SPSK_FPS_PREFETCH", "1") == "0" or self.training or torch.is_grad_enabled():
            return 0
        k = i + 1
        if k >= len(self.SA_modules) or self.layer_types[k] != "SA_Layer" or self.layer_inputs[k] != k or self.ctr_idx_list[k] != -1:
            return 0
        m = self.SA_modules[k]
        types, ranges, npoints = list(m.sample_type_list), list(m.sample_range_list), [n for n in m.npoint_list]
        if len(types) != 1 or ranges[0] != -1 or not ("D-FPS" in types[0] or "DFS" in types[0]) or "S-FPS" in types[0]:
            return 0
        if any(t in types[0] for t in ("cls", "ctr", "ss")):   # the dispatch order of _sample_one: score samplers win
            return 0
        return int(npoints[0]) if npoints[0] > 0 else 0

    def _prefetch_fps(self, new_xyz: torch.Tensor, npoint: int):
        if new_xyz.shape[1] <= npoint or not new_xyz.is_cuda:
            return None
        main = torch.cuda.current_stream()
        sides = self.__dict__.setdefault("_side_streams", {})
        side = sides.get(main.cuda_stream)
        if side is None:
            side = sides[main.cuda_stream] = torch.cuda.Stream(device=new_xyz.device)
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            idx = pointnet2_utils.furthest_point_sample(new_xyz, npoint)
            done = torch.cuda.Event()
            done.record(side)
        if not torch.cuda.is_current_stream_capturing():
            new_xyz.record_stream(side)
            idx.record_stream(main)
        return idx, done

    @staticmethod
    def break_up_pc(pc):
        batch_idx = pc[:, 0]
        xyz = pc[:, 1:4].contiguous()
        features = pc[:, 4:].contiguous() if pc.size(-1) > 4 else None
        return batch_idx, xyz, features

    def forward(self, batch_dict):
        """batch_dict['points']: (B*N, 1+3+C) rows [batch_idx, x, y, z, feat...], equal N per scene.

        fp16 range guard (inference, eager): the tensor-core path stores inter-layer activations as fp16; a checkpoint whose
        BN gains push them beyond 65504 would overflow where the reference's TF32 convolutions do not.  The kernels flag it,
        this wrapper polls the flag once per forward (one device synchronisation; SPSK_FP16_GUARD=0 or `_spsk_guard = False`
        turn it off, CUDA-graph capture skips it) and re-runs with the affected modules on the exact-fp32 kernels."""
        guard = (not self.training and batch_dict["points"].is_cuda and getattr(self, "_spsk_guard", True)
                 and pointnet2_modules.fp16_guard_enabled())
        if not guard:
            return self._forward_impl(batch_dict)
        out = self._forward_impl(dict(batch_dict))
        for _ in range(len(self.SA_modules)):
            if not pointnet2_modules.fp16_guard_check(list(self.SA_modules)):
                break
            out = self._forward_impl(dict(batch_dict))
        batch_dict.update(out)
        return batch_dict

    def _forward_impl(self, batch_dict):
        batch_size = batch_dict["batch_size"]
        points = batch_dict["points"]
        if points.shape[0] % batch_size != 0:
            raise RuntimeError("every scene must hold the same number of points (reference asserts min == max)")
        batch_idx, xyz, features = self.break_up_pc(points)
        stds = batch_dict.get("stds", None) if self._pass_stds else None
        xyz = xyz.view(batch_size, -1, 3)
        if features is not None:
            features = features.view(batch_size, -1, features.shape[-1]).permute(0, 2, 1).contiguous()
        bidx2d = batch_idx.view(batch_size, -1)

        # eager materialisation of the training-only extras when training or under autograd (reference behaviour), lazy otherwise
        lazy = (not self.training) and not torch.is_grad_enabled()

        def coords_of(t):
            f = lambda: torch.cat([bidx2d[:, :t.shape[1], None].float(), t.view(batch_size, -1, 3)], dim=-1)  # noqa: E731
            return f if lazy else f()

        def preds_of(t):
            f = lambda: torch.cat([bidx2d[:, :t.shape[1], None].float(), t.reshape(batch_size, -1, t.shape[-1])], dim=-1)  # noqa: E731
            return f if lazy else f()

        encoder_xyz, encoder_features, sa_ins_preds = [xyz], [features], LazyList()
        encoder_coords = LazyList([coords_of(xyz)])
        li_cls_pred = None
        centers = centers_origin = ctr_offsets = None
        surface = None  # (B, n_i, 60) point-major surface features of the points kept so far
        presampled = None
        for i, module in enumerate(self.SA_modules):
            xyz_input = encoder_xyz[self.layer_inputs[i]]
            feature_input = encoder_features[self.layer_inputs[i]]
            if self.layer_types[i] == "SA_Layer":
                ctr_xyz = encoder_xyz[self.ctr_idx_list[i]] if self.ctr_idx_list[i] != -1 else None
                kw = {"stds": stds} if self._pass_stds else {}
                if presampled is not None:
                    kw["_presampled"], presampled = presampled, None
                nxt = self._prefetchable(i)
                if nxt:
                    box = []
                    kw["_after_new_xyz"] = lambda c, n=nxt, b=box: b.append(self._prefetch_fps(c, n))
                li_xyz, li_features, li_cls_pred, sampled_idx, stds_out = module(xyz_input, feature_input, li_cls_pred, ctr_xyz=ctr_xyz, **kw)
                if nxt and box and box[0] is not None:
                    presampled = box[0]
                if self._pass_stds:
                    stds = stds_out
                if self._use_surface and i <= 4:
                    # reference PAGNet_backbone.py:153-157: extract once on the full cloud, then follow the samplers
                    # (kept point-major here: one row gather per layer instead of permute + gather_operation)
                    if i == 0:
                        surface = self.SF_extract(xyz)
                    surface = pointnet2_utils.gather_rows(surface.contiguous(), sampled_idx.int().contiguous())
            else:  # Vote_Layer
                kw = {"center_surface_futures": surface.permute(0, 2, 1).contiguous()} if surface is not None else {}
                li_xyz, li_features, xyz_select, ctr_offsets = module(xyz_input, feature_input, **kw)
                centers, centers_origin = li_xyz, xyz_select
                encoder_coords.append(coords_of(centers_origin))
            encoder_xyz.append(li_xyz)
            encoder_coords.append(coords_of(li_xyz))
            encoder_features.append(li_features)
            if li_cls_pred is not None:
                sa_ins_preds.append(preds_of(li_cls_pred))
            else:
                sa_ins_preds.append([])

        ctr_batch_idx = bidx2d[:, :li_xyz.shape[1]].contiguous().view(-1)
        batch_dict["ctr_offsets"] = torch.cat((ctr_batch_idx[:, None].float(), ctr_offsets.contiguous().view(-1, 3)), dim=1)
        batch_dict["centers"] = torch.cat((ctr_batch_idx[:, None].float(), centers.contiguous().view(-1, 3)), dim=1)
        batch_dict["centers_origin"] = torch.cat((ctr_batch_idx[:, None].float(), centers_origin.contiguous().view(-1, 3)), dim=1)
        last = encoder_features[-1]
        cf = last.permute(0, 2, 1).contiguous().view(-1, last.shape[1])
        hit = getattr(last, "_spsk_twin", None)
        if hit is not None and hit[1] == last._version and hit[2] > 0:
            # the producing layer's point-major fp16 rows [values | residuals] ARE the rows of centers_features:
            # hand them to the head (dense_head.IASSD_Head) so it does not convert again
            cf._spsk_rows16 = (hit[0].view(cf.shape[0], -1), cf._version, hit[2])
        batch_dict["centers_features"] = cf
        batch_dict["ctr_batch_idx"] = ctr_batch_idx
        batch_dict["encoder_xyz"] = encoder_xyz
        batch_dict["encoder_coords"] = encoder_coords
        batch_dict["sa_ins_preds"] = sa_ins_preds
        batch_dict["encoder_features"] = encoder_features
        return batch_dict


class PAGNet_Backbone(IASSD_Backbone):
    """SPSNet-IA backbone (reference PAGNet_backbone.py:7-237): IA-SSD's backbone threading the
    per-point stability `batch_dict['stds']` through every SA layer (PAGNet_backbone.py:117,150)."""

    _pass_stds = True

"""IA-SSD single-stage detector, inference only: backbone_3d -> point_head -> post-processing.

Mirror of the reference's `pcdet/models/detectors/IASSD.py:3-19` (and `PAGNet.py`, same forward) with the module
names Detector3DTemplate.build_networks gives them (`backbone_3d`, `point_head`: detector3d_template.py:19-22,101-120,
153-169), so a reference checkpoint's `model_state` keys load unchanged.  `forward` returns the reference's
(pred_dicts, recall_dicts); `forward_padded` is the sync-free form used by the CUDA-graph serving pipeline:
every tensor has a fixed shape, nothing is read back, nothing is allocated inside libspsk.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import backbone as bb
from . import dense_head as dh


class IASSD(nn.Module):
    def __init__(self, model_cfg=None, num_class: int = 3, input_channels: int = 4, backbone_cls=None):
        super().__init__()
        cfg = bb.Cfg(model_cfg) if model_cfg is not None else bb.Cfg({
            "BACKBONE_3D": bb.kitti_iassd_cfg(), "POINT_HEAD": dh.kitti_iassd_head_cfg(),
            "POST_PROCESSING": dh.KITTI_POST_PROCESSING})
        self.model_cfg = cfg
        self.num_class = num_class
        self.backbone_3d = (backbone_cls or bb.IASSD_Backbone)(cfg.BACKBONE_3D, num_class=num_class, input_channels=input_channels)
        self.point_head = dh.IASSD_Head(num_class, self.backbone_3d.num_point_features, cfg.POINT_HEAD,
                                        post_process_cfg=cfg.POST_PROCESSING)
        self.module_list = [self.backbone_3d, self.point_head]

    def forward_padded(self, batch_dict):
        """batch_dict with 'detections' (dense_head.Detections: padded boxes / scores / labels / index + counts)."""
        if self.training:
            raise NotImplementedError("training is out of scope (SURVEY.md §8); call .eval()")
        for m in self.module_list:
            batch_dict = m(batch_dict)
        det = dh.detections_padded(batch_dict, self.model_cfg.POST_PROCESSING)
        batch_dict["detections"] = det
        batch_dict["det_boxes"], batch_dict["det_scores"] = det.boxes, det.scores
        batch_dict["det_labels"], batch_dict["det_count"] = det.labels, det.count
        return batch_dict

    def forward(self, batch_dict):
        batch_dict = self.forward_padded(batch_dict)
        return dh.post_processing(batch_dict, self.model_cfg.POST_PROCESSING)


class _StabilityEncoding(nn.Module):
    """The role of the reference's `PAGNet_encoding` (pcdet/models/backbones_2d/map_to_bev/PAGNet_encoding.py:10-31): run
    the frozen stability generator and leave per-point `stds` in batch_dict.  Its point-deletion block (:33-69) is an
    ablation that needs ground-truth `fake_labels`; it is available separately as stability.delete_unstable_points."""

    def __init__(self, generator: nn.Module):
        super().__init__()
        self.generator = generator

    def forward(self, batch_dict):
        batch_dict = self.generator(batch_dict)
        for k in ("encoder_xyz", "encoder_coords", "sa_ins_preds", "soc_feature"):
            batch_dict.pop(k, None)
        return batch_dict


class SPSNetIA(IASSD):
    """SPSNet-IA detector as shipped (reference tools/cfgs/kitti_models/SPSNet.yaml:24-121): model NAME IASSD with
    MAP_TO_BEV = PAGNet_encoding (the stability generator -> per-point stds), BACKBONE_3D = PAGNet_Backbone (stability-aware
    sampling, USE_SURFACE) and POINT_HEAD = MLT_SSD_Head (eval-mode forward identical to IASSD_Head,
    MLT_SSD_head.py:788-841); modules run in the reference's topology order map_to_bev_module -> backbone_3d -> point_head
    (detector3d_template.py:23-26).  generator=None: batch_dict must already hold `stds`."""

    def __init__(self, model_cfg=None, num_class: int = 3, input_channels: int = 4, generator: nn.Module = None,
                 surface: bool = True):
        if model_cfg is None:
            model_cfg = {"BACKBONE_3D": bb.kitti_spsnet_surface_cfg() if surface else bb.kitti_spsnet_cfg(),
                         "POINT_HEAD": dh.kitti_iassd_head_cfg(), "POST_PROCESSING": dh.KITTI_POST_PROCESSING}
        super().__init__(model_cfg, num_class, input_channels, backbone_cls=bb.PAGNet_Backbone)
        if generator is not None:
            self.map_to_bev_module = _StabilityEncoding(generator)
            self.module_list = [self.map_to_bev_module, self.backbone_3d, self.point_head]

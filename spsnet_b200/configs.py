"""Model configurations, synthetic-weight helpers and nothing else: importing this module does NOT load libspsk.so.

Shared by the product modules (backbone.py, dense_head.py re-export these names) and by bench.py's reference arm, which
must build the reference's own modules with the same seeded weights WITHOUT mapping any of this repo's kernels.

Reference: tools/cfgs/kitti_models/IA-SSD.yaml:33-84,108-121, tools/cfgs/kitti_models/SPSNet.yaml:38-71,
tools/cfgs/waymo_models/IA-SSD.yaml:45-93,118-130.
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn


class Cfg(dict):
    """dict with attribute access and `.get`, enough for the reference's `model_cfg.SA_CONFIG.X` style."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return Cfg(v) if isinstance(v, dict) and not isinstance(v, Cfg) else v

    def get(self, k, default=None):
        v = super().get(k, default)
        return Cfg(v) if isinstance(v, dict) and not isinstance(v, Cfg) else v


# reference: tools/cfgs/kitti_models/IA-SSD.yaml:33-57
KITTI_IASSD_SA_CONFIG = {
    "NPOINT_LIST": [[4096], [1024], [512], [256], [-1], [256]],
    "SAMPLE_RANGE_LIST": [[-1], [-1], [-1], [-1], [-1], [-1]],
    "SAMPLE_METHOD_LIST": [["D-FPS"], ["D-FPS"], ["ctr_aware"], ["ctr_aware"], [], []],
    "RADIUS_LIST": [[0.2, 0.8], [0.8, 1.6], [1.6, 4.8], [], [], [4.8, 6.4]],
    "NSAMPLE_LIST": [[16, 32], [16, 32], [16, 32], [], [], [16, 32]],
    "MLPS": [[[16, 16, 32], [32, 32, 64]],
             [[64, 64, 128], [64, 96, 128]],
             [[128, 128, 256], [128, 256, 256]],
             [],
             [128],
             [[256, 256, 512], [256, 512, 1024]]],
    "LAYER_TYPE": ["SA_Layer", "SA_Layer", "SA_Layer", "SA_Layer", "Vote_Layer", "SA_Layer"],
    "DILATED_GROUP": [False, False, False, False, False, False],
    "AGGREGATION_MLPS": [[64], [128], [256], [256], [], [512]],
    "CONFIDENCE_MLPS": [[], [128], [256], [], [], []],
    "LAYER_INPUT": [0, 1, 2, 3, 4, 3],
    "CTR_INDEX": [-1, -1, -1, -1, -1, 5],
    "MAX_TRANSLATE_RANGE": [3.0, 3.0, 2.0],
}


def kitti_iassd_cfg() -> Cfg:
    """IA-SSD KITTI backbone: 16384 -> 4096 (D-FPS) -> 1024 (D-FPS) -> 512 (ctr) -> 256 (ctr) -> vote -> SA."""
    return Cfg({"SA_CONFIG": copy.deepcopy(KITTI_IASSD_SA_CONFIG)})


def kitti_spsnet_cfg() -> Cfg:
    """SPSNet-IA (reference tools/cfgs/kitti_models/SPSNet.yaml:38-71): stability-aware top-k in layers 2, 3.
    The surface-feature branch (USE_SURFACE) is out of scope, so layer 1 keeps IA-SSD's 64-wide MLP."""
    c = copy.deepcopy(KITTI_IASSD_SA_CONFIG)
    c["SAMPLE_METHOD_LIST"] = [["D-FPS"], ["D-FPS"], ["sss_aware"], ["sss_aware"], [], []]
    c["SS_RADIUS_LIST"] = [[0.05], [0.2], [], [], [], []]
    c["SS_NSAMPLE_LIST"] = [[16], [16], [], [], [], [1]]
    return Cfg({"SA_CONFIG": c})


def kitti_spsnet_surface_cfg() -> Cfg:
    """SPSNet-IA exactly as shipped (reference tools/cfgs/kitti_models/SPSNet.yaml:38-71): stability-aware top-k,
    USE_SURFACE: True (60 surface channels into the vote layer) and the 124-wide first MLP of SA layer 1."""
    c = copy.deepcopy(kitti_spsnet_cfg()["SA_CONFIG"])
    c["USE_SURFACE"] = True
    c["MLPS"][1] = [[124, 64, 128], [124, 96, 128]]
    return Cfg({"SA_CONFIG": c})


def waymo_iassd_cfg() -> Cfg:
    """reference tools/cfgs/waymo_models/IA-SSD.yaml:45-65: all point counts x4."""
    c = copy.deepcopy(KITTI_IASSD_SA_CONFIG)
    c["NPOINT_LIST"] = [[16384], [4096], [2048], [1024], [-1], [1024]]
    return Cfg({"SA_CONFIG": c})


# reference: tools/cfgs/kitti_models/IA-SSD.yaml:59-84,108-121
KITTI_IASSD_HEAD = {
    "NAME": "IASSD_Head",
    "CLS_FC": [256, 256],
    "REG_FC": [256, 256],
    "CLASS_AGNOSTIC": False,
    "TARGET_CONFIG": {
        "BOX_CODER": "PointResidual_BinOri_Coder",
        "BOX_CODER_CONFIG": {"angle_bin_num": 12, "use_mean_size": True,
                             "mean_size": [[3.9, 1.6, 1.56], [0.8, 0.6, 1.73], [1.76, 0.6, 1.73]]},
    },
}
KITTI_POST_PROCESSING = {
    "RECALL_THRESH_LIST": [0.3, 0.5, 0.7],
    "SCORE_THRESH": 0.1,
    "OUTPUT_RAW_SCORE": False,
    "NMS_CONFIG": {"MULTI_CLASSES_NMS": False, "NMS_TYPE": "nms_gpu", "NMS_THRESH": 0.01, "NMS_PRE_MAXSIZE": 4096,
                   "NMS_POST_MAXSIZE": 500},
}


def waymo_iassd_head_cfg() -> Cfg:
    """reference tools/cfgs/waymo_models/IA-SSD.yaml:69-93: same stacks, Waymo mean sizes."""
    c = copy.deepcopy(KITTI_IASSD_HEAD)
    c["TARGET_CONFIG"]["BOX_CODER_CONFIG"]["mean_size"] = [[4.7, 2.1, 1.7], [0.91, 0.86, 1.73], [1.78, 0.84, 1.78]]
    return Cfg(c)


def waymo_post_processing() -> dict:
    """reference tools/cfgs/waymo_models/IA-SSD.yaml:118-130 (NMS_THRESH 0.1)."""
    c = copy.deepcopy(KITTI_POST_PROCESSING)
    c["NMS_CONFIG"]["NMS_THRESH"] = 0.1
    return c


def kitti_iassd_head_cfg() -> Cfg:
    return Cfg(copy.deepcopy(KITTI_IASSD_HEAD))


def randomize_bn_stats(module: nn.Module, seed: int = 0) -> None:
    """Seeded non-trivial BN affine + running statistics so the BN fold is exercised."""
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            with torch.no_grad():
                n = m.num_features
                m.weight.copy_(torch.empty(n).uniform_(0.5, 1.5, generator=g))
                m.bias.copy_(torch.randn(n, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
                m.running_var.copy_(torch.empty(n).uniform_(0.5, 1.5, generator=g))

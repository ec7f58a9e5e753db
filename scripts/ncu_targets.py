#!/usr/bin/env python
"""One invocation of each widening kernel at its headline shape, for `ncu --set full` (scripts/gpu_ncu_widening.sh)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from spsnet_b200 import dense_head as H  # noqa: E402
from spsnet_b200 import scenes  # noqa: E402
from spsnet_b200 import surface_feature as SF  # noqa: E402

B, m = 16, 256
rng = np.random.default_rng(0)
cls = torch.from_numpy(rng.normal(0.0, 2.0, (B * m, 3)).astype(np.float32)).cuda()
reg = torch.from_numpy(rng.normal(0, 0.15, (B * m, 30)).astype(np.float32)).cuda()
ctr = torch.from_numpy(np.concatenate([scenes.make_boxes(b, m, n_objects=25)[:, :3] for b in range(B)])).cuda()
coder = H.PointResidual_BinOri_Coder(**H.KITTI_IASSD_HEAD["TARGET_CONFIG"]["BOX_CODER_CONFIG"])
H._detect_call(B, m, 3, 12, cls=cls, reg=reg, centers=ctr, mean_size=coder.mean_size, nms=H._nms_args(H.Cfg(H.KITTI_POST_PROCESSING)))
torch.manual_seed(0)
fe = SF.FeatureExtraction().cuda().eval()
xyz = torch.from_numpy(np.ascontiguousarray(scenes.make_batch(0, 16, 16384)[:, :, :3])).cuda()
with torch.no_grad():
    fe(xyz)
torch.cuda.synchronize()
print("ok")

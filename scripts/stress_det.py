#!/usr/bin/env python
"""Randomised parity sweep of the IoU / NMS / post-processing kernels against the reference's rebuilt CUDA kernels
(oracle/_ref): many seeds, sizes, thresholds and degenerate inputs; prints one JSON summary (mismatch counts)."""
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
sys.modules.setdefault("SharedArray", types.ModuleType("SharedArray"))
from pcdet.ops.iou3d_nms import iou3d_nms_cuda as RC  # noqa: E402
from pcdet.ops.iou3d_nms import iou3d_nms_utils as RU  # noqa: E402
from spsnet_b200 import iou3d_nms_utils as U  # noqa: E402
from spsnet_b200 import scenes  # noqa: E402


def boxes(rng, n, kind):
    b = scenes.make_boxes(int(rng.integers(0, 1 << 30)), n, n_objects=max(1, int(rng.integers(1, max(2, n // 4)))))
    if kind == "tiny":
        b[:, 3:6] *= 1e-3
    elif kind == "huge":
        b[:, 3:6] *= 30
    elif kind == "aligned":
        b[:, 6] = np.float32(np.pi / 2) * rng.integers(0, 4, n)
    elif kind == "same":
        b[:] = b[0]
        b[:, 0] += rng.normal(0, 1e-3, n).astype(np.float32)
    elif kind == "grid":
        b[:, 0] = np.round(b[:, 0])
        b[:, 1] = np.round(b[:, 1])
        b[:, 3:5] = 2.0
        b[:, 6] = 0
    return torch.from_numpy(b).cuda()


def main():
    rng = np.random.default_rng(2024)
    res = {"iou_cases": 0, "iou_mismatch_entries": 0, "iou3d_mismatch_entries": 0, "nms_cases": 0, "nms_mismatch": 0, "normal_mismatch": 0,
           "max_abs_iou_diff": 0.0}
    kinds = ["plain", "plain", "tiny", "huge", "aligned", "same", "grid"]
    for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 150):
        kind = kinds[it % len(kinds)]
        na, nb = int(rng.integers(1, 700)), int(rng.integers(1, 700))
        a, b = boxes(rng, na, kind), boxes(rng, nb, kind)
        ref = torch.zeros(na, nb, device="cuda")
        RC.boxes_iou_bev_gpu(a, b, ref)
        got = U.boxes_iou_bev(a, b)
        bad = (got != ref) & ~(torch.isnan(got) & torch.isnan(ref))
        res["iou_cases"] += 1
        res["iou_mismatch_entries"] += int(bad.sum())
        if bad.any():
            res["max_abs_iou_diff"] = max(res["max_abs_iou_diff"], float((got - ref)[bad].abs().max()))
        r3 = RU.boxes_iou3d_gpu(a, b)
        g3 = U.boxes_iou3d_gpu(a, b)
        res["iou3d_mismatch_entries"] += int(((g3 != r3) & ~(torch.isnan(g3) & torch.isnan(r3))).sum())
        n = int(rng.integers(1, 3000))
        bx = boxes(rng, n, kind)
        sc = torch.from_numpy(rng.permutation(n).astype(np.float32)).cuda()
        th = float(rng.choice([0.01, 0.1, 0.25, 0.5, 0.7]))
        k, _ = U.nms_gpu(bx, sc, th)
        rk, _ = RU.nms_gpu(bx, sc, th)
        res["nms_cases"] += 1
        res["nms_mismatch"] += int(not torch.equal(k, rk))
        k, _ = U.nms_normal_gpu(bx, sc, th)
        rk, _ = RU.nms_normal_gpu(bx, sc, th)
        res["normal_mismatch"] += int(not torch.equal(k, rk))
    print(json.dumps(res))


if __name__ == "__main__":
    main()

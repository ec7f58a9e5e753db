#!/usr/bin/env python
"""Generate tests/golden/iou3d_reference_cpu.npz by RUNNING the reference's own CPU rotated-IoU implementation
(pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp, compiled unmodified into oracle/_ref by oracle/build_ref.sh) in the build
container -- no GPU needed.  It is the pin for the iou3d part of the C oracle (tests/test_oracle_cpu.py demands
bit-equality) and, through the greedy loop replayed on the reference IoU matrix, for its NMS.

    python tests/golden/make_golden_iou3d.py          (needs /root/reference-built oracle/_ref)

Inputs are regenerated from seeds by spsnet_b200.scenes.make_boxes, so only OUTPUTS are stored.
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
sys.modules.setdefault("SharedArray", types.ModuleType("SharedArray"))  # absent, unrelated dependency of pcdet.utils
from spsnet_b200 import scenes  # noqa: E402

# (n_a, n_b, seed) -- shared with tests/test_oracle_cpu.py
IOU_CASES = [(1, 1, 0), (7, 5, 1), (64, 64, 2), (100, 37, 3), (257, 130, 4)]
NMS_CASES = [(1, 0.01, 10), (64, 0.01, 11), (65, 0.1, 12), (256, 0.01, 13), (300, 0.3, 14), (1000, 0.7, 15)]


def greedy(iou, thresh):
    """iou3d_nms.cpp:109-126 host loop, replayed on the reference IoU matrix."""
    n = iou.shape[0]
    dead = np.zeros(n, bool)
    keep = []
    for i in range(n):
        if dead[i]:
            continue
        keep.append(i)
        dead[i + 1:] |= iou[i, i + 1:] > np.float32(thresh)
    return np.asarray(keep, np.int64)


def main():
    from pcdet.ops.iou3d_nms import iou3d_nms_cuda as ref

    out = {}
    for na, nb, seed in IOU_CASES:
        ab = scenes.make_boxes(seed, na + nb)  # one draw: a and b share object locations, so many pairs overlap
        a, b = np.ascontiguousarray(ab[:na]), np.ascontiguousarray(ab[na:])
        iou = torch.zeros(na, nb)
        ref.boxes_iou_bev_cpu(torch.from_numpy(a), torch.from_numpy(b), iou)
        out[f"iou_{na}_{nb}_{seed}"] = iou.numpy()
    for n, thresh, seed in NMS_CASES:
        bx = scenes.make_boxes(seed, n)
        iou = torch.zeros(n, n)
        ref.boxes_iou_bev_cpu(torch.from_numpy(bx), torch.from_numpy(bx), iou)
        out[f"nms_{n}_{seed}"] = greedy(iou.numpy(), thresh)
    dst = Path(__file__).resolve().parent / "iou3d_reference_cpu.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

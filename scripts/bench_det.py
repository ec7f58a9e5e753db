#!/usr/bin/env python
"""Micro-benchmarks of the head / NMS widening on one GPU: ours vs the reference's rebuilt kernels (oracle/_ref).
CUDA-event timing for device-only ops, wall clock (with synchronize) for the reference's host-looping NMS.
    python scripts/bench_det.py > gpurun_out/det_ops.json
"""
import copy
import json
import sys
import time
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
sys.modules.setdefault("SharedArray", types.ModuleType("SharedArray"))
from spsnet_b200 import dense_head as H  # noqa: E402
from spsnet_b200 import iou3d_nms_utils as U  # noqa: E402
from spsnet_b200 import scenes  # noqa: E402

try:
    from pcdet.models.model_utils import model_nms_utils as RN
    from pcdet.ops.iou3d_nms import iou3d_nms_utils as RU
    from pcdet.utils import box_coder_utils as RC
except Exception as e:  # pragma: no cover
    print("reference modules unavailable:", e, file=sys.stderr)
    RU = RN = RC = None


def wall(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e6


def dev(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    out = {}
    for n in (256, 1024, 4096):
        a = torch.from_numpy(scenes.make_boxes(n, n)).cuda()
        r = {"ours_us": dev(lambda: U.boxes_iou_bev(a, a))}
        if RU:
            r["ref_us"] = dev(lambda: RU.boxes_iou_bev(a, a))
        out[f"boxes_iou_bev[{n}x{n}]"] = r
        r = {"ours_us": dev(lambda: U.boxes_iou3d_gpu(a, a))}
        if RU:
            r["ref_us"] = dev(lambda: RU.boxes_iou3d_gpu(a, a))
        out[f"boxes_iou3d[{n}x{n}]"] = r
        s = torch.linspace(1, 0, n, device="cuda")
        r = {"ours_us_wall": wall(lambda: U.nms_gpu(a, s, 0.1)), "ours_device_us": dev(lambda: U.nms_batch(a[None], 0.1))}
        if RU:
            r["ref_us_wall"] = wall(lambda: RU.nms_gpu(a, s, 0.1))
        out[f"nms_gpu[{n}]"] = r
    # batched post-processing, IA-SSD KITTI shape: 16 scenes x 256 centres
    B, m = 16, 256
    rng = np.random.default_rng(0)
    cls = torch.from_numpy(rng.normal(0.0, 2.0, (B * m, 3)).astype(np.float32)).cuda()
    reg = torch.from_numpy(rng.normal(0, 0.15, (B * m, 30)).astype(np.float32)).cuda()
    ctr = torch.from_numpy(np.concatenate([scenes.make_boxes(b, m, n_objects=25)[:, :3] for b in range(B)])).cuda()
    coder = H.PointResidual_BinOri_Coder(**H.KITTI_IASSD_HEAD["TARGET_CONFIG"]["BOX_CODER_CONFIG"])
    cfg = H.Cfg(H.KITTI_POST_PROCESSING)
    nms = H._nms_args(cfg)
    ms = coder.mean_size
    r = {"ours_us": dev(lambda: H._detect_call(B, m, 3, 12, cls=cls, reg=reg, centers=ctr, mean_size=ms, nms=nms))}
    _, _, _, det = H._detect_call(B, m, 3, 12, cls=cls, reg=reg, centers=ctr, mean_size=ms, nms=nms)
    r["detections_per_scene"] = float(det.count.float().mean())
    if RU:
        rc = RC.PointResidual_BinOri_Coder(**H.KITTI_IASSD_HEAD["TARGET_CONFIG"]["BOX_CODER_CONFIG"])

        def ref_chain():
            _, pc = cls.max(dim=-1)
            boxes = rc.decode_torch(reg, ctr, pc + 1)
            res = []
            for b in range(B):
                bp = boxes[b * m:(b + 1) * m]
                cp, lp = torch.max(torch.sigmoid(cls[b * m:(b + 1) * m]), dim=-1)
                sel, sc = RN.class_agnostic_nms(box_scores=cp, box_preds=bp, nms_config=cfg.NMS_CONFIG, score_thresh=cfg.SCORE_THRESH)
                res.append((bp[sel], sc, lp[sel] + 1))
            return res

        r["ref_us_wall"] = wall(ref_chain, iters=10)
    out["decode+postprocess[16x256]"] = r
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Surface-feature extractor at the SPSNet KITTI shape (16 x 16384 points): fused kernels vs the reference module."""
import json
import sys
import warnings
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
from spsnet_b200 import _lib, scenes  # noqa: E402
from spsnet_b200 import surface_feature as SF  # noqa: E402


def dev(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    out = {}
    torch.manual_seed(0)
    fe = SF.FeatureExtraction().cuda().eval()
    for B, N in ((16, 16384), (4, 16384)):
        xyz = torch.from_numpy(np.ascontiguousarray(scenes.make_batch(0, B, N)[:, :, :3])).cuda()
        with torch.no_grad():
            r = {"ours_ms": dev(lambda: fe(xyz))}
            # per-kernel split
            names = ("spsk_edge_conv_point", "spsk_edge_conv_aggregate", "spsk_ball_query_msg_grid", "spsk_ball_query_msg")
            recs, origs = [], {}
            for n in names:
                origs[n] = getattr(_lib.lib, n)

                def wrap(n=n, f=origs[n]):
                    def inner(*a):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(); rc = f(*a); e1.record(); recs.append((n, e0, e1)); return rc
                    return inner
                setattr(_lib.lib, n, wrap())
            fe(xyz)
            torch.cuda.synchronize()
            for n in names:
                setattr(_lib.lib, n, origs[n])
            split = {}
            for n, e0, e1 in recs:
                split.setdefault(n, []).append(round(e0.elapsed_time(e1), 4))
            r["kernels_ms"] = split
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    from pcdet.ops.pointnet2.pointnet2_batch import surface_feature as RS
                ref = RS.FeatureExtraction().cuda().eval()
                ref.load_state_dict(fe.state_dict())
                r["ref_ms"] = dev(lambda: ref(xyz), iters=3, warm=1)
            except Exception as e:  # pragma: no cover
                r["ref_error"] = str(e)[:200]
        out[f"FeatureExtraction[{B}x{N}]"] = r
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

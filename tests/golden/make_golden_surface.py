#!/usr/bin/env python
"""Generate tests/golden/surface_reference_b200.npz by RUNNING the reference's own FeatureExtraction
(pcdet/ops/pointnet2/pointnet2_batch/surface_feature.py, unmodified, on the rebuilt reference CUDA ops of oracle/_ref)
on a B200.  Its ball_query calls are RECORDED (inputs and outputs) while it runs, so the fixture pins both halves of the
oracle restatement: the neighbour lists the reference gets when it passes 24-wide features as xyz (bit-exact) and the
dense edge MLP (to fp32 round-off, with those lists teacher-forced).

    gpurun -- python tests/golden/make_golden_surface.py gpurun_out/golden      (then copy the .npz into tests/golden/)
"""
import sys
import warnings
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))

B, N, SEED = 2, 300, 5   # shared with tests/test_surface_cpu.py


def reference_state(seed=SEED):
    """Seeded weights, generated WITHOUT the reference (so the CPU test can rebuild them): dict name -> fp32 array."""
    rng = np.random.default_rng(seed)
    shapes = {}
    cin = 3
    for i in range(4):
        shapes[f"transforms.{i}.linear"] = (24, cin)
        shapes[f"convs.{i}.layer_first.linear"] = (12, 24 if i == 0 else 72)
        shapes[f"convs.{i}.layers.0.linear"] = (12, 36)
        shapes[f"convs.{i}.layer_last.linear"] = (12, 48)
        cin = 60
    sd = {}
    for k, (o, c) in shapes.items():
        sd[k + ".weight"] = (rng.standard_normal((o, c)) / np.sqrt(c) * 1.5).astype(np.float32)
        sd[k + ".bias"] = (rng.standard_normal(o) * 0.1).astype(np.float32)
    return sd


def main():
    out_dir = Path(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
    out_dir.mkdir(parents=True, exist_ok=True)
    from spsnet_b200 import scenes

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from pcdet.ops.pointnet2.pointnet2_batch import pointnet2_utils as RU
        from pcdet.ops.pointnet2.pointnet2_batch import surface_feature as RS
    fe = RS.FeatureExtraction().cuda().eval()
    fe.load_state_dict({k: torch.from_numpy(v) for k, v in reference_state().items()})
    xyz = torch.from_numpy(np.ascontiguousarray(scenes.make_batch(2000 + SEED, B, N)[:, :, :3])).cuda()
    rec = []
    orig = RU.ball_query

    def recording(radius, nsample, xyz_, new_xyz_):
        idx = orig(radius, nsample, xyz_, new_xyz_)
        rec.append((xyz_.detach().cpu().numpy().copy(), idx.cpu().numpy().copy()))
        return idx

    RU.ball_query = recording
    try:
        with torch.no_grad():
            y = fe(xyz)
    finally:
        RU.ball_query = orig
    out = {"out": y.cpu().numpy()}
    for i, (t, idx) in enumerate(rec):
        out[f"t{i}"] = t
        out[f"idx{i}"] = idx
    np.savez_compressed(out_dir / "surface_reference_b200.npz", **out)
    print("wrote", out_dir / "surface_reference_b200.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

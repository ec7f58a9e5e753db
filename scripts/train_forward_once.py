#!/usr/bin/env python
"""Two warm-up steps + ONE measured training forward of the KITTI layer-1 and layer-5 set-abstraction MLPs (B = 8) -- the command
the ncu launch list of the training path is taken over (profiles/r02_ncu_train_launches.txt): which kernels a fused training
forward launches and how long each takes."""
import copy
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "scripts"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench_train import LAYERS  # noqa: E402
from spsnet_b200 import pointnet2_modules as pm, pointnet2_utils as pu, scenes  # noqa: E402

B = 8
for name in ("L1", "L5"):
    c = LAYERS[name]
    torch.manual_seed(0)
    full = torch.from_numpy(np.ascontiguousarray(scenes.make_batch(5, B, 16384)[:, :, :3])).cuda()
    xyz = pu.gather_rows(full, pu.furthest_point_sample(full, c["n"])) if c["n"] < 16384 else full
    new_xyz = pu.gather_rows(xyz, pu.furthest_point_sample(xyz, c["m"])) if c["m"] < c["n"] else xyz.clone()
    feats = torch.randn(B, c["c"], c["n"], device="cuda")
    mod = pm.PointnetSAModuleMSG(npoint=c["m"], radii=c["radii"], nsamples=c["nsamples"], mlps=copy.deepcopy(c["mlps"]), use_xyz=True).cuda().train()
    for _ in range(3):
        out = mod(xyz, feats, new_xyz)[1]
    torch.cuda.synchronize()
    print(name, tuple(out.shape), float(out.abs().mean()))

#!/usr/bin/env python
"""Per-parameter gradient deviation of a set-abstraction module in train() against the unmodified reference module (oracle/_ref,
fp32): ours on the fused path, ours on the op-by-op composition, and the reference's own stock (cuDNN TF32) run -- which tensors
carry the deviation and whether it is specific to the fused path.   python scripts/diag_train_grads.py"""
import copy
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
import importlib, warnings  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

from spsnet_b200 import pointnet2_modules as pm, scenes  # noqa: E402

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    refm = importlib.import_module("pcdet.ops.pointnet2.pointnet2_batch.pointnet2_modules")


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def main():
    kw = dict(npoint_list=[256], sample_type_list=["D-FPS"], radii=[0.8, 1.6], nsamples=[16, 32], mlps=[[64, 64, 64, 128], [64, 64, 96, 128]],
              aggregation_mlp=[128], confidence_mlp=[128])
    for seed in (2, 3):
        torch.manual_seed(seed)
        ours = pm.PointnetSAModuleMSG_WithSampling(sample_range_list=[-1], num_class=3, **copy.deepcopy(kw)).cuda().train()
        for mod in ours.modules():
            if isinstance(mod, (nn.BatchNorm1d, nn.BatchNorm2d)):
                mod.weight.data.uniform_(0.5, 1.5); mod.bias.data.uniform_(-0.2, 0.2)
        ref = refm.PointnetSAModuleMSG_WithSampling(sample_range_list=[-1], num_class=3, **copy.deepcopy(kw)).cuda().train()
        ref.load_state_dict(ours.state_dict())
        xyz = torch.from_numpy(np.ascontiguousarray(scenes.make_batch(61, 2, 1024)[:, :, :3])).cuda()
        feats = torch.randn(2, 64, 1024, device="cuda")
        arms = {"ref_fp32": (ref, False, "0"), "ours_fused": (ours, False, "1"), "ours_composed": (copy.deepcopy(ours), False, "0"),
                "ref_stock_tf32": (copy.deepcopy(ref), True, "0")}
        grads, outs = {}, {}
        for name, (mod, tf32, flag) in arms.items():
            os.environ["SPSK_TRAIN_FUSED"] = flag
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            out = mod(xyz, feats.clone(), None)
            torch.manual_seed(13)
            loss = (out[1] * torch.randn_like(out[1])).sum() + (out[2] * torch.randn_like(out[2])).sum()
            mod.zero_grad(); loss.backward()
            grads[name] = {n: p.grad.clone() for n, p in mod.named_parameters()}
            outs[name] = (out[1].detach(), out[2].detach())
        print(f"seed {seed}: features / logits vs ref_fp32:", {k: (f"{rel(v[0], outs['ref_fp32'][0]):.1e}", f"{rel(v[1], outs['ref_fp32'][1]):.1e}") for k, v in outs.items() if k != "ref_fp32"})
        print(f"{'parameter':44s} {'fused':>9s} {'composed':>9s} {'stockTF32':>9s}")
        for n in grads["ref_fp32"]:
            e = [rel(grads[a][n], grads["ref_fp32"][n]) for a in ("ours_fused", "ours_composed", "ref_stock_tf32")]
            print(f"{n:44s} {e[0]:9.1e} {e[1]:9.1e} {e[2]:9.1e}")


if __name__ == "__main__":
    main()

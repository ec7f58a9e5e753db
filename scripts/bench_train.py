#!/usr/bin/env python
"""Training-mode grouped MLP (grouping -> [conv1x1 -> BatchNorm(batch statistics) -> ReLU] x 3 -> max-pool) of the KITTI IA-SSD
set-abstraction layers, forward and forward + backward, per batch of B scenes:

    fused      spsnet_b200/train_fused.py (3 statistics passes + 1 pooled pass on the tcgen05 kernel; recompute backward)
    composed   the same module on the op-by-op composition (SPSK_TRAIN_FUSED=0: libspsk grouping ops + torch/cuDNN)
    reference  the unmodified reference module + its CUDA ops (oracle/_ref), stock settings (cuDNN TF32)

with the peak device memory of one forward + backward.   python scripts/bench_train.py [--batch 8] [--out profiles/r02_train_mlp.json]"""
import argparse
import copy
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

LAYERS = {   # KITTI IA-SSD (tools/cfgs/kitti_models/IA-SSD.yaml:45-65): source points, centres, feature channels
    "L0": dict(n=16384, m=4096, c=1, radii=[0.2, 0.8], nsamples=[16, 32], mlps=[[1, 16, 16, 32], [1, 32, 32, 64]]),
    "L1": dict(n=4096, m=1024, c=64, radii=[0.8, 1.6], nsamples=[16, 32], mlps=[[64, 64, 64, 128], [64, 64, 96, 128]]),
    "L2": dict(n=1024, m=512, c=128, radii=[1.6, 4.8], nsamples=[16, 32], mlps=[[128, 128, 128, 256], [128, 128, 256, 256]]),
    "L5": dict(n=256, m=256, c=256, radii=[4.8, 6.4], nsamples=[16, 32], mlps=[[256, 256, 256, 512], [256, 256, 512, 1024]]),
}


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only-backbone", action="store_true")
    a = ap.parse_args()
    from spsnet_b200 import pointnet2_modules as pm
    from spsnet_b200 import pointnet2_utils as pu
    from spsnet_b200 import scenes

    try:
        import importlib
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            refm = importlib.import_module("pcdet.ops.pointnet2.pointnet2_batch.pointnet2_modules")
    except Exception as e:  # pragma: no cover
        print("reference modules unavailable:", e)
        refm = None
    B = a.batch
    rows = []
    for name, c in ({} if a.only_backbone else LAYERS).items():
        torch.manual_seed(0)
        base = scenes.make_batch(5, B, 16384)[:, :, :3]
        xyz = torch.from_numpy(np.ascontiguousarray(base[:, :c["n"]])).cuda()
        if c["n"] < 16384:   # deeper layers see FPS-thinned clouds: sample the sources like the backbone does
            full = torch.from_numpy(np.ascontiguousarray(base)).cuda()
            xyz = pu.gather_rows(full, pu.furthest_point_sample(full, c["n"]))
        new_xyz = pu.gather_rows(xyz, pu.furthest_point_sample(xyz, c["m"])) if c["m"] < c["n"] else xyz.clone()
        feats = torch.randn(B, c["c"], c["n"], device="cuda")
        ours = pm.PointnetSAModuleMSG(npoint=c["m"], radii=c["radii"], nsamples=c["nsamples"], mlps=copy.deepcopy(c["mlps"]), use_xyz=True).cuda().train()
        cout = sum(m[-1] for m in c["mlps"])
        gout = torch.randn(B, cout, c["m"], device="cuda")
        arms = [("fused", ours, "1"), ("composed", copy.deepcopy(ours), "0")]
        if refm is not None:
            ref = refm.PointnetSAModuleMSG(npoint=c["m"], radii=c["radii"], nsamples=c["nsamples"], mlps=copy.deepcopy(c["mlps"]), use_xyz=True).cuda().train()
            ref.load_state_dict(ours.state_dict())
            arms.append(("reference", ref, "0"))
        row = {"layer": name, "batch": B, "n": c["n"], "m": c["m"], "grouped_rows": B * c["m"] * sum(c["nsamples"])}
        for arm, mod, flag in arms:
            os.environ["SPSK_TRAIN_FUSED"] = flag

            def fwd(mod=mod):
                f = feats.clone().requires_grad_(True)
                return mod(xyz, f, new_xyz)[1], f

            def fwd_bwd(mod=mod):
                out, f = fwd(mod)
                mod.zero_grad(set_to_none=True)
                out.backward(gout)

            t_f = timed(lambda: fwd())
            t_fb = timed(fwd_bwd)
            torch.cuda.synchronize()
            torch.cuda.reset_peak_memory_stats()
            base_mem = torch.cuda.memory_allocated()
            fwd_bwd()
            torch.cuda.synchronize()
            row[arm] = {"fwd_ms": round(t_f, 3), "fwd_bwd_ms": round(t_fb, 3), "peak_mb": round((torch.cuda.max_memory_allocated() - base_mem) / 2 ** 20, 1)}
        rows.append(row)
        print(json.dumps(row), flush=True)
    # the whole IA-SSD SA stack in train(): one forward + backward per step, peak memory of a step
    from spsnet_b200 import backbone as bb
    from spsnet_b200.configs import kitti_iassd_cfg

    torch.cuda.empty_cache()   # the per-layer arms above leave multi-GB blocks cached; start the stack from a clean allocator
    torch.manual_seed(0)
    net = bb.IASSD_Backbone(kitti_iassd_cfg(), num_class=3, input_channels=4).cuda().train()
    pts = torch.from_numpy(np.ascontiguousarray(scenes.to_points(scenes.make_batch(7, B, 16384)))).cuda()
    row = {"layer": "backbone (KITTI IA-SSD SA stack)", "batch": B, "n": 16384}
    for arm, flag in (("fused", "1"), ("composed", "0")):
        os.environ["SPSK_TRAIN_FUSED"] = flag
        mod = copy.deepcopy(net)

        def step(mod=mod):
            out = mod({"batch_size": B, "points": pts})
            loss = out["centers_features"].square().mean() + out["ctr_offsets"][:, 1:].square().mean()
            mod.zero_grad(set_to_none=True)
            loss.backward()

        def fwd_only(mod=mod):
            return mod({"batch_size": B, "points": pts})["centers_features"]

        t_f = timed(fwd_only, reps=5)
        t_fb = timed(step, reps=5)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        base_mem = torch.cuda.memory_allocated()
        step()
        torch.cuda.synchronize()
        row[arm] = {"fwd_ms": round(t_f, 3), "fwd_bwd_ms": round(t_fb, 3), "peak_mb": round((torch.cuda.max_memory_allocated() - base_mem) / 2 ** 20, 1)}
        del mod
    rows.append(row)
    print(json.dumps(row), flush=True)
    if a.out:
        Path(a.out).write_text(json.dumps({"device": torch.cuda.get_device_name(0), "rows": rows}, indent=1))


if __name__ == "__main__":
    main()

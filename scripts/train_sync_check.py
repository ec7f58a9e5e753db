#!/usr/bin/env python
"""SyncBatchNorm on the fused training path, on real GPUs over NCCL (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/train_sync_check.py

Every rank holds 2 scenes of one 2*world-scene batch.  The module (KITTI layer-1 widths, converted with
nn.SyncBatchNorm.convert_sync_batchnorm like the reference's tools/train.py:122-123) runs its grouped MLPs
  (a) fused   -- statistics passes + one all-reduce of 2C+1 doubles per layer + pooled pass (spsnet_b200/train_fused.py),
  (b) composed with torch's own SyncBatchNorm (SPSK_TRAIN_FUSED=0: the reference's path),
and both are compared with (c) plain BatchNorm2d on the FULL batch in one process: pooled features, running statistics,
parameter gradients (summed over ranks).  Rank 0 prints one JSON line; exit code 1 on a miss."""
import copy
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from spsnet_b200 import pointnet2_modules as pm
    from spsnet_b200 import pointnet2_utils as pu
    from spsnet_b200 import scenes

    per, n, m = 2, 1024, 256
    B = per * world
    torch.manual_seed(0)
    plain = pm.PointnetSAModuleMSG(npoint=m, radii=[0.8, 1.6], nsamples=[16, 32], mlps=[[64, 64, 64, 128], [64, 64, 96, 128]], use_xyz=True)
    for mod in plain.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.uniform_(-0.2, 0.2)
    plain = plain.cuda().train()
    xyz = torch.from_numpy(np.ascontiguousarray(scenes.make_batch(91, B, n)[:, :, :3])).cuda()
    new_xyz = pu.gather_rows(xyz, pu.furthest_point_sample(xyz, m))
    feats = torch.randn(B, 64, n, device="cuda")
    out_shape_c = 256
    gout = torch.randn(B, out_shape_c, m, device="cuda")
    lo, hi = rank * per, (rank + 1) * per

    def run(mod, sl, fused):
        os.environ["SPSK_TRAIN_FUSED"] = "1" if fused else "0"
        f = feats[sl].clone().requires_grad_(True)
        c = new_xyz[sl].clone().requires_grad_(True)
        out, _ = mod._msg(xyz[sl].contiguous(), c, f)
        params = list(mod.parameters())
        g = torch.autograd.grad(out, [f, c] + params, gout[sl])
        bns = [b for b in mod.modules() if isinstance(b, (nn.BatchNorm2d, nn.SyncBatchNorm))]
        return out.detach(), g, [(b.running_mean.clone(), b.running_var.clone()) for b in bns]

    full = run(copy.deepcopy(plain), slice(0, B), fused=False)                       # (c) one process, whole batch
    sync_f = nn.SyncBatchNorm.convert_sync_batchnorm(copy.deepcopy(plain)).train()
    sync_c = nn.SyncBatchNorm.convert_sync_batchnorm(copy.deepcopy(plain)).train()
    res = {"world": world, "per_rank_scenes": per}
    ok = True
    for name, mod, fused in (("fused", sync_f, True), ("torch_syncbn", sync_c, False)):
        out, g, st = run(mod, slice(lo, hi), fused)
        e_out = rel(out, full[0][lo:hi])
        e_in = max(rel(g[0], full[1][0][lo:hi]), rel(g[1], full[1][1][lo:hi]))
        pg = [x.clone() for x in g[2:]]
        for x in pg:
            dist.all_reduce(x)
        e_pg = max(rel(a, b) for a, b in zip(pg, full[1][2:]))
        e_rm = max(rel(a[0], b[0]) for a, b in zip(st, full[2]))
        e_rv = max(rel(a[1], b[1]) for a, b in zip(st, full[2]))
        t = torch.tensor([e_out, e_in, e_pg, e_rm, e_rv], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_out, e_in, e_pg, e_rm, e_rv = [float(v) for v in t]
        res[name] = {"features": e_out, "input_grads": e_in, "param_grads_summed": e_pg, "running_mean": e_rm, "running_var": e_rv}
        ok = ok and e_out <= 1e-3 and e_in <= 5e-4 and e_pg <= 5e-4 and e_rm <= 1e-3 and e_rv <= 2e-3
    res["ok"] = bool(ok)
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

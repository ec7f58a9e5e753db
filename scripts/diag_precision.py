"""GPU diagnostic: error of the fused path vs the reference modules in fp64, next to the reference's own
fp32 (TF32 off) and stock (cudnn TF32 on) deviation from that truth.  python scripts/diag_precision.py"""
import copy, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
from helpers import rel_err
from test_gpu_modules import _sa_module, dev
from spsnet_b200 import scenes

def run(kind, n, seed):
    B = 2
    m, cin = _sa_module(kind, seed=seed); m = m.cuda()
    rng = np.random.default_rng(4 + seed)
    xyz = dev(np.ascontiguousarray(scenes.make_batch(60 + seed, B, n)[:, :, :3]))
    feats = dev(rng.standard_normal((B, cin, n)).astype(np.float32))
    cls = dev(scenes.make_cls_logits(19 + seed, B, n)) if kind in ("l2", "l3") else None
    with torch.no_grad():
        got = m(xyz, feats, cls)
        # truth: the composed (reference-style) path of the same module in float64 on the same indices
        m64 = copy.deepcopy(m).double()
        new_xyz = got[0].double()
        f64 = m64._msg_composed.__func__  # noqa
        import spsnet_b200.pointnet2_utils as pu
        outs = []
        for g, mlp in zip(m.groupers, m64.mlps):
            idx = pu.ball_query(g.radius, g.nsample, xyz, got[0].contiguous())
            gx = pu.grouping_operation(xyz.transpose(1, 2).contiguous(), idx).double() - new_xyz.transpose(1, 2).unsqueeze(-1)
            gf = pu.grouping_operation(feats, idx).double()
            outs.append(mlp(torch.cat([gx, gf], 1)).amax(-1))
        pooled = torch.cat(outs, 1)
        nf64 = m64.aggregation_layer(pooled)
        cls64 = m64.confidence_layers(nf64).transpose(1, 2) if m64.confidence_layers is not None else None
        res = {}
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            want = m._msg_composed(xyz, got[0], feats)
            nf = m.aggregation_layer(want)
            c = m.confidence_layers(nf).transpose(1, 2) if m.confidence_layers is not None else None
            res[tf32] = (rel_err(nf.cpu().numpy(), nf64.cpu().numpy()), rel_err(c.cpu().numpy(), cls64.cpu().numpy()) if c is not None else None)
        ours = (rel_err(got[1].cpu().numpy(), nf64.cpu().numpy()), rel_err(got[2].cpu().numpy(), cls64.cpu().numpy()) if cls64 is not None else None)
    print(f"{kind} n={n} seed={seed}: ours feat {ours[0]:.2e} cls {ours[1] if ours[1] is None else format(ours[1], '.2e')} | torch fp32 feat {res[False][0]:.2e} cls {res[False][1]} | "
          f"torch cudnn-TF32 feat {res[True][0]:.2e} cls {res[True][1]}  (cls range {float(cls64.abs().max()) if cls64 is not None else 0:.2f}, feat range {float(nf64.abs().max()):.2f})")

for kind, n in (("l1", 2048), ("l2", 1024)):
    for seed in (1, 2, 3):
        run(kind, n, seed)

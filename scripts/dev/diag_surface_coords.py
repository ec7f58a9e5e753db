import sys, numpy as np, torch
sys.path.insert(0, '.')
from spsnet_b200 import scenes, surface_feature as SF
torch.manual_seed(0)
fe = SF.FeatureExtraction().cuda().eval()
xyz = torch.from_numpy(np.ascontiguousarray(scenes.make_batch(0, 16, 16384)[:, :, :3])).cuda()
with torch.no_grad():
    out, idxs, ts = fe.fused_forward(xyz, return_idx=True)
for i, t in enumerate(ts):
    c = SF._as_ball_query_coords(t)          # (16, 16384, 3)
    for b in (0, 5):
        p = c[b]
        ext = (p.max(0)[0] - p.min(0)[0]).tolist()
        sd = p.std(0).tolist()
        q = torch.quantile(p, torch.tensor([0.01, 0.5, 0.99], device="cuda"), dim=0).t().tolist()
        zero = (p == 0).float().mean(0).tolist()
        print(f"unit {i} scene {b}: extent {[round(e,1) for e in ext]} std {[round(e,2) for e in sd]} q01/50/99 {[[round(v,2) for v in r] for r in q]} zeros {[round(z,2) for z in zero]}")

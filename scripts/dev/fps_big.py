import sys, numpy as np, torch
sys.path.insert(0, '.')
from spsnet_b200 import pointnet2_utils as pu, scenes
for B, N, M in [(4, 262144, 4096), (4, 200000, 1024), (4, 131072, 4096)]:
    xyz = torch.from_numpy(np.ascontiguousarray(scenes.make_batch(0, B, N, "waymo")[:, :, :3])).cuda()
    for _ in range(2): pu.furthest_point_sample(xyz, M)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pu.furthest_point_sample(xyz, M); e1.record(); torch.cuda.synchronize()
    print(f"fps B={B} N={N} M={M}: {e0.elapsed_time(e1):.2f} ms")

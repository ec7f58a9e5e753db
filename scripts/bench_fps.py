#!/usr/bin/env python
"""D-FPS micro-benchmark (KITTI / Waymo shapes) with the optional phase counters.  python scripts/bench_fps.py [--prof]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from spsnet_b200 import pointnet2_utils as pu, scenes
from spsnet_b200._lib import lib

def main():
    prof = "--prof" in sys.argv
    for B, N, M, kind in [(16, 16384, 4096, "kitti"), (16, 4096, 1024, "kitti"), (8, 65536, 16384, "waymo"), (8, 16384, 4096, "waymo")]:
        seed = 0 if "--bench-data" in sys.argv else 5
        xyz = torch.from_numpy(np.ascontiguousarray(scenes.make_batch(seed, B, max(N, 16384), kind)[:, :N, :3])).cuda().contiguous()
        for _ in range(2): pu.furthest_point_sample(xyz, M)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): pu.furthest_point_sample(xyz, M)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"fps {kind} B={B} N={N} M={M}: {ms:.3f} ms  {ms*1e3/(M-1):.3f} us/iter  {ms*1e6/(M-1)*1.965:.0f} cycles/iter")
        if prof and N <= 16384:
            buf = torch.zeros(6, dtype=torch.int64, device="cuda")
            lib.spsk_fps_set_profile(buf.data_ptr())
            pu.furthest_point_sample(xyz, M); torch.cuda.synchronize()
            lib.spsk_fps_set_profile(None)
            v = buf.cpu().numpy().astype(np.float64)
            it = B * (M - 1)
            print("   warp0 cycles/iter: " + "  ".join(f"{n}={x/it:.0f}" for n, x in zip(["query+bound", "buckets", "warp_argmax", "barrier", "block_argmax"], v[:5])) +
                  f"  sub-buckets visited/iter={v[5]/it:.1f} of {N//32}")

if __name__ == "__main__":
    main()

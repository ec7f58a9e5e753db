#!/usr/bin/env python
"""Micro-benchmark of the fused scale kernel (sa_mma.cu) on the IA-SSD KITTI shapes, B = 16, with the optional
per-role cycle counters.   python scripts/bench_sa_mma.py [names...] [--prof]"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch, ctypes as C
from spsnet_b200 import pointnet2_utils as pu, scenes
from spsnet_b200._lib import lib

# name: (N source pts, M centres, c_feat, nsample, radius, widths)
SHAPES = {
    "l0s1": (16384, 4096, 1, 16, 0.2, [16, 16, 32]), "l0s2": (16384, 4096, 1, 32, 0.8, [32, 32, 64]),
    "l1s1": (4096, 1024, 64, 16, 0.8, [64, 64, 128]), "l1s2": (4096, 1024, 64, 32, 1.6, [64, 96, 128]),
    "l2s1": (1024, 512, 128, 16, 1.6, [128, 128, 256]), "l2s2": (1024, 512, 128, 32, 4.8, [128, 256, 256]),
    "l5s1": (512, 256, 256, 16, 4.8, [256, 256, 512]), "l5s2": (512, 256, 256, 32, 6.4, [256, 512, 1024]),
}
NAMES = ["mma_total", "mma_wait_acc_empty", "mma_wait_w", "mma_wait_x", "prod_wait_stage", "prod_wait_hid", "epi_total", "epi_gather",
         "epi_wait_hid", "epi_work_hid", "epi_wait_pool", "epi_work_pool", "mma_issue", "mma_commit"]

def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    prof = "--prof" in sys.argv
    B = 16
    for name in (args or list(SHAPES)):
        N, M, cf, ns, radius, widths = SHAPES[name]
        rng = np.random.default_rng(1)
        xyz = torch.from_numpy(np.ascontiguousarray(scenes.make_batch(3, B, max(N, 1024))[:, :N, :3])).cuda().contiguous()
        feats = torch.randn(B, cf, N, device="cuda")
        sel = torch.stack([torch.randperm(N, device="cuda")[:M] for _ in range(B)]).int()
        new_xyz = pu.gather_rows(xyz, sel)
        idx = pu.ball_query(radius, ns, xyz, new_xyz)
        cin = cf + 3
        chain = []
        for w in widths:
            chain.append((torch.randn(cin, w, device="cuda") / np.sqrt(cin), torch.randn(w, device="cuda") * 0.1, True)); cin = w
        pk = pu.MmaChain(chain, cf, True)
        twin = pu.make_twin(feats, pk.cpad8) if (cf and not pk.split) else None
        out16 = torch.empty(B * M, (widths[-1] + 15) // 16 * 16, dtype=torch.float16, device="cuda")
        run = lambda: pu.sa_mma_forward(xyz=xyz, new_xyz=new_xyz, idx=idx, chain=pk, twin=twin, features=feats if pk.split else None, out16=out16)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        rows = B * M * ns
        flops = 2.0 * rows * sum(a * b for a, b in zip([cf + 3] + widths[:-1], widths))
        print(f"{name}: {us:8.1f} us  rows={rows} tiles={rows//128} ctas/SM={pk.ctas_per_sm} res={pk.resident} stages={pk.nstages} split={pk.split} "
              f"{flops/us/1e6:7.1f} TFLOP/s  {us*1e3/(rows/128)*148*pk.ctas_per_sm/ max(1,pk.ctas_per_sm):.0f} ns/tile/SM-slot")
        if prof:
            buf = torch.zeros(14, dtype=torch.int64, device="cuda")
            lib.spsk_sa_mma_set_profile(buf.data_ptr())
            run(); torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for _ in range(5): run()
            p1.record(); torch.cuda.synchronize()
            print(f"   profiling-kernel time {p0.elapsed_time(p1) * 200:8.1f} us  (SPSK_SA_ABL={os.environ.get('SPSK_SA_ABL', '0')})")
            buf.zero_(); run(); torch.cuda.synchronize()
            lib.spsk_sa_mma_set_profile(None)
            v = buf.cpu().numpy().astype(np.float64)
            ncta = min(rows // 256, 148) if pk.pair else min(rows // 128, 148 * pk.ctas_per_sm)   # persistent CTAs: one resident set (grid_mult = 1)
            print("   per-CTA kcycles: " + "  ".join(f"{n}={x/ncta/1e3:.1f}" for n, x in zip(NAMES, v)))

if __name__ == "__main__":
    main()

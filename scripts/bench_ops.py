#!/usr/bin/env python
"""Op micro-benchmark sweep (BASELINE.json configs[4]): FPS N = 16k..256k x npoint 512..16k; ball_query + group_points
nsample 16/32/64 x radius 0.2..4.8 -- libspsk vs the reference's own CUDA ops (oracle/_ref) on the same inputs, with
bit-exact index checks.   python scripts/bench_ops.py [--out profiles/r01_ops_sweep.json]"""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
import numpy as np, torch
from spsnet_b200 import pointnet2_utils as pu, scenes

try:
    import importlib, warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = importlib.import_module("pcdet.ops.pointnet2.pointnet2_batch.pointnet2_utils")
except Exception as e:  # pragma: no cover
    print("reference ops unavailable:", e); ref = None


def timeit(fn, reps, warm=3):
    """>= 3 warm-ups (kernel attributes, allocator blocks, caches), then `reps` timed launches between two CUDA events."""
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def _stream():
    return torch.cuda.current_stream().cuda_stream


def ours_group_into(feats, idx, out):
    """spsk_group_points straight through the C-ABI into a PRE-ALLOCATED output (no allocator in the timed region)."""
    from spsnet_b200._lib import check, lib
    B, C, N = feats.shape
    _, M, ns = idx.shape
    check(lib.spsk_group_points(B, C, N, M, ns, feats.data_ptr(), idx.data_ptr(), out.data_ptr(), _stream()), "group_points")
    return out


def ref_group_into(feats, idx, out):
    """the reference's pybind entry point (group_points_wrapper, group_points.cpp:30) into the same pre-allocated output."""
    B, C, N = feats.shape
    _, M, ns = idx.shape
    ref.pointnet2.group_points_wrapper(B, C, N, M, ns, feats, idx, out)
    return out


def main():
    out_path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    rows = []
    B = 4
    for N in (16384, 32768, 65536, 131072, 262144):
        kind = "kitti" if N == 16384 else "waymo"
        base = scenes.make_batch(11, B, min(N, 65536), kind)[:, :, :3]
        reps_n = (N + base.shape[1] - 1) // base.shape[1]
        xyz_np = np.concatenate([base + 0.013 * k for k in range(reps_n)], axis=1)[:, :N]
        xyz = torch.from_numpy(np.ascontiguousarray(xyz_np)).cuda()
        for M in (512, 1024, 4096, 16384):
            if M >= N: continue
            ms, idx = timeit(lambda: pu.furthest_point_sample(xyz, M), 2, warm=3 if N * M <= 65536 * 16384 else 1)
            r = {"op": "fps", "B": B, "N": N, "npoint": M, "ours_ms": ms, "us_per_iter": ms * 1e3 / (M - 1)}
            if ref is not None:
                rms, ridx = timeit(lambda: ref.furthest_point_sample(xyz, M), 1, warm=1)
                r.update(ref_ms=rms, speedup=rms / ms, bit_exact=bool(torch.equal(idx, ridx)))
            rows.append(r); print(r, flush=True)
    # ball query + grouping: KITTI layer shapes (N source points, M = N/4 centres)
    for N, M in ((16384, 4096), (4096, 1024)):
        xyz = torch.from_numpy(np.ascontiguousarray(scenes.make_batch(3, 16, 16384)[:, :N, :3])).cuda()
        sel = pu.furthest_point_sample(xyz, M)
        new_xyz = pu.gather_rows(xyz, sel)
        feats = torch.randn(16, 64, N, device="cuda")
        for radius in (0.2, 0.8, 1.6, 4.8):
            for ns in (16, 32, 64):
                ms, idx = timeit(lambda: pu.ball_query_msg([radius], [ns], xyz, new_xyz)[0], 5)
                gout = torch.empty((16, 64, M, ns), dtype=torch.float32, device="cuda")
                gms, g = timeit(lambda: ours_group_into(feats, idx, gout), 5)
                g = g.clone()
                r = {"op": "ball_query+group", "B": 16, "N": N, "M": M, "radius": radius, "nsample": ns, "ours_bq_ms": ms, "ours_group_ms": gms,
                     "group_GBps": (g.numel() * 4 + idx.numel() * 4) / gms / 1e6}
                if ref is not None:
                    rms, ridx = timeit(lambda: ref.ball_query(radius, ns, xyz, new_xyz), 3)
                    rgms, rg = timeit(lambda: ref_group_into(feats, ridx, gout), 3)
                    r.update(ref_bq_ms=rms, ref_group_ms=rgms, bq_speedup=rms / ms, bit_exact=bool(torch.equal(idx, ridx) and torch.equal(g, rg)))
                rows.append(r); print(r, flush=True)
    if out_path:
        Path(out_path).write_text(json.dumps({"device": torch.cuda.get_device_name(0), "rows": rows}, indent=1))


if __name__ == "__main__":
    main()

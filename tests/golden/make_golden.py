#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING the reference's own CUDA ops (oracle/_ref = pointnet2_batch rebuilt
unmodified for sm_100a by oracle/build_ref.sh) on a B200.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so these fixtures are the pin for the CPU oracle: tests/test_oracle_cpu.py replays the
same seeded inputs through oracle/oracle.c and demands bit-equality.

    gpurun -- python tests/golden/make_golden.py gpurun_out/golden      (then copy *.npz into tests/golden/)

Inputs are regenerated from seeds by `spsnet_b200.scenes`, so only the OUTPUTS are stored.
"""
import sys
import warnings
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
from spsnet_b200 import scenes  # noqa: E402

# shared with tests/test_oracle_cpu.py
FPS_CASES = [(2, 7, 5, 0), (2, 100, 64, 1), (2, 1000, 300, 2), (2, 1500, 700, 3), (2, 4096, 1024, 4), (2, 16384, 4096, 5)]
BQ_CASES = [(2, 1000, 77, 2.0, 16, 6), (2, 4096, 512, 0.8, 16, 7), (2, 4096, 512, 1.6, 32, 8), (1, 16384, 256, 0.2, 16, 9),
            (2, 1024, 256, 4.8, 32, 10)]
NN_CASES = [(2, 1000, 256, 8, 11), (1, 300, 1200, 4, 12)]
TOPK_CASES = [(2, 1024, 512, 13), (2, 512, 256, 14), (1, 4096, 2048, 15)]


def case_xyz(b, n, seed):
    return np.ascontiguousarray(scenes.make_batch(1000 + seed, b, n)[:, :, :3])


def lattice_xyz(b, n, seed):
    return np.random.default_rng(seed).integers(0, 6, (b, n, 3)).astype(np.float32)


def centres_for(xyz, m, seed):
    rng = np.random.default_rng(seed)
    sel = np.stack([rng.choice(xyz.shape[1], m, replace=False) for _ in range(xyz.shape[0])])
    c = np.ascontiguousarray(np.take_along_axis(xyz, sel[..., None], axis=1))
    c[:, -1] += 500.0  # empty ball
    return c


def main(out_dir):
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from pcdet.ops.pointnet2.pointnet2_batch import pointnet2_utils as ru

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).cuda()

    g = {}
    for b, n, m, seed in FPS_CASES:
        g[f"fps_{b}_{n}_{m}"] = ru.furthest_point_sample(dev(case_xyz(b, n, seed)), m).cpu().numpy()
        g[f"fpslat_{b}_{n}_{m}"] = ru.furthest_point_sample(dev(lattice_xyz(b, n, seed)), m).cpu().numpy()
    rng = np.random.default_rng(99)
    f = rng.standard_normal((2, 600, 6)).astype(np.float32)
    d = ((f[:, :, None, :] - f[:, None, :, :]) ** 2).sum(-1).astype(np.float32)
    g["fpsdist_2_600_200"] = ru.furthest_point_sample_with_dist(dev(d), 200).cpu().numpy()
    for b, n, m, r, ns, seed in BQ_CASES:
        xyz = case_xyz(b, n, seed)
        c = centres_for(xyz, m, seed)
        g[f"bq_{b}_{n}_{m}_{r}_{ns}"] = ru.ball_query(r, ns, dev(xyz), dev(c)).cpu().numpy()
        g[f"bqd_{b}_{n}_{m}_{r}_{ns}"] = ru.ball_query_dilated(r, 0.0, ns, dev(xyz), dev(c)).cpu().numpy()
        g[f"bqd2_{b}_{n}_{m}_{r}_{ns}"] = ru.ball_query_dilated(r, r * 0.5, ns, dev(xyz), dev(c)).cpu().numpy()
    for b, n, m, c, seed in NN_CASES:
        unknown = case_xyz(b, n, seed)
        known = case_xyz(b, m, seed + 100)
        known[:, 3] = known[:, 1]
        dist, idx = ru.three_nn(dev(unknown), dev(known))
        g[f"nn_dist_{b}_{n}_{m}"] = dist.cpu().numpy()
        g[f"nn_idx_{b}_{n}_{m}"] = idx.cpu().numpy()
        feats = np.random.default_rng(seed).standard_normal((b, c, m)).astype(np.float32)
        w = np.random.default_rng(seed + 1).uniform(0, 1, (b, n, 3)).astype(np.float32)
        g[f"interp_{b}_{n}_{m}"] = ru.three_interpolate(dev(feats), idx, dev(w)).cpu().numpy()
    for b, n, k, seed in TOPK_CASES:  # the torch-op chains of pointnet2_modules.py:287-303, run by torch on the GPU
        cls = scenes.make_cls_logits(seed, b, n)
        stds = scenes.make_stds(seed + 1, b, n)
        s1 = torch.sigmoid(dev(cls).max(dim=-1)[0])
        s2 = s1 * (1 - torch.sigmoid(dev(stds) / 8 - 3))
        for tag, s in (("ctr", s1), ("sss", s2)):
            v, i = torch.topk(s, k, dim=-1)
            g[f"topk_{tag}_score_{b}_{n}_{k}"] = s.cpu().numpy()
            g[f"topk_{tag}_idx_{b}_{n}_{k}"] = i.int().cpu().numpy()
    np.savez_compressed(out / "reference_ops_b200.npz", **g)
    print("wrote", out / "reference_ops_b200.npz", len(g), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")

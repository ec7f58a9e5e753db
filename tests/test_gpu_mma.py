"""-m gpu tests of the tcgen05 tensor-core scale kernel (sa_mma.cu) against the exact-fp32 FFMA path of the same
library and a float64 torch reference.  Tolerance: 1e-3 of the output range (BASELINE.json north_star);
the measured error is printed so it lands in the logs."""
import zlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import REL_TOL, rel_err  # noqa: E402
from spsnet_b200 import scenes  # noqa: E402

SCALES = {  # name: (c_feat, nsample, radius, widths)
    "l0s1": (1, 16, 0.2, [16, 16, 32]), "l0s2": (1, 32, 0.8, [32, 32, 64]),
    "l1s1": (64, 16, 0.8, [64, 64, 128]), "l1s2": (64, 32, 1.6, [64, 96, 128]),
    "l2s1": (128, 16, 1.6, [128, 128, 256]), "l2s2": (128, 32, 4.8, [128, 256, 256]),
    "l5s1": (256, 16, 4.8, [256, 256, 512]), "l5s2": (256, 32, 6.4, [256, 512, 1024]),
    "odd": (5, 8, 2.0, [24, 40]), "one": (3, 64, 3.0, [48]), "nofeat": (0, 16, 1.0, [16, 32]),
}


def _chain(c_feat, widths, seed):
    g = torch.Generator().manual_seed(seed)
    cin = c_feat + 3
    chain = []
    for w in widths:
        wt = (torch.randn(cin, w, generator=g) * (1.3 / np.sqrt(cin))).cuda()
        b = (torch.randn(w, generator=g) * 0.1).cuda()
        chain.append((wt, b, True))
        cin = w
    return chain


@pytest.mark.parametrize("name", list(SCALES))
@pytest.mark.parametrize("B,N,M", [(2, 1024, 200), (1, 700, 37)])
def test_mma_scale_vs_fp32(oracle, name, B, N, M):
    from spsnet_b200 import pointnet2_utils as pu

    c_feat, ns, radius, widths = SCALES[name]
    rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000)  # str hash() is salted per process
    xyz_np = np.ascontiguousarray(scenes.make_batch(3, B, N)[:, :, :3])
    xyz = torch.from_numpy(xyz_np).cuda()
    feats = torch.from_numpy(rng.standard_normal((B, c_feat, N)).astype(np.float32)).cuda() if c_feat else None
    sel = torch.from_numpy(np.stack([rng.choice(N, M, replace=False) for _ in range(B)]).astype(np.int32)).cuda()
    new_xyz = pu.gather_rows(xyz, sel)
    idx = pu.ball_query(radius, ns, xyz, new_xyz)
    chain = _chain(c_feat, widths, seed=len(name))
    cout = widths[-1]
    # exact-fp32 path of the same library
    ref = torch.zeros((B, cout + 5, M), device="cuda")
    rows = None
    for li, (wt, b, relu) in enumerate(chain):
        last = li == len(chain) - 1
        rows = pu.grouped_linear(xyz=xyz, new_xyz=new_xyz, features=feats, idx=idx, use_xyz=True, in_rows=rows, wt=wt, bias=b,
                                 relu=relu, pool=1 if last else 0, out_pooled=ref if last else None, co_off=3)
    # tensor-core path
    packed = pu.MmaChain(chain, c_feat, True)
    assert packed.ok
    twin = pu.make_twin(feats, packed.cpad8) if c_feat else None
    got = torch.full((B, cout + 5, M), -7.0, device="cuda")
    pu.sa_mma_forward(xyz=xyz, new_xyz=new_xyz, twin=twin, idx=idx, use_xyz=True, chain=packed, out_pooled=got, co_off=3)
    torch.cuda.synchronize()
    g, r = got.cpu().numpy(), ref.cpu().numpy()
    assert np.all(g[:, :3] == -7.0) and np.all(g[:, 3 + cout:] == -7.0), "wrote outside its channel window"
    e = rel_err(g[:, 3:3 + cout], r[:, 3:3 + cout])
    print(f"[mma] {name} B={B} N={N} M={M}: rel err vs fp32 path = {e:.2e}")
    assert e <= REL_TOL, f"{name}: tensor-core path error {e:.3e} > {REL_TOL}"


def test_make_twin():
    from spsnet_b200 import pointnet2_utils as pu

    f = torch.randn(3, 13, 777, device="cuda")
    t = pu.make_twin(f, 16)
    want = torch.zeros(3, 777, 16, device="cuda", dtype=torch.float16)
    want[:, :, :13] = f.transpose(1, 2).half()
    assert torch.equal(t, want)


def test_module_uses_mma_and_matches_ffma(monkeypatch):
    """The module-level fused path with SPSK_MLP=mma vs =ffma on a KITTI layer-2 shaped module."""
    from test_gpu_modules import _sa_module

    m, cin = _sa_module("l2", seed=3)
    m = m.cuda()
    rng = np.random.default_rng(8)
    xyz = torch.from_numpy(np.ascontiguousarray(scenes.make_batch(11, 2, 1024)[:, :, :3])).cuda()
    feats = torch.from_numpy(rng.standard_normal((2, cin, 1024)).astype(np.float32)).cuda()
    cls = torch.from_numpy(scenes.make_cls_logits(4, 2, 1024)).cuda()
    from spsnet_b200 import _lib

    with torch.no_grad():
        monkeypatch.setenv("SPSK_MLP", "ffma")
        a = m(xyz, feats, cls)
        monkeypatch.setenv("SPSK_MLP", "mma")
        n0 = _lib.lib.spsk_launch_count()
        b = m(xyz, feats, cls)
        n1 = _lib.lib.spsk_launch_count()
    assert torch.equal(a[3], b[3]) and torch.equal(a[0], b[0])
    e = rel_err(b[1].cpu().numpy(), a[1].cpu().numpy())
    print(f"[mma] module l2: rel err mma vs ffma = {e:.2e}, launches {n1 - n0}")
    assert e <= REL_TOL

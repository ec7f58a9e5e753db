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
    for split in ((None, False) if pu.MmaChain(chain, c_feat, True).split else (None,)):
        packed = pu.MmaChain(chain, c_feat, True, split=split)
        assert packed.ok
        twin = pu.make_twin(feats, packed.cpad8) if (c_feat and not packed.split) else None
        got = torch.full((B, cout + 5, M), -7.0, device="cuda")
        ld16 = cout + 24
        got16 = torch.full((B * M, ld16), -7.0, device="cuda", dtype=torch.float16)
        pu.sa_mma_forward(xyz=xyz, new_xyz=new_xyz, idx=idx, chain=packed, twin=twin, features=feats if packed.split else None,
                          out_pooled=got, co_off=3, out16=got16, co16=8)
        torch.cuda.synchronize()
        g, r = got.cpu().numpy(), ref.cpu().numpy()
        assert np.all(g[:, :3] == -7.0) and np.all(g[:, 3 + cout:] == -7.0), "wrote outside its channel window"
        e = rel_err(g[:, 3:3 + cout], r[:, 3:3 + cout])
        g16 = got16.float().cpu().numpy().reshape(B, M, ld16)
        assert np.all(g16[:, :, :8] == -7.0) and np.all(g16[:, :, 8 + cout:] == -7.0), "fp16 output wrote outside its column window"
        e16 = rel_err(g16[:, :, 8:8 + cout].transpose(0, 2, 1), r[:, 3:3 + cout])
        print(f"[mma] {name} B={B} N={N} M={M} split={packed.split} ctas/SM={packed.ctas_per_sm} resident={packed.resident}: "
              f"rel err vs fp32 path = {e:.2e} (fp16 rows {e16:.2e})")
        tol = 2e-5 if packed.split else REL_TOL
        assert e <= tol, f"{name}: tensor-core path error {e:.3e} > {tol}"
        assert e16 <= tol + 6e-4, f"{name}: fp16 point-major output error {e16:.3e}"


@pytest.mark.parametrize("rows,c_in,c_out,relu", [(4096, 1536, 512, True), (1000, 96, 64, True), (777, 256, 3, False),
                                                  (130, 128, 128, True), (65536, 96, 64, True), (64, 16, 200, False)])
def test_pw_mma_vs_fp32(rows, c_in, c_out, relu):
    """Point-wise tensor-core GEMM (pw_mma.cu) against torch fp64 on the fp16-rounded operands (exact up to fp32
    accumulation order) and against the un-rounded fp32 layer (1e-3 of the range)."""
    from spsnet_b200 import pointnet2_utils as pu

    g = torch.Generator().manual_seed(rows + c_in)
    B = 2 if rows % 2 == 0 else 1
    M = rows // B
    ldx = (c_in + 15) // 16 * 16 + 16
    x = torch.zeros(rows, ldx)
    x[:, :c_in] = torch.randn(rows, c_in, generator=g)
    wt = torch.randn(c_in, c_out, generator=g) * (1.3 / np.sqrt(c_in))
    b = torch.randn(c_out, generator=g) * 0.1
    layer = pu.PwLayer(wt.cuda(), b.cuda(), relu)
    x16 = x.half().cuda()
    out_cm = torch.full((B, c_out + 4, M), -7.0, device="cuda")
    out_pm = torch.full((B, M, c_out), -7.0, device="cuda")
    _, out16, _ = pu.pw_mma_forward(x16, layer, out_cm=out_cm, m=M, co_off=2, want16=True, out_pm=out_pm)
    torch.cuda.synchronize()
    act = (lambda t: t.clamp_min(0)) if relu else (lambda t: t)
    want_q = act(x16.double().cpu()[:, :c_in] @ wt.half().double() + b.double())
    want = act(x.double()[:, :c_in] @ wt.double() + b.double())
    got_pm = out_pm.cpu().double().reshape(rows, c_out)
    got_cm = out_cm.cpu().double()[:, 2:2 + c_out].permute(0, 2, 1).reshape(rows, c_out)
    assert torch.all(out_cm[:, :2] == -7.0) and torch.all(out_cm[:, 2 + c_out:] == -7.0)
    eq = rel_err(got_pm.numpy(), want_q.numpy())
    e = rel_err(got_pm.numpy(), want.numpy())
    print(f"[pw] rows={rows} {c_in}->{c_out}: vs rounded-operand fp64 {eq:.2e}, vs fp32 layer {e:.2e}")
    assert eq <= 1e-5 and e <= REL_TOL
    assert torch.equal(got_cm, got_pm)
    n16 = (c_out + 15) // 16 * 16
    assert out16.shape == (rows, n16)
    assert torch.equal(out16[:, :c_out].cpu(), out_pm.reshape(rows, c_out).half().cpu())
    assert torch.all(out16[:, c_out:] == 0)


@pytest.mark.parametrize("rows,c_in,c_out,relu", [(4096, 1536, 512, True), (1000, 96, 64, True), (515, 256, 3, False)])
def test_pw_mma_split_is_fp32_grade(rows, c_in, c_out, relu):
    """hi + lo fp16 operands (3 products): the tensor-core layer reproduces the fp32 layer to ~1e-6 of its range, and
    the [values | residuals] fp16 output carries the result to the same accuracy."""
    from spsnet_b200 import pointnet2_utils as pu

    g = torch.Generator().manual_seed(7 * rows + c_in)
    k = (c_in + 15) // 16 * 16
    x = torch.zeros(rows, k)
    x[:, :c_in] = torch.randn(rows, c_in, generator=g)
    xh = x.half()
    x16 = torch.cat([xh, (x - xh.float()).half()], dim=1).cuda().contiguous()
    wt = torch.randn(c_in, c_out, generator=g) * (1.3 / np.sqrt(c_in))
    b = torch.randn(c_out, generator=g) * 0.1
    layer = pu.PwLayer(wt.cuda(), b.cuda(), relu, split=True)
    out_pm = torch.full((1, rows, c_out), -7.0, device="cuda")
    _, out16, _ = pu.pw_mma_forward(x16, layer, xlo=k, want16_lo=True, out_pm=out_pm)
    torch.cuda.synchronize()
    want = x.double()[:, :c_in] @ wt.double() + b.double()
    if relu:
        want = want.clamp_min(0)
    e = rel_err(out_pm.cpu().double().reshape(rows, c_out).numpy(), want.numpy())
    n16 = layer.n16
    rec = out16[:, :n16].float() + out16[:, n16:].float()
    e16 = rel_err(rec[:, :c_out].cpu().double().numpy(), want.numpy())
    print(f"[pw split] rows={rows} {c_in}->{c_out}: vs fp64 {e:.2e}, hi+lo rows {e16:.2e}")
    assert e <= 2e-5 and e16 <= 2e-5
    assert torch.all(out16[:, c_out:n16] == 0) and torch.all(out16[:, n16 + c_out:] == 0)


@pytest.mark.parametrize("name", ["l2s2", "l5s1", "l5s2", "l1s2"])
def test_mma_pair_kernel_matches_single(name):
    """The CTA-pair kernel (tcgen05 cta_group::2, 256-row tiles; opt-in) reproduces the single-CTA kernel: same operands,
    same fp32 accumulation per output, so the pooled results agree to fp32 summation-order noise."""
    from spsnet_b200 import pointnet2_utils as pu

    c_feat, ns, radius, widths = SCALES[name]
    B, N, M = 3, 1500, 333
    rng = np.random.default_rng(21)
    xyz = torch.from_numpy(np.ascontiguousarray(scenes.make_batch(5, B, N)[:, :, :3])).cuda()
    feats = torch.from_numpy(rng.standard_normal((B, c_feat, N)).astype(np.float32)).cuda()
    sel = torch.from_numpy(np.stack([rng.choice(N, M, replace=False) for _ in range(B)]).astype(np.int32)).cuda()
    new_xyz = pu.gather_rows(xyz, sel)
    idx = pu.ball_query(radius, ns, xyz, new_xyz)
    chain = _chain(c_feat, widths, seed=3)
    outs = []
    for pair in (False, True):
        pk = pu.MmaChain(chain, c_feat, True, pair=pair)
        assert pk.ok and pk.pair == pair
        twin = pu.make_twin(feats, pk.cpad8)
        out = torch.full((B, widths[-1], M), -7.0, device="cuda")
        pu.sa_mma_forward(xyz=xyz, new_xyz=new_xyz, idx=idx, chain=pk, twin=twin, out_pooled=out)
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy())
    e = rel_err(outs[1], outs[0])
    print(f"[mma pair] {name}: pair vs single rel diff {e:.2e}")
    assert e <= 2e-6


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

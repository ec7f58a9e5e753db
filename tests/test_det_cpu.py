"""CPU tests of the head + NMS widening: the C oracle against golden vectors produced by the
reference's OWN CPU rotated IoU (tests/golden/iou3d_reference_cpu.npz, made by tests/golden/make_golden_iou3d.py),
geometric known answers, NMS properties, the decode restatement against plain torch ops, and the host-side packing of
the fused head (BN fold, stacked / block-diagonal layers, state_dict layout)."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest
import torch

GOLDEN = Path(__file__).parent / "golden" / "iou3d_reference_cpu.npz"


def _cases():
    spec = importlib.util.spec_from_file_location("make_golden_iou3d", Path(__file__).parent / "golden" / "make_golden_iou3d.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_golden_iou_matrix_bit_exact(oracle):
    from spsnet_b200 import scenes

    g, mg = np.load(GOLDEN), _cases()
    for na, nb, seed in mg.IOU_CASES:
        ab = scenes.make_boxes(seed, na + nb)
        got = oracle.boxes_matrix(ab[:na], ab[na:], "iou_bev")
        np.testing.assert_array_equal(got, g[f"iou_{na}_{nb}_{seed}"])


def test_golden_nms_keep_lists(oracle):
    from spsnet_b200 import scenes

    g, mg = np.load(GOLDEN), _cases()
    for n, thresh, seed in mg.NMS_CASES:
        np.testing.assert_array_equal(oracle.nms_sorted(scenes.make_boxes(seed, n), thresh), g[f"nms_{n}_{seed}"])


def test_iou_known_answers(oracle):
    a = np.array([[0, 0, 0, 4, 2, 1, 0.0]], np.float32)
    cases = [
        ([0, 0, 0, 4, 2, 1, 0.0], 8.0),                # identical
        ([2, 0, 0, 4, 2, 1, 0.0], 4.0),                # half shifted
        ([0, 0, 0, 2, 4, 1, np.pi / 2], 8.0),          # same rectangle described with a 90 degree heading
        ([10, 10, 0, 4, 2, 1, 0.3], 0.0),              # disjoint
        ([0, 0, 5, 2, 2, 1, np.pi / 4], 4.0),          # rotated square inside... clipped to |y|<=1: octagon-ish
    ]
    for b, area in cases[:4]:
        got = oracle.boxes_matrix(a, np.array([b], np.float32), "overlap")[0, 0]
        assert abs(got - area) < 1e-4, (b, got)
    # symmetric, bounded, self-IoU = 1
    from spsnet_b200 import scenes

    bx = scenes.make_boxes(3, 80)
    iou = oracle.boxes_matrix(bx, bx, "iou_bev")
    assert np.allclose(np.diag(iou), 1.0, atol=1e-5)
    assert np.allclose(iou, iou.T, atol=2e-5) and iou.min() >= 0 and iou.max() <= 1 + 1e-4
    # 3-D IoU: zero when heights are disjoint, equals the BEV-derived formula otherwise
    hi = bx.copy()
    hi[:, 2] += 100
    assert oracle.boxes_matrix(bx, hi, "iou3d").max() == 0
    i3 = oracle.boxes_matrix(bx, bx, "iou3d")
    assert np.allclose(np.diag(i3), 1.0, atol=1e-5) and (i3 <= iou + 1e-4).all()


def test_nms_properties(oracle):
    from spsnet_b200 import scenes

    for normal in (False, True):
        bx = scenes.make_boxes(21, 400)
        keep = oracle.nms_sorted(bx, 0.2, normal)
        assert keep[0] == 0 and np.all(np.diff(keep) > 0)
        if not normal:
            iou = oracle.boxes_matrix(bx, bx, "iou_bev")
            kk = iou[np.ix_(keep, keep)]
            assert (np.triu(kk, 1) <= 0.2).all()
            dead = np.setdiff1d(np.arange(400), keep)
            for j in dead:
                assert (iou[keep[keep < j], j] > 0.2).any()
        assert np.array_equal(oracle.nms_sorted(bx[keep], 0.2, normal), np.arange(keep.size))   # idempotent
    assert oracle.nms_sorted(np.zeros((0, 7), np.float32), 0.1).size == 0
    s = np.array([0.2, 0.9, 0.5], np.float32)
    far = np.array([[0, 0, 0, 1, 1, 1, 0], [10, 0, 0, 1, 1, 1, 0], [20, 0, 0, 1, 1, 1, 0]], np.float32)
    assert oracle.nms_gpu(far, s, 0.1).tolist() == [1, 2, 0]
    assert oracle.nms_gpu(far, s, 0.1, pre_maxsize=2).tolist() == [1, 2]


def test_decode_matches_torch_ops(oracle):
    """decode_bin_ori vs the torch-op chain of PointResidual_BinOri_Coder.decode_torch (box_coder_utils.py:288-319)
    written out with plain CPU torch ops (one rounding per op, like the reference on the GPU)."""
    rng = np.random.default_rng(0)
    n, bins = 500, 12
    enc = torch.from_numpy(rng.normal(0, 0.5, (n, 6 + 2 * bins)).astype(np.float32))
    pts = torch.from_numpy(rng.uniform(-40, 40, (n, 3)).astype(np.float32))
    cls = torch.from_numpy(rng.integers(1, 4, n))
    mean = torch.tensor([[3.9, 1.6, 1.56], [0.8, 0.6, 1.73], [1.76, 0.6, 1.73]])
    xt, yt, zt, dxt, dyt, dzt = torch.split(enc[..., :6], 1, dim=-1)
    xa, ya, za = torch.split(pts, 1, dim=-1)
    dxa, dya, dza = torch.split(mean[cls - 1], 1, dim=-1)
    diagonal = torch.sqrt(dxa ** 2 + dya ** 2)
    xg, yg, zg = xt * diagonal + xa, yt * diagonal + ya, zt * dza + za
    dxg, dyg, dzg = torch.exp(dxt) * dxa, torch.exp(dyt) * dya, torch.exp(dzt) * dza
    bin_inter = 2 * np.pi / bins
    _, bin_id = torch.max(enc[..., 6:6 + bins], dim=-1)
    bin_res = torch.sum(enc[..., 6 + bins:] * torch.nn.functional.one_hot(bin_id.long(), bins).float(), dim=-1)
    rg = bin_id.float() * bin_inter - np.pi + bin_inter / 2
    rg = (rg + bin_res * (bin_inter / 2)).unsqueeze(-1)
    want = torch.cat([xg, yg, zg, dxg, dyg, dzg, rg], dim=-1).numpy()
    got = oracle.decode_bin_ori(enc.numpy(), pts.numpy(), cls.numpy(), mean.numpy(), bins)
    np.testing.assert_allclose(got, want, rtol=2e-7, atol=1e-6)


def test_post_processing_oracle_semantics(oracle):
    from spsnet_b200 import scenes

    B, m = 3, 120
    rng = np.random.default_rng(4)
    cls = rng.normal(-1, 2, (B * m, 3)).astype(np.float32)
    boxes = np.concatenate([scenes.make_boxes(40 + b, m) for b in range(B)])
    out = oracle.post_processing(cls, boxes, B, 0.1, 0.1, 4096, 10)
    for b, d in enumerate(out):
        assert d["index"].size <= 10 and (d["pred_scores"] >= 0.1).all()
        assert (np.diff(d["pred_scores"]) <= 0).all()
        assert np.array_equal(d["pred_labels"], cls[b * m:(b + 1) * m][d["index"]].argmax(1) + 1)
    none = oracle.post_processing(np.full((m, 3), -20, np.float32), boxes[:m], 1, 0.1, 0.1, 4096, 10)
    assert none[0]["index"].size == 0


def test_fused_head_packing_matches_torch_stacks():
    """BN fold + stacked layer 0 + block-diagonal deeper layers == the two nn.Sequential stacks (fp64)."""
    from spsnet_b200 import backbone as bb
    from spsnet_b200 import dense_head as H

    torch.manual_seed(0)
    head = H.IASSD_Head(3, 64, H.Cfg({**H.KITTI_IASSD_HEAD, "CLS_FC": [32, 48], "REG_FC": [40, 24]}))
    bb.randomize_bn_stats(head, seed=1)
    head = head.double().eval()
    x = torch.randn(50, 64, dtype=torch.float64)
    dense, slices = H._fuse_dense([head.cls_center_layers, head.box_center_layers])
    assert [tuple(W.shape) for W, _, _ in dense] == [(72, 64), (72, 72), (33, 72)] and slices == [(0, 3), (3, 33)]
    y = x
    for W, b, relu in dense:
        y = y @ W.double().t() + b.double()
        y = y.relu() if relu else y
    with torch.no_grad():
        torch.testing.assert_close(y[:, 0:3], head.cls_center_layers(x), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(y[:, 3:33], head.box_center_layers(x), rtol=1e-5, atol=1e-6)
    with pytest.raises(RuntimeError):
        H._fuse_dense([head.cls_center_layers, torch.nn.Sequential(torch.nn.Linear(64, 3))])


def test_head_refuses_cpu_and_training():
    from spsnet_b200 import dense_head as H
    from spsnet_b200 import iou3d_nms_utils as U

    head = H.IASSD_Head(3, 512, H.kitti_iassd_head_cfg()).eval()
    assert head.box_coder.code_size == 30 and head.box_coder.bin_size == 12
    with pytest.raises(RuntimeError):
        head({"batch_size": 1, "centers_features": torch.zeros(4, 512), "centers": torch.zeros(4, 4)})
    with pytest.raises(RuntimeError):
        U.boxes_iou_bev(torch.zeros(2, 7), torch.zeros(2, 7))
    with pytest.raises(RuntimeError):
        U.boxes_bev_iou_cpu(torch.zeros(2, 7), torch.zeros(2, 7))
    with pytest.raises(NotImplementedError):
        H.IASSD_Head(3, 512, H.Cfg({**H.KITTI_IASSD_HEAD, "TARGET_CONFIG": {"BOX_CODER": "PointResidualCoder", "BOX_CODER_CONFIG": {}}}))


def test_head_state_dict_keys_match_reference_layout():
    from spsnet_b200 import dense_head as H

    head = H.IASSD_Head(3, 512, H.kitti_iassd_head_cfg())
    keys = list(head.state_dict().keys())
    want = []
    for stack in ("cls_center_layers", "box_center_layers"):
        for i in (0, 3):
            want += [f"{stack}.{i}.weight"] + [f"{stack}.{i + 1}.{k}" for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")]
        want += [f"{stack}.6.weight", f"{stack}.6.bias"]
    assert keys == want

"""GPU parity of the head + post-processing widening through the C-ABI:
rotated IoU / NMS against the C oracle and against the reference's own rebuilt CUDA kernels, the fused decode +
post-processing against the reference's torch chain, and the tensor-core head against the reference IASSD_Head."""
import copy

import numpy as np
import pytest
import torch

from helpers import assert_close

pytestmark = pytest.mark.gpu

IOU_CASES = [(1, 1, 0), (7, 5, 1), (64, 64, 2), (100, 37, 3), (257, 130, 4), (1000, 513, 5)]
NMS_CASES = [(1, 0.01), (2, 0.01), (63, 0.01), (64, 0.1), (65, 0.1), (256, 0.01), (300, 0.3), (1000, 0.7), (4096, 0.01), (5000, 0.25)]


def _boxes(seed, n):
    from spsnet_b200 import scenes

    return torch.from_numpy(scenes.make_boxes(seed, n)).cuda()


@pytest.mark.parametrize("na,nb,seed", IOU_CASES)
def test_iou_matrices_vs_oracle_and_reference(oracle, ref_det, na, nb, seed):
    from spsnet_b200 import iou3d_nms_utils as U

    ab = _boxes(seed, na + nb)
    a, b = ab[:na].contiguous(), ab[na:].contiguous()
    for name, fn in (("overlap", U.boxes_overlap_bev), ("iou_bev", U.boxes_iou_bev), ("iou3d", U.boxes_iou3d_gpu)):
        got = fn(a, b).cpu().numpy()
        want = oracle.boxes_matrix(a.cpu().numpy(), b.cpu().numpy(), name)
        # the oracle is FMA-free C, the kernels contract like the reference's nvcc build: last-bit differences only
        np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-5, err_msg=name)
        assert ((got > 0) == (want > 0)).mean() > 0.999
    if ref_det is not None:
        ref = torch.zeros(na, nb, device="cuda")
        ref_det.cuda.boxes_iou_bev_gpu(a, b, ref)
        got = U.boxes_iou_bev(a, b)
        assert torch.equal(got, ref), f"BEV IoU differs from the reference kernel in {(got != ref).sum().item()} entries, max {(got - ref).abs().max().item():.3e}"
        ref.zero_()
        ref_det.cuda.boxes_overlap_bev_gpu(a, b, ref)
        assert torch.equal(U.boxes_overlap_bev(a, b), ref)
        ref3 = ref_det.utils.boxes_iou3d_gpu(a, b)
        assert torch.equal(U.boxes_iou3d_gpu(a, b), ref3)


def test_iou_empty():
    from spsnet_b200 import iou3d_nms_utils as U

    a, e = _boxes(0, 5), torch.zeros((0, 7), device="cuda")
    assert U.boxes_iou_bev(a, e).shape == (5, 0) and U.boxes_iou_bev(e, a).shape == (0, 5)
    keep, _ = U.nms_gpu(e, torch.zeros(0, device="cuda"), 0.1)
    assert keep.numel() == 0


@pytest.mark.parametrize("n,thresh", NMS_CASES)
@pytest.mark.parametrize("normal", [False, True])
def test_nms_vs_oracle_and_reference(oracle, ref_det, n, thresh, normal):
    from spsnet_b200 import iou3d_nms_utils as U

    boxes = _boxes(100 + n, n)
    scores = torch.from_numpy(np.random.default_rng(n).permutation(n).astype(np.float32)).cuda() / n  # distinct
    fn = U.nms_normal_gpu if normal else U.nms_gpu
    keep, _ = fn(boxes, scores, thresh)
    want = oracle.nms_gpu(boxes.cpu().numpy(), scores.cpu().numpy(), thresh, normal=normal)
    assert keep.dtype == torch.int64
    assert np.array_equal(keep.cpu().numpy(), want), f"keep differs from the oracle: {keep.numel()} vs {want.size}"
    if ref_det is not None:
        rfn = ref_det.utils.nms_normal_gpu if normal else ref_det.utils.nms_gpu
        rkeep, _ = rfn(boxes, scores, thresh)
        assert torch.equal(keep, rkeep)
    if not normal and n >= 256:
        k2, _ = U.nms_gpu(boxes, scores, thresh, pre_maxsize=200)
        w2 = oracle.nms_gpu(boxes.cpu().numpy(), scores.cpu().numpy(), thresh, pre_maxsize=200)
        assert np.array_equal(k2.cpu().numpy(), w2)


def test_nms_batched_counts(oracle):
    from spsnet_b200 import iou3d_nms_utils as U

    B, N = 5, 300
    boxes = torch.stack([_boxes(200 + b, N) for b in range(B)])
    counts = torch.tensor([300, 0, 1, 64, 129], dtype=torch.int32, device="cuda")
    keep, num = U.nms_batch(boxes, 0.05, counts=counts)
    for b in range(B):
        c = int(counts[b])
        want = oracle.nms_sorted(boxes[b, :c].cpu().numpy(), 0.05)
        assert int(num[b]) == want.size
        assert np.array_equal(keep[b, :want.size].cpu().numpy(), want)


def test_nms_properties_large():
    """Size-independent properties at N = 4096 (the NMS_PRE_MAXSIZE of the KITTI config): survivors are mutually
    below the threshold, every suppressed box overlaps an earlier survivor, and NMS is idempotent."""
    from spsnet_b200 import iou3d_nms_utils as U

    n, thresh = 4096, 0.1
    boxes = _boxes(77, n)
    scores = torch.linspace(1, 0, n, device="cuda")
    keep, _ = U.nms_gpu(boxes, scores, thresh)
    kb = boxes[keep]
    iou = U.boxes_iou_bev(kb, kb)
    iou.fill_diagonal_(0)
    assert (torch.triu(iou, 1) <= thresh).all()
    dead = torch.ones(n, dtype=torch.bool, device="cuda")
    dead[keep] = False
    cross = U.boxes_iou_bev(boxes[dead], kb)          # suppressed x survivors
    earlier = keep[None, :] < torch.nonzero(dead).view(-1, 1)
    assert ((cross > thresh) & earlier).any(dim=1).all()
    keep2, _ = U.nms_gpu(kb, scores[keep], thresh)
    assert keep2.numel() == keep.numel()


def _ref_post(ref_det, cls, boxes, B, cfg):
    """detector3d_template.py:207-290 (class-agnostic branch) with the reference's own class_agnostic_nms + nms_gpu."""
    from spsnet_b200.backbone import Cfg

    m = cls.shape[0] // B
    out = []
    for b in range(B):
        box_preds = boxes[b * m:(b + 1) * m]
        cls_preds = torch.sigmoid(cls[b * m:(b + 1) * m])
        cls_preds, label_preds = torch.max(cls_preds, dim=-1)
        label_preds = label_preds + 1
        selected, selected_scores = ref_det.nms_utils.class_agnostic_nms(
            box_scores=cls_preds, box_preds=box_preds, nms_config=Cfg(cfg["NMS_CONFIG"]), score_thresh=cfg["SCORE_THRESH"])
        out.append({"pred_boxes": box_preds[selected], "pred_scores": selected_scores, "pred_labels": label_preds[selected]})
    return out


def _head_inputs(seed, B, m, num_class=3, code=30):
    """Distinct-score logits, box encodings that decode to clustered, overlapping boxes."""
    from spsnet_b200 import scenes

    rng = np.random.default_rng(seed)
    centers = np.zeros((B * m, 4), np.float32)
    reg = rng.normal(0, 0.15, (B * m, code)).astype(np.float32)
    for b in range(B):
        bx = scenes.make_boxes(seed * 31 + b, m, n_objects=max(2, m // 10))
        centers[b * m:(b + 1) * m, 0] = b
        centers[b * m:(b + 1) * m, 1:4] = bx[:, :3]
    reg[:, 6:18] = rng.normal(0, 1, (B * m, 12))
    cls = rng.normal(-1.0, 2.0, (B * m, num_class)).astype(np.float32)
    return torch.from_numpy(cls).cuda(), torch.from_numpy(reg).cuda(), torch.from_numpy(centers).cuda()


@pytest.mark.parametrize("B,m,seed", [(1, 1, 0), (2, 64, 1), (3, 100, 2), (16, 256, 3), (2, 1024, 4)])
def test_decode_and_postprocess_vs_reference_chain(oracle, ref_det, B, m, seed):
    from spsnet_b200 import dense_head as H

    cls, reg, centers = _head_inputs(seed, B, m)
    coder = H.PointResidual_BinOri_Coder(**H.KITTI_IASSD_HEAD["TARGET_CONFIG"]["BOX_CODER_CONFIG"])
    cfg = copy.deepcopy(H.KITTI_POST_PROCESSING)
    cfg["NMS_CONFIG"]["NMS_THRESH"] = 0.1
    boxes, scores, labels, det = H._detect_call(B, m, 3, 12, cls=cls, reg=reg, centers=centers[:, 1:4],
                                                mean_size=coder.mean_size, nms=H._nms_args(H.Cfg(cfg)))
    # decode vs oracle restatement of decode_torch (libm expf vs CUDA expf: last-bit differences)
    pred = cls.argmax(dim=1) + 1
    want = oracle.decode_bin_ori(reg.cpu().numpy(), centers[:, 1:4].cpu().numpy(), pred.cpu().numpy(), coder._mean_np)
    np.testing.assert_allclose(boxes.cpu().numpy(), want, rtol=3e-6, atol=3e-6)
    assert torch.equal(labels.long(), pred)
    # decode_torch drop-in (given classes) = fused decode
    assert torch.equal(coder.decode_torch(reg, centers[:, 1:4], pred), boxes)
    # post-processing vs oracle on the SAME decoded boxes
    o = oracle.post_processing(cls.cpu().numpy(), boxes.cpu().numpy(), B, cfg["SCORE_THRESH"], 0.1, 4096, 500)
    counts = det.count.tolist()
    for b in range(B):
        assert counts[b] == o[b]["index"].size, f"scene {b}: {counts[b]} vs {o[b]['index'].size} detections"
        assert np.array_equal(det.index[b, :counts[b]].cpu().numpy(), o[b]["index"])
        assert np.array_equal(det.labels[b, :counts[b]].cpu().numpy(), o[b]["pred_labels"])
        assert (det.index[b, counts[b]:] == -1).all() and (det.boxes[b, counts[b]:] == 0).all()
    if ref_det is not None:
        rcoder = ref_det.coder.PointResidual_BinOri_Coder(**H.KITTI_IASSD_HEAD["TARGET_CONFIG"]["BOX_CODER_CONFIG"])
        rboxes = rcoder.decode_torch(reg, centers[:, 1:4], pred)
        assert torch.equal(boxes, rboxes), f"decode differs from the reference coder: max {(boxes - rboxes).abs().max().item():.3e}"
        rp = _ref_post(ref_det, cls, rboxes, B, cfg)
        pd, _ = H.post_processing({"batch_size": B, "batch_cls_preds": cls, "batch_box_preds": boxes, "cls_preds_normalized": False}, cfg)
        for b in range(B):
            assert torch.equal(pd[b]["pred_boxes"], rp[b]["pred_boxes"]), f"scene {b}"
            assert torch.equal(pd[b]["pred_scores"], rp[b]["pred_scores"])
            assert torch.equal(pd[b]["pred_labels"], rp[b]["pred_labels"])


def test_class_agnostic_nms_dropin(ref_det):
    from spsnet_b200 import dense_head as H

    boxes = _boxes(9, 500)
    scores = torch.rand(500, device="cuda")
    cfg = H.Cfg(H.KITTI_POST_PROCESSING["NMS_CONFIG"])
    sel, sc = H.class_agnostic_nms(scores, boxes, cfg, score_thresh=0.3)
    assert (sc >= 0.3).all() and torch.equal(sc, scores[sel])
    if ref_det is not None:
        rsel, rsc = ref_det.nms_utils.class_agnostic_nms(scores, boxes, cfg, score_thresh=0.3)
        assert torch.equal(sel, rsel) and torch.equal(sc, rsc)


def _make_heads(ref_det, seed=0):
    from spsnet_b200 import backbone as bb
    from spsnet_b200 import dense_head as H

    torch.manual_seed(seed)
    head = H.IASSD_Head(3, 512, H.kitti_iassd_head_cfg(), post_process_cfg=H.KITTI_POST_PROCESSING)
    bb.randomize_bn_stats(head, seed=seed)
    head = head.cuda().eval()
    rhead = None
    if ref_det is not None:
        cfg = copy.deepcopy(H.KITTI_IASSD_HEAD)
        cfg["LOSS_CONFIG"] = {"LOSS_CLS": "WeightedCrossEntropy", "LOSS_REG": "WeightedSmoothL1Loss", "LOSS_INS": "WeightedCrossEntropy",
                              "CORNER_LOSS_REGULARIZATION": False, "CENTERNESS_REGULARIZATION": False,
                              "IOU3D_REGULARIZATION": False,
                              "LOSS_WEIGHTS": {"code_weights": [1.0] * 6}}
        rhead = ref_det.head.IASSD_Head(3, 512, H.Cfg(cfg)).cuda().eval()
        missing = rhead.load_state_dict(head.state_dict(), strict=False)
        assert not missing.unexpected_keys and all("loss" in k for k in missing.missing_keys), missing
    return head, rhead


def test_head_state_dict_layout():
    from spsnet_b200 import dense_head as H

    head = H.IASSD_Head(3, 512, H.kitti_iassd_head_cfg())
    keys = set(head.state_dict().keys())
    for k in ("cls_center_layers.0.weight", "cls_center_layers.1.running_mean", "cls_center_layers.6.bias",
              "box_center_layers.3.weight", "box_center_layers.6.weight"):
        assert k in keys
    assert head.box_center_layers[6].out_features == 30 and head.cls_center_layers[6].out_features == 3
    with pytest.raises(NotImplementedError):
        head.train().cuda()({"centers_features": torch.zeros(4, 512).cuda(), "centers": torch.zeros(4, 4).cuda(), "batch_size": 1})


@pytest.mark.parametrize("B,m", [(2, 64), (16, 256)])
def test_head_forward_vs_reference_head(oracle, ref_det, B, m):
    from spsnet_b200 import dense_head as H

    head, rhead = _make_heads(ref_det, seed=B)
    torch.manual_seed(5)
    feats = torch.randn(B * m, 512, device="cuda").relu() * 0.7
    _, _, centers = _head_inputs(11, B, m)
    bd = {"batch_size": B, "centers_features": feats, "centers": centers, "ctr_offsets": None, "centers_origin": None, "sa_ins_preds": []}
    out = head(dict(bd))
    # fp64 oracle of the two FC stacks
    cls_w = oracle.fc_stack(copy.deepcopy(head.cls_center_layers).cpu(), feats.cpu().numpy())
    reg_w = oracle.fc_stack(copy.deepcopy(head.box_center_layers).cpu(), feats.cpu().numpy())
    assert_close(out["batch_cls_preds"].cpu().numpy(), cls_w, 1e-3, "cls logits vs fp64 oracle")
    assert_close(head.forward_ret_dict["center_box_preds"].cpu().numpy(), reg_w, 1e-3, "box encodings vs fp64 oracle")
    assert out["batch_box_preds"].shape == (B * m, 7) and out["cls_preds_normalized"] is False
    assert torch.equal(out["batch_index"], centers[:, 0])
    # fused post-processing inside forward == post-processing of forward's own outputs through the generic entry
    pd, _ = H.post_processing(out, H.KITTI_POST_PROCESSING)
    out2 = dict(out)
    out2.pop("detections")
    pd2, _ = H.post_processing(out2, H.KITTI_POST_PROCESSING)
    for a, b in zip(pd, pd2):
        assert torch.equal(a["pred_boxes"], b["pred_boxes"]) and torch.equal(a["pred_labels"], b["pred_labels"])
    if rhead is not None:
        with torch.no_grad():
            old = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            try:
                rout = rhead({**bd, "ctr_offsets": centers, "centers_origin": centers})
            finally:
                torch.backends.cuda.matmul.allow_tf32 = old
        assert_close(out["batch_cls_preds"].cpu().numpy(), rout["batch_cls_preds"].cpu().numpy(), 1e-3, "cls vs reference head")
        assert_close(out["batch_box_preds"].cpu().numpy(), rout["batch_box_preds"].cpu().numpy(), 1e-3, "boxes vs reference head")
        # teacher-forced: the reference's post-processing on OUR logits/boxes gives the same detections
        rp = _ref_post(ref_det, out["batch_cls_preds"].contiguous(), out["batch_box_preds"], B, H.KITTI_POST_PROCESSING)
        for b in range(B):
            assert torch.equal(pd[b]["pred_boxes"], rp[b]["pred_boxes"])
            assert torch.equal(pd[b]["pred_scores"], rp[b]["pred_scores"])


def test_backbone_to_detections_end_to_end(oracle):
    """Backbone -> head -> post-processing on small scenes; the twin rows handed over by the backbone are used."""
    from helpers import make_backbone, small_sa_cfg
    from spsnet_b200 import dense_head as H
    from spsnet_b200 import scenes

    net = make_backbone(small_sa_cfg(), seed=3).cuda()
    head, _ = _make_heads(None, seed=4)
    B, N = 2, 2048
    pts = torch.from_numpy(scenes.to_points(scenes.make_batch(50, B, N))).cuda()
    with torch.no_grad():
        bd = net({"batch_size": B, "points": pts})
        assert getattr(bd["centers_features"], "_spsk_rows16", None) is not None
        out = head(bd)
    cls_w = oracle.fc_stack(copy.deepcopy(head.cls_center_layers).cpu(), bd["centers_features"].cpu().numpy())
    assert_close(out["batch_cls_preds"].cpu().numpy(), cls_w, 1e-3, "cls logits (twin rows) vs fp64 oracle")
    pd, _ = H.post_processing(out, H.KITTI_POST_PROCESSING)
    assert len(pd) == B
    for d in pd:
        assert d["pred_boxes"].shape[1] == 7 and d["pred_scores"].shape[0] == d["pred_boxes"].shape[0]
        assert (d["pred_scores"] >= 0.1).all() and ((d["pred_labels"] >= 1) & (d["pred_labels"] <= 3)).all()
        assert (d["pred_scores"][:-1] >= d["pred_scores"][1:]).all()


def test_multi_classes_nms_dropin(ref_det):
    """MULTI_CLASSES_NMS branch (not used by the IA-SSD / SPSNet configs): per-class device NMS == the reference function."""
    from spsnet_b200 import dense_head as H

    torch.manual_seed(3)
    boxes = _boxes(21, 400)
    scores = torch.rand(400, 3, device="cuda")
    cfg = H.Cfg({**H.KITTI_POST_PROCESSING["NMS_CONFIG"], "NMS_THRESH": 0.1, "NMS_POST_MAXSIZE": 50})
    sc, lab, bx = H.multi_classes_nms(scores, boxes, cfg, score_thresh=0.2)
    assert sc.shape[0] == lab.shape[0] == bx.shape[0] and (sc >= 0.2).all() and set(lab.tolist()) <= {0, 1, 2}
    if ref_det is not None:
        rsc, rlab, rbx = ref_det.nms_utils.multi_classes_nms(scores, boxes, cfg, score_thresh=0.2)
        assert torch.equal(sc, rsc) and torch.equal(lab, rlab) and torch.equal(bx, rbx)
    pp = {**H.KITTI_POST_PROCESSING, "NMS_CONFIG": {**H.KITTI_POST_PROCESSING["NMS_CONFIG"], "MULTI_CLASSES_NMS": True, "NMS_THRESH": 0.1}}
    logits = torch.randn(2 * 200, 3, device="cuda")
    pd, _ = H.post_processing({"batch_size": 2, "batch_cls_preds": logits, "batch_box_preds": boxes, "cls_preds_normalized": False}, pp)
    assert len(pd) == 2 and all(((d["pred_labels"] >= 1) & (d["pred_labels"] <= 3)).all() for d in pd)

"""oracle.py -- TEST INFRASTRUCTURE ONLY (checker + reported CPU baseline; never on the product path).

numpy/ctypes front-end of oracle/oracle.c (the plain-C restatement of the reference's pointnet2_batch
kernels) plus a CPU restatement of the torch-op chains of the reference's `pointnet2_modules.py`
(score top-k samplers, QueryAndGroup + shared MLP + pool, aggregation / confidence layers, Vote layer,
IA-SSD backbone loop).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module.

PARITY PIN: see the header of oracle/oracle.c and tests/golden/README.md.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle.so"


def build(force: bool = False) -> Path:
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < (_HERE / "oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.orc_fps_rank.restype = C.c_uint32
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().orc_num_threads())


# ---- the 11 native ops (reference src/*.cu; citations in oracle.c) ---------------------------------

def fps(xyz, npoint, temp=None):
    """reference FarthestPointSampling.forward (pointnet2_utils.py:12-29): temp starts at 1e10."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    t = np.full((B, N), 1e10, np.float32) if temp is None else _f32(temp).copy()
    idx = np.zeros((B, npoint), np.int32)
    lib().orc_farthest_point_sampling(B, N, npoint, _p(xyz), _p(t), _p(idx))
    return idx


def fps_with_dist(dist, npoint):
    dist = _f32(dist)
    B, N, _ = dist.shape
    t = np.full((B, N), 1e10, np.float32)
    idx = np.zeros((B, npoint), np.int32)
    lib().orc_furthest_point_sampling_with_dist(B, N, npoint, _p(dist), _p(t), _p(idx))
    return idx


def gather(points, idx):
    points, idx = _f32(points), _i32(idx)
    B, Cc, N = points.shape
    M = idx.shape[1]
    out = np.empty((B, Cc, M), np.float32)
    lib().orc_gather_points(B, Cc, N, M, _p(points), _p(idx), _p(out))
    return out


def gather_grad(grad_out, idx, n):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, Cc, M = grad_out.shape
    out = np.zeros((B, Cc, n), np.float32)
    lib().orc_gather_points_grad(B, Cc, n, M, _p(grad_out), _p(idx), _p(out))
    return out


def ball_query(radius, nsample, xyz, new_xyz):
    """reference BallQuery.forward (pointnet2_utils.py:231-249): idx pre-zeroed."""
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = np.zeros((B, M, nsample), np.int32)
    lib().orc_ball_query(B, N, M, C.c_float(radius), nsample, _p(new_xyz), _p(xyz), _p(idx))
    return idx


def ball_query_dilated(max_radius, min_radius, nsample, xyz, new_xyz):
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = np.zeros((B, M, nsample), np.int32)
    lib().orc_ball_query_dilated(B, N, M, C.c_float(max_radius), C.c_float(min_radius), nsample, _p(new_xyz), _p(xyz), _p(idx))
    return idx


def group(points, idx):
    points, idx = _f32(points), _i32(idx)
    B, Cc, N = points.shape
    _, M, S = idx.shape
    out = np.empty((B, Cc, M, S), np.float32)
    lib().orc_group_points(B, Cc, N, M, S, _p(points), _p(idx), _p(out))
    return out


def group_grad(grad_out, idx, n):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, Cc, M, S = grad_out.shape
    out = np.zeros((B, Cc, n), np.float32)
    lib().orc_group_points_grad(B, Cc, n, M, S, _p(grad_out), _p(idx), _p(out))
    return out


def three_nn(unknown, known):
    """returns (dist2, idx); the reference python wrapper returns sqrt(dist2) (pointnet2_utils.py:126)."""
    unknown, known = _f32(unknown), _f32(known)
    B, N, _ = unknown.shape
    M = known.shape[1]
    d2 = np.empty((B, N, 3), np.float32)
    idx = np.empty((B, N, 3), np.int32)
    lib().orc_three_nn(B, N, M, _p(unknown), _p(known), _p(d2), _p(idx))
    return d2, idx


def three_interpolate(points, idx, weight):
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    B, Cc, M = points.shape
    N = idx.shape[1]
    out = np.empty((B, Cc, N), np.float32)
    lib().orc_three_interpolate(B, Cc, M, N, _p(points), _p(idx), _p(weight), _p(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    B, Cc, N = grad_out.shape
    out = np.zeros((B, Cc, m), np.float32)
    lib().orc_three_interpolate_grad(B, Cc, N, m, _p(grad_out), _p(idx), _p(weight), _p(out))
    return out


def fps_rank(k: int, S: int) -> int:
    return int(lib().orc_fps_rank(int(k), int(S)))


# ---- torch-op chains of pointnet2_modules.py --------------------------------------------------------

def _sigmoid32(x):
    x = x.astype(np.float32)
    one = np.float32(1.0)
    return (one / (one + np.exp(-x, dtype=np.float32))).astype(np.float32)


def topk_scores(cls, stds=None):
    """ctr/cls-aware score (reference pointnet2_modules.py:288-289) or SPSNet's stability-weighted score
    (:297-302), each torch op rounding once to fp32."""
    s = _sigmoid32(_f32(cls).max(axis=-1))
    if stds is not None:
        t = (_f32(stds) * np.float32(0.125)).astype(np.float32) - np.float32(3.0)
        sta = (np.float32(1.0) - _sigmoid32(t)).astype(np.float32)
        s = (s * sta).astype(np.float32)
    return s


def score_topk(cls, npoint, stds=None):
    """indices by descending score, ties by ascending index; returns (idx int32 (B,npoint), scores)."""
    s = topk_scores(cls, stds)
    B, N = s.shape
    order = np.stack([np.lexsort((np.arange(N), -s[b].astype(np.float64))) for b in range(B)])
    idx = order[:, :npoint].astype(np.int32)
    return idx, np.take_along_axis(s, idx.astype(np.int64), axis=1)


def same_topk(idx_a, idx_b, scores_full, ulps: int = 4, atol: float = 3e-7) -> bool:
    """Tie-tolerant comparison of two top-k index lists: sequences must agree except
    where the scores of the differing entries are within `ulps` fp32 ulps of each other (a GPU expf vs
    libm expf last-bit difference, or an exact tie, may legally reorder them), and the selected SETS may
    differ only by such near-ties at the k-th boundary."""
    idx_a, idx_b = np.asarray(idx_a), np.asarray(idx_b)
    if idx_a.shape != idx_b.shape:
        return False
    for b in range(idx_a.shape[0]):
        a, c = idx_a[b].astype(np.int64), idx_b[b].astype(np.int64)
        diff = np.nonzero(a != c)[0]
        if diff.size == 0:
            continue
        sa, sc = scores_full[b][a[diff]], scores_full[b][c[diff]]
        # scores are products / differences of sigmoids in [0, 1]: a last-bit expf difference is an ABSOLUTE
        # error of a few ulp(1.0) (1 - sigmoid(t) cancels), so the tolerance has an absolute floor
        tol = np.maximum(ulps * np.spacing(np.maximum(np.abs(sa), np.abs(sc)).astype(np.float32)), atol)
        if not np.all(np.abs(sa.astype(np.float64) - sc.astype(np.float64)) <= tol):
            return False
    return True


def _t(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a))


def query_and_group(radius, nsample, xyz, new_xyz, features, use_xyz=True, idx=None):
    """reference QueryAndGroup.forward (pointnet2_utils.py:299-322) -> (B, 3+C, npoint, nsample)."""
    if idx is None:
        idx = ball_query(radius, nsample, xyz, new_xyz)
    g_xyz = group(np.ascontiguousarray(np.transpose(_f32(xyz), (0, 2, 1))), idx)
    g_xyz = (g_xyz - np.transpose(_f32(new_xyz), (0, 2, 1))[..., None]).astype(np.float32)
    if features is None:
        return g_xyz, idx
    g_f = group(features, idx)
    return (np.concatenate([g_xyz, g_f], axis=1) if use_xyz else g_f), idx


def sa_forward(module, xyz, features, cls_features=None, ctr_xyz=None, stds=None, forced_idx=None, dtype=None):
    """CPU restatement of PointnetSAModuleMSG_WithSampling.forward (reference pointnet2_modules.py:248-460)
    for the sampler types of the shipped configs (identity, ctr/cls-aware, ss/sss, D-FPS, F-FPS).
    `module` is any nn.Module with the reference's attribute layout, on CPU, eval mode.  numpy in/out.
    dtype: torch dtype for the conv/BN stack (float64 = clean truth, float32 = what the reference runs)."""
    import torch
    import torch.nn.functional as F

    dtype = dtype or torch.float64
    xyz = _f32(xyz)
    B = xyz.shape[0]
    sampled = []
    if ctr_xyz is None:
        if forced_idx is not None:
            sampled = _i32(forced_idx)
        else:
            picks = []
            for stype, srange, npoint in zip(module.sample_type_list, module.sample_range_list, module.npoint_list):
                if npoint <= 0:
                    continue
                assert srange == -1, "oracle covers the shipped [-1] ranges only"
                n = xyz.shape[1]
                if n <= npoint:
                    idx = np.tile(np.arange(n, dtype=np.int32), (B, 1))
                elif "cls" in stype or "ctr" in stype:
                    idx, _ = score_topk(cls_features, npoint)
                elif "ss" in stype or "sss" in stype:
                    idx, _ = score_topk(cls_features, npoint, stds=_f32(stds).reshape(B, -1))
                    stds = gather(_f32(stds).reshape(B, 1, -1), idx).reshape(B, -1)  # reference :305
                elif "D-FPS" in stype or "DFS" in stype:
                    idx = fps(xyz, npoint)
                    if stds is not None:
                        stds = gather(_f32(stds).reshape(B, 1, -1), idx).reshape(B, -1)  # reference :309-310
                else:
                    raise NotImplementedError(stype)
                picks.append(idx)
            sampled = np.concatenate(picks, axis=-1)
        new_xyz = np.ascontiguousarray(np.transpose(gather(np.ascontiguousarray(np.transpose(xyz, (0, 2, 1))), sampled), (0, 2, 1)))
    else:
        new_xyz = _f32(ctr_xyz)

    with torch.no_grad():
        if len(module.groupers) > 0:
            outs = []
            for grouper, mlp in zip(module.groupers, module.mlps):
                if hasattr(grouper, "radius_in"):
                    idx = ball_query_dilated(grouper.radius_in, grouper.radius_out, grouper.nsample, xyz, new_xyz)
                    grouped, _ = query_and_group(None, grouper.nsample, xyz, new_xyz, features, grouper.use_xyz, idx=idx)
                else:
                    grouped, _ = query_and_group(grouper.radius, grouper.nsample, xyz, new_xyz, features, grouper.use_xyz)
                y = mlp.to(dtype)(_t(grouped).to(dtype))
                if module.pool_method == "max_pool":
                    y = F.max_pool2d(y, kernel_size=[1, y.size(3)])
                else:
                    y = F.avg_pool2d(y, kernel_size=[1, y.size(3)])
                outs.append(y.squeeze(-1))
            nf = torch.cat(outs, dim=1)
            if module.aggregation_layer is not None:
                nf = module.aggregation_layer.to(dtype)(nf)
        else:
            nf = _t(gather(features, sampled)).to(dtype)
        cls_out = None
        if module.confidence_layers is not None:
            cls_out = module.confidence_layers.to(dtype)(nf).transpose(1, 2).contiguous().float().numpy()
    return new_xyz, nf.float().numpy(), cls_out, sampled, stds


def vote_forward(module, xyz, features, dtype=None):
    """CPU restatement of Vote_layer.forward (reference pointnet2_modules.py:485-516)."""
    import torch

    dtype = dtype or torch.float64
    with torch.no_grad():
        h = module.mlp_modules.to(dtype)(_t(_f32(features)).to(dtype))
        off = module.ctr_reg.to(dtype)(h).transpose(1, 2)[..., :3]
        if module.max_offset_limit is not None:
            lim = module.max_offset_limit.to(dtype).view(1, 1, 3)
            lim_off = torch.minimum(torch.maximum(off, -lim), lim)
        else:
            lim_off = off
        vote = _t(_f32(xyz)).to(dtype) + lim_off
    return vote.float().numpy(), off.float().numpy()


def backbone_forward(backbone, points_bnc, stds=None, dtype=None):
    """CPU restatement of IASSD_Backbone.forward / PAGNet_Backbone.forward (reference
    IASSD_backbone.py:93-168) on a (B, N, 3+C) array.  Returns a dict of per-layer outputs."""
    pts = _f32(points_bnc)
    xyz = np.ascontiguousarray(pts[:, :, :3])
    feats = np.ascontiguousarray(np.transpose(pts[:, :, 3:], (0, 2, 1))) if pts.shape[2] > 3 else None
    enc_xyz, enc_feat = [xyz], [feats]
    out = {"sample_idx": [], "cls": []}
    cls = None
    for i, m in enumerate(backbone.SA_modules):
        xi, fi = enc_xyz[backbone.layer_inputs[i]], enc_feat[backbone.layer_inputs[i]]
        if backbone.layer_types[i] == "SA_Layer":
            ctr = enc_xyz[backbone.ctr_idx_list[i]] if backbone.ctr_idx_list[i] != -1 else None
            lx, lf, cls, sidx, stds = sa_forward(m, xi, fi, cls, ctr_xyz=ctr, stds=stds, dtype=dtype)
            out["sample_idx"].append(sidx)
            out["cls"].append(cls)
        else:
            lx, off = vote_forward(m, xi, fi, dtype=dtype)
            lf = np.zeros((xi.shape[0], fi.shape[1], 0), np.float32)
            out["ctr_offsets"], out["centers_origin"], out["centers"] = off, xi, lx
        enc_xyz.append(lx)
        enc_feat.append(lf)
    out["encoder_xyz"], out["encoder_features"] = enc_xyz, enc_feat
    out["centers_features"] = np.ascontiguousarray(np.transpose(enc_feat[-1], (0, 2, 1))).reshape(-1, enc_feat[-1].shape[1])
    return out


# ---- iou3d_nms + IA-SSD head post-processing ---------------------------------------

def boxes_matrix(boxes_a, boxes_b, mode):
    """mode 'overlap' | 'iou_bev' | 'iou3d': reference boxes_overlap_bev_gpu / boxes_iou_bev_gpu
    (iou3d_nms_kernel.cu:236-265) / boxes_iou3d_gpu (iou3d_nms_utils.py:48-81)."""
    a, b = _f32(boxes_a), _f32(boxes_b)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    if out.size:
        lib().orc_boxes_matrix(a.shape[0], _p(a), b.shape[0], _p(b), _p(out), {"overlap": 0, "iou_bev": 1, "iou3d": 2}[mode])
    return out


def nms_sorted(boxes, thresh, normal=False):
    """Greedy NMS over boxes sorted by descending score (iou3d_nms.cpp:90-188): positions kept, ascending."""
    bx = _f32(boxes)
    keep = np.zeros(max(bx.shape[0], 1), np.int64)
    lib().orc_nms.restype = C.c_int
    n = lib().orc_nms(bx.shape[0], _p(bx), C.c_float(float(thresh)), int(bool(normal)), _p(keep))
    return keep[:n].copy()


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, normal=False):
    """reference iou3d_nms_utils.nms_gpu / nms_normal_gpu (iou3d_nms_utils.py:84-116); ties by ascending index."""
    s = _f32(scores)
    order = np.lexsort((np.arange(s.shape[0]), -s.astype(np.float64)))
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    keep = nms_sorted(_f32(boxes)[order], thresh, normal)
    return order[keep]


def decode_bin_ori(box_encodings, points, pred_classes, mean_size, bin_size=12):
    """PointResidual_BinOri_Coder.decode_torch (box_coder_utils.py:279-319), each torch op rounded to fp32 once.
    pred_classes in 1..num_class; mean_size (num_class, 3) or None."""
    enc, pts = _f32(box_encodings), _f32(points)
    f = np.float32
    if mean_size is not None:
        anchor = _f32(mean_size)[np.asarray(pred_classes) - 1]
        dxa, dya, dza = anchor[:, 0], anchor[:, 1], anchor[:, 2]
        diag = np.sqrt((dxa * dxa).astype(f) + (dya * dya).astype(f), dtype=f)
        xg = (enc[:, 0] * diag).astype(f) + pts[:, 0]
        yg = (enc[:, 1] * diag).astype(f) + pts[:, 1]
        zg = (enc[:, 2] * dza).astype(f) + pts[:, 2]
        dxg = np.exp(enc[:, 3], dtype=f) * dxa
        dyg = np.exp(enc[:, 4], dtype=f) * dya
        dzg = np.exp(enc[:, 5], dtype=f) * dza
    else:
        xg, yg, zg = enc[:, 0] + pts[:, 0], enc[:, 1] + pts[:, 1], enc[:, 2] + pts[:, 2]
        dxg, dyg, dzg = (np.exp(enc[:, k], dtype=f) for k in (3, 4, 5))
    bin_inter = 2 * np.pi / bin_size
    bins = enc[:, 6:6 + bin_size]
    bin_id = bins.argmax(axis=1)
    res = enc[np.arange(enc.shape[0]), 6 + bin_size + bin_id]
    rg = (bin_id.astype(f) * f(bin_inter)).astype(f) - f(np.pi)
    rg = (rg + f(bin_inter / 2)).astype(f)
    rg = (rg + (res * f(bin_inter / 2)).astype(f)).astype(f)
    return np.stack([xg, yg, zg, dxg, dyg, dzg, rg], axis=1).astype(f)


def fc_stack(seq, x, dtype=None):
    """point_head_template.make_fc_layers stack (Linear/BN1d/ReLU ... Linear+bias) on CPU in `dtype`."""
    import torch

    dtype = dtype or torch.float64
    with torch.no_grad():
        return seq.to(dtype)(_t(_f32(x)).to(dtype)).float().numpy()


def head_forward(head, centers_features, centers, dtype=None):
    """IASSD_Head.forward in eval mode (IASSD_head.py:788-840): cls / box stacks + generate_predicted_boxes
    (point_head_template.py:193-207).  `head` is any object with cls_center_layers, box_center_layers (nn.Sequential),
    mean_size (num_class,3 array or None), bin_size.  centers (R,4) [bs,x,y,z]."""
    cls = fc_stack(head.cls_center_layers, centers_features, dtype)
    reg = fc_stack(head.box_center_layers, centers_features, dtype)
    pred_classes = cls.argmax(axis=1) + 1
    boxes = decode_bin_ori(reg, _f32(centers)[:, 1:4], pred_classes, head.mean_size, head.bin_size)
    return cls, reg, boxes


def post_processing(cls_logits, box_preds, batch_size, score_thresh, nms_thresh, pre_max, post_max, normal=False):
    """Detector3DTemplate.post_processing, class-agnostic branch (detector3d_template.py:207-290) +
    class_agnostic_nms (model_nms_utils.py:6-27), equal number of centres per scene.  Returns a list of dicts
    with pred_boxes / pred_scores / pred_labels / index (centre index within the scene)."""
    cls, boxes = _f32(cls_logits), _f32(box_preds)
    m = cls.shape[0] // batch_size
    out = []
    for b in range(batch_size):
        c, bx = cls[b * m:(b + 1) * m], boxes[b * m:(b + 1) * m]
        sc = _sigmoid32(c.max(axis=1))
        lab = c.argmax(axis=1) + 1
        ok = np.nonzero(sc >= np.float32(score_thresh))[0]
        sel = np.zeros(0, np.int64)
        if ok.size:
            s_ok = sc[ok]
            order = np.lexsort((np.arange(ok.size), -s_ok.astype(np.float64)))[:min(pre_max, ok.size)]
            keep = nms_sorted(bx[ok][order], nms_thresh, normal)
            sel = ok[order[keep[:post_max]]]
        out.append({"pred_boxes": bx[sel], "pred_scores": sc[sel], "pred_labels": lab[sel].astype(np.int64), "index": sel})
    return out


# ---- SPSNet surface-feature extractor ------------------------------------------------

def as_ball_query_coords(pos):
    """What the reference ball-query kernel reads when DenseEdgeConv hands it a (B, N, d) FEATURE tensor as xyz
    (surface_feature.py:79 with pos = x in dynamic-graph mode, :170-173): 3 floats per point from the flat buffer
    (ball_query_gpu.cu:18-20), i.e. the first B*N*3 floats viewed as (B, N, 3)."""
    pos = _f32(pos)
    B, N = pos.shape[:2]
    return np.ascontiguousarray(pos.reshape(-1)[:B * N * 3].reshape(B, N, 3))


def dense_edge_conv(conv, x_t, idx):
    """DenseEdgeConv.forward (surface_feature.py:73-115), literally (grouped tensors, concatenations) in x_t's dtype.
    x_t: torch (B, N, d); idx: (B, N, K) integer array of neighbour indices."""
    import torch

    B, N, d = x_t.shape
    ii = torch.from_numpy(np.asarray(idx).astype(np.int64))
    knn_feat = torch.stack([x_t[b][ii[b]] for b in range(B)])            # (B, N, K, d)
    x_tiled = x_t.unsqueeze(-2).expand_as(knn_feat)
    edge = knn_feat - x_tiled if conv.relative_feat_only else torch.cat([x_tiled, knn_feat, knn_feat - x_tiled], dim=3)
    y = torch.cat([conv.layer_first(edge), x_tiled], dim=-1)
    for layer in conv.layers:
        y = torch.cat([layer(y), y], dim=-1)
    y = torch.cat([conv.layer_last(y), y], dim=-1)
    assert conv.aggr.oper == "max"
    return y.max(dim=-2)[0]


def surface_feature_extraction(fe, x, dtype=None, forced_idx=None, forced_t=None):
    """FeatureExtraction.forward (surface_feature.py:118-187) on CPU.  `fe` is a module with the reference's layout
    (transforms[i], convs[i]); pass a deep copy, it is cast to `dtype`.  forced_idx[i]: use these neighbour lists
    (teacher-forcing); forced_t[i]: use these fp32 coordinates for the ball query of unit i.
    Returns (out (B, N, 60) fp32, [idx_i], [t_i fp32])."""
    import torch

    dtype = dtype or torch.float64
    fe = fe.to(dtype)
    cur = _t(_f32(x)).to(dtype)
    pos0 = _f32(x)
    idxs, ts = [], []
    with torch.no_grad():
        for i in range(len(fe.convs)):
            conv = fe.convs[i]
            t = fe.transforms[i](cur)
            t32 = t.float().numpy()
            if forced_idx is not None:
                idx = np.asarray(forced_idx[i])
            else:
                pos = (forced_t[i] if forced_t is not None else t32) if fe.dynamic_graph else pos0
                coords = as_ball_query_coords(pos)
                idx = ball_query(conv.group.radius, conv.knn, coords, coords)
            idxs.append(idx)
            ts.append(t32)
            cur = dense_edge_conv(conv, t, idx)
    return cur.float().numpy(), idxs, ts

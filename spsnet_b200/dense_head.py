"""The consumer of the SA path: IA-SSD detection head (inference) and post-processing.

Drop-ins, same names / arguments / `state_dict` layout as the reference:
  * `PointResidual_BinOri_Coder`   pcdet/utils/box_coder_utils.py:224-319 (decode; encode is training-only)
  * `IASSD_Head`                   pcdet/models/dense_heads/IASSD_head.py:10-38,788-840 (eval-mode forward) with
                                   `make_fc_layers` / `generate_predicted_boxes` of point_head_template.py:36-47,193-207
  * `class_agnostic_nms`           pcdet/models/model_utils/model_nms_utils.py:6-27
  * `post_processing`              Detector3DTemplate.post_processing, class-agnostic branch
                                   (pcdet/models/detectors/detector3d_template.py:186-292)

B200 design.  The reference runs, per batch, 2 x (Linear -> BN1d -> ReLU) x 2 + 2 Linear as ~14 torch kernels, ~25 decode
kernels, then PER SCENE ~15 torch ops, a cudaMalloc, the mask kernel, a blocking D2H copy of the mask and a host loop.
Here the whole head is 6 launches for the batch, with no allocation inside the library and no host synchronisation:
  3 x `spsk_pw_mma_forward` (tcgen05): the cls and box stacks share their input, so layer k of both is ONE GEMM -- weights
      concatenated (layer 0) or block-diagonal (deeper layers), BN folded, fp32-grade hi + lo fp16 arithmetic, rows =
      the point-major fp16 twin the last SA layer already produced;
  3 x `spsk_detect_postprocess`: decode + sigmoid + argmax + score sort per scene, the rotated-IoU suppression bit mask
      (upper triangle only), the greedy pass + gather of the final boxes, all on the device (csrc/iou3d_nms.cu).
`post_processing` returns the reference's list of dicts (ONE read of the per-scene counts for the whole batch);
`detections_padded` is the sync-free, CUDA-graph-capturable form.  Training (target assignment, losses) is out of scope.
"""
from __future__ import annotations

import ctypes as C
from typing import List, NamedTuple, Optional

import numpy as np
import torch
import torch.nn as nn

from . import iou3d_nms_utils
from . import pointnet2_utils as pu
from ._lib import DetectDesc, check, lib
from .configs import Cfg

__all__ = ["PointResidual_BinOri_Coder", "IASSD_Head", "MLT_SSD_Head", "detections_padded", "Detections", "class_agnostic_nms", "multi_classes_nms", "post_processing",
           "kitti_iassd_head_cfg", "waymo_iassd_head_cfg", "waymo_post_processing", "KITTI_POST_PROCESSING"]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ceil(a: int, b: int) -> int:
    return (a + b - 1) // b * b


from .configs import (KITTI_IASSD_HEAD, KITTI_POST_PROCESSING, kitti_iassd_head_cfg, waymo_iassd_head_cfg,  # noqa: F401,E402  (re-exported)
                      waymo_post_processing)


class Detections(NamedTuple):
    """Padded per-scene detections (first count[b] rows of scene b valid, the rest zero / index -1)."""
    boxes: torch.Tensor    # (B, P, 7)
    scores: torch.Tensor   # (B, P)
    labels: torch.Tensor   # (B, P) int64, 1..num_class
    index: torch.Tensor    # (B, P) int64 centre index within the scene
    count: torch.Tensor    # (B,) int32


def _detect_call(batch, m, num_class, bin_size, *, cls=None, reg=None, centers=None, ld_centers=3, mean_size=None,
                 box_preds=None, labels=None, nms=None, device=None):
    """One spsk_detect_postprocess call.  nms = None (decode only) or (score_thresh, nms_thresh, normal, pre_max, post_max).
    Returns (box_preds (R,7), scores (R,), labels (R,) int32, Detections | None)."""
    R = batch * m
    device = device if device is not None else (cls if cls is not None else reg).device
    d = DetectDesc()
    d.batch, d.m, d.num_class, d.bin_size = batch, m, num_class, bin_size
    keep_alive = []
    if cls is not None:
        if cls.dtype != torch.float32 or cls.dim() != 2 or cls.shape[0] != R or cls.stride(1) != 1 or not cls.is_cuda:
            raise RuntimeError(f"class logits must be CUDA float32 ({R}, >= {num_class}) rows, got {tuple(cls.shape)}")
        d.cls, d.ld_cls = cls.data_ptr(), cls.stride(0)
    if reg is not None:
        if reg.dtype != torch.float32 or reg.dim() != 2 or reg.shape[0] != R or reg.stride(1) != 1 or not reg.is_cuda:
            raise RuntimeError(f"box encodings must be CUDA float32 ({R}, >= {6 + 2 * bin_size}) rows, got {tuple(reg.shape)}")
        if reg.shape[1] < 6 + 2 * bin_size:
            raise RuntimeError(f"box encodings have {reg.shape[1]} columns, the coder needs {6 + 2 * bin_size}")
        if centers.dtype != torch.float32 or centers.shape[0] != R or centers.stride(1) != 1:
            raise RuntimeError("centers must be float32 rows, one per box encoding")
        d.reg, d.ld_reg = reg.data_ptr(), reg.stride(0)
        d.centers, d.ld_centers = centers.data_ptr(), centers.stride(0)
        if mean_size is not None:
            d.mean_size = mean_size.data_ptr()
        box_preds = torch.empty((R, 7), dtype=torch.float32, device=device)
    else:
        if box_preds is None or box_preds.shape != (R, 7) or box_preds.dtype != torch.float32:
            raise RuntimeError(f"box_preds must be float32 ({R}, 7)")
        box_preds = box_preds.contiguous()
    scores = torch.empty((R,), dtype=torch.float32, device=device) if cls is not None else None
    if labels is None:
        labels = torch.empty((R,), dtype=torch.int32, device=device)
    d.box_preds, d.labels = box_preds.data_ptr(), labels.data_ptr()
    if scores is not None:
        d.scores = scores.data_ptr()
    det = None
    if nms is not None:
        score_thresh, nms_thresh, normal, pre_max, post_max = nms
        P = max(1, min(int(post_max), m))
        d.score_thresh = float(score_thresh) if score_thresh is not None else float("-inf")
        d.nms_thresh, d.nms_normal, d.pre_max, d.post_max = float(nms_thresh), int(bool(normal)), max(1, int(pre_max)), P
        det = Detections(torch.empty((batch, P, 7), dtype=torch.float32, device=device),
                         torch.empty((batch, P), dtype=torch.float32, device=device),
                         torch.empty((batch, P), dtype=torch.int64, device=device),
                         torch.empty((batch, P), dtype=torch.int64, device=device),
                         torch.empty((batch,), dtype=torch.int32, device=device))
        d.out_boxes, d.out_scores, d.out_labels = det.boxes.data_ptr(), det.scores.data_ptr(), det.labels.data_ptr()
        d.out_index, d.out_count = det.index.data_ptr(), det.count.data_ptr()
        need = int(lib.spsk_detect_workspace_bytes(batch, m))
        ws = torch.empty(max(need, 8), dtype=torch.uint8, device=device)
        keep_alive.append(ws)
        d.workspace, d.workspace_bytes = ws.data_ptr(), need
    with torch.cuda.device(device):
        check(lib.spsk_detect_postprocess(C.byref(d), _stream()), "detect_postprocess")
    return box_preds, scores, labels, det


class PointResidual_BinOri_Coder(object):
    """reference pcdet/utils/box_coder_utils.py:224-319.  `decode_torch` runs the fused decode kernel."""

    def __init__(self, code_size=8, use_mean_size=True, **kwargs):
        self.bin_size = kwargs.get("bin_size", 12)
        self.code_size = 6 + 2 * self.bin_size
        self.bin_inter = 2 * np.pi / self.bin_size
        self.use_mean_size = use_mean_size
        self._mean_np = None
        self._mean_dev = {}
        if self.use_mean_size:
            self._mean_np = np.asarray(kwargs["mean_size"], dtype=np.float32)
            assert self._mean_np.min() > 0

    def mean_size_on(self, device) -> Optional[torch.Tensor]:
        """(num_class, 3) fp32 on `device` (cached: the reference pins it to the current GPU at construction, :233)."""
        if self._mean_np is None:
            return None
        key = str(device)
        if key not in self._mean_dev:
            self._mean_dev[key] = torch.from_numpy(self._mean_np).to(device)
        return self._mean_dev[key]

    @property
    def mean_size(self):
        return self.mean_size_on(torch.device("cuda", torch.cuda.current_device()))

    def encode_torch(self, gt_boxes, points, gt_classes=None):
        raise NotImplementedError("box encoding is used by training-time target assignment only (out of scope, SURVEY.md §8)")

    def decode_torch(self, box_encodings, points, pred_classes=None):
        """(N, 6 + 2*bin) encodings, (N, 3) points, (N,) classes in 1..num_class -> (N, 7) boxes."""
        if not box_encodings.is_cuda:
            raise RuntimeError("spsnet_b200 has no CPU path: box_encodings must be a CUDA tensor")
        enc = box_encodings.reshape(-1, box_encodings.shape[-1]).float().contiguous()
        pts = points.reshape(-1, 3).float().contiguous()
        mean = self.mean_size_on(enc.device)
        ncls = mean.shape[0] if mean is not None else 1
        if pred_classes is None:
            if mean is not None:
                raise RuntimeError("use_mean_size needs pred_classes")
            labels = torch.ones(enc.shape[0], dtype=torch.int32, device=enc.device)
        else:
            labels = pred_classes.reshape(-1).to(torch.int32).contiguous()
        out = torch.empty((enc.shape[0], 7), dtype=torch.float32, device=enc.device)
        for s in range(0, enc.shape[0], 4096):  # the kernel takes up to SPSK_DETECT_MAX_M rows per "scene"
            e = min(s + 4096, enc.shape[0])
            out[s:e] = _detect_call(1, e - s, ncls, self.bin_size, reg=enc[s:e], centers=pts[s:e], mean_size=mean,
                                    labels=labels[s:e].contiguous())[0]
        return out.view(*box_encodings.shape[:-1], 7)


def _fold_fc(seq: nn.Sequential):
    """[(W (out,in) fp32, bias (out,), relu)] of a make_fc_layers stack with eval-mode BatchNorm1d folded in."""
    mods = list(seq)
    out, i = [], 0
    while i < len(mods):
        lin = mods[i]
        if not isinstance(lin, nn.Linear):
            raise RuntimeError(f"unexpected module {type(lin).__name__} in an FC stack")
        W = lin.weight.detach().float()
        b = lin.bias.detach().float() if lin.bias is not None else torch.zeros(W.shape[0], device=W.device)
        i += 1
        if i < len(mods) and isinstance(mods[i], nn.BatchNorm1d):
            bn = mods[i]
            scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            W = W * scale[:, None]
            b = (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
            i += 1
        relu = i < len(mods) and isinstance(mods[i], nn.ReLU)
        if relu:
            i += 1
        out.append((W, b, relu))
    return out


def _fuse_dense(stacks: List[nn.Sequential]):
    """Dense fused layers [(W (out,in), bias, relu)] of several equally deep FC stacks that read the same rows, plus
    the column slice of the last layer's output that belongs to each stack."""
    folded = [_fold_fc(s) for s in stacks]
    depth = {len(f) for f in folded}
    if len(depth) != 1:
        raise RuntimeError("FC stacks of different depth cannot be fused")  # caller falls back to one stack per call
    layers, out_slices = [], []
    for k in range(depth.pop()):
        Ws = [f[k][0] for f in folded]
        relu = folded[0][k][2]
        if any(f[k][2] != relu for f in folded):
            raise RuntimeError("FC stacks disagree on the activation of a layer")
        bias = torch.cat([f[k][1] for f in folded])
        # layer 0: shared input, stack the output rows; deeper: inputs = outputs of layer k-1 laid out stack after stack
        W = torch.cat(Ws, dim=0) if k == 0 else torch.block_diag(*Ws)
        layers.append((W, bias, relu))
        if k == len(folded[0]) - 1:
            o = 0
            for w in Ws:
                out_slices.append((o, o + w.shape[0]))
                o += w.shape[0]
    return layers, out_slices


class _FusedStacks:
    """Several FC stacks that read the same rows, packed so that layer k of ALL stacks is one tensor-core GEMM."""

    def __init__(self, stacks: List[nn.Sequential]):
        dense, self.out_slices = _fuse_dense(stacks)
        self.layers = [pu.PwLayer(W.t().contiguous(), bias, relu, split=True) for W, bias, relu in dense]

    def __call__(self, x16: torch.Tensor, xlo: int) -> torch.Tensor:
        for i, layer in enumerate(self.layers):
            if i < len(self.layers) - 1:
                _, x16, _ = pu.pw_mma_forward(x16, layer, xlo=xlo, want16_lo=True)
                xlo = layer.n16
            else:
                out = torch.empty((x16.shape[0], layer.c_out), dtype=torch.float32, device=x16.device)
                pu.pw_mma_forward(x16, layer, xlo=xlo, out_pm=out)
                return out
        raise RuntimeError("empty FC stack")


def _rows16(x: torch.Tensor):
    """Point-major fp16 rows [values | residuals] of an (R, C) fp32 tensor: the copy the backbone attached, or a fresh one."""
    hit = getattr(x, "_spsk_rows16", None)
    if hit is not None and hit[1] == x._version and hit[0].shape[0] == x.shape[0] and hit[2] >= x.shape[1]:
        return hit[0], hit[2]
    c16 = _ceil(x.shape[1], 16)
    rows = torch.zeros((x.shape[0], 2 * c16), dtype=torch.float16, device=x.device)
    hi = x.half()
    rows[:, :x.shape[1]] = hi
    rows[:, c16:c16 + x.shape[1]] = (x - hi.float()).half()
    return rows, c16


class IASSD_Head(nn.Module):
    """A simple point-based detect head, used for IA-SSD (reference IASSD_head.py:10-38).  Inference only.

    `post_process_cfg` (optional, the model's POST_PROCESSING block): when given, forward() also runs score filtering +
    NMS inside the same fused call and leaves the padded result in batch_dict['detections'] (see `post_processing`)."""

    def __init__(self, num_class, input_channels, model_cfg, predict_boxes_when_training=False, post_process_cfg=None, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg if isinstance(model_cfg, Cfg) else Cfg(model_cfg)
        self.num_class = num_class
        self.predict_boxes_when_training = predict_boxes_when_training
        self.post_process_cfg = Cfg(post_process_cfg) if post_process_cfg is not None else None
        target_cfg = self.model_cfg.TARGET_CONFIG
        coder = target_cfg.BOX_CODER
        if coder != "PointResidual_BinOri_Coder":
            raise NotImplementedError(f"box coder {coder}: only PointResidual_BinOri_Coder (IA-SSD / SPSNet-IA configs) is provided")
        self.box_coder = PointResidual_BinOri_Coder(**target_cfg.BOX_CODER_CONFIG)
        detector_dim = self.model_cfg.get("INPUT_DIM", input_channels)
        self.cls_center_layers = self.make_fc_layers(self.model_cfg.CLS_FC, detector_dim, num_class)
        self.box_center_layers = self.make_fc_layers(self.model_cfg.REG_FC, detector_dim, self.box_coder.code_size)
        self.box_iou3d_layers = self.make_fc_layers(self.model_cfg.IOU_FC, detector_dim, 1) \
            if self.model_cfg.get("IOU_FC", None) is not None else None
        self.forward_ret_dict = None

    @staticmethod
    def make_fc_layers(fc_cfg, input_channels, output_channels):
        """reference point_head_template.py:36-47."""
        fc_layers = []
        c_in = input_channels
        for k in range(len(fc_cfg)):
            fc_layers.extend([nn.Linear(c_in, fc_cfg[k], bias=False), nn.BatchNorm1d(fc_cfg[k]), nn.ReLU()])
            c_in = fc_cfg[k]
        fc_layers.append(nn.Linear(c_in, output_channels, bias=True))
        return nn.Sequential(*fc_layers)

    def invalidate_cache(self) -> None:
        """Call after changing weights in place (load_state_dict does it automatically)."""
        self.__dict__.pop("_fused", None)

    def load_state_dict(self, *a, **k):
        self.invalidate_cache()
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self.invalidate_cache()
        return super()._apply(fn, *a, **k)

    def _fused_stacks(self):
        f = self.__dict__.get("_fused")
        if f is None:
            stacks = [self.cls_center_layers, self.box_center_layers]
            if self.box_iou3d_layers is not None:
                stacks.append(self.box_iou3d_layers)
            try:
                f = [(_FusedStacks(stacks), None)]
            except RuntimeError:
                f = [(_FusedStacks([s]), None) for s in stacks]
            self.__dict__["_fused"] = f
        return f

    def _fc_outputs(self, center_features: torch.Tensor):
        """cls / box / iou3d predictions as column views of the packed GEMM output."""
        x16, xlo = _rows16(center_features)
        outs = []
        for stack, _ in self._fused_stacks():
            y = stack(x16, xlo)
            outs.extend(y[:, a:b] for a, b in stack.out_slices)
        return outs

    def generate_predicted_boxes(self, points, point_cls_preds, point_box_preds):
        """reference point_head_template.py:193-207."""
        _, pred_classes = point_cls_preds.max(dim=-1)
        return point_cls_preds, self.box_coder.decode_torch(point_box_preds, points, pred_classes + 1)

    def forward(self, batch_dict):
        """reference IASSD_head.py:788-840 (eval branch).  batch_dict: batch_size, centers_features (R, C), centers (R, 4)
        [bs, x, y, z] with the same number of centres per scene, ctr_offsets, centers_origin, sa_ins_preds."""
        if self.training:
            raise NotImplementedError("IASSD_Head: target assignment / losses are out of scope (SURVEY.md §8); call .eval()")
        center_features = batch_dict["centers_features"]
        center_coords = batch_dict["centers"]
        if not center_features.is_cuda:
            raise RuntimeError("spsnet_b200 has no CPU path: centers_features must be a CUDA tensor")
        with torch.no_grad():
            outs = self._fc_outputs(center_features)
            center_cls_preds, center_box_preds = outs[0], outs[1]
            box_iou3d_preds = outs[2] if len(outs) > 2 else None
            B = int(batch_dict["batch_size"])
            R = center_features.shape[0]
            if R % B != 0:
                raise RuntimeError("IASSD_Head needs the same number of centres in every scene")
            nms = _nms_args(self.post_process_cfg) if self.post_process_cfg is not None and R // B <= 4096 else None
            centers_xyz = center_coords[:, 1:4]
            box_preds, scores, labels, det = _detect_call(
                B, R // B, self.num_class, self.box_coder.bin_size, cls=center_cls_preds, reg=center_box_preds,
                centers=centers_xyz, mean_size=self.box_coder.mean_size_on(center_features.device), nms=nms) \
                if R // B <= 4096 else self._decode_large(center_cls_preds, center_box_preds, centers_xyz)
        ret_dict = {"center_cls_preds": center_cls_preds, "center_box_preds": center_box_preds,
                    "ctr_offsets": batch_dict.get("ctr_offsets"), "centers": batch_dict["centers"],
                    "centers_origin": batch_dict.get("centers_origin"), "sa_ins_preds": batch_dict.get("sa_ins_preds"),
                    "box_iou3d_preds": box_iou3d_preds, "point_box_preds": box_preds}
        batch_dict["batch_cls_preds"] = center_cls_preds
        batch_dict["batch_box_preds"] = box_preds
        batch_dict["box_iou3d_preds"] = box_iou3d_preds
        batch_dict["batch_index"] = center_coords[:, 0]
        batch_dict["cls_preds_normalized"] = False
        batch_dict["point_scores"], batch_dict["point_labels"] = scores, labels
        if det is not None:
            batch_dict["detections"] = det
            batch_dict["detections_cfg"] = nms
        self.forward_ret_dict = ret_dict
        return batch_dict

    def _decode_large(self, cls, reg, xyz):
        _, pred = cls.max(dim=-1)
        boxes = self.box_coder.decode_torch(reg, xyz, pred + 1)
        return boxes, None, None, None


# SPSNet-IA's head (reference MLT_SSD_head.py:10-41,788-841): same layers, same eval-mode forward; it differs from
# IASSD_Head only in training-time target assignment / losses, which are out of scope.
MLT_SSD_Head = IASSD_Head


def _nms_args(cfg):
    n = cfg.NMS_CONFIG
    if n.get("MULTI_CLASSES_NMS", False):
        return None
    t = n.get("NMS_TYPE", "nms_gpu")
    if t not in ("nms_gpu", "nms_normal_gpu"):
        raise NotImplementedError(f"NMS_TYPE {t}")
    return (cfg.get("SCORE_THRESH", None), n.NMS_THRESH, t == "nms_normal_gpu", n.NMS_PRE_MAXSIZE, n.NMS_POST_MAXSIZE)


def class_agnostic_nms(box_scores, box_preds, nms_config, score_thresh=None):
    """reference model_nms_utils.py:6-27, same returns (selected indices into the inputs, their scores)."""
    src_box_scores = box_scores
    if score_thresh is not None:
        scores_mask = box_scores >= score_thresh
        box_scores = box_scores[scores_mask]
        box_preds = box_preds[scores_mask]
    selected = []
    if box_scores.shape[0] > 0:
        box_scores_nms, indices = torch.topk(box_scores, k=min(nms_config.NMS_PRE_MAXSIZE, box_scores.shape[0]))
        boxes_for_nms = box_preds[indices]
        keep_idx, _ = getattr(iou3d_nms_utils, nms_config.NMS_TYPE)(boxes_for_nms[:, 0:7].contiguous(), box_scores_nms,
                                                                  nms_config.NMS_THRESH, **nms_config)
        selected = indices[keep_idx[:nms_config.NMS_POST_MAXSIZE]]
    if score_thresh is not None:
        original_idxs = scores_mask.nonzero().view(-1)
        selected = original_idxs[selected]
    return selected, src_box_scores[selected]


def multi_classes_nms(cls_scores, box_preds, nms_config, score_thresh=None):
    """reference model_nms_utils.py:30-66: NMS per class on the device kernels (one count read per class).
    cls_scores (N, num_class), box_preds (N, 7 + C) -> (pred_scores, pred_labels, pred_boxes), classes concatenated."""
    nms_fn = getattr(iou3d_nms_utils, nms_config.NMS_TYPE)
    out_scores, out_labels, out_boxes = [], [], []
    for k in range(cls_scores.shape[1]):
        scores_k, boxes_k = cls_scores[:, k], box_preds
        if score_thresh is not None:
            keep_mask = scores_k >= score_thresh
            scores_k, boxes_k = scores_k[keep_mask], box_preds[keep_mask]
        picked = scores_k.new_zeros((0,), dtype=torch.long)
        if scores_k.shape[0] > 0:
            top_scores, top_idx = torch.topk(scores_k, k=min(nms_config.NMS_PRE_MAXSIZE, scores_k.shape[0]))
            keep_idx, _ = nms_fn(boxes_k[top_idx][:, 0:7].contiguous(), top_scores, nms_config.NMS_THRESH, **nms_config)
            picked = top_idx[keep_idx[:nms_config.NMS_POST_MAXSIZE]]
        out_scores.append(scores_k[picked])
        out_labels.append(torch.full((picked.shape[0],), k, dtype=torch.long, device=cls_scores.device))
        out_boxes.append(boxes_k[picked])
    return torch.cat(out_scores, dim=0), torch.cat(out_labels, dim=0), torch.cat(out_boxes, dim=0)


def detections_padded(batch_dict, post_process_cfg) -> Detections:
    """Sync-free post-processing of a whole batch (class-agnostic NMS): reuses what IASSD_Head.forward already computed
    for the same configuration, else runs the fused kernels on batch_cls_preds / batch_box_preds."""
    cfg = post_process_cfg if isinstance(post_process_cfg, Cfg) else Cfg(post_process_cfg)
    nms = _nms_args(cfg)
    if nms is None:
        raise NotImplementedError("MULTI_CLASSES_NMS is not fused; use post_processing()")
    if batch_dict.get("detections") is not None and batch_dict.get("detections_cfg") == nms:
        return batch_dict["detections"]
    cls, boxes = batch_dict["batch_cls_preds"], batch_dict["batch_box_preds"]
    if batch_dict.get("cls_preds_normalized", False):
        raise NotImplementedError("fused post-processing expects raw logits (cls_preds_normalized = False)")
    B = int(batch_dict["batch_size"])
    if cls.dim() == 3:
        cls, boxes = cls.reshape(-1, cls.shape[-1]), boxes.reshape(-1, boxes.shape[-1])
    R = cls.shape[0]
    if R % B:
        raise RuntimeError("fused post-processing needs the same number of boxes in every scene")
    cls = cls if cls.stride(1) == 1 else cls.contiguous()
    _, _, _, det = _detect_call(B, R // B, cls.shape[1], 12, cls=cls.float(), box_preds=boxes[:, :7].float().contiguous(), nms=nms)
    return det


def post_processing(batch_dict, post_process_cfg, num_class: Optional[int] = None):
    """Detector3DTemplate.post_processing (reference detector3d_template.py:186-292), class-agnostic NMS branch, for
    batches with the same number of boxes per scene.  Returns (pred_dicts, recall_dict) like the reference; recall
    bookkeeping needs ground truth and is evaluation code (out of scope): recall_dict is returned empty."""
    cfg = post_process_cfg if isinstance(post_process_cfg, Cfg) else Cfg(post_process_cfg)
    if cfg.NMS_CONFIG.get("MULTI_CLASSES_NMS", False):
        # reference detector3d_template.py:232-257 (single head): per scene, per class NMS; labels are 1-based
        B = int(batch_dict["batch_size"])
        cls = batch_dict["batch_cls_preds"].reshape(B, -1, batch_dict["batch_cls_preds"].shape[-1])
        boxes = batch_dict["batch_box_preds"].reshape(B, cls.shape[1], -1)
        if not batch_dict.get("cls_preds_normalized", False):
            cls = torch.sigmoid(cls)
        pred_dicts = []
        for b in range(B):
            sc, lab, bx = multi_classes_nms(cls[b], boxes[b], cfg.NMS_CONFIG, cfg.get("SCORE_THRESH", None))
            pred_dicts.append({"pred_boxes": bx, "pred_scores": sc, "pred_labels": lab + 1})
        return pred_dicts, {}
    det = detections_padded(batch_dict, cfg)
    counts = det.count.tolist()  # the one host read of the batch
    raw = None
    if cfg.get("OUTPUT_RAW_SCORE", False):
        raw = batch_dict["batch_cls_preds"].reshape(int(batch_dict["batch_size"]), -1, batch_dict["batch_cls_preds"].shape[-1]).max(dim=-1)[0]
    pred_dicts = []
    for b, n in enumerate(counts):
        scores = det.scores[b, :n]
        if raw is not None:
            scores = raw[b][det.index[b, :n]]
        pred_dicts.append({"pred_boxes": det.boxes[b, :n], "pred_scores": scores, "pred_labels": det.labels[b, :n]})
    return pred_dicts, {}

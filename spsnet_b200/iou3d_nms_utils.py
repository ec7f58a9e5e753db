"""Drop-in for the reference's `pcdet/ops/iou3d_nms/iou3d_nms_utils.py`.

Same public names, arguments and return values (reference iou3d_nms_utils.py:31-116); every native call goes through
the C-ABI of include/spsk.h section 3 (libspsk.so, csrc/iou3d_nms.cu) on torch's current stream.  Differences, all
deliberate:
  * `nms_gpu` / `nms_normal_gpu` run the greedy suppression on the device.  The reference copies the N x N/64 bit
    mask to the host and loops there (src/iou3d_nms.cpp:103-126); here the only host interaction is reading ONE int
    (the number kept) because the reference API returns a variable-length tensor.  `nms_batch` is the batched,
    sync-free form (padded keep lists + counts) used by the fused post-processing;
  * `boxes_iou3d_gpu` is a single launch instead of one kernel + ~14 torch ops;
  * there is no CPU path (`boxes_bev_iou_cpu` raises: the product has no CPU fallback by design -- the CPU
    restatement lives in oracle/ and is test infrastructure only);
  * preconditions raise instead of `assert` / `exit(-1)`.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._lib import check, lib

__all__ = ["boxes_iou_bev", "boxes_overlap_bev", "boxes_iou3d_gpu", "nms_gpu", "nms_normal_gpu", "nms_batch",
           "boxes_bev_iou_cpu"]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _boxes(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dim() != 2 or t.shape[1] != 7:
        raise RuntimeError(f"{name} must be (N, 7) [x, y, z, dx, dy, dz, heading], got {tuple(t.shape)}")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _matrix(fn, boxes_a: torch.Tensor, boxes_b: torch.Tensor, what: str) -> torch.Tensor:
    a, b = _boxes(boxes_a, "boxes_a"), _boxes(boxes_b, "boxes_b")
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(fn(a.shape[0], a.data_ptr(), b.shape[0], b.data_ptr(), out.data_ptr(), _stream()), what)
    return out


def boxes_bev_iou_cpu(boxes_a, boxes_b):
    """reference iou3d_nms_utils.py:12-28.  Not provided: this package has no CPU implementation of anything."""
    raise RuntimeError("spsnet_b200 has no CPU path; move the boxes to the GPU and call boxes_iou_bev")


def boxes_overlap_bev(boxes_a: torch.Tensor, boxes_b: torch.Tensor) -> torch.Tensor:
    """(N,7),(M,7) -> (N,M) BEV intersection areas; reference iou3d_nms_cuda.boxes_overlap_bev_gpu."""
    return _matrix(lib.spsk_boxes_overlap_bev, boxes_a, boxes_b, "boxes_overlap_bev")


def boxes_iou_bev(boxes_a: torch.Tensor, boxes_b: torch.Tensor) -> torch.Tensor:
    """(N,7),(M,7) -> (N,M) rotated BEV IoU; reference iou3d_nms_utils.py:31-45."""
    return _matrix(lib.spsk_boxes_iou_bev, boxes_a, boxes_b, "boxes_iou_bev")


def boxes_iou3d_gpu(boxes_a: torch.Tensor, boxes_b: torch.Tensor) -> torch.Tensor:
    """(N,7),(M,7) -> (N,M) 3-D IoU; reference iou3d_nms_utils.py:48-81."""
    return _matrix(lib.spsk_boxes_iou3d, boxes_a, boxes_b, "boxes_iou3d")


def nms_batch(boxes: torch.Tensor, thresh: float, counts: Optional[torch.Tensor] = None, normal: bool = False,
              workspace: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Batched greedy NMS without any host synchronisation.

    boxes (B, N, 7) already sorted by descending score per scene; counts (B,) int32 = boxes present per scene (None: N).
    Returns keep (B, N) int64 (first num_keep[b] entries valid: positions into the sorted order) and num_keep (B,) int32.
    """
    if boxes.dim() != 3 or boxes.shape[2] != 7 or not boxes.is_cuda or boxes.dtype != torch.float32:
        raise RuntimeError(f"boxes must be a CUDA float32 (B, N, 7) tensor, got {tuple(boxes.shape)} {boxes.dtype}")
    boxes = boxes.contiguous()
    B, N, _ = boxes.shape
    if counts is not None:
        if counts.dtype != torch.int32 or counts.shape != (B,) or not counts.is_cuda:
            raise RuntimeError("counts must be a CUDA int32 (B,) tensor")
        counts = counts.contiguous()
    need = int(lib.spsk_nms_workspace_bytes(B, N))
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(max(need, 8), dtype=torch.uint8, device=boxes.device)
    keep = torch.empty((B, N), dtype=torch.int64, device=boxes.device)
    num = torch.empty((B,), dtype=torch.int32, device=boxes.device)
    with torch.cuda.device(boxes.device):
        check(lib.spsk_nms(B, N, boxes.data_ptr(), counts.data_ptr() if counts is not None else None, float(thresh),
                           int(bool(normal)), keep.data_ptr(), num.data_ptr(), workspace.data_ptr(),
                           workspace.numel() * workspace.element_size(), _stream()), "nms")
    return keep, num


def _nms(boxes: torch.Tensor, scores: torch.Tensor, thresh: float, pre_maxsize: Optional[int], normal: bool):
    boxes = _boxes(boxes, "boxes")
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    if order.numel() == 0:
        return order.contiguous(), None
    keep, num = nms_batch(boxes[order].unsqueeze(0), thresh, normal=normal)
    n = int(num.item())  # the reference API returns a variable-length tensor: one int crosses to the host
    return order[keep[0, :n]].contiguous(), None


def nms_gpu(boxes: torch.Tensor, scores: torch.Tensor, thresh: float, pre_maxsize: Optional[int] = None, **kwargs):
    """reference iou3d_nms_utils.py:84-99: rotated-IoU NMS; returns (indices into `boxes` ordered by score, None)."""
    return _nms(boxes, scores, thresh, pre_maxsize, False)


def nms_normal_gpu(boxes: torch.Tensor, scores: torch.Tensor, thresh: float, **kwargs):
    """reference iou3d_nms_utils.py:102-116: axis-aligned BEV NMS."""
    return _nms(boxes, scores, thresh, None, True)

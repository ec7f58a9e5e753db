"""`iou3d_nms_cuda` -- the GPU functions the reference binds with pybind11 (pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:9-15),
same names, arguments and return values, over include/spsk.h section 3.  `boxes_iou_bev_cpu` (the reference's CPU
implementation, iou3d_cpu.cpp) has no counterpart here on purpose: this library has no CPU path.

nms_gpu / nms_normal_gpu fill the caller's CPU LongTensor `keep` (iou3d_nms_utils.py:97,114) and return the count, like
the reference; the suppression loop itself runs on the device (one count read-back instead of the reference's cudaMalloc +
blocking mask copy + host loop, src/iou3d_nms.cpp:90-140).
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import torch

_LIB_PATH = os.environ.get("SPSK_LIB") or str(Path(__file__).resolve().parents[1] / "_C" / "libspsk.so")
if not os.path.exists(_LIB_PATH):
    raise ImportError(f"iou3d_nms_cuda shim: {_LIB_PATH} not found (build with `python -m spsnet_b200.build` or set SPSK_LIB)")
_l = ctypes.CDLL(_LIB_PATH)
_l.spsk_last_error.restype = ctypes.c_char_p
_vp, _i, _f, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong
_l.spsk_boxes_overlap_bev.argtypes = _l.spsk_boxes_iou_bev.argtypes = [_i, _vp, _i, _vp, _vp, _vp]
_l.spsk_nms_workspace_bytes.argtypes = [_i, _i]
_l.spsk_nms_workspace_bytes.restype = _ll
_l.spsk_nms.argtypes = [_i, _i, _vp, _vp, _f, _i, _vp, _vp, _vp, _ll, _vp]


def _s():
    return torch.cuda.current_stream().cuda_stream


def _ok(rc):
    if rc != 0:
        msg = _l.spsk_last_error()
        raise RuntimeError(f"libspsk status {rc}: {msg.decode() if msg else ''}")
    return 1


def _chk(t):
    if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
        raise RuntimeError("iou3d_nms_cuda: boxes must be contiguous CUDA float32 tensors")   # reference: CHECK_INPUT (iou3d_nms.cpp:12-14)
    return t.data_ptr()


# iou3d_nms.cpp:47-67
def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    with torch.cuda.device(boxes_a.device):
        return _ok(_l.spsk_boxes_overlap_bev(boxes_a.size(0), _chk(boxes_a), boxes_b.size(0), _chk(boxes_b), _chk(ans_overlap), _s()))


# iou3d_nms.cpp:69-88
def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    with torch.cuda.device(boxes_a.device):
        return _ok(_l.spsk_boxes_iou_bev(boxes_a.size(0), _chk(boxes_a), boxes_b.size(0), _chk(boxes_b), _chk(ans_iou), _s()))


def _nms(boxes, keep, nms_overlap_thresh, normal):
    n = boxes.size(0)
    if n == 0:
        return 0
    with torch.cuda.device(boxes.device):
        ws = torch.empty(max(int(_l.spsk_nms_workspace_bytes(1, n)), 8), dtype=torch.uint8, device=boxes.device)
        k = torch.empty(n, dtype=torch.int64, device=boxes.device)
        num = torch.empty(1, dtype=torch.int32, device=boxes.device)
        _ok(_l.spsk_nms(1, n, _chk(boxes), None, float(nms_overlap_thresh), normal, k.data_ptr(), num.data_ptr(), ws.data_ptr(),
                        ws.numel(), _s()))
    m = int(num.item())
    keep[:m] = k[:m].cpu()
    return m


# iou3d_nms.cpp:90-140: boxes (N, 7) sorted by descending score; keep = CPU LongTensor(N); returns the number kept
def nms_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, 0)


# iou3d_nms.cpp:143-188
def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, 1)

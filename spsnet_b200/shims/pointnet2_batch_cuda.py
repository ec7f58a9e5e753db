"""`pointnet2_batch_cuda` -- the 11 functions the reference binds with pybind11 (src/pointnet2_api.cpp:10-26), same
names, argument order and return values, over the C-ABI of include/spsk.h section 1.

The reference's Python layer keeps allocating its legacy buffers and they are honoured: FPS reads the caller's `temp`
(pre-filled with 1e10, pointnet2_utils.py:26) and writes the final minima back; ball query leaves the rows of empty balls
untouched (the caller pre-zeroes `idx`, pointnet2_utils.py:246).  Differences, deliberate: launches go to torch's CURRENT
stream (the reference uses the legacy default stream), and a failed precondition raises RuntimeError where the reference
`fprintf`s and `exit(-1)`s (ball_query.cpp:19-20, sampling_gpu.cu:248-252).  This file has no dependency on the rest of
the package besides the library path: it can be copied next to the reference's `pointnet2_utils.py` as is (set SPSK_LIB).
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import torch

_LIB_PATH = os.environ.get("SPSK_LIB") or str(Path(__file__).resolve().parents[1] / "_C" / "libspsk.so")
if not os.path.exists(_LIB_PATH):
    raise ImportError(f"pointnet2_batch_cuda shim: {_LIB_PATH} not found (build with `python -m spsnet_b200.build` or set SPSK_LIB); "
                      "there is no CPU fallback")
_l = ctypes.CDLL(_LIB_PATH)
_l.spsk_last_error.restype = ctypes.c_char_p
_vp, _i, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
for _name, _args in {
    "spsk_farthest_point_sampling": [_i, _i, _i, _vp, _vp, _vp, _vp],
    "spsk_furthest_point_sampling_with_dist": [_i, _i, _i, _vp, _vp, _vp, _vp],
    "spsk_gather_points": [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "spsk_gather_points_grad": [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "spsk_ball_query": [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp],
    "spsk_ball_query_dilated": [_i, _i, _i, _f, _f, _i, _vp, _vp, _vp, _vp],
    "spsk_group_points": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "spsk_group_points_grad": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "spsk_three_nn": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "spsk_three_interpolate": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "spsk_three_interpolate_grad": [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
}.items():
    getattr(_l, _name).argtypes = _args
    getattr(_l, _name).restype = _i


def _s():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
        raise RuntimeError("pointnet2_batch_cuda: every tensor argument must be a contiguous CUDA tensor")
    return t.data_ptr()


def _ok(rc, ret=1):
    if rc != 0:
        msg = _l.spsk_last_error()
        raise RuntimeError(f"libspsk status {rc}: {msg.decode() if msg else ''}")
    return ret


def _on(t):
    return torch.cuda.device(t.device)


# sampling.cpp:34-43 -> int
def farthest_point_sampling_wrapper(b, n, m, points_tensor, temp_tensor, idx_tensor):
    with _on(points_tensor):
        return _ok(_l.spsk_farthest_point_sampling(b, n, m, _p(points_tensor), _p(temp_tensor), _p(idx_tensor), _s()))


# sampling.cpp:46-56 -> 2 (sic)
def furthest_point_sampling_with_dist_wrapper(b, n, m, points_tensor, temp_tensor, idx_tensor):
    with _on(points_tensor):
        return _ok(_l.spsk_furthest_point_sampling_with_dist(b, n, m, _p(points_tensor), _p(temp_tensor), _p(idx_tensor), _s()), 2)


# sampling.cpp:11-20
def gather_points_wrapper(b, c, n, npoints, points_tensor, idx_tensor, out_tensor):
    with _on(points_tensor):
        return _ok(_l.spsk_gather_points(b, c, n, npoints, _p(points_tensor), _p(idx_tensor), _p(out_tensor), _s()))


# sampling.cpp:22-32
def gather_points_grad_wrapper(b, c, n, npoints, grad_out_tensor, idx_tensor, grad_points_tensor):
    with _on(grad_out_tensor):
        return _ok(_l.spsk_gather_points_grad(b, c, n, npoints, _p(grad_out_tensor), _p(idx_tensor), _p(grad_points_tensor), _s()))


# ball_query.cpp:32-42
def ball_query_wrapper(b, n, m, radius, nsample, new_xyz_tensor, xyz_tensor, idx_tensor):
    with _on(xyz_tensor):
        return _ok(_l.spsk_ball_query(b, n, m, float(radius), nsample, _p(new_xyz_tensor), _p(xyz_tensor), _p(idx_tensor), _s()))


# ball_query.cpp:45-56
def ball_query_dilated_wrapper(b, n, m, max_radius, min_radius, nsample, new_xyz_tensor, xyz_tensor, idx_tensor):
    with _on(xyz_tensor):
        return _ok(_l.spsk_ball_query_dilated(b, n, m, float(max_radius), float(min_radius), nsample, _p(new_xyz_tensor),
                                              _p(xyz_tensor), _p(idx_tensor), _s()))


# group_points.cpp:30-40
def group_points_wrapper(b, c, n, npoints, nsample, points_tensor, idx_tensor, out_tensor):
    with _on(points_tensor):
        return _ok(_l.spsk_group_points(b, c, n, npoints, nsample, _p(points_tensor), _p(idx_tensor), _p(out_tensor), _s()))


# group_points.cpp:18-28
def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out_tensor, idx_tensor, grad_points_tensor):
    with _on(grad_out_tensor):
        return _ok(_l.spsk_group_points_grad(b, c, n, npoints, nsample, _p(grad_out_tensor), _p(idx_tensor), _p(grad_points_tensor), _s()))


# interpolate.cpp:21-30 -> void
def three_nn_wrapper(b, n, m, unknown_tensor, known_tensor, dist2_tensor, idx_tensor):
    with _on(unknown_tensor):
        _ok(_l.spsk_three_nn(b, n, m, _p(unknown_tensor), _p(known_tensor), _p(dist2_tensor), _p(idx_tensor), _s()))


# interpolate.cpp:32-44 -> void
def three_interpolate_wrapper(b, c, m, n, points_tensor, idx_tensor, weight_tensor, out_tensor):
    with _on(points_tensor):
        _ok(_l.spsk_three_interpolate(b, c, m, n, _p(points_tensor), _p(idx_tensor), _p(weight_tensor), _p(out_tensor), _s()))


# interpolate.cpp:46-58 -> void
def three_interpolate_grad_wrapper(b, c, n, m, grad_out_tensor, idx_tensor, weight_tensor, grad_points_tensor):
    with _on(grad_out_tensor):
        _ok(_l.spsk_three_interpolate_grad(b, c, n, m, _p(grad_out_tensor), _p(idx_tensor), _p(weight_tensor), _p(grad_points_tensor), _s()))

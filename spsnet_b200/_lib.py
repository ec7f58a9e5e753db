"""ctypes binding of libspsk.so (include/spsk.h) -- the only way the Python layer reaches the kernels.

There is NO fallback: if the CUDA library is missing or does not export a symbol declared in
include/spsk.h, importing this module raises.  Build it with `python -m spsnet_b200.build`
(or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_ROOT = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("SPSK_LIB", _ROOT / "_C" / "libspsk.so"))

_f = C.c_float
_i = C.c_int
_p = C.c_void_p

# name -> argtypes (all return int unless listed in _RESTYPE)
SIGNATURES = {
    "spsk_last_error": [],
    "spsk_abi_version": [],
    "spsk_built_for_sm": [],
    "spsk_launch_count": [],
    "spsk_farthest_point_sampling": [_i, _i, _i, _p, _p, _p, _p],
    "spsk_fps_set_profile": [_p],
    "spsk_furthest_point_sampling_with_dist": [_i, _i, _i, _p, _p, _p, _p],
    "spsk_gather_points": [_i, _i, _i, _i, _p, _p, _p, _p],
    "spsk_gather_points_grad": [_i, _i, _i, _i, _p, _p, _p, _p],
    "spsk_ball_query": [_i, _i, _i, _f, _i, _p, _p, _p, _p],
    "spsk_ball_query_dilated": [_i, _i, _i, _f, _f, _i, _p, _p, _p, _p],
    "spsk_group_points": [_i, _i, _i, _i, _i, _p, _p, _p, _p],
    "spsk_group_points_grad": [_i, _i, _i, _i, _i, _p, _p, _p, _p],
    "spsk_three_nn": [_i, _i, _i, _p, _p, _p, _p, _p],
    "spsk_three_interpolate": [_i, _i, _i, _i, _p, _p, _p, _p, _p],
    "spsk_three_interpolate_grad": [_i, _i, _i, _i, _p, _p, _p, _p, _p],
    "spsk_scatter_grad_workspace_bytes": [_i, _i, C.c_longlong],
    "spsk_scatter_grad": [_i, _i, _i, C.c_longlong, _i, _i, _p, _p, _p, _p, _p, C.c_longlong, _p],
    "spsk_score_topk": [_i, _i, _i, _i, _p, _p, _p, _p, _p],
    "spsk_gather_rows": [_i, _i, _i, _i, _p, _p, _p, _p],
    "spsk_ball_query_msg": [_i, _i, _i, _i, C.POINTER(_f), C.POINTER(_i), _p, _p, C.POINTER(_p), _p],
    "spsk_ball_query_grid_workspace_bytes": [_i, _i],
    "spsk_ball_query_msg_grid": [_i, _i, _i, _i, C.POINTER(_f), C.POINTER(_i), _p, _p, C.POINTER(_p), _p, C.c_longlong, _p],
    "spsk_grouped_linear": [_p, _i, _p, _i, _p, _p, _i, _i, _i, _p, _p, _i, _i, _p],
    "spsk_pointwise_linear": [_i, _i, _p, _i, _p, _p, _i, _i, _p, _p],
    "spsk_make_twin": [_i, _i, _i, _i, _p, _p, _p],
    "spsk_sa_mma_config": [_p, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)],
    "spsk_sa_mma_forward": [_p, _p],
    "spsk_sa_mma_stats_parts": [_p, C.POINTER(_i)],
    "spsk_sa_mma_schedule": [_p, C.POINTER(_i), _p, _p],
    "spsk_sa_pack_layer": [_p, _i, _i, _p, _i, _i, _i, _i, _i, _i, _p, _p],
    "spsk_bn_stats_reduce": [_p, _i, _i, _i, C.c_double, _p, _p],
    "spsk_bn_stats_finalize": [_p, _i, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p],
    "spsk_bn_stats_reduce_finalize": [_p, _i, _i, _i, C.c_double, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p],
    "spsk_sa_mma_set_profile": [_p],
    "spsk_pw_mma_forward": [_p, _p],
    "spsk_fp16_overflow_poll": [C.POINTER(C.c_uint), _i],
    "spsk_boxes_overlap_bev": [_i, _p, _i, _p, _p, _p],
    "spsk_boxes_iou_bev": [_i, _p, _i, _p, _p, _p],
    "spsk_boxes_iou3d": [_i, _p, _i, _p, _p, _p],
    "spsk_nms_workspace_bytes": [_i, _i],
    "spsk_nms": [_i, _i, _p, _p, _f, _i, _p, _p, _p, C.c_longlong, _p],
    "spsk_detect_workspace_bytes": [_i, _i],
    "spsk_detect_postprocess": [_p, _p],
    "spsk_edge_conv_point": [_p, _i, _p, _i, _p, _p, _p],
    "spsk_edge_conv_aggregate": [_p, _i, _i, _i, _p, _p, _p, _p, _i, _p],
}
_RESTYPE = {"spsk_last_error": C.c_char_p, "spsk_launch_count": C.c_ulonglong,
            "spsk_ball_query_grid_workspace_bytes": C.c_longlong, "spsk_nms_workspace_bytes": C.c_longlong, "spsk_scatter_grad_workspace_bytes": C.c_longlong,
            "spsk_detect_workspace_bytes": C.c_longlong}


class GroupDesc(C.Structure):
    """struct spsk_group_desc (include/spsk.h)."""

    _fields_ = [
        ("b", _i), ("n", _i), ("m", _i), ("nsample", _i),
        ("c_feat", _i), ("use_xyz", _i),
        ("xyz", _p), ("new_xyz", _p), ("features", _p), ("idx", _p),
    ]


class SaMmaDesc(C.Structure):
    """struct spsk_sa_mma_desc (include/spsk.h)."""

    _fields_ = [
        ("b", _i), ("n", _i), ("m", _i), ("nsample", _i),
        ("xyz", _p), ("new_xyz", _p), ("idx", _p),
        ("use_xyz", _i), ("c_feat", _i),
        ("twin", _p), ("ldtwin", _i),
        ("features", _p), ("split", _i),
        ("nlayers", _i), ("kpad", _i * 4), ("cpad", _i * 4),
        ("wtiles", _p), ("bias", _p), ("cout_last", _i),
        ("out_cm", _p), ("c_total", _i), ("co_off", _i),
        ("out16", _p), ("ld16", _i), ("co16", _i), ("n16", _i), ("o16lo", _i), ("l0_fused", _i), ("pair", _i), ("ovf_tag", _i),
        ("stats", _p), ("stats_parts", _i),
    ]


class PwDesc(C.Structure):
    """struct spsk_pw_desc (include/spsk.h)."""

    _fields_ = [
        ("rows", _i), ("k", _i), ("ldx", _i), ("n", _i), ("relu", _i), ("split", _i), ("xlo", _i),
        ("x", _p), ("wtiles", _p), ("bias", _p),
        ("out_cm", _p), ("m", _i), ("c_total", _i), ("co_off", _i),
        ("out16", _p), ("ld16", _i), ("n16", _i), ("o16lo", _i),
        ("out_pm", _p), ("ldpm", _i), ("ovf_tag", _i),
    ]


class DetectDesc(C.Structure):
    """struct spsk_detect_desc (include/spsk.h)."""

    _fields_ = [
        ("batch", _i), ("m", _i), ("num_class", _i), ("bin_size", _i),
        ("cls", _p), ("ld_cls", _i),
        ("reg", _p), ("ld_reg", _i),
        ("centers", _p), ("ld_centers", _i),
        ("mean_size", _p),
        ("score_thresh", _f), ("nms_thresh", _f),
        ("nms_normal", _i), ("pre_max", _i), ("post_max", _i),
        ("box_preds", _p), ("scores", _p), ("labels", _p),
        ("out_boxes", _p), ("out_scores", _p), ("out_labels", _p), ("out_index", _p), ("out_count", _p),
        ("workspace", _p), ("workspace_bytes", C.c_longlong),
    ]


class EdgePointWeights(C.Structure):
    """struct spsk_edge_point_weights (include/spsk.h)."""

    _fields_ = [("cin", _i), ("relu", _i), ("wt", _f * (64 * 24)), ("bt", _f * 24), ("m", _f * (24 * 48)), ("c", _f * 48)]


class EdgeAggrWeights(C.Structure):
    """struct spsk_edge_aggr_weights (include/spsk.h)."""

    _fields_ = [("w2a", _f * 144), ("w3a", _f * 144), ("w3b", _f * 144)]


def declared_symbols(header: Path | None = None) -> list[str]:
    """Every function name include/spsk.h declares (used by the CPU tests to check the exports)."""
    import re

    header = header or (_ROOT.parent / "include" / "spsk.h")
    txt = header.read_text()
    return sorted(set(re.findall(r"SPSK_API\s+[\w\s\*]+?\b(spsk_\w+)\s*\(", txt)))


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"spsnet_b200: CUDA library {LIB_PATH} not found. There is no CPU fallback; "
            "build it with `python -m spsnet_b200.build`."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, argtypes in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # pragma: no cover
            raise ImportError(f"spsnet_b200: {LIB_PATH} does not export {name}") from e
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, _i)
    return lib


lib = _load()


class SpskError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    """Turn a negative spsk_status into an exception (the reference exit(-1)s instead)."""
    if rc != 0:
        msg = lib.spsk_last_error()
        raise SpskError(f"{what or 'spsk'} failed with status {rc}: {msg.decode() if msg else ''}")

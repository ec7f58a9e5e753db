"""Shared test helpers (CPU-only imports at module level)."""
import copy

import numpy as np
import torch

REL_TOL = 1e-3  # BASELINE.json north_star: grouped features / MLP outputs within 1e-3 relative


def rel_err(got, want):
    """max |got - want| relative to the largest magnitude of the reference tensor."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    scale = max(np.abs(want).max(), 1e-30)
    return float(np.abs(got - want).max() / scale)


def assert_close(got, want, tol=REL_TOL, what=""):
    assert np.asarray(got).shape == np.asarray(want).shape, f"{what}: shape {np.asarray(got).shape} vs {np.asarray(want).shape}"
    e = rel_err(got, want)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"


def small_sa_cfg(npoints=(512, 128, 64, 32)):
    """IA-SSD KITTI SA_CONFIG (same radii / nsample / widths) with fewer points, for second-scale tests."""
    from spsnet_b200.backbone import KITTI_IASSD_SA_CONFIG, Cfg

    c = copy.deepcopy(KITTI_IASSD_SA_CONFIG)
    c["NPOINT_LIST"] = [[npoints[0]], [npoints[1]], [npoints[2]], [npoints[3]], [-1], [npoints[3]]]
    return Cfg({"SA_CONFIG": c})


def make_backbone(cfg, input_channels=4, num_class=3, seed=0, cls=None):
    from spsnet_b200 import backbone as bb

    torch.manual_seed(seed)
    net = (cls or bb.IASSD_Backbone)(cfg, num_class=num_class, input_channels=input_channels)
    bb.randomize_bn_stats(net, seed=seed)
    return net.eval()

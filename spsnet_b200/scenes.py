"""Synthetic LiDAR-shaped scenes: the workload generator shared by tests and bench.

KITTI-shaped: N = 16384 rows [x, y, z, intensity], range x in [0, 70.4], y in [-40, 40], z in [-3, 1]
(reference tools/cfgs/dataset_configs/kitti_dataset.yaml:4); half uniform clutter, half a LiDAR-like
radial profile (range ~ 1/u, ground plane near z = -1.7, a few dozen car-sized boxes of dense returns);
~3 % exact duplicate rows, mimicking `sample_points` padding short clouds by re-drawing existing points
(reference pcdet/datasets/processor/data_processor.py:243-247) -- duplicates exercise every tie-break;
rows shuffled.  Waymo-shaped: N = 65536, 5 columns, x, y in [-75.2, 75.2], z in [-2, 4]
(waymo_dataset.yaml:5).  Deterministic: seed = 1234 + scene_id.
"""
from __future__ import annotations

import numpy as np

KITTI_RANGE = (0.0, -40.0, -3.0, 70.4, 40.0, 1.0)
WAYMO_RANGE = (-75.2, -75.2, -2.0, 75.2, 75.2, 4.0)


def make_scene(scene_id: int, n: int = 16384, kind: str = "kitti", dup_frac: float = 0.03) -> np.ndarray:
    rng = np.random.default_rng(1234 + scene_id)
    x0, y0, z0, x1, y1, z1 = KITTI_RANGE if kind == "kitti" else WAYMO_RANGE
    n_feat = 1 if kind == "kitti" else 2
    n_dup = int(round(n * dup_frac))
    n_base = n - n_dup
    n_uni = n_base // 2
    n_lidar = n_base - n_uni
    # uniform clutter
    uni = np.stack([rng.uniform(x0, x1, n_uni), rng.uniform(y0, y1, n_uni), rng.uniform(z0, z1, n_uni)], axis=1)
    # LiDAR-like: range ~ 1/u (dense near the sensor), ground plane + boxes
    n_box_pts = n_lidar // 3
    n_ground = n_lidar - n_box_pts
    rmax = 0.5 * np.hypot(x1 - x0, y1 - y0)
    r = np.minimum(2.5 / rng.uniform(0.03, 1.0, n_ground), rmax)
    if kind == "kitti":
        phi = rng.uniform(-0.25 * np.pi, 0.25 * np.pi, n_ground)
        gx, gy = r * np.cos(phi), r * np.sin(phi)
    else:
        phi = rng.uniform(-np.pi, np.pi, n_ground)
        gx, gy = r * np.cos(phi), r * np.sin(phi)
    gz = -1.7 + 0.05 * rng.standard_normal(n_ground)
    ground = np.stack([gx, gy, gz], axis=1)
    n_boxes = 40
    centres = np.stack([rng.uniform(x0 + 5, x1 - 5, n_boxes), rng.uniform(y0 + 5, y1 - 5, n_boxes),
                        np.full(n_boxes, -0.9)], axis=1)
    which = rng.integers(0, n_boxes, n_box_pts)
    half = np.array([2.0, 0.8, 0.75])
    boxes = centres[which] + rng.uniform(-1.0, 1.0, (n_box_pts, 3)) * half
    pts = np.concatenate([uni, ground, boxes], axis=0)
    pts[:, 0] = np.clip(pts[:, 0], x0, x1)
    pts[:, 1] = np.clip(pts[:, 1], y0, y1)
    pts[:, 2] = np.clip(pts[:, 2], z0, z1)
    feat = rng.uniform(0.0, 1.0, (n_base, n_feat))
    rows = np.concatenate([pts, feat], axis=1).astype(np.float32)
    if n_dup:
        rows = np.concatenate([rows, rows[rng.integers(0, n_base, n_dup)]], axis=0)
    rng.shuffle(rows, axis=0)
    return np.ascontiguousarray(rows)


def make_batch(first_scene: int, batch: int, n: int = 16384, kind: str = "kitti") -> np.ndarray:
    """(batch, n, 3 + C) float32."""
    return np.stack([make_scene(first_scene + i, n, kind) for i in range(batch)], axis=0)


def to_points(batch_arr: np.ndarray) -> np.ndarray:
    """(B, N, 3+C) -> OpenPCDet `points` layout (B*N, 1+3+C) with the batch index prepended
    (reference pcdet/datasets/dataset.py collate_batch)."""
    B, N, C = batch_arr.shape
    bidx = np.repeat(np.arange(B, dtype=np.float32), N)[:, None]
    return np.concatenate([bidx, batch_arr.reshape(B * N, C)], axis=1)


def make_cls_logits(seed: int, b: int, n: int, num_class: int = 3) -> np.ndarray:
    """N(0, 2) logits for stand-alone top-k tests."""
    return (np.random.default_rng(seed).standard_normal((b, n, num_class)) * 2.0).astype(np.float32)


def make_stds(seed: int, b: int, n: int) -> np.ndarray:
    """8 * exp(N(0, 0.5)): the stability generator sums 8 exp(0.5*logvar) terms
    (reference stability_generate/model.py:577)."""
    return (8.0 * np.exp(0.5 * np.random.default_rng(seed).standard_normal((b, n)))).astype(np.float32)


def make_boxes(seed: int, n: int, n_objects: int | None = None, extent: float = 35.0) -> np.ndarray:
    """Detection-shaped boxes (n, 7) [x, y, z, dx, dy, dz, heading] for the IoU / NMS tests and benches: centres
    cluster around `n_objects` object locations (a detector proposes many near-duplicate boxes per object), sizes are
    jittered KITTI mean sizes (car / pedestrian / cyclist, reference IA-SSD.yaml:79-83), headings arbitrary, plus a few
    exact duplicates and a few axis-aligned boxes (degenerate clipping cases).  Deterministic in (seed, n)."""
    rng = np.random.default_rng(9000 + seed)
    if n == 0:
        return np.zeros((0, 7), np.float32)
    n_objects = n_objects or max(1, n // 8)
    means = np.array([[3.9, 1.6, 1.56], [0.8, 0.6, 1.73], [1.76, 0.6, 1.73]], np.float32)
    obj_xy = rng.uniform(-extent, extent, (n_objects, 2))
    obj_cls = rng.integers(0, 3, n_objects)
    obj_ang = rng.uniform(-np.pi, np.pi, n_objects)
    which = rng.integers(0, n_objects, n)
    boxes = np.zeros((n, 7), np.float64)
    boxes[:, :2] = obj_xy[which] + rng.normal(0.0, 0.6, (n, 2))
    boxes[:, 2] = rng.uniform(-1.5, 0.0, n)
    boxes[:, 3:6] = means[obj_cls[which]] * rng.uniform(0.8, 1.25, (n, 3))
    boxes[:, 6] = obj_ang[which] + rng.normal(0.0, 0.25, n) + np.pi * rng.integers(0, 2, n)
    k = max(1, n // 16)
    boxes[rng.integers(0, n, k)] = boxes[rng.integers(0, n, k)]       # exact duplicates
    boxes[rng.integers(0, n, k), 6] = 0.0                            # axis-aligned
    boxes[rng.integers(0, n, k), 6] = np.float32(np.pi / 2)
    return boxes.astype(np.float32)

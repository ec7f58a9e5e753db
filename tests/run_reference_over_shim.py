"""Helper process of tests/test_gpu_shim.py: the reference's UNMODIFIED python layer (oracle/_ref/pcdet: pointnet2_utils.py,
pointnet2_modules.py, IASSD_backbone.py, iou3d_nms_utils.py) running on libspsk.so through spsnet_b200/shims.  It runs in its
own process so that the rebuilt reference extension is never imported next to the shim.

    python tests/run_reference_over_shim.py OUT.npz
"""
import sys
import types
import warnings
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))


def main(out_path):
    from spsnet_b200 import configs, scenes, shims

    shims.install()
    sys.modules.setdefault("SharedArray", types.ModuleType("SharedArray"))   # absent dependency of pcdet.utils.common_utils, unused
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from pcdet.models.backbones_3d import IASSD_backbone
        from pcdet.ops.iou3d_nms import iou3d_nms_utils
        from pcdet.ops.pointnet2.pointnet2_batch import pointnet2_batch_cuda, pointnet2_utils
    assert pointnet2_batch_cuda.__file__.endswith("shims/pointnet2_batch_cuda.py"), pointnet2_batch_cuda.__file__
    loaded = [ln.split()[-1] for ln in open("/proc/self/maps") if ln.rstrip().endswith(".so") and ("oracle/_ref" in ln or "libspsk" in ln)]
    assert not any("oracle/_ref" in p for p in loaded), f"the rebuilt reference extension is mapped: {set(loaded)}"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    out = {}
    B, N = 2, 4096
    cfg = configs.Cfg({"SA_CONFIG": {**configs.KITTI_IASSD_SA_CONFIG, "NPOINT_LIST": [[1024], [256], [128], [64], [-1], [64]]}})
    torch.manual_seed(4)
    net = IASSD_backbone.IASSD_Backbone(cfg, num_class=3, input_channels=4)
    configs.randomize_bn_stats(net, seed=4)
    net = net.cuda().eval()
    pts = torch.from_numpy(scenes.to_points(scenes.make_batch(90, B, N))).cuda()
    with torch.no_grad():
        res = net({"batch_size": B, "points": pts})
    for i, t in enumerate(res["encoder_xyz"]):
        out[f"encoder_xyz_{i}"] = t.cpu().numpy()
    out["centers_features"] = res["centers_features"].cpu().numpy()
    out["centers"] = res["centers"].cpu().numpy()
    # op level, through the reference's autograd Functions (legacy torch.cuda.*Tensor allocations included)
    xyz = pts[:, 1:4].reshape(B, N, 3).contiguous()
    idx = pointnet2_utils.furthest_point_sample(xyz, 300)
    out["fps"] = idx.cpu().numpy()
    new_xyz = pointnet2_utils.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    out["ball"] = pointnet2_utils.ball_query(0.8, 16, xyz, new_xyz).cpu().numpy()
    out["ball_dilated"] = pointnet2_utils.ball_query_dilated(1.6, 0.8, 16, xyz, new_xyz).cpu().numpy()
    f = torch.randn(B, 6, N, generator=torch.Generator().manual_seed(1)).cuda().requires_grad_(True)
    g = pointnet2_utils.grouping_operation(f, torch.from_numpy(out["ball"]).cuda())
    g.sum().backward()
    out["group"], out["group_grad"] = g.detach().cpu().numpy(), f.grad.cpu().numpy()
    d, i3 = pointnet2_utils.three_nn(xyz, new_xyz)
    out["three_nn_d"], out["three_nn_i"] = d.cpu().numpy(), i3.cpu().numpy()
    w = torch.softmax(-d, dim=-1).contiguous()
    out["three_interp"] = pointnet2_utils.three_interpolate(f.detach()[:, :, :300].contiguous(), i3, w).cpu().numpy()
    dist = torch.cdist(xyz[:, :512], xyz[:, :512]).pow(2).contiguous()
    out["ffps"] = pointnet2_utils.furthest_point_sample_with_dist(dist, 64).cpu().numpy()
    boxes = torch.from_numpy(scenes.make_boxes(3, 300)).cuda()
    scores = torch.linspace(1, 0, 300).cuda()
    keep, _ = iou3d_nms_utils.nms_gpu(boxes, scores, 0.1)
    out["nms_keep"] = keep.cpu().numpy()
    out["iou_bev"] = iou3d_nms_utils.boxes_iou_bev(boxes[:50], boxes[50:120]).cpu().numpy()
    out["iou3d"] = iou3d_nms_utils.boxes_iou3d_gpu(boxes[:50], boxes[50:120]).cpu().numpy()
    np.savez(out_path, **out)
    print("ok", sorted(set(loaded)))


if __name__ == "__main__":
    main(sys.argv[1])

"""Op-level drop-ins: pure-Python modules with the EXACT names and signatures of the reference's pybind11 extensions,
bound to libspsk.so through ctypes.  Put (or inject) one in place of the compiled extension and the reference's own,
unmodified Python layer runs on the sm_100a kernels:

    pointnet2_batch_cuda  <-  pcdet/ops/pointnet2/pointnet2_batch/src/pointnet2_api.cpp:10-26   (11 functions)
    iou3d_nms_cuda        <-  pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:9-15                    (4 GPU functions)

`install()` registers them in sys.modules under the reference's module paths (tests/test_gpu_shim.py runs the reference's
`pointnet2_utils.py` / `pointnet2_modules.py` / `IASSD_backbone.py` that way).
"""
from __future__ import annotations

import sys


def install(pointnet2: bool = True, iou3d: bool = True) -> None:
    """Make `from pcdet.ops.pointnet2.pointnet2_batch import pointnet2_batch_cuda` (pointnet2_utils.py:7) and
    `from pcdet.ops.iou3d_nms import iou3d_nms_cuda` (iou3d_nms_utils.py:9) resolve to the shims.  Call before the
    reference packages are imported."""
    if pointnet2:
        from . import pointnet2_batch_cuda

        sys.modules["pcdet.ops.pointnet2.pointnet2_batch.pointnet2_batch_cuda"] = pointnet2_batch_cuda
    if iou3d:
        from . import iou3d_nms_cuda

        sys.modules["pcdet.ops.iou3d_nms.iou3d_nms_cuda"] = iou3d_nms_cuda

import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _require_ref() -> bool:
    """The `if ref_ops is not None` comparisons against the rebuilt reference must not vanish silently: with a CUDA device
    present the fixtures FAIL when oracle/_ref is missing or broken (SPSK_REQUIRE_REF=0 opts out; =1 forces it anywhere)."""
    v = os.environ.get("SPSK_REQUIRE_REF")
    if v is not None:
        return v == "1"
    return _has_gpu()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def ref_ops():
    """The reference's own pointnet2_batch (rebuilt unmodified for sm_100a by oracle/build_ref.sh), or None."""
    ref_root = ROOT / "oracle" / "_ref"
    so = ref_root / "pcdet" / "ops" / "pointnet2" / "pointnet2_batch" / "pointnet2_batch_cuda.so"
    if not so.exists():
        if _require_ref():
            pytest.fail(f"{so} is missing: the reference comparison is mandatory on a GPU box (build it with "
                        "oracle/build_ref.sh where /root/reference exists, or set SPSK_REQUIRE_REF=0 to skip it knowingly)")
        return None
    import importlib
    import warnings

    if str(ref_root) not in sys.path:
        sys.path.insert(0, str(ref_root))
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            utils = importlib.import_module("pcdet.ops.pointnet2.pointnet2_batch.pointnet2_utils")
            modules = importlib.import_module("pcdet.ops.pointnet2.pointnet2_batch.pointnet2_modules")
    except Exception as e:  # pragma: no cover
        if _require_ref():
            pytest.fail(f"reference extension in oracle/_ref is not importable: {e}")
        print("reference ext not importable:", e)
        return None

    class R:
        pass

    R.utils, R.modules = utils, modules
    return R


@pytest.fixture(scope="session")
def ref_det():
    """The reference's own iou3d_nms extension, IASSD_Head, box coder and class_agnostic_nms (oracle/_ref, unmodified,
    rebuilt by oracle/build_ref.sh), or None.  `SharedArray` -- an absent dependency of pcdet.utils.common_utils that
    nothing on this path uses -- is satisfied with an empty module."""
    ref_root = ROOT / "oracle" / "_ref"
    if not (ref_root / "pcdet" / "ops" / "iou3d_nms" / "iou3d_nms_cuda.so").exists():
        if _require_ref():
            pytest.fail("oracle/_ref/pcdet/ops/iou3d_nms/iou3d_nms_cuda.so is missing (oracle/build_ref.sh); SPSK_REQUIRE_REF=0 skips knowingly")
        return None
    import importlib
    import types
    import warnings

    if str(ref_root) not in sys.path:
        sys.path.insert(0, str(ref_root))
    sys.modules.setdefault("SharedArray", types.ModuleType("SharedArray"))
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            R = types.SimpleNamespace(
                cuda=importlib.import_module("pcdet.ops.iou3d_nms.iou3d_nms_cuda"),
                utils=importlib.import_module("pcdet.ops.iou3d_nms.iou3d_nms_utils"),
                nms_utils=importlib.import_module("pcdet.models.model_utils.model_nms_utils"),
                head=importlib.import_module("pcdet.models.dense_heads.IASSD_head"),
                coder=importlib.import_module("pcdet.utils.box_coder_utils"),
            )
    except Exception as e:  # pragma: no cover
        if _require_ref():
            pytest.fail(f"reference detection modules in oracle/_ref are not importable: {e}")
        print("reference detection modules not importable:", e)
        return None
    return R

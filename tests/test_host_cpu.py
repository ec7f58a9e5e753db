"""CPU tests of the host-side mirror of the reference interface: constructors, state_dict layout, BN folding,
config plumbing, the synthetic workload generator, the reference-arm JSON line."""
import copy
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from helpers import make_backbone, small_sa_cfg

ROOT = Path(__file__).resolve().parents[1]
KEYS = ROOT / "tests" / "golden" / "iassd_backbone_state_dict_keys.json"


def test_state_dict_layout_matches_reference():
    """Keys + shapes of the reference IASSD_Backbone (KITTI cfg), recorded by instantiating the reference's own
    class (tests/golden/make_state_dict_keys.py) -> checkpoints load unchanged."""
    from spsnet_b200 import backbone as bb

    net = make_backbone(bb.kitti_iassd_cfg())
    want = json.loads(KEYS.read_text())
    got = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert list(got.keys()) == list(want.keys())
    assert got == want
    assert sum(p.numel() for p in net.parameters()) == 2293897
    assert [sum(p.numel() for p in m.parameters()) for m in net.SA_modules] == [10688, 90435, 381699, 0, 33411, 1777664]


def test_reference_class_keys_if_present():
    ref_root = ROOT / "oracle" / "_ref"
    if not (ref_root / "pcdet" / "models" / "backbones_3d" / "IASSD_backbone.py").exists():
        pytest.skip("oracle/_ref not installed")
    import importlib
    import warnings

    sys.path.insert(0, str(ref_root))
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mod = importlib.import_module("pcdet.models.backbones_3d.IASSD_backbone")
    except Exception as e:
        pytest.skip(f"reference module not importable here: {e}")
    from spsnet_b200 import backbone as bb

    ref = mod.IASSD_Backbone(bb.kitti_iassd_cfg(), num_class=3, input_channels=4)
    mine = make_backbone(bb.kitti_iassd_cfg())
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    mine.load_state_dict(ref.state_dict())


def test_bn_fold_matches_torch():
    from spsnet_b200 import backbone as bb
    from spsnet_b200.pointnet2_modules import _Folded, _conv_bn_relu_2d

    torch.manual_seed(0)
    seq = _conv_bn_relu_2d([7, 12, 5])
    bb.randomize_bn_stats(seq, seed=3)
    seq.eval()
    x = torch.randn(3, 7, 11, 4)
    with torch.no_grad():
        want = seq(x)
        h = x.permute(0, 2, 3, 1).reshape(-1, 7)
        for wt, bias, relu in _Folded().get(seq):
            h = h @ wt + bias
            if relu:
                h = torch.relu(h)
        got = h.reshape(3, 11, 4, 5).permute(0, 3, 1, 2)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    # cache invalidation on in-place weight updates
    f = _Folded()
    a = f.get(seq)
    assert f.get(seq) is a
    with torch.no_grad():
        seq[0].weight.mul_(2.0)
    assert f.get(seq) is not a


def test_constructor_quirks_preserved():
    from spsnet_b200 import pointnet2_modules as pm

    spec = [[4, 8, 16]]
    pm.PointnetSAModuleMSG(npoint=8, radii=[1.0], nsamples=[4], mlps=spec)
    assert spec[0][0] == 7  # the reference mutates the caller's list (pointnet2_modules.py:117-118)
    v = pm.Vote_layer(mlp_list=[64, 32], pre_channel=16, max_translate_range=[3.0, 3.0, 2.0])
    assert v.mlp_modules[0].in_channels == 64 and v.mlp_modules[0].out_channels == 32  # only the last entry survives
    m = pm.PointnetSAModuleMSG_WithSampling(npoint_list=[4], sample_range_list=[-1], sample_type_list=["nope"],
                                            radii=[], nsamples=[], mlps=[], aggregation_mlp=None, confidence_mlp=None, num_class=3)
    with pytest.raises(NotImplementedError):
        m._sample_one("nope", 2, None, torch.zeros(1, 5, 3), None, None, None, None)


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: the op layer fails loudly on non-CUDA inputs."""
    from spsnet_b200 import pointnet2_utils as pu

    with pytest.raises(RuntimeError, match="CUDA"):
        pu.furthest_point_sample(torch.zeros(1, 8, 3), 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        pu.ball_query(1.0, 4, torch.zeros(1, 8, 3), torch.zeros(1, 2, 3))


def test_scenes_deterministic_with_duplicates():
    from spsnet_b200 import scenes

    a = scenes.make_scene(3)
    b = scenes.make_scene(3)
    assert a.shape == (16384, 4) and a.dtype == np.float32
    np.testing.assert_array_equal(a, b)
    assert not np.array_equal(a, scenes.make_scene(4))
    uniq = np.unique(a, axis=0).shape[0]
    assert 0.02 * 16384 < 16384 - uniq < 0.04 * 16384  # ~3 % exact duplicate rows
    assert a[:, 0].min() >= 0 and a[:, 0].max() <= 70.4 and a[:, 2].min() >= -3 and a[:, 2].max() <= 1
    w = scenes.make_scene(0, 65536, "waymo")
    assert w.shape == (65536, 5)
    pts = scenes.to_points(scenes.make_batch(0, 2, 128))
    assert pts.shape == (256, 5) and set(pts[:, 0]) == {0.0, 1.0}


def test_cfg_access():
    from spsnet_b200.backbone import Cfg, kitti_spsnet_cfg, waymo_iassd_cfg

    c = Cfg({"SA_CONFIG": {"A": [1], "B": {"C": 2}}})
    assert c.SA_CONFIG.A == [1] and c.SA_CONFIG.B.C == 2 and c.SA_CONFIG.get("Z", 5) == 5
    assert kitti_spsnet_cfg().SA_CONFIG.SAMPLE_METHOD_LIST[2] == ["sss_aware"]
    assert waymo_iassd_cfg().SA_CONFIG.NPOINT_LIST[0] == [16384]


def test_oracle_backbone_runs_and_is_deterministic(oracle):
    net = make_backbone(small_sa_cfg((256, 64, 32, 16)), seed=5)
    from spsnet_b200 import scenes

    pts = scenes.make_batch(3, 2, 1024)
    a = oracle.backbone_forward(copy.deepcopy(net), pts)
    b = oracle.backbone_forward(copy.deepcopy(net), pts, dtype=torch.float32)
    assert a["centers_features"].shape == (2 * 16, 512)
    np.testing.assert_array_equal(a["encoder_xyz"][1], b["encoder_xyz"][1])
    assert np.abs(a["encoder_features"][1] - b["encoder_features"][1]).max() < 1e-3


def test_bench_reference_cpu_line():
    """bench.py --impl reference-cpu prints one well-formed JSON line (bounded sample of the workload)."""
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference-cpu", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
              "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in line
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_stability_generator_state_dict_layout():
    """Generator checkpoints (reference stability_generate/model.py) address feature_extract.SA_modules.*,
    feature_encoder.fc{1,2}.* and obj_encoder.*: same key layout here."""
    from spsnet_b200 import stability as st

    gen = st.Generate_center(st.sf_unc_cfg())
    keys = list(gen.state_dict().keys())
    assert "feature_extract.SA_modules.0.mlps.0.0.weight" in keys and "feature_extract.SA_modules.0.mlps.1.7.running_var" in keys
    assert "feature_extract.SA_modules.0.aggregation_layer.0.weight" in keys
    for k in ("feature_encoder.fc1.weight", "feature_encoder.fc2.bias", "obj_encoder.fc1.weight", "obj_encoder.fc_ce2.weight", "global_step"):
        assert k in keys
    assert gen.feature_encoder.fc2.weight.shape == (8, 64) and gen.obj_encoder.fc1.weight.shape == (64, 72)
    with __import__("pytest").raises(NotImplementedError):
        gen.train()({"batch_size": 1, "points": None})


def test_torch_library_ops_registered_with_fake_impls():
    """every op of the thin custom-op layer is visible to the dispatcher (torch.ops.spsk.*), infers its
    output shapes/dtypes on meta tensors (what torch.compile / FakeTensor tracing use) and refuses CPU tensors loudly."""
    import pytest
    import torch

    import spsnet_b200.torch_ops as T

    for name in T.OPS:
        assert hasattr(torch.ops.spsk, name), name
    xyz = torch.empty(2, 100, 3, device="meta")
    ctr = torch.empty(2, 10, 3, device="meta")
    f = torch.empty(2, 8, 100, device="meta")
    i2 = torch.empty(2, 10, dtype=torch.int32, device="meta")
    i3 = torch.empty(2, 10, 16, dtype=torch.int32, device="meta")
    o = torch.ops.spsk.furthest_point_sample(xyz, 10)
    assert o.shape == (2, 10) and o.dtype == torch.int32
    assert torch.ops.spsk.gather_points(f, i2).shape == (2, 8, 10)
    assert torch.ops.spsk.gather_rows(xyz, i2).shape == (2, 10, 3)
    assert torch.ops.spsk.group_points(f, i3).shape == (2, 8, 10, 16)
    assert torch.ops.spsk.ball_query(0.5, 16, xyz, ctr).shape == (2, 10, 16)
    assert torch.ops.spsk.ball_query_dilated(0.5, 0.1, 16, xyz, ctr).shape == (2, 10, 16)
    d, i = torch.ops.spsk.three_nn(ctr, xyz)
    assert d.shape == (2, 10, 3) and i.dtype == torch.int32
    assert torch.ops.spsk.three_interpolate(f, torch.empty(2, 10, 3, dtype=torch.int32, device="meta"), d).shape == (2, 8, 10)
    assert torch.ops.spsk.score_topk(torch.empty(2, 100, 3, device="meta"), 5).shape == (2, 5)
    with pytest.raises(RuntimeError):
        torch.ops.spsk.furthest_point_sample(torch.zeros(2, 100, 3), 10)   # no CPU fallback


def test_shims_export_the_reference_pybind_names():
    """spsnet_b200/shims: exactly the names of pointnet2_api.cpp:10-26 and the GPU names of iou3d_nms_api.cpp:12-15."""
    from spsnet_b200.shims import iou3d_nms_cuda, pointnet2_batch_cuda

    want = {"ball_query_wrapper", "ball_query_dilated_wrapper", "group_points_wrapper", "group_points_grad_wrapper", "gather_points_wrapper",
            "gather_points_grad_wrapper", "farthest_point_sampling_wrapper", "furthest_point_sampling_with_dist_wrapper", "three_nn_wrapper",
            "three_interpolate_wrapper", "three_interpolate_grad_wrapper"}
    assert {n for n in dir(pointnet2_batch_cuda) if n.endswith("_wrapper")} == want
    for n in ("boxes_overlap_bev_gpu", "boxes_iou_bev_gpu", "nms_gpu", "nms_normal_gpu"):
        assert callable(getattr(iou3d_nms_cuda, n))


def test_configs_module_does_not_load_the_library():
    """bench.py's reference arm imports spsnet_b200.configs / scenes only: neither may dlopen libspsk.so."""
    import subprocess
    import sys as _sys

    code = ("import sys; sys.path.insert(0, %r); import spsnet_b200.configs, spsnet_b200.scenes; "
            "assert 'spsnet_b200._lib' not in sys.modules; "
            "assert not any('libspsk' in l for l in open('/proc/self/maps')); print('clean')" % str(ROOT))
    r = subprocess.run([_sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "clean" in r.stdout, r.stderr

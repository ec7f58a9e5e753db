"""parity.py -- TEST INFRASTRUCTURE ONLY: the teacher-forced parity check of a whole SA backbone against the UNMODIFIED
reference (oracle/_ref: its pointnet2_batch CUDA ops rebuilt for sm_100a + its own pointnet2_modules.py and
IASSD_backbone.py / PAGNet_backbone.py).  Used by tests/test_gpu_timed_path.py and by `bench.py --verify-only` (the
verdict `bench.py` prints as `"verified"`), never by the product.

Why teacher-forced: the two D-FPS layers are compared end to end (bit-exact).  From the first score-based layer on, the
sampled ORDER depends on the last bits of the confidence logits (cuDNN vs our GEMM summation order), so an end-to-end index
comparison is ill-posed for ANY fp32 implementation; every layer is therefore ALSO fed exactly the tensors the reference
module received (reference IASSD_backbone.py:128-148) and must return bit-identical sample indices / new_xyz and features
within the tolerance (north_star: 1e-3 relative).
"""
from __future__ import annotations

import importlib
import sys
import warnings
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
REF_ROOT = ROOT / "oracle" / "_ref"
REL_TOL = 1e-3


def rel_err(got, want) -> float:
    """max |got - want| relative to the largest magnitude of the reference tensor (range-relative)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


def elem_err(got, want) -> float:
    """Elementwise bar with a floor: max over elements of |got - want| / (|want| + rms(want)).  Small channels are not
    hidden behind the tensor's largest value; the rms floor keeps exact zeros (ReLU outputs) from dividing by zero."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    rms = max(float(np.sqrt(np.mean(want * want))), 1e-30)
    return float((np.abs(got - want) / (np.abs(want) + rms)).max())


def reference_backbone(cls_name: str, cfg, input_channels: int, num_class: int = 3):
    """Instantiate the reference's own backbone class from oracle/_ref (no weights loaded)."""
    so = REF_ROOT / "pcdet" / "ops" / "pointnet2" / "pointnet2_batch" / "pointnet2_batch_cuda.so"
    if not so.exists():
        raise RuntimeError(f"{so} missing: run oracle/build_ref.sh where /root/reference exists")
    if str(REF_ROOT) not in sys.path:
        sys.path.insert(0, str(REF_ROOT))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module("pcdet.models.backbones_3d." + cls_name.replace("_Backbone", "_backbone"))
    return getattr(mod, cls_name)(cfg, num_class=num_class, input_channels=input_channels)


class _Fp32Reference:
    """cuDNN / cuBLAS TF32 off: the reference evaluated in true fp32 is the truth the tolerance is stated against."""

    def __enter__(self):
        self.old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

    def __exit__(self, *a):
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = self.old


def teacher_forced_check(net, ref, batch_size: int, points: torch.Tensor, extra: dict | None = None,
                         feat_tol: float = REL_TOL, logit_tol: float = REL_TOL, fps_layers=(1, 2), min_overlap: float = 0.9) -> dict:
    """`net` (ours) and `ref` (reference backbone) hold the same weights, both on the GPU in eval mode.  Raises
    AssertionError on the first violated bar; returns the measured errors."""
    extra = extra or {}
    captured = {}

    def mk_hook(i):
        def hook(mod, args, kwargs, output):
            captured[i] = (args, kwargs, output)
        return hook

    hooks = [m.register_forward_hook(mk_hook(i), with_kwargs=True) for i, m in enumerate(ref.SA_modules)]
    report = {"layers": {}, "fps_layers_bit_exact": [], "sampled_set_overlap": {}}
    try:
        with _Fp32Reference(), torch.no_grad():
            want = ref({"batch_size": batch_size, "points": points.clone(), **extra})
            got = net({"batch_size": batch_size, "points": points.clone(), **extra})
            for i in fps_layers:   # D-FPS layers: exact end to end
                a, b = got["encoder_xyz"][i], want["encoder_xyz"][i]
                assert a.shape == b.shape and torch.equal(a, b), f"encoder_xyz[{i}] (D-FPS sampling) differs from the reference"
                report["fps_layers_bit_exact"].append(i)
            # end to end after score-based sampling: same point SET up to a few near-tie swaps
            for i in range(max(fps_layers) + 1, len(want["encoder_xyz"])):
                if i - 1 < len(net.layer_types) and net.layer_types[i - 1] != "SA_Layer":
                    continue
                if net.ctr_idx_list[i - 1] != -1:
                    continue
                ga = got["encoder_xyz"][i].reshape(batch_size, -1, 3)
                wa = want["encoder_xyz"][i].reshape(batch_size, -1, 3)
                inter = tot = 0
                for s in range(batch_size):
                    a = {tuple(r) for r in ga[s].cpu().numpy().round(4).tolist()}
                    b = {tuple(r) for r in wa[s].cpu().numpy().round(4).tolist()}
                    inter += len(a & b)
                    tot += len(b)
                report["sampled_set_overlap"][i] = inter / max(tot, 1)
                assert inter >= min_overlap * tot, f"encoder_xyz[{i}]: sampled sets diverge ({inter}/{tot})"
            # teacher-forced, layer by layer
            for i, mod in enumerate(net.SA_modules):
                args, kwargs, out = captured[i]
                mine = mod(*args, **kwargs)
                rec = report["layers"].setdefault(i, {})
                for j, (g, w) in enumerate(zip(mine, out)):
                    if not (isinstance(w, torch.Tensor) and w.numel() > 0):
                        continue
                    if w.dtype in (torch.int32, torch.int64):
                        assert torch.equal(g.to(w.dtype), w), f"layer {i} output {j} (sampled indices) differs from the reference"
                        rec[f"out{j}_indices"] = "bit-exact"
                    elif j == 0 and net.layer_types[i] == "SA_Layer":
                        assert torch.equal(g, w), f"layer {i} new_xyz differs from the reference"
                        rec["new_xyz"] = "bit-exact"
                    else:
                        e = rel_err(g.cpu().numpy(), w.cpu().numpy())
                        ee = elem_err(g.cpu().numpy(), w.cpu().numpy())
                        is_logits = net.layer_types[i] == "SA_Layer" and j == 2
                        tol = logit_tol if is_logits else feat_tol
                        rec[f"out{j}_{'logits' if is_logits else 'float'}"] = {"range_rel": e, "elementwise": ee}
                        assert e <= tol, f"layer {i} output {j}: relative error {e:.3e} > {tol:.1e}"
    finally:
        for h in hooks:
            h.remove()
    return report

"""Drop-in for the reference's `pcdet/ops/pointnet2/pointnet2_batch/pointnet2_modules.py`.

Same classes, keyword-only constructors, forward signatures, return tuples and `state_dict` layout
(`mlps.{i}.{0,3,6}.weight`, `mlps.{i}.{1,4,7}.*`, `aggregation_layer.*`, `confidence_layers.*`,
`mlp_modules.*`, `ctr_reg.*`; reference pointnet2_modules.py:84-125,128-460,462-516,519-587,590-763) so
IA-SSD / SPSNet-IA checkpoints load unchanged and `IASSD_Backbone` / `PAGNet_Backbone` pick it up.

Two execution paths per module:
  * inference (`not self.training`, inputs not requiring grad): the fused sm_100a path -- one launch for
    the sampler (FPS / score top-k), one row-gather for new_xyz, ONE multi-radius ball query for all MSG
    scales, and GEMM kernels whose A-loader does the grouping and whose epilogue does BN(folded)+ReLU
    [+max-pool]; the (B,3+C,npoint,nsample) grouped tensor and the conv/BN/ReLU intermediates of the
    reference are never materialised in the reference's form;
  * training (`self.training`): the grouped shared MLPs run on the same fused kernel with BATCH-statistics BatchNorm
    (train_fused.py: one statistics pass per layer + one pooled pass; running statistics updated like torch does;
    SyncBatchNorm = one small all-reduce per layer); the backward recomputes the reference composition from the saved
    inputs.  Conv1d stacks (aggregation / confidence / vote) stay on torch modules in training.
  * autograd in eval mode, or anything the fused kernels do not cover (avg-pool, GroupAll, SPSK_TRAIN_FUSED=0): the
    reference's op-by-op composition (grouping -> torch Conv/BN/ReLU -> pool) on top of the same CUDA ops.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils as pu
from . import train_fused

__all__ = [
    "PointnetSAModuleMSG", "PointnetSAModuleMSG_WithSampling", "Vote_layer", "PointnetSAModule",
    "PointnetFPModule", "PointnetSampling",
]


# ---------------------------------------------------------------------------------------------------
# builders (define the state_dict layout)
# ---------------------------------------------------------------------------------------------------

def _conv_bn_relu_2d(spec: List[int]) -> nn.Sequential:
    layers: List[nn.Module] = []
    for cin, cout in zip(spec[:-1], spec[1:]):
        layers += [nn.Conv2d(cin, cout, kernel_size=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU()]
    return nn.Sequential(*layers)


def _conv_bn_relu_1d(cin: int, widths: List[int]):
    layers: List[nn.Module] = []
    for cout in widths:
        layers += [nn.Conv1d(cin, cout, kernel_size=1, bias=False), nn.BatchNorm1d(cout), nn.ReLU()]
        cin = cout
    return layers, cin


def _make_grouper(radii, i, nsample, use_xyz, dilated, has_npoint):
    if not has_npoint:
        return pu.GroupAll(use_xyz)
    if dilated:
        inner = 0.0 if i == 0 else radii[i - 1]
        return pu.QueryDilatedAndGroup(radii[i], inner, nsample, use_xyz=use_xyz)
    return pu.QueryAndGroup(radii[i], nsample, use_xyz=use_xyz)


# ---------------------------------------------------------------------------------------------------
# BN folding for the fused inference path
# ---------------------------------------------------------------------------------------------------

class _Folded:
    """Caches (W^T with BN folded, bias, relu) per layer of a Conv(1x1)[+BN][+ReLU] Sequential.
    y = relu(x @ wt + bias) reproduces conv -> BN(eval, eps) -> ReLU (reference :204-211, 216-243)."""

    def __init__(self):
        self._key = None
        self._layers = None

    @staticmethod
    def _sig(seq: nn.Sequential):
        sig = []
        for t in list(seq.parameters()) + list(seq.buffers()):
            sig.append((t.data_ptr(), t._version, t.device))
        return tuple(sig)

    def get(self, seq: nn.Sequential):
        key = self._sig(seq)
        if key == self._key:
            return self._layers
        layers = []
        with torch.no_grad():
            for mod in seq:
                if isinstance(mod, (nn.Conv1d, nn.Conv2d)):
                    w = mod.weight.detach().float().reshape(mod.out_channels, mod.in_channels)
                    b = mod.bias.detach().float() if mod.bias is not None else torch.zeros(mod.out_channels, device=w.device)
                    layers.append([w, b, False])
                elif isinstance(mod, (nn.BatchNorm1d, nn.BatchNorm2d)):
                    w, b, _ = layers[-1]
                    inv = torch.rsqrt(mod.running_var.float() + mod.eps)
                    g = mod.weight.float() * inv if mod.affine else inv
                    beta = mod.bias.float() if mod.affine else torch.zeros_like(inv)
                    layers[-1][0] = w * g[:, None]
                    layers[-1][1] = (b - mod.running_mean.float()) * g + beta
                elif isinstance(mod, nn.ReLU):
                    layers[-1][2] = True
                else:  # pragma: no cover
                    raise RuntimeError(f"cannot fold {type(mod).__name__}")
            out = [(w.t().contiguous(), b.contiguous(), relu) for w, b, relu in layers]
        self._key, self._layers = key, out
        return out


def mlp_backend(module: nn.Module | None = None) -> str:
    """'mma' (default): grouped shared MLPs on the tcgen05 tensor cores (fp16 operands, fp32 accumulate);
    'ffma': the exact-fp32 CUDA-core kernels.  Set SPSK_MLP=ffma to force the latter everywhere; a module whose
    `_spsk_exact` attribute is set (the fp16 range guard found its activations beyond 65504, see
    `fp16_guard_check`) uses the exact kernels on its own."""
    if module is not None and getattr(module, "_spsk_exact", False):
        return "ffma"
    return os.environ.get("SPSK_MLP", "mma").lower()


def _tag(module: nn.Module) -> None:
    """Every tensor-core launch issued from here on carries this module's fp16-range-guard tag (pointnet2_utils)."""
    pu.current_ovf_tag = int(getattr(module, "_spsk_tag", 0)) & 31


def fp16_guard_enabled() -> bool:
    """The eager, inference-mode forward of the backbones polls the fp16 range-guard word once (ONE device synchronisation
    per forward).  Off inside CUDA-graph capture (BackbonePipeline checks while it warms up, before it captures) and with
    SPSK_FP16_GUARD=0."""
    return os.environ.get("SPSK_FP16_GUARD", "1") != "0" and not torch.cuda.is_current_stream_capturing()


def fp16_guard_check(modules) -> bool:
    """Poll the fp16 range-guard word (synchronises).  When a tensor-core kernel stored a value beyond the fp16 range since
    the last poll, switch the modules whose tag bit is set (all of them if the untagged bit 0 is set) to the exact-fp32
    kernels for good and return True: the caller re-runs its forward, which then reproduces the reference's finite fp32
    result (pointnet2_modules.py:203-211 of the reference keeps the fp32 exponent range under TF32)."""
    mask = pu.fp16_overflow(clear=True)
    if not mask:
        return False
    hit = False
    for m in modules:
        t = int(getattr(m, "_spsk_tag", 0)) & 31
        if (mask >> t) & 1 or (mask & 1):
            if not getattr(m, "_spsk_exact", False):
                m._spsk_exact = True
                hit = True
    if not hit:   # every flagged module is already exact: the flag came from somewhere else (e.g. a stale word)
        return False
    import warnings

    warnings.warn("spsnet_b200: activations beyond the fp16 range (65504) on the tensor-core path; the affected set-abstraction "
                  "modules now run on the exact-fp32 kernels", RuntimeWarning)
    return True


def _fused_ok(module: nn.Module, *tensors) -> bool:
    if module.training:
        return False
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        return False
    return True


def _pool_code(pool_method: str) -> int:
    if pool_method == "max_pool":
        return 1
    if pool_method == "avg_pool":
        return 2
    raise NotImplementedError


def _ceil(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _get_twin(features: torch.Tensor, mult: int = 8):
    """Point-major fp16 twin of a channel-major (B, C, N) fp32 tensor, as (twin (B, N, ld), lo_off): the copy the
    producing fused layer left on the tensor object when there is one (and the tensor was not modified since) -- that
    copy also carries the fp16 residuals at column `lo_off` (> 0) -- else a fresh conversion (lo_off = 0)."""
    B, Cc, N = features.shape
    hit = getattr(features, "_spsk_twin", None)
    if hit is not None:
        tw, ver, lo = hit
        width = lo if lo else tw.shape[2]
        if ver == features._version and tw.shape[0] == B and tw.shape[1] == N and width >= _ceil(Cc, mult) and tw.device == features.device:
            return tw, lo
    return pu.make_twin(features.contiguous(), _ceil(Cc, mult)), 0


def _split_rows(features: torch.Tensor):
    """Fresh point-major fp16 rows [values | residuals] of a channel-major (B, C, N) fp32 tensor: ((B, N, 2*c16), c16)."""
    B, Cc, N = features.shape
    c16 = _ceil(Cc, 16)
    pm = features.permute(0, 2, 1)
    rows = torch.zeros((B, N, 2 * c16), dtype=torch.float16, device=features.device)
    hi = pm.half()
    rows[:, :, :Cc] = hi
    rows[:, :, c16:c16 + Cc] = (pm - hi.float()).half()
    return rows, c16


def _set_twin(features: torch.Tensor, twin: torch.Tensor, lo_off: int = 0) -> None:
    features._spsk_twin = (twin.view(features.shape[0], features.shape[2], -1), features._version, lo_off)


class _PwCache:
    """Packed tensor-core weights (pu.PwLayer) of a folded Conv1d stack, rebuilt when the folded chain changes;
    kept per arithmetic (split = hi + lo fp16 inputs, fp32-grade; plain = fp16 inputs)."""

    def __init__(self):
        self._chain = None
        self._layers = {}

    def get(self, chain, split: bool):
        if self._chain is not chain:
            self._layers = {}
            self._chain = chain
        if split not in self._layers:
            self._layers[split] = [pu.PwLayer(wt, bias, relu, split=split) for wt, bias, relu in chain]
        return self._layers[split]


def _pw_stack(layers, x16: torch.Tensor, xlo: int, B: int, M: int, final: str):
    """Run a packed Conv1d stack on point-major fp16 rows (hi at column 0, residuals at column `xlo` for split layers).
    Intermediate layers produce fp16 rows only; the last one produces, per `final`:
      "cm+16" -> ((B, C, M) fp32 channel-major, (B*M, ld) fp16 rows, lo_off);   "pm" -> (B, M, C) fp32."""
    split = layers[0].split
    for i, layer in enumerate(layers):
        if i < len(layers) - 1:
            _, x16, _ = pu.pw_mma_forward(x16, layer, xlo=xlo, want16=not split, want16_lo=split)
            xlo = layer.n16 if split else 0
            continue
        if final == "cm+16":
            out_cm = torch.empty((B, layer.c_out, M), dtype=torch.float32, device=x16.device)
            _, out16, _ = pu.pw_mma_forward(x16, layer, xlo=xlo, out_cm=out_cm, m=M, want16=not split, want16_lo=split)
            return out_cm, out16, (layer.n16 if split else 0)
        out_pm = torch.empty((B, M, layer.c_out), dtype=torch.float32, device=x16.device)
        pu.pw_mma_forward(x16, layer, xlo=xlo, out_pm=out_pm)
        return out_pm
    raise RuntimeError("empty Conv1d stack")


# ---------------------------------------------------------------------------------------------------
# base: grouping + shared MLP + pooling over all scales
# ---------------------------------------------------------------------------------------------------

class _PointnetSAModuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None
        self.pool_method = "max_pool"

    # reference: pointnet2_modules.py:19-43
    def calc_square_dist(self, a, b, norm=True):
        a_sq = torch.sum(a.unsqueeze(2) * a.unsqueeze(2), dim=-1)  # (bs, n, 1)
        b_sq = torch.sum(b.unsqueeze(1) * b.unsqueeze(1), dim=-1)  # (bs, 1, m)
        a_sq = a_sq.repeat((1, 1, b.shape[1]))
        b_sq = b_sq.repeat((1, a.shape[1], 1))
        coor = torch.matmul(a, b.transpose(1, 2))
        return a_sq + b_sq - (2.0 if norm else 2) * coor

    def _folded(self, name: str, seq: nn.Sequential):
        cache = self.__dict__.setdefault("_fold_cache", {})
        if name not in cache:
            cache[name] = _Folded()
        return cache[name].get(seq)

    def _mma_chain(self, si: int, chain, c_feat: int, use_xyz: bool):
        """Packed tensor-core weights of scale `si`, rebuilt only when the folded chain object changes."""
        cache = self.__dict__.setdefault("_mma_cache", {})
        hit = cache.get(si)
        if hit is None or hit[0] is not chain or hit[1] != (c_feat, use_xyz):
            hit = (chain, (c_feat, use_xyz), pu.MmaChain(chain, c_feat, use_xyz))
            cache[si] = hit
        return hit[2]

    # -- reference composition (training / autograd): reference :62-79, 429-445
    def _msg_composed(self, xyz, new_xyz, features):
        outs = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            nf = mlp(grouper(xyz, new_xyz, features))  # (B, mlp[-1], npoint, nsample)
            if self.pool_method == "max_pool":
                nf = F.max_pool2d(nf, kernel_size=[1, nf.size(3)])
            elif self.pool_method == "avg_pool":
                nf = F.avg_pool2d(nf, kernel_size=[1, nf.size(3)])
            else:
                raise NotImplementedError
            outs.append(nf.squeeze(-1))
        return torch.cat(outs, dim=1)

    def _fusable_groupers(self) -> bool:
        return all(isinstance(g, (pu.QueryAndGroup, pu.QueryDilatedAndGroup)) for g in self.groupers)

    # -- fused inference path
    def _msg_fused(self, xyz, new_xyz, features, want16=False):
        """All MSG scales.  Returns (out_cm, out16): the pooled features as the reference's (B, C_total, M) fp32 tensor,
        or -- when `want16` and every scale runs on the tensor cores -- only as (B*M, 2*ceil16(C_total)) fp16 point-major
        rows [values | residuals], the operand layout of the aggregation GEMM (the fp32 tensor is never materialised)."""
        xyz = xyz.contiguous()
        new_xyz = new_xyz.contiguous()
        if features is not None:
            features = features.contiguous()
        B, M = new_xyz.shape[0], new_xyz.shape[1]
        pool = _pool_code(self.pool_method)
        chains = [self._folded(f"mlps.{i}", mlp) for i, mlp in enumerate(self.mlps)]
        c_total = sum(ch[-1][0].shape[1] for ch in chains)
        if all(isinstance(g, pu.QueryAndGroup) for g in self.groupers):
            idxs = pu.ball_query_msg([g.radius for g in self.groupers], [g.nsample for g in self.groupers], xyz, new_xyz)
        else:
            idxs = []
            for g in self.groupers:
                if isinstance(g, pu.QueryDilatedAndGroup):
                    idxs.append(pu.ball_query_dilated(g.radius_in, g.radius_out, g.nsample, xyz, new_xyz))
                else:
                    idxs.append(pu.ball_query(g.radius, g.nsample, xyz, new_xyz))
        c_feat = features.shape[1] if features is not None else 0
        packs = [None] * len(chains)
        _tag(self)
        if mlp_backend(self) == "mma" and pool == 1:
            for si, (g, chain, idx) in enumerate(zip(self.groupers, chains, idxs)):
                ns = idx.shape[2]
                if ns <= 128 and (ns & (ns - 1)) == 0:
                    pk = self._mma_chain(si, chain, c_feat, g.use_xyz)
                    packs[si] = pk if pk.ok else None
        out_cm = out16 = None
        c16 = _ceil(c_total, 16)
        if want16 and all(pk is not None for pk in packs):
            # [hi (c16) | lo (c16)]: pooled features as fp16 values + fp16 residuals, so the aggregation GEMM stays fp32-grade
            alloc = torch.empty if c16 == c_total else torch.zeros
            out16 = alloc((B * M, 2 * c16), dtype=torch.float16, device=xyz.device)
        else:
            out_cm = torch.zeros((B, c_total, M), dtype=torch.float32, device=xyz.device)
        co = 0
        twin = None
        for si, (g, chain, idx) in enumerate(zip(self.groupers, chains, idxs)):
            pk = packs[si]
            if pk is not None:
                if c_feat and not pk.split and twin is None:
                    twin = _get_twin(features)[0]
                pu.sa_mma_forward(xyz=xyz, new_xyz=new_xyz, idx=idx, chain=pk, twin=twin, features=features if pk.split else None,
                                  out_pooled=out_cm, co_off=co, out16=out16, co16=co, o16lo=c16 if out16 is not None else 0)
            else:
                rows = None
                for li, (wt, bias, relu) in enumerate(chain):
                    last = li == len(chain) - 1
                    rows = pu.grouped_linear(xyz=xyz, new_xyz=new_xyz, features=features, idx=idx, use_xyz=g.use_xyz,
                                             in_rows=rows, wt=wt, bias=bias, relu=relu, pool=pool if last else 0,
                                             out_pooled=out_cm if last else None, co_off=co)
            co += chain[-1][0].shape[1]
        return out_cm, out16

    def _msg(self, xyz, new_xyz, features, want16=False):
        if _fused_ok(self, xyz, features, new_xyz) and self._fusable_groupers() and len(self.mlps) > 0 \
                and all(len(m) > 0 for m in self.mlps):
            return self._msg_fused(xyz, new_xyz, features, want16)
        if self.training and train_fused.enabled():
            # training-mode BatchNorm on the fused kernel: L statistics passes + one pooled pass (train_fused.py)
            out = train_fused.msg_train(self, xyz, new_xyz, features)
            if out is not None:
                return out, None
        return self._msg_composed(xyz, new_xyz, features), None

    def _pw_layers(self, name: str, seq: nn.Sequential, split: bool):
        cache = self.__dict__.setdefault("_pw_cache", {})
        if name not in cache:
            cache[name] = _PwCache()
        return cache[name].get(self._folded(name, seq), split)

    def _seq_1d(self, name: str, seq: nn.Sequential, x: torch.Tensor) -> torch.Tensor:
        """Conv1d/BN1d/ReLU stack on a channel-major tensor: exact-fp32 point-wise GEMMs at inference (SPSK_MLP=ffma),
        torch modules otherwise."""
        if not _fused_ok(self, x):
            return seq(x)
        x = x.contiguous()
        for wt, bias, relu in self._folded(name, seq):
            x = pu.pointwise_linear(x, wt, bias, relu)
        return x

    def _aggregate(self, out_cm, out16, B, M):
        """aggregation_layer over the pooled MSG features (reference :447-449).  Tensor-core path: fp16 rows (values +
        residuals) in, (B, C, M) fp32 + its fp16 point-major twin out (the twin feeds the confidence GEMMs and the next
        layer's gather); hi + lo arithmetic keeps this layer fp32-grade."""
        if _fused_ok(self, out_cm) and mlp_backend(self) == "mma":
            if out16 is not None:
                xlo = out16.shape[1] // 2
            else:
                out16, xlo = pu.make_twin(out_cm, _ceil(out_cm.shape[1], 16)).view(B * M, -1), 0
            layers = self._pw_layers("aggregation_layer", self.aggregation_layer, split=xlo > 0)
            new_features, twin, lo = _pw_stack(layers, out16, xlo, B, M, "cm+16")
            _set_twin(new_features, twin, lo)
            return new_features
        return self._seq_1d("aggregation_layer", self.aggregation_layer, out_cm)

    def _confidence(self, new_features):
        """confidence_layers -> (B, npoint, num_class) (reference :454-455, incl. the transpose)."""
        if _fused_ok(self, new_features) and mlp_backend(self) == "mma":
            B, _, M = new_features.shape
            tw, lo = _get_twin(new_features, 16)
            layers = self._pw_layers("confidence_layers", self.confidence_layers, split=lo > 0)
            return _pw_stack(layers, tw.view(B * M, -1), lo, B, M, "pm")
        return self._seq_1d("confidence_layers", self.confidence_layers, new_features).transpose(1, 2)

    # reference: pointnet2_modules.py:45-81
    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None, new_xyz=None):
        if new_xyz is None and self.npoint is not None:
            idx = pu.farthest_point_sample(xyz.contiguous(), self.npoint)
            if _fused_ok(self, xyz):
                new_xyz = pu.gather_rows(xyz.contiguous(), idx)
            else:
                new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        return new_xyz, self._msg(xyz, new_xyz, features)[0]


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Pointnet set abstraction layer with multiscale grouping (reference :84-125)."""

    def __init__(self, *, npoint: int, radii: List[float], nsamples: List[int], mlps: List[List[int]], bn: bool = True,
                 use_xyz: bool = True, pool_method="max_pool", **kwargs):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for i in range(len(radii)):
            self.groupers.append(_make_grouper(radii, i, nsamples[i], use_xyz, False, npoint is not None))
            spec = mlps[i]
            if use_xyz:
                spec[0] += 3  # mutates the caller's list, like the reference (:117-118)
            self.mlps.append(_conv_bn_relu_2d(spec))
        self.pool_method = pool_method


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale set abstraction layer (reference :519-536)."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True, pool_method="max_pool"):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz,
                         pool_method=pool_method)


# ---------------------------------------------------------------------------------------------------
# samplers (reference dispatch: substring match in this order, pointnet2_modules.py:284-419)
# ---------------------------------------------------------------------------------------------------

def _gather_stds(stds, idx):
    """stds (B,1,N) -> (B,npoint) at idx (reference :305,310 `gather_operation(...).squeeze()`)."""
    B = stds.shape[0]
    return pu.gather_operation(stds.reshape(B, 1, -1).contiguous(), idx).reshape(B, -1)


def _gather_features(features, idx, fused):
    """SA layer without groupers: new_features = features[:, :, idx] (reference :451-452); the cached fp16 twin of the
    source, if any, is gathered along (as 4-byte words) so that the next fused layer does not have to convert."""
    features = features.contiguous()
    out = pu.gather_operation(features, idx).contiguous()
    hit = getattr(features, "_spsk_twin", None) if fused else None
    if hit is not None and hit[1] == features._version and hit[0].shape[2] % 2 == 0 and mlp_backend() == "mma":
        g = pu.gather_rows(hit[0].view(torch.float32), idx.contiguous())
        _set_twin(out, g.view(torch.float16), hit[2])
    return out


def _sector_fps(xyz_tmp, npoint, key_fn, part_num=4):
    """'ds_FPS' / 'ry_FPS' (reference :372-419): sort each scene by a scalar key, split into 4 equal
    parts, FPS npoint/4 in each part, map back to original indices."""
    B = xyz_tmp.shape[0]
    xyz_div, idx_div = [], []
    for per_xyz in xyz_tmp:
        _, order = key_fn(per_xyz).sort(dim=0, descending=False)
        xyz_div.append(per_xyz[order].view(part_num, -1, 3))
        idx_div.append(order.view(part_num, -1))
    xyz_div = torch.cat(xyz_div, dim=0).contiguous()
    idx_div = torch.cat(idx_div, dim=0)
    picked = pu.furthest_point_sample(xyz_div, npoint // part_num)
    mapped = torch.gather(idx_div, 1, picked.long())
    return mapped.reshape(B, npoint).int()


class PointnetSAModuleMSG_WithSampling(_PointnetSAModuleBase):
    """Set abstraction layer with a per-layer down-sampling policy and multiscale grouping
    (reference pointnet2_modules.py:128-460)."""

    def __init__(self, *,
                 npoint_list: List[int],
                 sample_range_list: List[int],
                 sample_type_list: List[str],
                 radii: List[float],
                 nsamples: List[int],
                 mlps: List[List[int]],
                 use_xyz: bool = True,
                 dilated_group=False,
                 pool_method="max_pool",
                 aggregation_mlp: List[int],
                 confidence_mlp: List[int],
                 num_class,
                 **kwargs):
        super().__init__()
        self.sample_type_list = sample_type_list
        self.sample_range_list = sample_range_list
        self.dilated_group = dilated_group
        # SPSNet stable sampling ('S-FPS'): first entry of each list (reference :161-168)
        if kwargs.get("ss_radii", None) is not None and len(kwargs["ss_radii"]) > 0:
            self.ss_radii = kwargs["ss_radii"][0]
            self.ss_nsamples = kwargs["ss_nsamples"][0]

        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint_list = npoint_list
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        out_channels = 0
        for i in range(len(radii)):
            self.groupers.append(_make_grouper(radii, i, nsamples[i], use_xyz, self.dilated_group, npoint_list is not None))
            spec = mlps[i]
            if use_xyz:
                spec[0] += 3
            self.mlps.append(_conv_bn_relu_2d(spec))
            out_channels += spec[-1]
        self.pool_method = pool_method

        if aggregation_mlp is not None and len(aggregation_mlp) != 0 and len(self.mlps) > 0:
            layers, out_channels = _conv_bn_relu_1d(out_channels, aggregation_mlp)
            self.aggregation_layer = nn.Sequential(*layers)
        else:
            self.aggregation_layer = None

        if confidence_mlp is not None and len(confidence_mlp) != 0:
            layers, out_channels = _conv_bn_relu_1d(out_channels, confidence_mlp)
            layers.append(nn.Conv1d(out_channels, num_class, kernel_size=1, bias=True))
            self.confidence_layers = nn.Sequential(*layers)
        else:
            self.confidence_layers = None

    # -- one entry of (sample_type_list, sample_range_list, npoint_list)
    def _sample_one(self, sample_type, npoint, xyz, xyz_tmp, feature_tmp, cls_tmp, stds, xyz_flipped):
        B, n_tmp = xyz_tmp.shape[0], xyz_tmp.shape[1]
        if n_tmp <= npoint:  # no down-sampling (reference :284-285)
            return torch.arange(n_tmp, device=xyz_tmp.device, dtype=torch.int32).repeat(B, 1), stds

        if ("cls" in sample_type) or ("ctr" in sample_type):  # reference :287-291
            return pu.score_topk(cls_tmp.contiguous(), npoint), stds

        if ("ss" in sample_type) or ("sss" in sample_type):  # SPSNet, reference :293-305
            if stds is None:
                raise NotImplementedError
            idx = pu.score_topk(cls_tmp.contiguous(), npoint, stds=stds.reshape(B, -1).contiguous())
            return idx, _gather_stds(stds, idx)

        if "D-FPS" in sample_type or "DFS" in sample_type:  # reference :307-310
            idx = pu.furthest_point_sample(xyz_tmp.contiguous(), npoint)
            if stds is not None:
                stds = _gather_stds(stds, idx)
            return idx, stds

        if "S-FPS" in sample_type or "SFS" in sample_type:  # SPSNet, reference :314-353
            if stds is None:
                raise NotImplementedError
            fps_idx = pu.furthest_point_sample(xyz_tmp.contiguous(), npoint)
            centres = pu.gather_rows(xyz.contiguous(), fps_idx)
            nbr = pu.ball_query(self.ss_radii, self.ss_nsamples, xyz.contiguous(), centres)
            s3 = stds.reshape(B, 1, -1).contiguous()
            grouped = pu.grouping_operation(s3, nbr).reshape(B, npoint, -1)
            stable = torch.argmin(grouped, dim=-1, keepdim=True)
            idx = torch.gather(nbr, 2, stable).view(B, -1).contiguous()
            new_stds = _gather_stds(s3, idx)
            if idx[0].unique().shape[0] < 3500:  # the reference's fallback (host sync included)
                idx = pu.furthest_point_sample(xyz_tmp.contiguous(), npoint)
            return idx, new_stds

        if "F-FPS" in sample_type or "FFS" in sample_type:  # reference :357-361
            f = torch.cat([xyz_tmp, feature_tmp], dim=-1)
            return pu.furthest_point_sample_with_dist(self.calc_square_dist(f, f).contiguous(), npoint), stds

        if sample_type == "FS":  # reference :363-369
            f = torch.cat([xyz_tmp, feature_tmp], dim=-1)
            i1 = pu.furthest_point_sample_with_dist(self.calc_square_dist(f, f).contiguous(), npoint)
            i2 = pu.furthest_point_sample(xyz_tmp.contiguous(), npoint)
            return torch.cat([i1, i2], dim=-1), stds

        if "Rand" in sample_type:  # reference :370-371 (one permutation shared by the batch)
            return torch.randperm(n_tmp, device=xyz_tmp.device)[None, :npoint].int().repeat(B, 1), stds

        if sample_type in ("ds_FPS", "ds-FPS"):  # reference :372-395
            return _sector_fps(xyz_tmp, npoint, lambda p: p.norm(dim=-1) - 5), stds

        if sample_type in ("ry_FPS", "ry-FPS"):  # reference :397-419
            return _sector_fps(xyz_tmp, npoint, lambda p: torch.atan(p[:, 0] / p[:, 1])), stds

        raise NotImplementedError(f"unknown sample type {sample_type!r}")

    def _sample(self, xyz, features, cls_features, stds, xyz_flipped):
        picked = []
        start = 0
        for sample_type, sample_range, npoint in zip(self.sample_type_list, self.sample_range_list, self.npoint_list):
            if npoint <= 0:
                continue
            if sample_range == -1:
                sl = slice(start, None)
            else:
                sl = slice(start, sample_range)
            xyz_tmp = xyz[:, sl, :]
            feature_tmp = features.transpose(1, 2)[:, sl, :] if features is not None else None
            cls_tmp = cls_features[:, sl, :] if cls_features is not None else None
            if sample_range != -1:
                start += sample_range  # (sic) the reference advances by the range END (:282)
            idx, stds = self._sample_one(sample_type, npoint, xyz, xyz_tmp, feature_tmp, cls_tmp, stds, xyz_flipped)
            picked.append(idx)
        return torch.cat(picked, dim=-1).contiguous(), stds

    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None, cls_features: torch.Tensor = None,
                new_xyz=None, ctr_xyz=None, **kwargs):
        """
        :param xyz: (B, N, 3); features: (B, C, N); cls_features: (B, N, num_class); ctr_xyz: (B, M, 3) or None
        :return: new_xyz (B, npoint, 3), new_features (B, C_out, npoint), cls_features (B, npoint, num_class) or
                 None, sampled_idx_list (B, npoint) int32 ([] when ctr_xyz is given), stds
        """
        B = xyz.shape[0]
        stds = kwargs.get("stds", None)
        if stds is not None:
            stds = stds.reshape(B, 1, -1).contiguous()
        fused = _fused_ok(self, xyz, features)
        sampled_idx_list: object = []
        if ctr_xyz is None:
            xyz_flipped = None
            pre = kwargs.get("_presampled")
            if pre is not None:
                # D-FPS of this layer was launched on a side stream as soon as the previous layer had its centres
                # (backbone.IASSD_Backbone._prefetch_fps): join that stream, then proceed as if sampled here
                torch.cuda.current_stream().wait_event(pre[1])
                sampled_idx_list = pre[0]
                if stds is not None:
                    stds = _gather_stds(stds, sampled_idx_list)
            else:
                sampled_idx_list, stds = self._sample(xyz, features, cls_features, stds, xyz_flipped)
            if fused:
                new_xyz = pu.gather_rows(xyz.contiguous(), sampled_idx_list)
            else:
                new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), sampled_idx_list).transpose(1, 2).contiguous()
        else:
            new_xyz = ctr_xyz
        hook = kwargs.get("_after_new_xyz")
        if hook is not None:
            hook(new_xyz)

        if len(self.groupers) > 0:
            B_, M_ = new_xyz.shape[0], new_xyz.shape[1]
            out_cm, out16 = self._msg(xyz, new_xyz, features, want16=self.aggregation_layer is not None)
            if self.aggregation_layer is not None:
                new_features = self._aggregate(out_cm, out16, B_, M_)
            else:
                new_features = out_cm
        else:
            new_features = _gather_features(features, sampled_idx_list, fused)

        if self.confidence_layers is not None:
            cls_features = self._confidence(new_features)
        else:
            cls_features = None
        return new_xyz, new_features, cls_features, sampled_idx_list, stds


class Vote_layer(nn.Module):
    """Light voting module with a limited translation range (reference pointnet2_modules.py:462-516)."""

    def __init__(self, mlp_list, pre_channel, max_translate_range):
        super().__init__()
        self.mlp_list = mlp_list
        if len(mlp_list) > 0:
            # the reference rebuilds `shared_mlps` inside the loop, so only the LAST entry survives (:467-477)
            for width in mlp_list:
                shared = [nn.Conv1d(pre_channel, width, kernel_size=1, bias=False), nn.BatchNorm1d(width), nn.ReLU()]
                pre_channel = width
            self.mlp_modules = nn.Sequential(*shared)
        else:
            self.mlp_modules = None
        self.ctr_reg = nn.Conv1d(pre_channel, 3, kernel_size=1)
        self.max_offset_limit = torch.tensor(max_translate_range).float() if max_translate_range is not None else None

    def _folded(self, name, seq):
        cache = self.__dict__.setdefault("_fold_cache", {})
        if name not in cache:
            cache[name] = _Folded()
        return cache[name].get(seq)

    def forward(self, xyz, features, **kwargs):
        xyz_select = xyz
        features_select = features
        if kwargs.get("center_surface_futures", None) is not None:
            self.center_surface_futures = kwargs["center_surface_futures"]
        if self.mlp_modules is None:
            raise RuntimeError("Vote_layer without mlp_list is undefined in the reference (uses an unbound name)")
        if hasattr(self, "center_surface_futures"):
            features_select = torch.cat([self.center_surface_futures, features_select], dim=1)

        _tag(self)
        if _fused_ok(self, features_select) and mlp_backend(self) == "mma":
            B_, _, M_ = features_select.shape
            caches = self.__dict__.setdefault("_pw_cache", {"mlp_modules": _PwCache(), "ctr_reg": _PwCache()})
            tw, lo = _get_twin(features_select.contiguous(), 16)
            if lo == 0:
                # no producer rows with residuals (concatenated surface features, a producer on the exact-fp32 kernels, a
                # caller-supplied tensor): build [values | residuals] here so the offsets stay fp32-grade
                tw, lo = _split_rows(features_select)
            layers = caches["mlp_modules"].get(self._folded("mlp_modules", self.mlp_modules), lo > 0) + \
                caches["ctr_reg"].get(self._folded("ctr_reg", nn.Sequential(self.ctr_reg)), lo > 0)
            ctr_offsets = _pw_stack(layers, tw.view(B_ * M_, -1), lo, B_, M_, "pm")  # (B, npoint, 3 [+ extra])
        elif _fused_ok(self, features_select):
            h = features_select.contiguous()
            for wt, bias, relu in self._folded("mlp_modules", self.mlp_modules):
                h = pu.pointwise_linear(h, wt, bias, relu)
            for wt, bias, relu in self._folded("ctr_reg", nn.Sequential(self.ctr_reg)):
                ctr_offsets = pu.pointwise_linear(h, wt, bias, relu)
            ctr_offsets = ctr_offsets.transpose(1, 2)
        else:
            ctr_offsets = self.ctr_reg(self.mlp_modules(features_select)).transpose(1, 2)

        new_features = ctr_offsets[..., 3:]
        ctr_offsets = ctr_offsets[..., :3]
        if self.max_offset_limit is not None:
            # device copy cached once (a pageable H2D copy per forward would also break CUDA-graph capture)
            cached = self.__dict__.get("_limit_dev")
            if cached is None or cached.device != xyz_select.device:
                cached = self.max_offset_limit.to(xyz_select.device).view(1, 1, 3)
                self.__dict__["_limit_dev"] = cached
            limit = cached
            if _fused_ok(self, ctr_offsets):
                # one clamp kernel instead of the reference's compare / where / neg / compare / where chain (:504-513): same
                # values (NaN passes through both), `limited` is only an intermediate
                neg = self.__dict__.get("_neg_limit_dev")
                if neg is None or neg.device != limit.device:
                    neg = self.__dict__["_neg_limit_dev"] = -limit
                vote_xyz = xyz_select + torch.clamp(ctr_offsets, min=neg, max=limit)
            else:
                limited = torch.where(ctr_offsets > limit, limit, ctr_offsets)
                limited = torch.where(limited < -limit, -limit, limited)
                vote_xyz = xyz_select + limited
        else:
            vote_xyz = xyz_select + ctr_offsets
        return vote_xyz, new_features, xyz_select, ctr_offsets


class PointnetFPModule(nn.Module):
    """Feature propagation: 3-NN inverse-distance interpolation + shared MLP (reference :539-587)."""

    def __init__(self, *, mlp: List[int], bn: bool = True):
        super().__init__()
        self.mlp = _conv_bn_relu_2d(mlp)

    def forward(self, unknown: torch.Tensor, known: torch.Tensor, unknow_feats: torch.Tensor,
                known_feats: torch.Tensor) -> torch.Tensor:
        if known is not None:
            dist, idx = pu.three_nn(unknown.contiguous(), known.contiguous())
            dist_recip = 1.0 / (dist + 1e-8)
            norm = torch.sum(dist_recip, dim=2, keepdim=True)
            weight = dist_recip / norm
            interpolated = pu.three_interpolate(known_feats.contiguous(), idx, weight.contiguous())
        else:
            interpolated = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
        if unknow_feats is not None:
            new_features = torch.cat([interpolated, unknow_feats], dim=1)  # (B, C2 + C1, n)
        else:
            new_features = interpolated
        if _fused_ok(self, new_features):
            x = new_features.contiguous()
            cache = self.__dict__.setdefault("_fold_cache", _Folded())
            for wt, bias, relu in cache.get(self.mlp):
                x = pu.pointwise_linear(x, wt, bias, relu)
            return x
        return self.mlp(new_features.unsqueeze(-1)).squeeze(-1)


class PointnetSampling(_PointnetSAModuleBase):
    """SA layer of the SPSNet stability generator: identity / D-FPS sampling + MSG grouping + aggregation,
    no confidence head, 3-tuple return (reference pointnet2_modules.py:590-763)."""

    def __init__(self, *,
                 npoint_list: List[int],
                 sample_range_list: List[int],
                 sample_type_list: List[str],
                 radii: List[float],
                 nsamples: List[int],
                 mlps: List[List[int]],
                 use_xyz: bool = True,
                 dilated_group=False,
                 pool_method="max_pool",
                 aggregation_mlp: List[int]):
        super().__init__()
        self.sample_type_list = sample_type_list
        self.sample_range_list = sample_range_list
        self.dilated_group = dilated_group
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint_list = npoint_list
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        out_channels = 0
        for i in range(len(radii)):
            self.groupers.append(_make_grouper(radii, i, nsamples[i], use_xyz, self.dilated_group, npoint_list is not None))
            spec = mlps[i]
            if use_xyz:
                spec[0] += 3
            self.mlps.append(_conv_bn_relu_2d(spec))
            out_channels += spec[-1]
        self.pool_method = pool_method
        if aggregation_mlp is not None and len(aggregation_mlp) != 0 and len(self.mlps) > 0:
            layers, out_channels = _conv_bn_relu_1d(out_channels, aggregation_mlp)
            self.aggregation_layer = nn.Sequential(*layers)
        else:
            self.aggregation_layer = None
        self.confidence_layers = None

    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None, cls_features: torch.Tensor = None,
                new_xyz=None, ctr_xyz=None, **kwargs):
        B = xyz.shape[0]
        fused = _fused_ok(self, xyz, features)
        sampled_idx_list: object = []
        if ctr_xyz is None:
            picked = []
            start = 0
            for sample_type, sample_range, npoint in zip(self.sample_type_list, self.sample_range_list, self.npoint_list):
                if npoint <= 0:
                    continue
                sl = slice(start, None) if sample_range == -1 else slice(start, sample_range)
                xyz_tmp = xyz[:, sl, :]
                if sample_range != -1:
                    start += sample_range
                if xyz_tmp.shape[1] <= npoint:
                    idx = torch.arange(xyz_tmp.shape[1], device=xyz.device, dtype=torch.int32).repeat(B, 1)
                elif "D-FPS" in sample_type or "DFS" in sample_type:
                    idx = pu.furthest_point_sample(xyz_tmp.contiguous(), npoint)
                else:
                    raise NotImplementedError(f"PointnetSampling supports identity / D-FPS only, got {sample_type!r}")
                picked.append(idx)
            sampled_idx_list = torch.cat(picked, dim=-1).contiguous()
            if fused:
                new_xyz = pu.gather_rows(xyz.contiguous(), sampled_idx_list)
            else:
                new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), sampled_idx_list).transpose(1, 2).contiguous()
        else:
            new_xyz = ctr_xyz
        if len(self.groupers) > 0:
            out_cm, out16 = self._msg(xyz, new_xyz, features, want16=self.aggregation_layer is not None)
            if self.aggregation_layer is not None:
                new_features = self._aggregate(out_cm, out16, new_xyz.shape[0], new_xyz.shape[1])
            else:
                new_features = out_cm
        else:
            new_features = _gather_features(features, sampled_idx_list, fused)
        return new_xyz, new_features, sampled_idx_list

"""Training-mode forward of the grouped shared MLP on the fused tensor-core kernel: Conv2d(1x1) -> BatchNorm2d with BATCH
statistics -> ReLU, repeated, -> max-pool over the neighbours (reference pointnet2_modules.py:203-211 built, :429-445 run,
with the module in `train()`; SyncBatchNorm after `tools/train.py:122-123`), SURVEY.md section 8(f) rank 4.

The reference materialises the grouped tensor (B, C, npoint, nsample) and every Conv / BN / ReLU result and lets cuDNN reduce
the batch statistics.  Here nothing of that size exists in the forward:

  pass l = 0 .. L-1   `spsk_sa_mma_forward` in its statistics mode on the chain truncated after conv l -- layers < l carry the
                      batch statistics already found (folded into W and bias exactly like running statistics are at inference),
                      conv l runs raw -- and the epilogue of conv l reduces sum z and sum z^2 per channel over all
                      B*npoint*nsample rows (per-thread cells, fp64 across tiles, no atomics: reproducible).  A (2, C) fp64 sum
                      over the per-CTA slices [+ ONE all-reduce of 2C+1 doubles per layer for SyncBatchNorm] gives mean and biased
                      variance; the running statistics get the reference's momentum update (unbiased variance).
  pass L              the inference kernel with all L layers folded on the batch statistics -> pooled features.

Around the passes everything stays on the device and launch-light (csrc/train_bn.cu): one kernel packs a layer's fp16 weight
tiles from the Conv2d weight (x the BN scale), one reduces the per-CTA slices, one turns the sums into (scale, bias) for the next
pass and updates the running statistics -- about 7 launches per layer, no host arithmetic, no synchronisation (`TrainPlan`).
SPSK_TRAIN_PACK=torch runs the same algebra with torch ops instead (the CPU-tested formulas below; slower, host-bound).

L + 1 launches that recompute the (cheap, on-chip) chain instead of L round trips of the (B, C, npoint, nsample) tensors through
HBM.  The backward recomputes the reference composition from the saved inputs (xyz, new_xyz, features, idx -- a few MB) with
torch autograd, so gradients are those of the fp32 reference function, BatchNorm's dependence on the batch statistics included;
the activations (hundreds of MB per layer at KITTI sizes) are never kept between forward and backward.  Hand-written backward
GEMMs are the step after this one.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils as pu
from ._lib import SaMmaDesc, check, lib

_BN_TYPES = (nn.BatchNorm2d, nn.SyncBatchNorm)


def enabled() -> bool:
    """SPSK_TRAIN_FUSED=0 keeps `module.train()` on the reference's op-by-op composition."""
    return os.environ.get("SPSK_TRAIN_FUSED", "1") != "0"


# ---------------------------------------------------------------------------------------------------
# pure-torch pieces (run on CPU too: tests/test_train_fused_cpu.py)
# ---------------------------------------------------------------------------------------------------

def split_layers(seq: nn.Sequential) -> Optional[List[Tuple[nn.Conv2d, nn.Module]]]:
    """[(conv, bn), ...] when `seq` is (Conv2d 1x1 without bias, BatchNorm2d | SyncBatchNorm (affine, in training mode), ReLU)
    repeated -- the only form the reference builds with bn=True (pointnet2_modules.py:204-211) -- else None."""
    mods = list(seq)
    if not mods or len(mods) % 3:
        return None
    out = []
    for i in range(0, len(mods), 3):
        conv, bn, act = mods[i:i + 3]
        if not isinstance(conv, nn.Conv2d) or conv.kernel_size != (1, 1) or conv.bias is not None or conv.groups != 1:
            return None
        if not isinstance(bn, _BN_TYPES) or not bn.affine or not bn.training:
            return None
        if not isinstance(act, nn.ReLU):
            return None
        out.append((conv, bn))
    return out


def sync_group(bn: nn.Module):
    """The process group a SyncBatchNorm layer reduces over, or None when the statistics stay local (plain BatchNorm2d, no
    initialised process group, or a world of one)."""
    if not isinstance(bn, nn.SyncBatchNorm):
        return None
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return None
    group = bn.process_group if bn.process_group is not None else dist.group.WORLD
    return group if dist.get_world_size(group) > 1 else None


def bn_moments(sums: torch.Tensor, count: int, group=None) -> Tuple[torch.Tensor, torch.Tensor, float]:
    """(C, 2) fp64 [sum z, sum z^2] over `count` local rows -> (mean, biased variance, total count), fp64; with `group` the three
    quantities are summed over the ranks first: ONE all-reduce of 2C + 1 doubles (what SyncBatchNorm's all_gather of
    (mean, invstd, count) amounts to, torch/nn/modules/_functions.py)."""
    c = sums.shape[0]
    if group is not None:
        import torch.distributed as dist

        buf = torch.cat([sums.reshape(-1), sums.new_tensor([float(count)])])
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        sums, total = buf[:2 * c].reshape(c, 2), float(buf[-1].item())
    else:
        total = float(count)
    mean = sums[:, 0] / total
    var = (sums[:, 1] / total - mean * mean).clamp_min_(0.0)
    return mean, var, total


def bn_fold(weight: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, mean: torch.Tensor, var: torch.Tensor, eps: float):
    """conv (cout, cin) + BN on (mean, var) -> (wt (cin, cout), bias (cout)) fp32 with y = x @ wt + bias; folded in fp64."""
    g = gamma.double() * torch.rsqrt(var.double() + eps)
    wt = (weight.double() * g[:, None]).t().contiguous().float()
    bias = (beta.double() - mean.double() * g).float()
    return wt, bias


def bn_update_running(bn: nn.Module, mean: torch.Tensor, var: torch.Tensor, total: float) -> None:
    """The reference's running-statistics update (torch _BatchNorm.forward + batch_norm kernels): exponential average with
    `momentum` (cumulative average when it is None), unbiased variance."""
    if not bn.track_running_stats or bn.running_mean is None:
        return
    bn.num_batches_tracked.add_(1)
    mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked.item())
    unbiased = var * (total / max(total - 1.0, 1.0))
    bn.running_mean.mul_(1.0 - mom).add_(mean.to(bn.running_mean.dtype), alpha=mom)
    bn.running_var.mul_(1.0 - mom).add_(unbiased.to(bn.running_var.dtype), alpha=mom)


def _bn_train(z: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, group):
    """Batch-statistics BN as a differentiable torch expression (no running-statistics side effect).  Local: F.batch_norm.
    Synchronised: two differentiable all-reduces (sum -> mean, centred sum of squares -> variance)."""
    if group is None:
        return F.batch_norm(z, None, None, gamma, beta, True, 0.0, eps)
    import torch.distributed.nn.functional as dfn

    c = z.shape[1]
    dims = (0, 2, 3)
    first = dfn.all_reduce(torch.cat([z.sum(dims), z.new_tensor([z.numel() / c])]), group=group)
    total = first[-1]
    mean = (first[:c] / total).view(1, c, 1, 1)
    zc = z - mean
    var = (dfn.all_reduce((zc * zc).sum(dims), group=group) / total).view(1, c, 1, 1)
    return zc * (gamma.view(1, c, 1, 1) * torch.rsqrt(var + eps)) + beta.view(1, c, 1, 1)


def mlp_recompute(grouped: torch.Tensor, params: Sequence[torch.Tensor], eps: Sequence[float], groups: Sequence) -> torch.Tensor:
    """The reference composition on a grouped tensor (B, C, npoint, nsample): [conv1x1 -> BN(batch stats) -> ReLU] x L -> max
    over nsample, as a torch graph (reference pointnet2_modules.py:431-436)."""
    x = grouped
    if x.is_cuda and os.environ.get("SPSK_TRAIN_CHANNELS_LAST", "1") != "0":
        # NHWC: the 1x1 convolutions become plain GEMMs over (B*npoint*nsample, C) for cuDNN in both directions, and BatchNorm /
        # ReLU / max-pool follow the format; same fp32 math, measured faster than NCHW for these few-channel, many-row tensors
        x = x.contiguous(memory_format=torch.channels_last)
    for l in range(len(params) // 3):
        w, gamma, beta = params[3 * l:3 * l + 3]
        x = F.conv2d(x, w)
        x = F.relu(_bn_train(x, gamma, beta, eps[l], groups[l]))
    return F.max_pool2d(x, kernel_size=[1, x.size(3)]).squeeze(-1)


# ---------------------------------------------------------------------------------------------------
# the fused forward
# ---------------------------------------------------------------------------------------------------

def _ceil(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class _PlanChain:
    """What pu.sa_mma_forward needs from a chain (cf. pu.MmaChain), over the plan's shared weight / bias buffers."""

    def __init__(self, plan: "TrainPlan", nlayers: int):
        self.plan, self.nlayers = plan, nlayers
        self.c_feat, self.use_xyz, self.split, self.cpad8 = plan.c_feat, plan.use_xyz, plan.split, plan.cpad8
        self.kpad = plan.kpad[:nlayers]
        self.cpad = plan.cp_hidden[:nlayers - 1] + [_ceil(plan.couts[nlayers - 1], 128)]
        self.cout_last = plan.couts[nlayers - 1]
        v = [C.c_int(0) for _ in range(4)]
        self.ok = lib.spsk_sa_mma_config(C.byref(self._desc()), *[C.byref(x) for x in v]) == 0

    def _desc(self) -> SaMmaDesc:
        d = SaMmaDesc()
        d.nlayers = self.nlayers
        for l in range(self.nlayers):
            d.kpad[l], d.cpad[l] = self.kpad[l], self.cpad[l]
        d.split = 1 if self.split else 0
        d.ovf_tag = pu.current_ovf_tag
        d.c_feat, d.use_xyz, d.cout_last = self.c_feat, 1 if self.use_xyz else 0, self.cout_last
        d.wtiles, d.bias = self.plan.wbuf.data_ptr(), self.plan.bbuf.data_ptr()
        return d


class TrainPlan:
    """Device buffers and launch shapes of one MSG scale's training forward, built once per (shapes, device) and reused every
    step.  ONE weight buffer and ONE bias buffer serve all L + 1 passes: pass l reads [folded 0 .. folded l-1 | raw l]; once its
    statistics are in, `folded l` is packed over `raw l` (same offset, never larger) for the passes that follow."""

    def __init__(self, shapes: Sequence[Tuple[int, int]], c_feat: int, use_xyz: bool, split: bool, device):
        self.c_feat, self.use_xyz, self.split = c_feat, use_xyz, split
        self.couts = [co for co, _ in shapes]
        self.cins = [ci for _, ci in shapes]
        self.L = len(shapes)
        self.cpad8 = _ceil(c_feat, 8) if c_feat > 0 else 0
        k0 = 16 if split else _ceil(max(self.cpad8 + (8 if use_xyz else 0), 16), 16)
        self.cp_hidden = [_ceil(co, 16) for co in self.couts]
        self.kpad = [k0] + self.cp_hidden[:-1]
        wk = [(2 if split else 1) * k for k in self.kpad]
        self.w_off, self.b_off = [], []
        wo = bo = wcap = bcap = 0
        for l in range(self.L):
            self.w_off.append(wo)
            self.b_off.append(bo)
            wcap = max(wcap, wo + wk[l] * _ceil(self.couts[l], 128) * 2)
            bcap = max(bcap, bo + _ceil(self.couts[l], 128))
            wo += wk[l] * self.cp_hidden[l] * 2
            bo += self.cp_hidden[l]
        self.wbuf = torch.zeros(wcap, dtype=torch.uint8, device=device)
        self.bbuf = torch.zeros(bcap, dtype=torch.float32, device=device)       # padding entries stay zero for good
        self.sums = [torch.zeros(2 * co + 1, dtype=torch.float64, device=device) for co in self.couts]
        self.scale = [torch.zeros(co, dtype=torch.float32, device=device) for co in self.couts]
        self.chains = [_PlanChain(self, l + 1) for l in range(self.L)]           # chains[L-1] is also the final (pooled) pass
        self.ok = all(ch.ok for ch in self.chains) and max(self.kpad) <= 1024 and max(self.cp_hidden[:-1] + [0]) <= 1024
        self._parts = {}

    def parts(self, l: int, idx: torch.Tensor, n: int) -> torch.Tensor:
        key = (l, tuple(idx.shape), n)
        buf = self._parts.get(key)
        if buf is None:
            nparts = pu.sa_mma_stats_parts(idx, n, self.chains[l])
            buf = self._parts[key] = torch.empty((nparts, self.chains[l].cpad[-1], 2), dtype=torch.float64, device=idx.device)
        return buf   # spsk_sa_mma_forward zeroes the slices it accumulates into

    def pack(self, l: int, weight: torch.Tensor, scale: Optional[torch.Tensor], last: bool) -> None:
        cp = _ceil(self.couts[l], 128) if last else self.cp_hidden[l]
        check(lib.spsk_sa_pack_layer(weight.data_ptr(), self.couts[l], self.cins[l], scale.data_ptr() if scale is not None else None,
                                     1 if l == 0 else 0, self.c_feat, 1 if self.use_xyz else 0, self.kpad[l], cp, 1 if self.split else 0,
                                     self.wbuf.data_ptr() + self.w_off[l], pu._stream()), "sa_pack_layer")

    def finalize(self, l: int, parts: torch.Tensor, count: int, bn: nn.Module, gamma: torch.Tensor, beta: torch.Tensor, group) -> None:
        co = self.couts[l]
        st = pu._stream()
        track = bn.track_running_stats and bn.running_mean is not None
        if group is None and (not track or bn.momentum is not None):
            # statistics local to this rank: reduce + finalize + num_batches_tracked in ONE launch
            check(lib.spsk_bn_stats_reduce_finalize(parts.data_ptr(), parts.shape[0], parts.shape[1], co, float(count), gamma.data_ptr(), beta.data_ptr(),
                                                    float(bn.eps), float(bn.momentum) if track else -1.0,
                                                    bn.running_mean.data_ptr() if track else None, bn.running_var.data_ptr() if track else None,
                                                    bn.num_batches_tracked.data_ptr() if track else None, self.scale[l].data_ptr(),
                                                    self.bbuf.data_ptr() + 4 * self.b_off[l], self.sums[l].data_ptr(), st), "bn_stats_reduce_finalize")
            return
        check(lib.spsk_bn_stats_reduce(parts.data_ptr(), parts.shape[0], parts.shape[1], co, float(count), self.sums[l].data_ptr(), st), "bn_stats_reduce")
        if group is not None:
            import torch.distributed as dist

            dist.all_reduce(self.sums[l], op=dist.ReduceOp.SUM, group=group)      # SyncBatchNorm: 2c + 1 doubles
        mom = -1.0
        if track:
            bn.num_batches_tracked.add_(1)
            mom = float(bn.momentum) if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked.item())
        check(lib.spsk_bn_stats_finalize(self.sums[l].data_ptr(), co, gamma.data_ptr(), beta.data_ptr(), float(bn.eps), mom,
                                         bn.running_mean.data_ptr() if track else None, bn.running_var.data_ptr() if track else None,
                                         self.scale[l].data_ptr(), self.bbuf.data_ptr() + 4 * self.b_off[l], None, st), "bn_stats_finalize")


def _forward_kernels(meta, common, c_feat: int, params) -> torch.Tensor:
    """L statistics passes + the pooled pass with the device-side pack / reduce / finalize kernels (TrainPlan)."""
    plan, bns, groups = meta["plan"], meta["bns"], meta["groups"]
    idx = common["idx"]
    B, M, ns = idx.shape
    N = common["xyz"].shape[1]
    for l, bn in enumerate(bns):
        w, gamma, beta = params[3 * l:3 * l + 3]
        if not (w.is_contiguous() and gamma.is_contiguous() and beta.is_contiguous() and w.dtype == gamma.dtype == beta.dtype == torch.float32):
            raise RuntimeError("train_fused: Conv2d / BatchNorm parameters must be contiguous fp32")
        plan.pack(l, w, None, last=True)                                   # raw conv l closes the truncated chain
        parts = plan.parts(l, idx, N)
        pu.sa_mma_forward(chain=plan.chains[l], stats=parts, **common)
        plan.finalize(l, parts, B * M * ns, bn, gamma, beta, groups[l])    # -> scale, bias (in place), running statistics
        plan.pack(l, w, plan.scale[l], last=l == len(bns) - 1)             # folded conv l for every later pass
    out = torch.empty((B, plan.couts[-1], M), dtype=torch.float32, device=idx.device)
    pu.sa_mma_forward(chain=plan.chains[-1], out_pooled=out, co_off=0, **common)
    return out


def _forward_torch(meta, common, c_feat: int, params) -> torch.Tensor:
    """The same passes with the host algebra in torch ops (SPSK_TRAIN_PACK=torch; the formulas tests/test_train_fused_cpu.py pins)."""
    use_xyz, split, bns, eps, groups = meta["use_xyz"], meta["split"], meta["bns"], meta["eps"], meta["groups"]
    idx = common["idx"]
    B, M, ns = idx.shape
    N = common["xyz"].shape[1]
    folded = []
    for l, bn in enumerate(bns):
        w = params[3 * l].reshape(params[3 * l].shape[0], -1).float()
        gamma, beta = params[3 * l + 1].float(), params[3 * l + 2].float()
        raw = (w.t().contiguous(), torch.zeros(w.shape[0], dtype=torch.float32, device=w.device), True)
        pk = pu.MmaChain(folded + [raw], c_feat, use_xyz, split=split, pair=False)
        if not pk.ok:
            raise RuntimeError("train_fused: chain does not fit the fused kernel (probe and launch disagree)")
        parts = torch.zeros((pu.sa_mma_stats_parts(idx, N, pk), pk.cpad[-1], 2), dtype=torch.float64, device=idx.device)
        pu.sa_mma_forward(chain=pk, stats=parts, **common)
        mean, var, total = bn_moments(parts.sum(dim=0)[:w.shape[0]], B * M * ns, groups[l])
        bn_update_running(bn, mean, var, total)
        wt, bias = bn_fold(w, gamma, beta, mean, var, eps[l])
        folded.append((wt, bias, True))
    pk = pu.MmaChain(folded, c_feat, use_xyz, split=split, pair=False)
    out = torch.empty((B, pk.cout_last, M), dtype=torch.float32, device=idx.device)
    pu.sa_mma_forward(chain=pk, out_pooled=out, co_off=0, **common)
    return out


class FusedTrainMLP(torch.autograd.Function):
    """One MSG scale.  apply(meta, xyz, new_xyz, features | None, idx, W0, gamma0, beta0, W1, ...) -> (B, C_last, npoint).
    `meta`: dict(use_xyz, split, bns=[BatchNorm modules], eps=[...], groups=[process group | None, ...], plan=TrainPlan | None)."""

    @staticmethod
    def forward(ctx, meta, xyz, new_xyz, features, idx, *params):
        c_feat = features.shape[1] if features is not None else 0
        with torch.no_grad(), torch.cuda.device(idx.device):
            twin = None
            if c_feat and not meta["split"]:
                twin = pu.make_twin(features.contiguous(), (c_feat + 7) // 8 * 8)
            common = dict(xyz=xyz, new_xyz=new_xyz, idx=idx, twin=twin, features=features if meta["split"] else None)
            run = _forward_kernels if meta.get("plan") is not None else _forward_torch
            out = run(meta, common, c_feat, [p.detach() for p in params])
        ctx.meta = meta
        ctx.has_feat = features is not None
        ctx.save_for_backward(xyz, new_xyz, features if features is not None else xyz.new_empty(0), idx, *params)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        meta = ctx.meta
        xyz, new_xyz, features, idx, *params = ctx.saved_tensors
        need = ctx.needs_input_grad   # (meta, xyz, new_xyz, features, idx, *params)
        with torch.enable_grad():
            xyz_ = xyz.detach().requires_grad_(need[1])
            new_xyz_ = new_xyz.detach().requires_grad_(need[2])
            feat_ = features.detach().requires_grad_(need[3]) if ctx.has_feat else None
            params_ = [p.detach().requires_grad_(need[5 + i]) for i, p in enumerate(params)]
            grouped = pu._group(xyz_, new_xyz_, feat_, idx, meta["use_xyz"])
            out = mlp_recompute(grouped, params_, meta["eps"], meta["groups"])
            cand = [xyz_, new_xyz_, feat_] + params_
            wanted = [t for t in cand if t is not None and t.requires_grad]
            grads = iter(torch.autograd.grad(out, wanted, grad_out.contiguous(), allow_unused=True)) if wanted else iter(())
        res = [next(grads) if (t is not None and t.requires_grad) else None for t in cand]
        return (None, res[0], res[1], res[2], None, *res[3:])


def msg_train(module, xyz: torch.Tensor, new_xyz: torch.Tensor, features: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """All MSG scales of a set-abstraction module in training mode on the fused path; None when some scale is outside what the
    kernel covers (the caller then runs the reference composition)."""
    if not xyz.is_cuda or module.pool_method != "max_pool" or len(module.mlps) == 0:
        return None
    if torch.is_autocast_enabled() or xyz.dtype != torch.float32 or (features is not None and features.dtype != torch.float32):
        return None   # mixed-precision training keeps torch's own casting rules: the composition handles it
    if not all(isinstance(g, (pu.QueryAndGroup, pu.QueryDilatedAndGroup)) for g in module.groupers):
        return None
    c_feat = features.shape[1] if features is not None else 0
    scales = []
    use_kernels = os.environ.get("SPSK_TRAIN_PACK", "kernel") != "torch"
    cache = module.__dict__.setdefault("_train_plans", {})
    for si, (g, mlp) in enumerate(zip(module.groupers, module.mlps)):
        ns = g.nsample
        if ns > 128 or (ns & (ns - 1)):
            return None
        layers = split_layers(mlp)
        if layers is None or len(layers) > 4 or (features is None and not g.use_xyz):
            return None
        key = (si, c_feat, g.use_xyz, tuple(conv.weight.shape for conv, _ in layers), layers[0][0].weight.device)
        if cache.get(si, (None,))[0] != key:
            shapes = [(conv.out_channels, conv.in_channels) for conv, _ in layers]
            # split (hi + lo fp16, fp32-grade) arithmetic for narrow chains, the rule of pu.MmaChain
            split = c_feat <= 8 and all(_ceil(co, 16) <= 64 for co, _ in shapes[:-1])
            plan = TrainPlan(shapes, c_feat, g.use_xyz, split, layers[0][0].weight.device)
            cache[si] = (key, plan.ok and shapes[0][1] == c_feat + (3 if g.use_xyz else 0), split, plan)
        _, ok, split, plan = cache[si]
        if not ok:
            return None
        scales.append((g, layers, split, plan if use_kernels else None))
    xyz = xyz.contiguous()
    new_xyz = new_xyz.contiguous()
    if features is not None:
        features = features.contiguous()
    with torch.no_grad():
        if all(isinstance(g, pu.QueryAndGroup) for g in module.groupers):
            idxs = pu.ball_query_msg([g.radius for g in module.groupers], [g.nsample for g in module.groupers], xyz, new_xyz)
        else:
            idxs = [pu.ball_query_dilated(g.radius_in, g.radius_out, g.nsample, xyz, new_xyz) if isinstance(g, pu.QueryDilatedAndGroup)
                    else pu.ball_query(g.radius, g.nsample, xyz, new_xyz) for g in module.groupers]
    outs = []
    for (g, layers, split, plan), idx in zip(scales, idxs):
        meta = dict(use_xyz=g.use_xyz, split=split, bns=[bn for _, bn in layers], eps=[float(bn.eps) for _, bn in layers],
                    groups=[sync_group(bn) for _, bn in layers], plan=plan)
        params = []
        for conv, bn in layers:
            params += [conv.weight, bn.weight, bn.bias]
        outs.append(FusedTrainMLP.apply(meta, xyz, new_xyz, features, idx, *params))
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)

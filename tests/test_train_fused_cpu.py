"""Host logic of the training-mode fused path (spsnet_b200/train_fused.py) on CPU: the statistics -> (mean, var) -> folded
weights -> running-statistics algebra against torch's own BatchNorm2d in train(), the recompute graph of the backward against
nn.Sequential autograd, and the SyncBatchNorm variant over a world_size-2 gloo group against the full-batch result
(reference: pointnet2_modules.py:203-211 in train(), tools/train.py:122-123)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from spsnet_b200 import pointnet2_modules as pm
from spsnet_b200 import train_fused as tf


def _seq(spec):
    torch.manual_seed(5)
    seq = pm._conv_bn_relu_2d(spec)
    for m in seq:
        if isinstance(m, nn.BatchNorm2d):
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.uniform_(-0.3, 0.3)
    return seq.train()


def test_split_layers_accepts_only_the_reference_form():
    seq = _seq([7, 16, 32])
    layers = tf.split_layers(seq)
    assert [(c.out_channels, type(b)) for c, b in layers] == [(16, nn.BatchNorm2d), (32, nn.BatchNorm2d)]
    seq[1].eval()                                     # a frozen BN layer keeps its running statistics: not this path
    assert tf.split_layers(seq) is None
    assert tf.split_layers(nn.Sequential(nn.Conv2d(4, 8, 1, bias=True), nn.BatchNorm2d(8), nn.ReLU())) is None
    assert tf.split_layers(nn.Sequential(nn.Conv2d(4, 8, 1, bias=False), nn.ReLU())) is None
    assert tf.split_layers(nn.Sequential(nn.Conv2d(4, 8, 1, bias=False), nn.BatchNorm2d(8, affine=False), nn.ReLU())) is None


def _fused_algebra(seq, x):
    """What FusedTrainMLP.forward does, with the statistics kernel replaced by a float64 torch reduction."""
    layers = tf.split_layers(seq)
    rows = x.permute(0, 2, 3, 1).reshape(-1, x.shape[1])
    count = rows.shape[0]
    for conv, bn in layers:
        w = conv.weight.detach().reshape(conv.out_channels, -1)
        z = rows.double() @ w.double().t()
        sums = torch.stack([z.sum(0), (z * z).sum(0)], dim=1)
        mean, var, total = tf.bn_moments(sums, count)
        tf.bn_update_running(bn, mean, var, total)
        wt, bias = tf.bn_fold(w, bn.weight.detach(), bn.bias.detach(), mean, var, bn.eps)
        rows = torch.relu(rows @ wt + bias)
    B, _, M, ns = x.shape
    return rows.reshape(B, M, ns, -1).max(dim=2)[0].permute(0, 2, 1)


@pytest.mark.parametrize("momentum", [0.1, None])
def test_moments_fold_and_running_stats_match_torch_batchnorm(momentum):
    import copy

    seq = _seq([7, 16, 32, 24])
    for m in seq:
        if isinstance(m, nn.BatchNorm2d):
            m.momentum = momentum
    ref = copy.deepcopy(seq)
    for step in range(2):                              # two steps: the running statistics chain correctly
        x = torch.randn(3, 7, 10, 8) * (1.0 + step) + 0.5
        got = _fused_algebra(seq, x)
        want = ref(x).max(dim=3)[0]
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-5)
    for a, b in zip(seq, ref):
        if isinstance(a, nn.BatchNorm2d):
            assert torch.allclose(a.running_mean, b.running_mean, rtol=1e-5, atol=1e-6)
            assert torch.allclose(a.running_var, b.running_var, rtol=1e-5, atol=1e-6)
            assert int(a.num_batches_tracked) == int(b.num_batches_tracked) == 2


def _params(seq):
    out = []
    for conv, bn in tf.split_layers(seq):
        out += [conv.weight, bn.weight, bn.bias]
    return out


def test_recompute_graph_matches_sequential_autograd():
    seq = _seq([6, 16, 16, 32])
    x = torch.randn(2, 6, 9, 4, requires_grad=True)
    g = torch.randn(2, 32, 9)
    before = [m.running_mean.clone() for m in seq if isinstance(m, nn.BatchNorm2d)]
    params = _params(seq)
    out = tf.mlp_recompute(x, params, [1e-5] * 3, [None] * 3)
    grads = torch.autograd.grad(out, [x] + params, g)
    after = [m.running_mean for m in seq if isinstance(m, nn.BatchNorm2d)]
    assert all(torch.equal(a, b) for a, b in zip(before, after))     # the recompute never touches the running statistics
    x2 = x.detach().clone().requires_grad_(True)
    want = seq(x2).max(dim=3)[0]
    wgrads = torch.autograd.grad(want, [x2] + params, g)
    assert torch.allclose(out, want, rtol=1e-5, atol=1e-6)
    for a, b in zip(grads, wgrads):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sync_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seq = _seq([5, 16, 24])                        # same seed on every rank: replicated weights
        sync = nn.SyncBatchNorm.convert_sync_batchnorm(seq).train()
        layers = tf.split_layers(sync)
        groups = [tf.sync_group(bn) for _, bn in layers]
        assert all(g is not None for g in groups)
        torch.manual_seed(11)
        full = torch.randn(4, 5, 6, 8) + 0.25
        gfull = torch.randn(4, 24, 6)
        # ranks hold UNEQUAL shards (3 + 1 scenes): the count must travel with the sums
        lo, hi = (0, 3) if rank == 0 else (3, 4)
        x = full[lo:hi].clone().requires_grad_(True)
        # forward statistics of layer 0 as the kernel would deliver them
        conv0, bn0 = layers[0]
        rows = x.detach().permute(0, 2, 3, 1).reshape(-1, 5).double()
        z = rows @ conv0.weight.detach().reshape(16, 5).double().t()
        mean, var, total = tf.bn_moments(torch.stack([z.sum(0), (z * z).sum(0)], dim=1), rows.shape[0], groups[0])
        tf.bn_update_running(bn0, mean, var, total)
        params = []
        for conv, bn in layers:
            params += [conv.weight, bn.weight, bn.bias]
        y = tf.mlp_recompute(x, params, [bn.eps for _, bn in layers], groups)
        grads = torch.autograd.grad(y, [x] + params, gfull[lo:hi])
        # gradients of replicated parameters are summed over ranks (what DDP's all-reduce does, up to its 1/world)
        pg = [g.clone() for g in grads[1:]]
        for g in pg:
            dist.all_reduce(g)
        # numpy payloads: a tensor in a Queue travels as a shared-memory handle that dies with this process
        out.put((rank, mean.float().numpy(), var.float().numpy(), total, bn0.running_mean.numpy().copy(), bn0.running_var.numpy().copy(),
                 y.detach().numpy(), grads[0].numpy(), [g.numpy() for g in pg]))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_sync_batchnorm_equals_full_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sync_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process truth on the full batch
    seq = _seq([5, 16, 24])
    torch.manual_seed(11)
    full = (torch.randn(4, 5, 6, 8) + 0.25).requires_grad_(True)
    gfull = torch.randn(4, 24, 6)
    params = _params(seq)
    want = seq(full).max(dim=3)[0]
    wg = torch.autograd.grad(want, [full] + params, gfull)
    bn0 = seq[1]
    z = torch.nn.functional.conv2d(full.detach(), seq[0].weight.detach())
    for rank, *payload in got:
        mean, var, total, rmean, rvar, y, gx, pg = [torch.from_numpy(t) if hasattr(t, "dtype") else t for t in payload]
        pg = [torch.from_numpy(t) for t in pg]
        assert total == 4 * 6 * 8
        assert torch.allclose(mean, z.mean((0, 2, 3)), rtol=1e-5, atol=1e-6)
        assert torch.allclose(var, z.var((0, 2, 3), unbiased=False), rtol=1e-4, atol=1e-6)
        assert torch.allclose(rmean, bn0.running_mean, rtol=1e-5, atol=1e-6)
        assert torch.allclose(rvar, bn0.running_var, rtol=1e-4, atol=1e-6)
        lo, hi = (0, 3) if rank == 0 else (3, 4)
        assert torch.allclose(y, want[lo:hi], rtol=1e-4, atol=1e-5)
        assert torch.allclose(gx, wg[0][lo:hi], rtol=1e-3, atol=1e-5)
        for a, b in zip(pg, wg[1:]):
            assert torch.allclose(a, b, rtol=1e-3, atol=1e-5)


def test_sync_group_is_none_without_a_process_group():
    bn = nn.SyncBatchNorm(8)
    assert tf.sync_group(bn) is None and tf.sync_group(nn.BatchNorm2d(8)) is None


PLAN_CHAINS = [(1, [16, 16, 32]), (1, [32, 32, 64]), (64, [64, 64, 128]), (64, [64, 96, 128]), (128, [128, 256, 256]), (256, [256, 512, 1024]),
               (0, [16, 32]), (5, [17, 33]), (20, [40])]


@pytest.mark.parametrize("c_feat,widths", PLAN_CHAINS, ids=[f"c{c}-" + "x".join(map(str, w)) for c, w in PLAN_CHAINS])
def test_train_plan_layout_equals_the_inference_packing(c_feat, widths):
    """TrainPlan's shared weight / bias buffers: every truncated chain (pass l) and the full chain have exactly the shapes, offsets
    and arithmetic pu.MmaChain gives the same layers at inference -- the kernels see one layout."""
    from spsnet_b200 import pointnet2_utils as pu

    cin = c_feat + 3
    shapes, chain = [], []
    for co in widths:
        shapes.append((co, cin))
        chain.append((torch.randn(cin, co), torch.zeros(co), True))
        cin = co
    split = c_feat <= 8 and all(tf._ceil(co, 16) <= 64 for co, _ in shapes[:-1])
    plan = tf.TrainPlan(shapes, c_feat, True, split, "cpu")
    assert plan.ok
    for l in range(len(widths)):
        want = pu.MmaChain(chain[:l + 1], c_feat, True, split=split, pair=False)
        got = plan.chains[l]
        assert want.split == split and want.ok == got.ok
        assert (want.kpad, want.cpad, want.cout_last, want.cpad8) == (got.kpad, got.cpad, got.cout_last, got.cpad8)
        wk = [(2 if split else 1) * k for k in want.kpad]
        assert plan.w_off[l] == sum(wk[j] * want.cpad[j] * 2 for j in range(l))
        assert plan.b_off[l] == sum(want.cpad[:l])
        assert want.wtiles.numel() * 2 <= plan.wbuf.numel() and want.bias.numel() <= plan.bbuf.numel()


def test_msg_train_declines_what_it_does_not_cover():
    """CPU tensors, avg-pool, GroupAll, non-power-of-two nsample and frozen BatchNorm go to the reference composition (None)."""
    m = pm.PointnetSAModuleMSG(npoint=8, radii=[0.5], nsamples=[16], mlps=[[4, 8, 16]], use_xyz=True).train()
    xyz, new_xyz, feats = torch.randn(1, 32, 3), torch.randn(1, 8, 3), torch.randn(1, 4, 32)
    assert tf.msg_train(m, xyz, new_xyz, feats) is None                      # CPU: the fused path needs the device
    for bad in (pm.PointnetSAModuleMSG(npoint=8, radii=[0.5], nsamples=[16], mlps=[[4, 8]], pool_method="avg_pool"),
                pm.PointnetSAModuleMSG(npoint=None, radii=[0.5], nsamples=[16], mlps=[[4, 8]]),
                pm.PointnetSAModuleMSG(npoint=8, radii=[0.5], nsamples=[12], mlps=[[4, 8]])):
        bad.train()
        fake = type("T", (), {"is_cuda": True, "dtype": torch.float32, "shape": xyz.shape})()   # device checks pass, structure must not
        assert tf.msg_train(bad, fake, new_xyz, None if bad.npoint is None else feats) is None
    m.mlps[0][1].eval()
    assert tf.split_layers(m.mlps[0]) is None

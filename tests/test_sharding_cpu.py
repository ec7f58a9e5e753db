"""N > 1 host logic on CPU: world_size-2 gloo process group -- scene sharding is a partition, and the
reporting reduction (the only communication of the path) sums scenes and takes the MAX of the per-rank times."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spsnet_b200 import sharding


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 16, 17, 128):
        for world in (1, 2, 3, 4, 8):
            cuts = [sharding.shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for (a0, a1), (b0, b1) in zip(cuts, cuts[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_scenes, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_range(n_scenes, rank, world)
        # every rank "processes" its scenes: a per-scene checksum stands in for the independent per-scene work
        local = torch.arange(lo, hi, dtype=torch.float64)
        checksum = (local * local + 1).sum()
        total, ms = sharding.reduce_report(hi - lo, 10.0 + 5.0 * rank)
        dist.all_reduce(checksum, op=dist.ReduceOp.SUM)
        out.put((rank, lo, hi, total, ms, float(checksum)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_sharded_report():
    world, n_scenes = 2, 33
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_scenes, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, tot0, ms0, c0), (r1, lo1, hi1, tot1, ms1, c1) = got
    assert (lo0, hi0, lo1, hi1) == (0, 17, 17, 33)                      # contiguous partition, remainder to rank 0
    assert tot0 == tot1 == n_scenes and ms0 == ms1 == 15.0              # SUM of scenes, MAX of times on every rank
    want = float(sum(i * i + 1 for i in range(n_scenes)))
    assert c0 == c1 == want                                             # the shards cover every scene exactly once
    assert sharding.throughput(tot0, ms0) == pytest.approx(33 / 0.015)


def test_reduce_report_without_group_is_identity():
    assert sharding.reduce_report(16, 2.5) == (16, 2.5)


def test_strong_scaling_plan_partitions_the_host_batch():
    """bench.py --scaling strong: the per-rank shards of one 128-scene host batch (shard_range) tile it exactly, in whole
    sub-batches of 16 scenes, for every world size of the scaling run."""
    import types

    import bench

    bench.set_workload("kitti")
    for world in (1, 2, 4, 8):
        seen = []
        for rank in range(world):
            lo, per_step = bench.strong_plan(types.SimpleNamespace(total_scenes=128), rank, world)
            seen += list(range(lo, lo + per_step * bench.BATCH))
        assert seen == list(range(128))
    with pytest.raises(SystemExit):
        bench.strong_plan(types.SimpleNamespace(total_scenes=100), 0, 8)

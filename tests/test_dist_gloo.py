"""world_size-2 gloo test of the N>1 host logic: batch sharding, the loss-boundary all-gather, max-over-ranks timing."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graph_weather_b200.dist import BoundaryGather, all_gather_batch, max_over_ranks, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(total, rank, world)
    full = torch.arange(total * 3 * 2, dtype=torch.float32).reshape(total, 3, 2)
    got = all_gather_batch(full[a:b] * 1.0, total)
    mx = max_over_ranks(10.0 + rank, "cpu")
    # the loss-boundary object the bench and training loops use: on CPU tensors it is the gloo all-gather, same interface
    gather = BoundaryGather(total, "cpu")
    got2 = gather(full[a:b] * 1.0)
    gather.wait()
    q.put((rank, bool(torch.equal(got, full)) and bool(torch.equal(got2, full)) and gather.mode == "gloo", mx))
    dist.destroy_process_group()


def _run(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(timeout=60) for p in ps]
    assert all(ok for _, ok, _ in res), res
    assert all(mx == 11.0 for _, _, mx in res), res


def test_shard_range():
    assert [shard_range(8, r, 2) for r in range(2)] == [(0, 4), (4, 8)]
    assert [shard_range(7, r, 3) for r in range(3)] == [(0, 3), (3, 5), (5, 7)]
    assert shard_range(1, 1, 2) == (1, 1)


def test_all_gather_even():
    _run(8)


def test_all_gather_uneven():
    _run(7)


def _loss_worker(rank, world, port, q):
    """The scalar exchange of NormalizedMSELoss.forward(group=...): each rank reduces its batch shard to one double, the sums are
    all-reduced and divided by the global row count.  On CPU the per-shard reduction (the CUDA kernel's job) is stood in for by
    the oracle's arithmetic; everything else is the product code path."""
    import numpy as np

    from graph_weather_b200.losses import NormalizedMSELoss, node_weights
    from oracle import restate

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lat_lons = [(float(a), float(b)) for a in range(-90, 90, 30) for b in range(0, 360, 30)]
    g = torch.Generator().manual_seed(3)
    total, F = 5, 6
    pred, target = torch.randn(total, len(lat_lons), F, generator=g), torch.randn(total, len(lat_lons), F, generator=g)
    var = [0.5 + 0.25 * i for i in range(F)]
    crit = NormalizedMSELoss(var, lat_lons, normalize=True)
    w = torch.from_numpy(node_weights(lat_lons, len(lat_lons))).double()

    def cpu_local_sum(p, t):  # stands in for gw_normalized_mse_loss_sum
        return ((((p - t) ** 2) / torch.tensor(var)).mean(-1).double() * w).sum().reshape(1)

    crit.local_sum = cpu_local_sum
    a, b = shard_range(total, rank, world)
    got = float(crit(pred[a:b], target[a:b], group=dist.group.WORLD, total_batch=total))
    got_derived = float(crit(pred[a:b], target[a:b], group=dist.group.WORLD))  # global batch derived from the shards
    ref = float(restate.normalized_mse_loss(pred, target, var, lat_lons, True))
    q.put((rank, abs(got - ref) <= 1e-6 * abs(ref) and abs(got_derived - ref) <= 1e-6 * abs(ref), got))
    dist.destroy_process_group()


def test_loss_scalar_exchange_uneven_shards():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_loss_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(timeout=60) for p in ps]
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == res[1][2]  # every rank holds the same loss

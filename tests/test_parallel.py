"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: query sharding with result gather
(no data-path collective) and the flat-bucket gradient all-reduce of data-parallel map training."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gs_localization_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, out), nprocs=world, join=True)
    return dict(out)


def _queries(rank, world):
    mine = parallel.shard_queries(11)
    local = {q: torch.tensor([q, 10.0 * q, rank], dtype=torch.float64) for q in mine}
    table = parallel.gather_query_results(local, 11, 3)
    return mine, None if table is None else table.tolist()


def test_query_sharding_and_gather():
    out = _spawn(_queries)
    assert sorted(out[0][0] + out[1][0]) == list(range(11))          # every query exactly once
    assert abs(len(out[0][0]) - len(out[1][0])) <= 1                  # balanced
    table = out[0][1]
    assert out[1][1] is None
    for q in range(11):
        assert table[q][0] == q and table[q][1] == 10.0 * q and table[q][2] == q % 2


def _grads(rank, world):
    torch.manual_seed(0)
    shapes = [(50, 3), (50, 1, 3), (50, 15, 3), (50, 1), (50, 3), (50, 4)]   # xyz, f_dc, f_rest, opacity, scaling, rotation
    params = [torch.zeros(s, requires_grad=True) for s in shapes]
    for i, p in enumerate(params):
        p.grad = torch.full(p.shape, float(rank + 1) * (i + 1))
    params[2].grad = None                                               # a parameter that got no gradient this step
    bucket = parallel.GradientBucket(params)
    assert bucket.nbytes() == 59 * 50 * 4                               # 59 floats per Gaussian at SH degree 3
    bucket.allreduce()
    return [float(p.grad.flatten()[0]) for p in params]


def test_gradient_bucket_allreduce():
    out = _spawn(_grads)
    want = [3.0 * (i + 1) for i in range(6)]
    want[2] = 0.0
    assert out[0] == want and out[1] == want


def test_view_sharding_covers_all_views():
    seen = [parallel.shard_views(10, step, r, 4) for step in range(5) for r in range(4)]
    assert sorted(set(seen)) == list(range(10))


def test_single_process_defaults():
    assert parallel.shard_queries(5) == [0, 1, 2, 3, 4]
    t = parallel.gather_query_results({1: torch.ones(2)}, 3, 2)
    assert t.shape == (3, 2) and t[1, 0] == 1 and torch.isnan(t[0, 0])


def _sparse(rank, world):
    torch.manual_seed(0)
    P = 500
    like = [torch.zeros(P, 3), torch.zeros(P, 4, 3), torch.zeros(P, 1), torch.zeros(P, 3), torch.zeros(P, 4)]
    g = torch.Generator().manual_seed(100 + rank)
    vis = torch.rand(P, generator=g) < (0.1 if rank == 0 else 0.3)       # different row counts per rank
    grads = [torch.randn(t.shape, generator=g) * vis.view(-1, *([1] * (t.dim() - 1))) for t in like]
    dense = [x.clone() for x in grads]
    for x in dense:
        dist.all_reduce(x)
    ex = parallel.SparseGradientExchange(like, granularity=64)
    out = ex.exchange(grads, vis)
    err = max(float((a - b).abs().max()) for a, b in zip(out, dense))
    return err, ex.last_rows, ex.F


def test_sparse_gradient_exchange_equals_dense_allreduce():
    out = _spawn(_sparse)
    for r in (0, 1):
        err, rows, F = out[r]
        assert err < 1e-6 and F == 3 + 12 + 1 + 3 + 4
        assert rows % 64 == 0 and 64 <= rows <= 256                       # padded to the larger rank's visible count

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


class _Replica:
    """The part of GaussianModel that DataParallelTrainer.resync touches."""

    def __init__(self, rank):
        g = torch.Generator().manual_seed(5)
        mk = lambda *s: torch.randn(*s, generator=g)
        self.p = [mk(40, 3), mk(40, 16, 3), mk(40, 1), mk(40, 3), mk(40, 4)]
        self._state = {"m": [mk(*t.shape) for t in self.p], "v": [mk(*t.shape).abs() for t in self.p]}
        self.max_radii2D, self.xyz_gradient_accum, self.denom = mk(40).abs(), mk(40, 1).abs(), mk(40, 1).abs()
        if rank:                                   # the drift of a replica: last-ulp differences everywhere
            for t in self.p + self._state["m"] + self._state["v"] + [self.max_radii2D, self.xyz_gradient_accum, self.denom]:
                t.mul_(1.0 + 1.2e-7 * rank)
        self.refreshed = 0

    def _params(self):
        return self.p

    def _refresh_activations(self):
        self.refreshed += 1


def _resync(rank, world):
    import types
    opt = types.SimpleNamespace(densify_until_iter=15_000, densify_from_iter=500, densification_interval=100, opacity_reset_interval=3000)
    model = _Replica(rank)
    trainer = parallel.DataParallelTrainer(model, opt, mode="sparse")
    structural = [it for it in (100, 500, 600, 650, 3000, 14_900, 15_000, 15_100) if trainer.structural_step(it)]
    before = model.p[1].clone()
    trainer.resync()
    flat = torch.cat([t.reshape(-1) for t in model.p + model._state["m"] + model._state["v"] + [model.max_radii2D, model.xyz_gradient_accum, model.denom]])
    return structural, bool((before != model.p[1]).any()), flat.tolist(), model.refreshed


def test_resync_makes_replicas_identical_before_structural_steps():
    """Three ranks whose replicas differ in the last ulp (what the visible-row exchange leaves at >= 3 ranks) are bit-identical
    after DataParallelTrainer.resync; structural_step names exactly the densification / opacity-reset iterations."""
    out = _spawn(_resync, world=3)
    assert out[0][0] == [600, 3000, 14_900]
    assert out[0][1] is False and out[1][1] is True and out[2][1] is True      # rank 0 is the source
    assert out[0][2] == out[1][2] == out[2][2]
    assert out[0][3] == out[1][3] == 1


def test_auto_mode_follows_the_measured_crossovers():
    """DataParallelTrainer.choose_mode on the cases measured on B200 (DESIGN section 8, C4-sized map: 3M Gaussians, 63-float table
    rows, 59-float arena rows): the peer pull wins at 2 GPUs up to the largest view and at 8 GPUs up to ~9 % visible rows;
    all-gather + add loses to the all-reduce earlier."""
    P, W, A = 3_000_000, 63, 3_000_000 * 59
    choose = parallel.DataParallelTrainer.choose_mode
    assert choose(1, int(0.19 * P), W, A, True) == "sparse"
    assert choose(2, int(0.19 * P), W, A, True) == "sparse"          # 0.70 ms vs 1.29 ms
    assert choose(2, int(0.19 * P), W, A, False) == "sparse"         # 1.26 ms vs 1.29 ms (all-gather path, barely)
    assert choose(4, int(0.19 * P), W, A, True) == "sparse"          # bench train_dp at 4 GPUs: 1106 vs 864 views/s
    assert choose(8, int(0.095 * P), W, A, True) == "sparse"         # 1.62 ms vs 1.74 ms
    assert choose(8, int(0.19 * P), W, A, True) == "dense"           # 2.85 ms vs 1.74 ms
    assert choose(8, int(0.095 * P), W, A, False) == "dense"         # 2.57 ms vs 1.74 ms
    assert choose(8, int(0.035 * P), W, A, False) == "sparse"        # 1.15 ms vs 1.74 ms

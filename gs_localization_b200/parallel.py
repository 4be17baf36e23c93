"""Multi-GPU plumbing for the two paths that shard naturally.

* Localization: queries are independent (each owns its camera, Adam state and loop,
  gs_localization/pipelines/7scenes_localize_full_dslam.py:352-365) and the map is read-only, so
  queries are split across one process per GPU with NO collective on the data path; only the
  per-query results (a few floats) are gathered at the end.
* Map training: data-parallel over views on a replicated map; one all-reduce (sum) per step of
  the per-Gaussian gradients, packed into a single flat fp32 bucket so that NCCL sees one large
  message over NVLink/NVSwitch instead of six small ones.

Everything here is torch.distributed host logic: `nccl` on the GPU box, `gloo` in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_queries(num_queries: int, rank: int | None = None, world_size: int | None = None) -> List[int]:
    """Round-robin split: query q goes to rank q % world.  Ranks differ by at most one query."""
    if rank is None or world_size is None:
        rank, world_size = world()
    return list(range(rank, num_queries, world_size))


def shard_views(num_views: int, step: int, rank: int | None = None, world_size: int | None = None) -> int:
    """View rendered by this rank at a given data-parallel training step (rank r takes views r::world)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    return (step * world_size + rank) % num_views


def gather_query_results(local: Dict[int, torch.Tensor], num_queries: int, width: int) -> torch.Tensor | None:
    """Collect per-query result vectors ([width] each, keyed by query id) on rank 0 as [num_queries, width].
    One small all_gather of a dense table; not on the per-iteration path."""
    rank, ws = world()
    dev = next(iter(local.values())).device if local else torch.device("cpu")
    table = torch.full((num_queries, width), float("nan"), dtype=torch.float64, device=dev)
    for q, v in local.items():
        table[q] = v.to(torch.float64)
    if ws == 1:
        return table
    parts = [torch.empty_like(table) for _ in range(ws)]
    dist.all_gather(parts, table)
    if rank != 0:
        return None
    out = parts[0]
    for p in parts[1:]:
        out = torch.where(torch.isnan(out), p, out)
    return out


class GradientBucket:
    """Flat fp32 bucket for the map-training gradient exchange.

    `params` are the per-Gaussian parameter tensors (xyz, f_dc, f_rest, opacity, scaling, rotation:
    59 floats per Gaussian at SH degree 3, gaussian_splatting/scene/gaussian_model.py:108-111).
    `allreduce()` packs their .grad into one contiguous buffer, sums it over ranks with a single
    collective and unpacks in place."""

    def __init__(self, params: Sequence[torch.Tensor]):
        self.params = list(params)
        self.sizes = [p.numel() for p in self.params]
        total = sum(self.sizes)
        p0 = self.params[0]
        self.flat = torch.zeros(total, dtype=torch.float32, device=p0.device)
        self.views = []
        off = 0
        for p, n in zip(self.params, self.sizes):
            self.views.append(self.flat[off:off + n].view_as(p))
            off += n

    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def allreduce(self, average: bool = False):
        rank, ws = world()
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        if ws > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if average:
                self.flat.div_(ws)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        return self.flat

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


class SparseGradientExchange:
    """Visible-rows-only gradient exchange for data-parallel map training.

    One view touches a few percent of the map (3 % at config C4), so the dense 59-float-per-Gaussian all-reduce
    moves 30x more bytes than the gradients that exist.  Here every rank packs the rows of its visible Gaussians
    (`radii > 0`; every other row of the rasterizer's gradients is exactly zero) into one [K, 1 + F] table — row
    index (as float bits) + the F gradient floats of all tensors — the tables are all-gathered at a common padded
    length and every rank adds all tables into its dense gradients.  Sums are identical to the dense all-reduce up
    to float addition order.  Two collectives: a MAX of the row counts (tiny) and the all-gather."""

    def __init__(self, like: Sequence[torch.Tensor], granularity: int = 4096):
        self.shapes = [tuple(t.shape[1:]) for t in like]
        self.widths = [int(torch.tensor(s).prod()) if len(s) else 1 for s in self.shapes]
        self.F = sum(self.widths)
        self.granularity = granularity
        self.last_rows = 0
        self.last_bytes = 0

    def _is_map_layout(self, grads) -> bool:
        """xyz [P,3], features [P,M,3], opacity [P(,1)], scaling [P,3], rotation [P,4], contiguous fp32: the fused kernels apply."""
        w = self.widths
        return (len(grads) == 5 and w[0] == 3 and w[1] % 3 == 0 and w[2] == 1 and w[3] == 3 and w[4] == 4 and
                all(g.is_contiguous() and g.dtype == torch.float32 for g in grads))

    def exchange(self, grads: Sequence[torch.Tensor], visible: torch.Tensor) -> List[torch.Tensor]:
        """grads: dense per-Gaussian gradient tensors [P, ...] (modified in place and returned summed over ranks);
        visible: bool [P] mask of the rows that may be non-zero on this rank."""
        rank, ws = world()
        if ws == 1:
            return list(grads)
        dev = grads[0].device
        idx = visible.nonzero(as_tuple=True)[0]
        n = torch.tensor([idx.numel()], dtype=torch.int64, device=dev)
        dist.all_reduce(n, op=dist.ReduceOp.MAX)
        K = (int(n.item()) + self.granularity - 1) // self.granularity * self.granularity
        K = max(K, self.granularity)
        fused = dev.type == "cuda" and self._is_map_layout(grads)
        k = idx.numel()
        if fused:
            import ctypes as C

            from . import _lib
            lib = _lib.load()
            M = self.widths[1] // 3
            stream = torch.cuda.current_stream(dev).cuda_stream
            tab = (C.c_void_p * 5)(*[g.data_ptr() for g in grads])
            table = torch.empty(K, 1 + self.F, dtype=torch.float32, device=dev)
            _lib.check(lib.gsr_pack_gradient_rows(idx.data_ptr(), k, K, M, tab, table.data_ptr(), stream), "gsr_pack_gradient_rows")
        else:
            table = torch.zeros(K, 1 + self.F, dtype=torch.float32, device=dev)
            table[:k, 0] = idx.to(torch.int32).view(torch.float32)              # row id, bit-cast
            table[k:, 0] = torch.tensor(-1, dtype=torch.int32).view(torch.float32)   # padding rows
            off = 1
            for g, w in zip(grads, self.widths):
                table[:k, off:off + w] = g.index_select(0, idx).reshape(k, w)
                off += w
        gathered = torch.empty(ws * K, 1 + self.F, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(gathered, table)
        self.last_rows, self.last_bytes = K, gathered.numel() * 4
        # add the other ranks' rows (ours are already in place)
        for r in range(ws):
            if r == rank:
                continue
            part = gathered[r * K:(r + 1) * K]
            if fused:
                _lib.check(lib.gsr_add_gradient_rows(part.data_ptr(), K, M, tab, stream), "gsr_add_gradient_rows")
                continue
            rows = part[:, 0].contiguous().view(torch.int32).to(torch.int64)
            keep = rows >= 0
            rows = rows[keep]
            off = 1
            for g, w in zip(grads, self.widths):
                g.view(g.shape[0], -1).index_add_(0, rows, part[keep, off:off + w])
                off += w
        return list(grads)

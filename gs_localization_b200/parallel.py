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


def _backend() -> str:
    return dist.get_backend() if dist.is_available() and dist.is_initialized() else ""


def _broadcast(t: torch.Tensor, src: int = 0):
    """dist.broadcast in place; with the gloo backend CUDA tensors go through the host."""
    if t.is_cuda and _backend() == "gloo":
        h = t.cpu()
        dist.broadcast(h, src)
        t.copy_(h)
    else:
        dist.broadcast(t, src)


def _all_reduce(t: torch.Tensor, op=None):
    """dist.all_reduce; with the gloo backend (CPU tests, two processes sharing one GPU) CUDA tensors go through the host."""
    op = dist.ReduceOp.SUM if op is None else op
    if t.is_cuda and _backend() == "gloo":
        h = t.cpu()
        dist.all_reduce(h, op=op)
        t.copy_(h)
    else:
        dist.all_reduce(t, op=op)
    return t


def _all_gather_into(out: torch.Tensor, t: torch.Tensor):
    if t.is_cuda and _backend() == "gloo":
        ho = torch.empty(out.shape, dtype=out.dtype)
        dist.all_gather_into_tensor(ho, t.cpu())
        out.copy_(ho)
    else:
        dist.all_gather_into_tensor(out, t)
    return out


def gradient_arena(grads: Sequence[torch.Tensor]):
    """The single allocation the rasterizer's backward carved its gradient tensors from (binding/torch_binding.cpp: one
    arena, 128-byte aligned slices, padding zero-filled), or None if the tensors do not share one."""
    base = None
    for g in grads:
        if g is None:
            continue
        b = g._base if g._base is not None else g
        if base is None:
            base = b
        elif b is not base:
            return None
    return base


def allreduce_gradients(grads: Sequence[torch.Tensor], average: bool = False) -> int:
    """Dense data-parallel exchange, in place and without a copy: the gradient tensors of one backward already live in
    ONE arena, so the all-reduce runs on the arena itself (the pack into / unpack from a separate flat bucket that
    GradientBucket does costs two extra passes over 59 floats per Gaussian).  Returns the bytes reduced."""
    rank, ws = world()
    arena = gradient_arena(grads)
    if arena is None:                      # tensors from different allocations (ctypes binding): one collective each
        n = 0
        for g in grads:
            if g is not None:
                if ws > 1:
                    _all_reduce(g)
                    if average:
                        g.div_(ws)
                n += g.numel() * 4
        return n
    if ws > 1:
        _all_reduce(arena)
        if average:
            arena.div_(ws)
    return arena.numel() * arena.element_size()


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
                _lib.check(lib.gsr_add_gradient_rows(part.data_ptr(), K, M, int(grads[0].shape[0]), tab, stream), "gsr_add_gradient_rows")
                continue
            rows = part[:, 0].contiguous().view(torch.int32).to(torch.int64)
            keep = rows >= 0
            rows = rows[keep]
            off = 1
            for g, w in zip(grads, self.widths):
                g.view(g.shape[0], -1).index_add_(0, rows, part[keep, off:off + w])
                off += w
        return list(grads)


class VisibleRowExchange:
    """Visible-rows gradient exchange that never makes the host wait (SparseGradientExchange pays two round trips per step:
    the row count for `nonzero`, and the MAX of the counts over ranks).

    The number of visible rows of a view is known as soon as its FORWARD has run (`radii`), a whole backward before the
    gradients exist.  `begin(radii)` — called right after the forward — counts them on the device, all-reduces the MAX
    over ranks on a side stream and copies it to pinned memory, all while the backward runs; by the time `exchange()` is
    called the count has long landed, so the table size is exact (no overflow case) and reading it costs no stall.
    Every rank then appends the rows of its visible Gaussians through a device counter into its table
    (gsr_pack_visible_rows), the tables are all-gathered and added on the device (gsr_add_counted_rows), which also
    accumulates the densification statistics of ALL views of the step, so that the replicas take identical
    densification decisions."""

    def __init__(self, P: int, M: int, device, granularity: int = 4096, peer_memory: bool | None = None):
        self.P, self.M, self.dev = int(P), int(M), torch.device(device)
        self.W = 1 + 11 + 3 * self.M + 3
        self.granularity = int(granularity)
        # Peer-memory pull (NVLink / NVSwitch): the tables live in symmetric memory and every rank's add kernel reads the
        # peers' tables in place — gather and add are ONE kernel per peer, no gathered copy, no NCCL call on the data path.
        # None = use it when the process group runs NCCL on CUDA devices and symmetric memory can be set up
        # (GSR_DP_PEER=0 forces the all-gather path).
        import os
        self.peer_memory = peer_memory if peer_memory is not None else os.environ.get("GSR_DP_PEER", "1") != "0"
        self._symm, self._symm_hdl, self._symm_floats = None, None, 0
        self.used_peer_memory = False
        self.count = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._side = torch.cuda.Stream(self.dev) if self.dev.type == "cuda" else None
        self._max_rows = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._pinned = torch.zeros(1, dtype=torch.int32).pin_memory() if self.dev.type == "cuda" else torch.zeros(1, dtype=torch.int32)
        self._ready = None
        self._buf = None
        self.last_rows, self.last_bytes = 0, 0

    def resize(self, P: int):
        """After densification / pruning the map has a different number of rows."""
        self.P = int(P)

    def begin(self, radii: torch.Tensor):
        """Start the (tiny) exchange of the row count; call right after the forward, before the backward is queued."""
        rank, ws = world()
        cur = torch.cuda.current_stream(self.dev)
        self._max_rows.copy_((radii > 0).sum(dtype=torch.int32).reshape(1))
        fork = torch.cuda.Event()
        fork.record(cur)
        with torch.cuda.stream(self._side):
            self._side.wait_event(fork)
            if ws > 1:
                _all_reduce(self._max_rows, dist.ReduceOp.MAX)
            self._pinned.copy_(self._max_rows, non_blocking=True)
            self._ready = torch.cuda.Event()
            self._ready.record(self._side)

    def max_rows(self, radii: torch.Tensor) -> int:
        """Largest visible-row count of this step over the ranks (waits for begin()'s copy, which has long landed)."""
        if self._ready is None:
            self.begin(radii)                       # caller did not overlap the count exchange: do it now
        self._ready.synchronize()
        return int(self._pinned[0])

    def exchange(self, grads: Sequence[torch.Tensor], radii: torch.Tensor, dL_dmeans2D: torch.Tensor | None = None, stats=None):
        """grads: (xyz [P,3], features [P,M,3], opacity [P,1], scaling [P,3], rotation [P,4]) dense fp32, summed over ranks in
        place.  stats = (max_radii2D [P], xyz_gradient_accum [P,1], denom [P,1]) are updated with every rank's view."""
        import ctypes as C

        from . import _lib
        lib = _lib.load()
        rank, ws = world()
        rows = self.max_rows(radii)
        self._ready = None
        g_ = self.granularity
        cap = max(1, min(self.P, (rows + g_ - 1) // g_ * g_))
        W, dev = self.W, self.dev
        stream = torch.cuda.current_stream(dev).cuda_stream
        tab = (C.c_void_p * 5)(*[g.data_ptr() for g in grads])
        # persistent buffers that only grow: a table whose size follows the view would otherwise ask the allocator for a new
        # block (a cudaMalloc, i.e. a device synchronisation) almost every step
        need = (cap + 1) * W
        sp = [None, None, None] if stats is None else [t.data_ptr() for t in stats]
        hdl = self._peer_tables(need) if (ws > 1 and self.peer_memory) else None
        self.used_peer_memory = hdl is not None
        if hdl is not None:
            # every rank packs into its own symmetric table; after the barrier each rank's add kernels read the peers'
            # tables over NVLink (the kernel takes a pointer: a peer's table is just another address)
            table = self._symm[:need]
            _lib.check(lib.gsr_pack_visible_rows(radii.data_ptr(), self.P, self.M, tab, None if dL_dmeans2D is None else dL_dmeans2D.data_ptr(),
                                                 table.data_ptr(), cap, self.count.data_ptr(), stream), "gsr_pack_visible_rows")
            hdl.barrier(channel=0)                      # all tables of the step are complete
            for k in range(ws):
                r = (rank + k) % ws                     # start with our own, then the peers in a rotated order (no hot spot)
                if r == rank and stats is None:
                    continue
                part = table if r == rank else hdl.get_buffer(r, (need,), torch.float32)
                _lib.check(lib.gsr_add_counted_rows(part.data_ptr(), cap, self.M, self.P, tab, int(r != rank), sp[0], sp[1], sp[2], stream),
                           "gsr_add_counted_rows")
            hdl.barrier(channel=1)                      # nobody refills its table before every peer has read it
            self.last_rows, self.last_bytes = cap, int(need * 4 * ws)
            return list(grads)
        if self._buf is None or self._buf.numel() < need * (ws + 1):
            self._buf = torch.empty(int(need * (ws + 1) * 1.25), dtype=torch.float32, device=dev)
        table = self._buf[:need]
        _lib.check(lib.gsr_pack_visible_rows(radii.data_ptr(), self.P, self.M, tab, None if dL_dmeans2D is None else dL_dmeans2D.data_ptr(),
                                             table.data_ptr(), cap, self.count.data_ptr(), stream), "gsr_pack_visible_rows")
        if ws > 1:
            gathered = self._buf[need:need * (ws + 1)]
            _all_gather_into(gathered, table)
        else:
            gathered = table
        for r in range(ws):
            part = gathered[r * (cap + 1) * W:(r + 1) * (cap + 1) * W]
            if r == rank and stats is None:
                continue
            _lib.check(lib.gsr_add_counted_rows(part.data_ptr(), cap, self.M, self.P, tab, int(r != rank), sp[0], sp[1], sp[2], stream),
                       "gsr_add_counted_rows")
        self.last_rows, self.last_bytes = cap, int(gathered.numel() * 4)
        return list(grads)

    def peer_path_expected(self) -> bool:
        """Will `exchange` pull over peer memory?  (Known for sure after the first exchange; before it, from the setup.)"""
        return bool(self.peer_memory) and self._symm_hdl is not False and self.dev.type == "cuda" and _backend() == "nccl"

    def _peer_tables(self, need: int):
        """Symmetric-memory table of at least `need` floats on every rank (torch.distributed._symmetric_memory: one
        allocation per rank, mapped into every peer).  `need` follows the all-reduced row count, so all ranks grow in the
        same step; growing is a collective (rendezvous), which is why the buffer only ever grows, by half again each time.
        Returns the handle, or None when symmetric memory is not available here (then the all-gather path runs)."""
        if self.dev.type != "cuda" or _backend() != "nccl":
            return None
        if self._symm_hdl is not None and self._symm_floats >= need:
            return self._symm_hdl
        if self._symm_hdl is False:
            return None
        try:
            import torch.distributed._symmetric_memory as symm
            floats = min(int(need * 1.5) + 1024, (self.P + 1) * self.W + 1024)
            floats = max(floats, need)
            buf = symm.empty(floats, dtype=torch.float32, device=self.dev)
            hdl = symm.rendezvous(buf, dist.group.WORLD)
        except Exception as ex:      # no peer access / unsupported build: every rank takes this branch together
            import warnings
            warnings.warn(f"symmetric memory unavailable ({ex!r}); visible-row exchange falls back to all_gather")
            self._symm_hdl = False
            return None
        self._symm, self._symm_hdl, self._symm_floats = buf, hdl, floats
        return hdl


class DataParallelTrainer:
    """Synchronous data-parallel map training over views (loop of gs/7scenes_gs_full_dslam.py:145-241):
    every rank renders its own view of the step on its replica of the map, the per-Gaussian gradients are summed over
    ranks — `mode="dense"`: one all-reduce on the gradient arena; `mode="sparse"`: VisibleRowExchange — together with the
    inputs of the densification statistics, and every rank applies the same fused optimiser / densification step, so
    the replicas stay identical (same torch.manual_seed on every rank for the split samples).  The all-reduce of step t
    cannot overlap the forward of step t+1: that forward reads the parameters the optimiser step of t writes."""

    # An all-reduce moves 2 (N-1)/N x the arena whatever the views see, and on NVSwitch it is reduced inside the switch
    # (measured 8 x B200: 708 MB in 1.74 ms, 712 GB/s bus; 2 x B200: 549 GB/s).  The visible-row exchange moves (N-1) x the
    # view's table into every GPU: by peer-memory pull at about 400 GB/s (8 x B200: 7 x 72 MB in 1.6 ms, 7 x 145 MB in
    # 2.85 ms; 2 x B200: 145 MB in 0.70 ms), by all-gather + add at about 0.4 of the all-reduce's rate (8 x B200: 7 x 145 MB
    # in 4.6 ms).  "auto" compares the two per step with the row count that is known before the exchange starts: sparse
    # wins while N x visible fraction is small (2 GPUs at C4: 0.5-0.7 ms vs 1.3 ms), dense once the views of a step cover
    # most of the map (8 GPUs at C4, largest view 19 %: 1.7 ms vs 2.9 ms).  tests/tools/exchange_peer_probe.py.
    ALLGATHER_RATE_VS_ALLREDUCE = 0.4
    PEER_PULL_RATE_VS_ALLREDUCE = 0.6

    def __init__(self, model, opt, mode: str = "auto", extent: float = 1.0):
        assert mode in ("dense", "sparse", "auto")
        self.model, self.opt, self.mode, self.extent = model, opt, mode, float(extent)
        self.exchange = None
        self.exchange_bytes = 0
        self.last_choice = None

    @classmethod
    def choose_mode(cls, ws: int, rows: int, row_floats: int, arena_floats: int, peer_pull: bool) -> str:
        """"sparse" or "dense" for one step: (N-1) tables of `rows` rows into every GPU against an all-reduce of the arena
        (2 (N-1)/N of its bytes on the wire), weighted by the measured rates of the two paths."""
        if ws <= 1:
            return "sparse"
        sparse_bytes = (ws - 1) * rows * row_floats * 4
        dense_bytes = 2.0 * (ws - 1) / ws * arena_floats * 4
        rate = cls.PEER_PULL_RATE_VS_ALLREDUCE if peer_pull else cls.ALLGATHER_RATE_VS_ALLREDUCE
        return "sparse" if sparse_bytes < dense_bytes * rate else "dense"

    def _sparse(self):
        m = self.model
        P, M = int(m._xyz.shape[0]), int(m._features.shape[1])
        if self.exchange is None:
            self.exchange = VisibleRowExchange(P, M, m.device)
        elif self.exchange.P != P:
            self.exchange.resize(P)
        return self.exchange

    def reduce(self, g, dL_dmeans2D, radii, want_stats: bool):
        """Sum the gradients over ranks and (if the step updates them) the densification statistics of all views."""
        m = self.model
        rank, ws = world()
        stats = (m.max_radii2D, m.xyz_gradient_accum, m.denom) if want_stats else None
        mode = self.mode
        if mode == "auto":
            ex = self._sparse()
            rows = ex.max_rows(radii)                  # exact, already on the host (exchanged under the backward)
            mode = self.choose_mode(ws, rows, ex.W, sum(t.numel() for t in g), ex.peer_path_expected())
        self.last_choice = mode
        if mode == "sparse":
            ex = self._sparse()
            ex.exchange(list(g), radii, dL_dmeans2D, stats)
            self.exchange_bytes = ex.last_bytes
        else:
            if self.exchange is not None:
                self.exchange._ready = None            # the row count of this step was not consumed by an exchange
            inc = rad = None
            if want_stats:
                # The per-view increments are taken BEFORE the all-reduce: dL_dmeans2D lives in the same arena as the
                # parameter gradients and is summed over ranks with them, and the statistic is the sum over views of the
                # per-view norms (gaussian_model.py:405-407), not the norm of the summed gradient.
                # No boolean indexing (it would make the host wait for the visible count): masked dense increments.
                vis = (radii > 0).to(torch.float32)
                inc = torch.stack([torch.linalg.vector_norm(dL_dmeans2D[:, :2], dim=-1) * vis, vis], dim=1)
                rad = radii.clamp_min(0).to(torch.float32)
            self.exchange_bytes = allreduce_gradients(g)
            if want_stats:
                if ws > 1:
                    _all_reduce(inc)
                    _all_reduce(rad, dist.ReduceOp.MAX)
                torch.maximum(m.max_radii2D, rad, out=m.max_radii2D)
                m.xyz_gradient_accum[:, 0] += inc[:, 0]
                m.denom[:, 0] += inc[:, 1]
        return g

    def after_forward(self, radii):
        """Hook for GaussianModel.compute_gradients: the sparse exchange starts its row-count collective under the backward."""
        if self.mode in ("sparse", "auto"):
            self._sparse().begin(radii)

    def structural_step(self, iteration: int) -> bool:
        """Does `apply_gradients` clone / split / prune or reset opacities at this iteration (gaussian_model.apply_gradients)?"""
        opt = self.opt
        if iteration >= opt.densify_until_iter:
            return False
        return (iteration > opt.densify_from_iter and iteration % opt.densification_interval == 0) or iteration % opt.opacity_reset_interval == 0

    def resync(self):
        """Rank 0's replica (parameters, Adam moments, densification statistics) to every rank.
        The dense all-reduce gives every rank the same bits, and so does the visible-row exchange at two ranks; with three
        or more ranks it adds the tables in a different order on each rank, so replicas differ in the last ulp — harmless for
        the optimiser, but densification compares per-Gaussian statistics with thresholds, and ONE Gaussian deciding
        differently on one rank changes P there.  Called before every structural step (a few GB once per
        densification interval), it makes those decisions identical by construction."""
        rank, ws = world()
        if ws <= 1:
            return
        m = self.model
        tensors = list(m._params()) + list(m._state["m"]) + list(m._state["v"]) + [m.max_radii2D, m.xyz_gradient_accum, m.denom]
        for t in tensors:
            _broadcast(t, 0)
        m._refresh_activations()

    def before_apply(self, iteration: int):
        """Call between `reduce` and `GaussianModel.apply_gradients`: re-synchronises the replicas when this iteration is a
        structural step and the ranks may differ in the last ulp (see `resync`)."""
        if self.structural_step(iteration) and world()[1] > 2 and self.last_choice == "sparse":
            self.resync()

    def step(self, cam, gt_image, bg, iteration: int, pseudo_depth=None, gt_depth=None):
        m, opt = self.model, self.opt
        loss, g, g2d, out = m.compute_gradients(cam, gt_image, bg, opt, iteration, pseudo_depth, gt_depth, after_forward=self.after_forward)
        want_stats = iteration < opt.densify_until_iter
        self.reduce(g, g2d, out["radii"], want_stats)
        self.before_apply(iteration)
        out["densify"] = m.apply_gradients(g, None, None, opt, iteration, self.extent, stats_done=True)
        return loss, out
